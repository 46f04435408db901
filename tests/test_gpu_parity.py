"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (via the Python mirror of the reference
interface), against the committed golden fixtures, the oracle on small seeded inputs, and size-independent
properties at the BASELINE sizes.  Stated fp32 tolerances: images 2e-5 absolute; per-voxel gradients
max-abs <= 2e-2 * ||g||_inf and relative L2 <= 5e-3 (see tests/test_golden.py for why)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def gold(name):
    return dict(np.load(os.path.join(GOLD, name + '.npz')))


@pytest.fixture(scope='module')
def sdfrt_mod(built):
    import sdfrt
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return sdfrt


def build_scene(sdfrt, kind, R, W, integ, albedo_res=None, warp='default', H=None):
    from oracle import sdf_oracle as O  # data + camera only; the renderer under test is the CUDA path
    H = W if H is None else H
    osc = O.make_scene(kind, res=R, width=W, height=H, albedo_grid_res=albedo_res)
    sdf = sdfrt.Grid3d(osc.sdf.data.float().numpy(), device=0)
    integrator = sdfrt.integrators.create(integ, sdf=sdf)
    if warp == 'default':
        wf = sdfrt.WarpField2D(sdf, weight_strategy=6, edge_eps=0.01)
        wf.clamping_thresh = 0.05
        integrator.warp_field = wf
    elif warp == 'dummy':
        integrator.warp_field = sdfrt.DummyWarpField(sdf)
    elif isinstance(warp, dict):
        wf = sdfrt.WarpField2D(sdf, weight_strategy=warp.get('weight_strategy', 6), edge_eps=warp.get('edge_eps', 0.01))
        wf.clamping_thresh = warp.get('clamping_thresh', 0.05)
        wf.max_reparam_depth = warp.get('max_reparam_depth', -1)
        wf.normalize_warp_field = warp.get('normalize_warp_field', True)
        integrator.warp_field = wf
    alb = None if albedo_res is None else osc.albedo_grid.float().cuda().contiguous()
    scene = sdfrt.Scene(integrator, [sdfrt.Sensor(osc.sensor.to_world, W, H)], albedo_grid=alb)
    return scene, osc


def test_library_is_the_cuda_build(sdfrt_mod):
    ctx = sdfrt_mod.Context(0)
    assert ctx.lib.sdfrt_version() == 100 and os.path.basename(sdfrt_mod._capi.LIB_PATH) == 'libsdfrt.so'


@pytest.mark.parametrize('name', ['sphere_direct', 'knot_direct', 'blobs_tex_direct', 'knot_silhouette', 'sphere_simple'])
def test_golden_image_gradient_and_rays(sdfrt_mod, name):
    from golden import make_golden as MG
    kind, R, W, spp, integ, ar = MG.CASES[name]
    g = gold(name)
    scene, _ = build_scene(sdfrt_mod, kind, R, W, integ, ar)
    if ar is not None:
        scene.albedo_grid.requires_grad_(True)
    sdf = scene.integrator().sdf
    sdf.data.requires_grad_(True)
    img = sdfrt_mod.render(scene, sensor=0, seed=MG.SEED, spp=spp, seed_grad=MG.SEED_GRAD, spp_grad=spp)
    assert np.abs(img.detach().cpu().numpy() - g['image']).max() < 2e-5
    (img * torch.from_numpy(MG.grad_image(W)).float().cuda()).sum().backward()
    gs = sdf.data.grad[..., 0].cpu().numpy()
    assert np.abs(gs - g['grad_sdf']).max() < 2e-2 * np.abs(g['grad_sdf']).max()
    assert np.linalg.norm(gs - g['grad_sdf']) < 5e-3 * np.linalg.norm(g['grad_sdf'])
    assert np.abs(sdf.p_grad.cpu().numpy() - g['grad_p']).max() < 2e-2 * max(1.0, np.abs(g['grad_p']).max())
    if ar is not None:
        ga = scene.albedo_grid.grad.cpu().numpy()
        assert np.abs(ga - g['grad_albedo']).max() < 1e-3 * np.abs(g['grad_albedo']).max()
    # the developed image of the adjoint re-render
    img_adj = sdf.ctx.develop(scene.integrator().last_film).cpu().numpy()
    assert np.abs(img_adj - g['image_adj']).max() < 2e-5
    # fixed rays through SDFBase.ray_intersect
    ray = sdfrt_mod.Ray3f(torch.from_numpy(g['ray_o']).cuda(), torch.from_numpy(g['ray_d']).cuda(), torch.from_numpy(g['ray_maxt']).cuda())
    its_t, warp_t, warp_t_d, ww, wwd = sdf.ray_intersect(ray, warp=scene.integrator().warp_field)
    its_t, warp_t = its_t.cpu().numpy(), warp_t.cpu().numpy()
    assert (np.isfinite(its_t) == np.isfinite(g['its_t'])).mean() > 0.99
    both = np.isfinite(its_t) & np.isfinite(g['its_t'])
    assert np.abs(its_t[both] - g['its_t'][both]).max() < 5e-5
    wf = np.isfinite(warp_t) & np.isfinite(g['warp_t'])
    assert np.quantile(np.abs(warp_t[wf] - g['warp_t'][wf]) / np.maximum(1, np.abs(g['warp_t'][wf])), 0.99) < 1e-3
    assert np.quantile(np.abs(ww.cpu().numpy()[wf] - g['warp_weight'][wf]), 0.99) < 1e-3


def test_eval_sdf_against_oracle(sdfrt_mod):
    from oracle import sdf_oracle as O
    scene, osc = build_scene(sdfrt_mod, 'blobs', 32, 8, 'sdf_direct_reparam')
    sdf = scene.integrator().sdf
    x = np.random.default_rng(0).uniform(-0.1, 1.1, size=(4096, 3))
    v, g, H = osc.sdf._taps(torch.from_numpy(x), 2)
    v2, _, g2, _, H2 = sdf.eval_all(x)
    assert (v2.cpu().double() - v).abs().max() < 2e-6
    assert (g2.cpu().double() - g).abs().max() < 2e-5
    assert (H2.cpu().double() - H).abs().max() < 1e-5 * max(1.0, H.abs().max().item())
    assert (sdf.eval(x).cpu().double() - v).abs().max() < 2e-6
    assert (sdf.eval_grad(x).cpu().double() - g).abs().max() < 2e-5


VARIANTS = [('dummy', 'sdf_direct_reparam'), (None, 'sdf_direct_reparam'),
            (dict(max_reparam_depth=0), 'sdf_direct_reparam'),                        # configs.WarpPrimary
            (dict(normalize_warp_field=False), 'sdf_direct_reparam'),                 # configs.WarpNotNormalized
            (dict(weight_strategy=4, edge_eps=0.05, clamping_thresh=0.0), 'sdf_silhouette_reparam')]  # WarpField2D ctor defaults


@pytest.mark.parametrize('warp,integ', VARIANTS)
def test_variants_against_oracle(sdfrt_mod, warp, integ):
    """Configurations of python/configs.py beyond the default `Warp`: DummyWarpField / None (shading gradients only),
    WarpPrimary, WarpNotNormalized, and the WarpField2D constructor defaults."""
    from oracle import sdf_oracle as O
    W, spp = 20, 4
    scene, osc = build_scene(sdfrt_mod, 'knot', 32, W, integ, warp=warp)
    if isinstance(warp, dict):
        owf = O.WarpField2D(osc.sdf, weight_strategy=warp.get('weight_strategy', 6), edge_eps=warp.get('edge_eps', 0.01))
        owf.clamping_thresh = warp.get('clamping_thresh', 0.05)
        owf.max_reparam_depth = warp.get('max_reparam_depth', -1)
        owf.normalize_warp_field = warp.get('normalize_warp_field', True)
    else:
        owf = O.DummyWarpField(osc.sdf) if warp == 'dummy' else None
    g_in = np.random.default_rng(1).uniform(-1, 1, size=(W, W, 3))
    img_o, g_o, _, _ = O.render_backward(osc, owf, integ, torch.from_numpy(g_in), seed=21, spp=spp)
    sdf = scene.integrator().sdf
    sdf.data.requires_grad_(True)
    img = sdfrt_mod.render(scene, sensor=0, seed=21, spp=spp, seed_grad=21, spp_grad=spp)
    assert (img.detach().cpu().double() - img_o).abs().max() < 2e-5
    (img * torch.from_numpy(g_in).float().cuda()).sum().backward()
    gs = sdf.data.grad[..., 0].cpu().double()
    assert (gs - g_o).abs().max() < 2e-2 * g_o.abs().max()
    assert (gs - g_o).norm() < 5e-3 * g_o.norm()


def test_non_square_film_and_non_pow2_spp(sdfrt_mod):
    from oracle import sdf_oracle as O
    scene, osc = build_scene(sdfrt_mod, 'sphere', 24, 28, 'sdf_direct_reparam', H=12)
    owf = O.WarpField2D(osc.sdf, weight_strategy=6, edge_eps=0.01)
    owf.clamping_thresh = 0.05
    with torch.no_grad():
        img_o = O.render(osc, owf, 'sdf_direct_reparam', seed=2, spp=3)
    img = scene.integrator().render(scene, sensor=0, seed=2, spp=3)
    assert img.shape == (12, 28, 3)
    assert (img.cpu().double() - img_o).abs().max() < 2e-5


def test_error_behaviour(sdfrt_mod):
    ctx = sdfrt_mod.Context(0)
    film = ctx.new_film()
    with pytest.raises(sdfrt_mod.SdfrtError):          # no SDF bound
        ctx.render_primal(0, 4, film)
    sdf = torch.from_numpy(sdfrt_mod.create_sphere_sdf((16, 16, 16), redistance=False)).cuda()
    ctx.set_sdf(sdf)
    ctx.set_sensor(np.eye(4), 16, 16)
    film = ctx.new_film()
    with pytest.raises(sdfrt_mod.SdfrtError):          # backprop before accumulate
        ctx.adjoint_backprop(film, torch.ones(16, 16, 3, device='cuda'), torch.zeros_like(sdf))
    with pytest.raises(sdfrt_mod.SdfrtError):
        ctx.set_integrator(7)
    with pytest.raises(sdfrt_mod.SdfrtError):          # python/integrators/reparam.py:48-50
        ctx.set_sensor(np.eye(4), 8192, 8192)
        ctx.render_primal(0, 4096, ctx.new_film())
    with pytest.raises(sdfrt_mod.SdfrtError):          # CPU tensors are rejected, never silently copied
        ctx.set_sdf(sdf.cpu())


def test_empty_and_inside_cases(sdfrt_mod):
    """Edge cases: an SDF with no surface (all positive), and a camera inside the bounding box."""
    ctx = sdfrt_mod.Context(0)
    ctx.set_sdf(torch.full((16, 16, 16), 0.5, device='cuda'))
    cam = sdfrt_mod.get_regular_cameras(6, resx=16, resy=16)[0]
    ctx.set_sensor(cam.to_world, 16, 16)
    ctx.set_integrator('sdf_direct_reparam')
    img = ctx.develop(ctx.render_primal(0, 4))
    assert torch.allclose(img, torch.ones_like(img), atol=1e-6)     # background radiance only
    g = torch.zeros(16, 16, 16, device='cuda')
    film = ctx.adjoint_accumulate(1, 4)
    ctx.adjoint_backprop(film, torch.ones(16, 16, 3, device='cuda'), g)
    assert torch.isfinite(g).all()
    ctx.set_sdf(torch.from_numpy(sdfrt_mod.create_sphere_sdf((16, 16, 16), redistance=False)).cuda())
    ctx.set_sensor(sdfrt_mod.look_at([0.5, 0.5, 0.02], [0.5, 0.5, 0.5], [0, 1, 0]), 16, 16)
    img = ctx.develop(ctx.render_primal(0, 4))
    assert torch.isfinite(img).all() and img[8, 8].sum() < 3.0       # the sphere is seen from inside the bbox


# ---- size-independent properties at the BASELINE sizes (256^3, 512x512, 64 spp adjoint) -----------------
@pytest.fixture(scope='module')
def big(sdfrt_mod):
    from sdfrt import synthetic as syn
    ctx = sdfrt_mod.Context(0)
    sdf = syn.make_sdf_device('blobs', 256, 'cuda')
    ctx.set_sdf(sdf)
    cam = sdfrt_mod.get_regular_cameras(12, resx=512, resy=512)[0]
    ctx.set_sensor(cam.to_world, 512, 512)
    ctx.set_integrator('sdf_direct_reparam')
    ctx.set_warp()
    return ctx, sdf


def test_full_size_partition_invariance_and_counts(big):
    ctx, sdf = big
    ctx.stats(reset=True)
    ctx.set_partition(0, 1)
    film = ctx.render_primal(5, 64).clone()
    st = ctx.stats(reset=True)
    assert st['n_rays'] == ctx.wavefront_size(64) == 516 * 516 * 64
    parts = torch.zeros_like(film)
    for r in range(4):
        ctx.set_partition(r, 4)
        parts += ctx.render_primal(5, 64)
    st4 = ctx.stats(reset=True)
    ctx.set_partition(0, 1)
    assert st4['n_rays'] == st['n_rays'] and st4['n_evals_value'] == st['n_evals_value'] and st4['n_hits'] == st['n_hits']
    assert (parts - film).abs().max() < 1e-3 * film.abs().max()       # atomics reorder the same addends
    w = film[2:-2, 2:-2, 3]
    assert (w > 0).all()                                               # every pixel received samples


def test_full_size_adjoint_linearity_and_host_api(big):
    ctx, sdf = big
    film = ctx.adjoint_accumulate(9, 64)
    rng = torch.Generator(device='cuda').manual_seed(0)
    a = torch.randn(512, 512, 3, device='cuda', generator=rng)
    b = torch.randn(512, 512, 3, device='cuda', generator=rng)
    ga, gb, gab = (torch.zeros_like(sdf) for _ in range(3))
    ctx.adjoint_backprop(film, a, ga)
    ctx.adjoint_backprop(film, b, gb)
    ctx.adjoint_backprop(film, (2 * a - 3 * b).contiguous(), gab)
    ref = 2 * ga - 3 * gb
    assert torch.isfinite(gab).all() and ga.abs().max() > 0
    assert (gab - ref).abs().max() < 1e-3 * ref.abs().max()
    # host-buffer entry point == device API (same seeds)
    ones = torch.ones(512, 512, 3, device='cuda')
    g_dev = torch.zeros_like(sdf)
    img_dev = ctx.develop(ctx.render_primal(3, 16))
    ctx.adjoint_backprop(ctx.adjoint_accumulate(4, 16), ones, g_dev)
    img_h, g_h = ctx.render_view_host(sdf.cpu().pin_memory(), 3, 16, 4, 16)
    ctx.set_sdf(sdf)
    assert (img_h.cuda() - img_dev).abs().max() < 1e-4
    assert (g_h.cuda() - g_dev).abs().max() < 1e-3 * g_dev.abs().max()


def test_gradient_matches_finite_differences_of_the_loss(sdfrt_mod):
    """The reference's own validation method (figures/result_utils.py:126-161: AD vs finite differences of a scene
    parameter), applied to the optimisation loss of python/shape_opt.py:78-83: for the uniform offset sdf.data += c (which
    shrinks / grows the object) dL/dc = sum of dL/d(sdf.data) must match the central difference of the primal loss.  The sum
    averages the Monte Carlo noise of the per-voxel gradient; on the oracle the two agree to 4 %."""
    sdfrt = sdfrt_mod
    R, W, spp = 48, 64, 64
    sensors = sdfrt.get_regular_cameras(6, resx=W, resy=W)

    def make(radius):
        sdf = sdfrt.Grid3d(sdfrt.create_sphere_sdf((R, R, R), radius=radius, redistance=False), device=0)
        integ = sdfrt.integrators.create('sdf_direct_reparam', sdf=sdf)
        wf = sdfrt.WarpField2D(sdf, weight_strategy=6, edge_eps=0.01)
        wf.clamping_thresh = 0.05
        integ.warp_field = wf
        return sdfrt.Scene(integ, sensors), sdf, integ

    tscene, _, _ = make(0.22)
    targets = [sdfrt.render(tscene, sensor=i, seed=100 + i, spp=256).clone() for i in range(3)]
    scene, sdf, integ = make(0.3)
    base = sdf.data.detach().clone()

    def loss_at(offset, spp_primal):
        with torch.no_grad():
            sdf.data.copy_(base + offset)
        integ.parameters_changed()
        with torch.no_grad():
            return sum(((sdfrt.render(scene, sensor=i, seed=7 + i, spp=spp_primal) - targets[i]).abs().mean() / 3).item() for i in range(3))

    sdf.data.requires_grad_(True)
    for i in range(3):
        img = sdfrt.render(scene, sensor=i, seed=7 + i, spp=256, seed_grad=20 + i, spp_grad=spp)
        ((img - targets[i]).abs().mean() / 3).backward()
    dL_dc = sdf.data.grad.sum().item()
    sdf.data.requires_grad_(False)
    delta = 0.01
    fd = (loss_at(delta, 1024) - loss_at(-delta, 1024)) / (2 * delta)
    assert fd < 0 and dL_dc < 0                       # shrinking the sphere towards the target lowers the loss
    assert abs(dL_dc - fd) < 0.15 * abs(fd), (dL_dc, fd)


def test_upsample_sdf_matches_oracle(sdfrt_mod):
    """variables.upsample_sdf (python/variables.py:18-23): cubic lookup at the cell centres of the x2 grid."""
    from oracle import sdf_oracle as O
    scene, osc = build_scene(sdfrt_mod, 'blobs', 16, 8, 'sdf_direct_reparam')
    up = sdfrt_mod.variables.upsample_sdf(scene.integrator().sdf)
    assert up.shape == (32, 32, 32, 1)
    c = (torch.arange(32, dtype=torch.float64) + 0.5) / 32
    Z, Y, X = torch.meshgrid(c, c, c, indexing='ij')
    ref = osc.sdf.eval(torch.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], -1)).reshape(32, 32, 32)
    assert (up[..., 0].cpu().double() - ref).abs().max() < 2e-6


def test_pipelined_view_equals_separate_sweeps(big):
    """sdfrt_render_view (primal + adjoint sweep 1 in one pipelined call) == sdfrt_render_primal + sdfrt_adjoint_accumulate."""
    ctx, sdf = big
    fp = ctx.render_primal(21, 32).clone()
    fa = ctx.adjoint_accumulate(22, 16).clone()
    g_sep = torch.zeros_like(sdf)
    gi = torch.ones(512, 512, 3, device='cuda')
    ctx.adjoint_backprop(fa, gi, g_sep)
    fp2, fa2 = ctx.render_view(21, 32, 22, 16)
    g_cmb = torch.zeros_like(sdf)
    ctx.adjoint_backprop(fa2, gi, g_cmb)
    assert (fp2 - fp).abs().max() < 1e-3 * fp.abs().max()
    assert (fa2 - fa).abs().max() < 1e-3 * fa.abs().max()
    assert (g_cmb - g_sep).abs().max() < 1e-3 * g_sep.abs().max()


def test_antithetic_sampling_against_oracle(sdfrt_mod):
    """prop antithetic_sampling through the wavefront pipeline: doubled work list, mirrored film positions, cloned samplers."""
    from oracle import sdf_oracle as O
    W, spp = 16, 4
    osc = O.make_scene('knot', res=32, width=W, height=W)
    sdf = sdfrt_mod.Grid3d(osc.sdf.data.float().numpy(), device=0)
    integ = sdfrt_mod.integrators.create('sdf_direct_reparam', dict(antithetic_sampling=True), sdf=sdf)
    wf = sdfrt_mod.WarpField2D(sdf, weight_strategy=6, edge_eps=0.01)
    wf.clamping_thresh = 0.05
    integ.warp_field = wf
    scene = sdfrt_mod.Scene(integ, [sdfrt_mod.Sensor(osc.sensor.to_world, W, W)])
    owf = O.WarpField2D(osc.sdf, weight_strategy=6, edge_eps=0.01)
    owf.clamping_thresh = 0.05
    with torch.no_grad():
        img_o = O.render(osc, owf, 'sdf_direct_reparam', seed=3, spp=spp, antithetic=True).numpy()
    sdf.data.requires_grad_(True)
    img = sdfrt_mod.render(scene, sensor=0, seed=3, spp=spp, seed_grad=11, spp_grad=spp)
    assert np.abs(img.detach().cpu().numpy() - img_o).max() < 2e-5
    g_in = np.random.default_rng(5).uniform(-1, 1, size=(W, W, 3))
    _, g_o, _, _ = O.render_backward(osc, owf, 'sdf_direct_reparam', torch.from_numpy(g_in), seed=11, spp=spp, antithetic=True)
    (img * torch.from_numpy(g_in).float().cuda()).sum().backward()
    gs, g_o = sdf.data.grad[..., 0].cpu().numpy(), g_o.numpy()
    assert np.abs(gs - g_o).max() < 2e-2 * np.abs(g_o).max() and np.linalg.norm(gs - g_o) < 5e-3 * np.linalg.norm(g_o)
    assert sdf.ctx.stats()['n_rays'] > 0
    integ.antithetic_sampling = False        # and the switch goes back off
    img_plain = integ.render(scene, sensor=0, seed=3, spp=spp)
    with torch.no_grad():
        assert np.abs(img_plain.cpu().numpy() - O.render(osc, owf, 'sdf_direct_reparam', seed=3, spp=spp).numpy()).max() < 2e-5


def test_sdf_to_world_transform_against_oracle(sdfrt_mod):
    """Grid3d(data, transform): general affine to_world through the wavefront pipeline (empty-space test off, world AABB,
    pulled-back gradients / Hessians, grid-space scatter) and through sdf_prb_reparam."""
    from oracle import sdf_oracle as O
    from test_hostmath_parity import _affine
    W, spp, M = 16, 4, _affine()
    osc = O.make_scene('knot', res=32, width=W, height=W)
    osc.sdf = O.Grid3d(osc.sdf.data, to_world=M)
    owf = O.WarpField2D(osc.sdf, weight_strategy=6, edge_eps=0.01)
    owf.clamping_thresh = 0.05
    sdf = sdfrt_mod.Grid3d(osc.sdf.data.float().numpy(), transform=M, device=0)
    assert sdf.has_transform and min(np.subtract(*sdf.bbox()[::-1])) > 1.1
    g_in = np.random.default_rng(5).uniform(-1, 1, size=(W, W, 3))
    for name in ('sdf_direct_reparam', 'sdf_prb_reparam'):
        integ = sdfrt_mod.integrators.create(name, dict(max_depth=3), sdf=sdf)
        wf = sdfrt_mod.WarpField2D(sdf, weight_strategy=6, edge_eps=0.01)
        wf.clamping_thresh = 0.05
        integ.warp_field = wf
        scene = sdfrt_mod.Scene(integ, [sdfrt_mod.Sensor(osc.sensor.to_world, W, W)])
        sdf.data.requires_grad_(True)
        sdf.data.grad = None
        img = sdfrt_mod.render(scene, sensor=0, seed=3, spp=spp, seed_grad=11, spp_grad=spp)
        (img * torch.from_numpy(g_in).float().cuda()).sum().backward()
        if name == 'sdf_prb_reparam':
            cfg = O.PrbConfig(max_depth=3)
            img_o = O.render_prb(osc, owf, seed=3, spp=spp, cfg=cfg).numpy()
            g_o = O.render_backward_prb(osc, owf, torch.from_numpy(g_in), seed=11, spp=spp, cfg=cfg)[1].numpy()
        else:
            with torch.no_grad():
                img_o = O.render(osc, owf, name, seed=3, spp=spp).numpy()
            g_o = O.render_backward(osc, owf, name, torch.from_numpy(g_in), seed=11, spp=spp)[1].numpy()
        assert np.abs(img.detach().cpu().numpy() - img_o).max() < 2e-5, name
        gs = sdf.data.grad[..., 0].cpu().numpy()
        assert np.abs(gs - g_o).max() < 2e-2 * np.abs(g_o).max() and np.linalg.norm(gs - g_o) < 5e-3 * np.linalg.norm(g_o), name
    x = np.random.default_rng(0).uniform(-0.1, 1.1, size=(256, 3))
    v, g, H = osc.sdf._taps(torch.from_numpy(x), 2)
    v2, _, g2, _, H2 = sdf.eval_all(x)
    assert (v2.cpu().double() - v).abs().max() < 2e-6 and (g2.cpu().double() - g).abs().max() < 2e-5
    assert (H2.cpu().double() - H).abs().max() < 1e-5 * max(1.0, H.abs().max().item())


@pytest.mark.parametrize('prop', ['detach_indirect_si', 'decouple_reparam'])
def test_direct_integrator_shadow_origin_variants(sdfrt_mod, prop):
    from oracle import sdf_oracle as O
    W, spp = 16, 4
    osc = O.make_scene('knot', res=32, width=W, height=W)
    setattr(osc, prop, True)
    owf = O.WarpField2D(osc.sdf, weight_strategy=6, edge_eps=0.01)
    owf.clamping_thresh = 0.05
    sdf = sdfrt_mod.Grid3d(osc.sdf.data.float().numpy(), device=0)
    integ = sdfrt_mod.integrators.create('sdf_direct_reparam', {prop: True}, sdf=sdf)
    wf = sdfrt_mod.WarpField2D(sdf, weight_strategy=6, edge_eps=0.01)
    wf.clamping_thresh = 0.05
    integ.warp_field = wf
    scene = sdfrt_mod.Scene(integ, [sdfrt_mod.Sensor(osc.sensor.to_world, W, W)])
    g_in = np.random.default_rng(5).uniform(-1, 1, size=(W, W, 3))
    g_o = O.render_backward(osc, owf, 'sdf_direct_reparam', torch.from_numpy(g_in), seed=11, spp=spp)[1].numpy()
    integ.render_backward(scene, None, torch.from_numpy(g_in).float().cuda(), sensor=0, seed=11, spp=spp)
    gs = sdf.data.grad[..., 0].cpu().numpy()
    assert np.abs(gs - g_o).max() < 2e-2 * np.abs(g_o).max() and np.linalg.norm(gs - g_o) < 5e-3 * np.linalg.norm(g_o)


def test_render_op_holds_several_views_and_reflectance_only_gradients(sdfrt_mod):
    """`render()` is the mi.render replacement (python/shape_opt.py:78-83).  (i) A loss summed over five views rendered before one
    backward: the context keeps the lane records of a few views (view slots), the op re-runs sweep 1 of a view whose slot was
    reused (sdfrt_adjoint_token) -- the gradient equals the sum of the per-view gradients.  (ii) Texture-only
    optimisation (SDF frozen): the reflectance grid is a differentiable input of the op and gets its gradient."""
    sdfrt = sdfrt_mod
    from oracle import sdf_oracle as O
    W, spp = 24, 8
    osc = O.make_scene('blobs', res=32, width=W, height=W, albedo_grid_res=8)
    sdf = sdfrt.Grid3d(osc.sdf.data.float().numpy(), device=0)
    integ = sdfrt.integrators.create('sdf_direct_reparam', sdf=sdf)
    wf = sdfrt.WarpField2D(sdf, weight_strategy=6, edge_eps=0.01)
    wf.clamping_thresh = 0.05
    integ.warp_field = wf
    cams = sdfrt.get_regular_cameras(6, resx=W, resy=W)
    alb = osc.albedo_grid.float().cuda().contiguous()
    scene = sdfrt.Scene(integ, cams[:5], albedo_grid=alb)
    ws = [torch.from_numpy(np.random.default_rng(i).uniform(-1, 1, (W, W, 3))).float().cuda() for i in range(5)]
    w0 = ws[0]
    assert sdf.ctx.num_view_slots() < 5   # five views before one backward: more than there are view slots -> one sweep is re-run

    def grads(order):
        sdf.data.grad = None
        alb.grad = None
        sdf.data.requires_grad_(True)
        alb.requires_grad_(True)
        if order == 'per_view':
            for i, w in enumerate(ws):
                (sdfrt.render(scene, sensor=i, seed=3 + i, spp=spp, seed_grad=11 + i, spp_grad=spp) * w).sum().backward()
        else:
            imgs = [sdfrt.render(scene, sensor=i, seed=3 + i, spp=spp, seed_grad=11 + i, spp_grad=spp) for i in range(5)]
            sum((im * w).sum() for im, w in zip(imgs, ws)).backward()
        return sdf.data.grad.clone(), alb.grad.clone()

    g1, a1 = grads('per_view')
    g2, a2 = grads('summed')
    assert g1.abs().max() > 0 and (g1 - g2).abs().max() < 1e-4 * g1.abs().max()       # atomics order only
    assert (a1 - a2).abs().max() < 1e-4 * a1.abs().max()
    # (ii) SDF frozen, reflectance grid optimised
    sdf.data.grad = None
    alb.grad = None
    sdf.data.requires_grad_(False)
    alb.requires_grad_(True)
    (sdfrt.render(scene, sensor=0, seed=3, spp=spp, seed_grad=11, spp_grad=spp) * w0).sum().backward()
    only = alb.grad.clone()
    assert sdf.data.grad is None
    sdf.data.requires_grad_(True)
    alb.grad = None
    (sdfrt.render(scene, sensor=0, seed=3, spp=spp, seed_grad=11, spp_grad=spp) * w0).sum().backward()
    assert only.abs().max() > 0 and (only - alb.grad).abs().max() < 1e-4 * only.abs().max()


@pytest.mark.parametrize('env', [{'SDFRT_ROWS8': '0'}, {'SDFRT_BOX': '10'}, {'SDFRT_BOX': '10', 'SDFRT_ROWS8': '0'}], ids=['rows4', 'tma_box', 'tma_box_rows4'])
def test_marcher_variants_give_the_same_films_and_gradients(sdfrt_mod, env):
    """Developer variants of k_march read the SAME taps by other routes -- the 8x row-expanded copy with 256-bit loads (SDFRT_ROWS8)
    and TMA-staged shared-memory boxes for the shadow segments (SDFRT_BOX) -- so films, gradients and evaluation counts must not
    change (films/gradients up to the order of the atomics)."""
    import baseline_cases as B
    sdfrt = sdfrt_mod
    kind, R, AR, W, spp, mult, ns = B.CONFIGS['c2']
    sdf = torch.from_numpy(B.grid(kind, R)).cuda()
    cam = sdfrt.get_regular_cameras(ns, resx=W, resy=W)[0]
    G = torch.ones((W, W, 3), device='cuda')

    def run(e):
        old = {k: os.environ.get(k) for k in e}
        os.environ.update(e)
        try:
            ctx = sdfrt.Context(0)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        ctx.set_sdf(sdf)
        ctx.set_sensor(cam.to_world, W, W)
        ctx.set_shading()
        ctx.set_integrator('sdf_direct_reparam')
        ctx.set_warp()
        ctx.stats(reset=True)
        fp, fa = ctx.render_view(5, spp * mult, 18, spp)
        g = torch.zeros_like(sdf)
        ctx.adjoint_backprop(fa, G, g)
        return fp.clone(), fa.clone(), g, ctx.stats(reset=True)

    base = {'SDFRT_DEFER': '0'}     # (the staged boxes and the deferred differentiable march exclude each other: compare like with like)
    fp0, fa0, g0, st0 = run(base)
    fp1, fa1, g1, st1 = run({**base, **env})
    for k in ('n_evals_value', 'n_evals_hess', 'n_evals_grad', 'n_hits', 'n_scatter_sites'):   # (n_tap_loads depends on which segment a slot held before)
        assert st0[k] == st1[k], (k, st0[k], st1[k])
    assert (fp0 - fp1).abs().max() < 1e-5 * fp0.abs().max() and (fa0 - fa1).abs().max() < 1e-5 * fa0.abs().max()
    assert (g0 - g1).abs().max() < 1e-4 * g0.abs().max()
    if 'SDFRT_BOX' in env:
        assert st1['n_box_loads'] > 0.05 * st1['n_tap_loads'], st1
    else:
        assert st1['n_box_loads'] == 0


@pytest.mark.parametrize('kind,albedo_res', [('sphere', None), ('knot', 8)])
def test_use_mis_against_oracle(sdfrt_mod, kind, albedo_res):
    """`use_mis=True` (python/integrators/reparam.py:17; sdf_direct_reparam.py:76-79,87-107) through the wavefront pipeline: second
    region-compacted queue of BSDF-sampled rays, the weight channel splatted once per sample, nine more record rows."""
    from oracle import sdf_oracle as O
    sdfrt = sdfrt_mod
    W, spp = 20, 4
    osc = O.make_scene(kind, res=32, width=W, height=W, albedo_grid_res=albedo_res)
    osc.use_mis = True
    owf = O.WarpField2D(osc.sdf, weight_strategy=6, edge_eps=0.01)
    owf.clamping_thresh = 0.05
    sdf = sdfrt.Grid3d(osc.sdf.data.float().numpy(), device=0)
    integ = sdfrt.integrators.create('sdf_direct_reparam', {'use_mis': True}, sdf)
    wf = sdfrt.WarpField2D(sdf, weight_strategy=6, edge_eps=0.01)
    wf.clamping_thresh = 0.05
    integ.warp_field = wf
    alb = None if albedo_res is None else osc.albedo_grid.float().cuda().contiguous().requires_grad_(True)
    scene = sdfrt.Scene(integ, [sdfrt.Sensor(osc.sensor.to_world, W, W)], albedo_grid=alb)
    g_in = np.random.default_rng(7).uniform(-1, 1, size=(W, W, 3))
    img_o, g_o, _, ga_o = O.render_backward(osc, owf, 'sdf_direct_reparam', torch.from_numpy(g_in), seed=11, spp=spp)
    with torch.no_grad():
        img_p = O.render(osc, owf, 'sdf_direct_reparam', seed=5, spp=spp)
    sdf.data.requires_grad_(True)
    img = sdfrt.render(scene, sensor=0, seed=5, spp=spp, seed_grad=11, spp_grad=spp)
    assert (img.detach().cpu().double() - img_p).abs().max() < 2e-5
    (img * torch.from_numpy(g_in).float().cuda()).sum().backward()
    img_adj = sdf.ctx.develop(integ.last_film).cpu().double()
    assert (img_adj - img_o).abs().max() < 2e-5
    gs = sdf.data.grad[..., 0].cpu().double()
    assert (gs - g_o).abs().max() < 2e-2 * g_o.abs().max() and (gs - g_o).norm() < 5e-3 * g_o.norm()
    if alb is not None:
        assert (alb.grad.cpu().double() - ga_o).abs().max() < 1e-3 * ga_o.abs().max()


@pytest.mark.parametrize('use_mis,albedo_res', [(False, None), (True, 8)])
def test_envmap_emitter_against_oracle(sdfrt_mod, use_mis, albedo_res):
    """`envmap` emitter (oracle/ref_semantics.md S18) through the wavefront pipeline: escaped camera rays, importance-sampled emitter
    directions (emitter_val / ds.pdf) and the MIS branch's escaped BSDF rays look the bitmap up with the attached directions
    (python/integrators/sdf_direct_reparam.py:31,62-70,105).  High-contrast bitmap: dLe/dd is a few percent of the gradient."""
    from oracle import sdf_oracle as O
    sdfrt = sdfrt_mod
    W, spp = 20, 4
    osc = O.make_scene('knot', res=32, width=W, height=W, albedo_grid_res=albedo_res)
    bm = np.random.default_rng(1).uniform(0.05, 4.0, size=(24, 48, 3)).astype(np.float32)
    R = None
    if use_mis:                       # the second case also rotates the emitter (<transform name="to_world"> of the scene file)
        R, rr = np.linalg.qr(np.random.default_rng(5).normal(size=(3, 3)))
        R = R * np.sign(np.diag(rr))
        if np.linalg.det(R) < 0:
            R[:, 0] = -R[:, 0]
    osc.envmap = O.EnvMap(bm.astype(np.float64), to_world=R)
    osc.use_mis = use_mis
    owf = O.WarpField2D(osc.sdf, weight_strategy=6, edge_eps=0.01)
    owf.clamping_thresh = 0.05
    sdf = sdfrt.Grid3d(osc.sdf.data.float().numpy(), device=0)
    integ = sdfrt.integrators.create('sdf_direct_reparam', {'use_mis': use_mis}, sdf)
    wf = sdfrt.WarpField2D(sdf, weight_strategy=6, edge_eps=0.01)
    wf.clamping_thresh = 0.05
    integ.warp_field = wf
    alb = None if albedo_res is None else osc.albedo_grid.float().cuda().contiguous().requires_grad_(True)
    scene = sdfrt.Scene(integ, [sdfrt.Sensor(osc.sensor.to_world, W, W)], albedo_grid=alb)
    scene.set_envmap(bm, to_world=R)
    g_in = np.random.default_rng(7).uniform(-1, 1, size=(W, W, 3))
    img_o, g_o, _, ga_o = O.render_backward(osc, owf, 'sdf_direct_reparam', torch.from_numpy(g_in), seed=11, spp=spp)
    with torch.no_grad():
        img_p = O.render(osc, owf, 'sdf_direct_reparam', seed=5, spp=spp)
    try:
        sdf.data.requires_grad_(True)
        img = sdfrt.render(scene, sensor=0, seed=5, spp=spp, seed_grad=11, spp_grad=spp)
        err = (img.detach().cpu().double() - img_p).abs()
        # a sample on a CDF boundary may pick the neighbouring cell in fp32: a few pixels may differ visibly
        assert err.flatten().quantile(0.99) < 1e-4 and err.max() < 0.2, (err.flatten().quantile(0.99), err.max())
        (img * torch.from_numpy(g_in).float().cuda()).sum().backward()
        gs = sdf.data.grad[..., 0].cpu().double()
        assert (gs - g_o).abs().max() < 2e-2 * g_o.abs().max() and (gs - g_o).norm() < 5e-3 * g_o.norm()
        if alb is not None:
            assert (alb.grad.cpu().double() - ga_o).abs().max() < 2e-3 * ga_o.abs().max()
        # the path-replay integrator has the constant emitter only: it must refuse rather than render something else
        prb = sdfrt.integrators.create('sdf_prb_reparam', {}, sdf)
        prb.warp_field = wf
        scene2 = sdfrt.Scene(prb, scene.sensors())
        scene2.set_envmap(bm, to_world=R)
        with pytest.raises(RuntimeError, match='envmap'):
            sdfrt.render(scene2, sensor=0, seed=1, spp=1)
    finally:
        sdf.ctx.set_envmap(None)
        sdf.ctx._envmap_key = None


@pytest.mark.parametrize('integ_name,use_mis,n_aux', [('sdf_silhouette_reparam', False, 4), ('sdf_direct_reparam', False, 4), ('sdf_direct_reparam', True, 2)])
def test_convolution_warp_against_oracle(sdfrt_mod, integ_name, use_mis, n_aux):
    """WarpFieldConvolution (python/warp.py:126-176, python/warp_conv.py; oracle/ref_semantics.md S19) through `k_backward<CONV>`:
    the images are those of a render without warp field, the gradient is the closed-form adjoint of the reference's CustomOp
    `backward` with the auxiliary rays marched inside the adjoint sweep.  sdf_prb_reparam refuses the warp field."""
    from oracle import sdf_oracle as O
    sdfrt = sdfrt_mod
    W, spp = 16, 4
    osc = O.make_scene('knot', res=24, width=W, height=W)
    osc.use_mis = use_mis
    owf = O.WarpFieldConvolution(osc.sdf, n_aux_rays=n_aux)
    sdf = sdfrt.Grid3d(osc.sdf.data.float().numpy(), device=0)
    integ = sdfrt.integrators.create(integ_name, {'use_mis': use_mis}, sdf)
    integ.warp_field = sdfrt.WarpFieldConvolution(sdf, n_aux_rays=n_aux)
    scene = sdfrt.Scene(integ, [sdfrt.Sensor(osc.sensor.to_world, W, W)])
    g_in = np.random.default_rng(7).uniform(-1, 1, size=(W, W, 3))
    img_o, g_o, gp_o, _ = O.render_backward(osc, owf, integ_name, torch.from_numpy(g_in), seed=11, spp=spp)
    sdf.data.requires_grad_(True)
    img = sdfrt.render(scene, sensor=0, seed=11, spp=spp, seed_grad=11, spp_grad=spp)
    assert (img.detach().cpu().double() - img_o).abs().max() < 2e-5
    (img * torch.from_numpy(g_in).float().cuda()).sum().backward()
    gs = sdf.data.grad[..., 0].cpu().double()
    assert (gs - g_o).norm() < 2e-3 * g_o.norm() and (gs - g_o).abs().max() < 2e-2 * g_o.abs().max()
    assert (sdf.p_grad.cpu().double() - gp_o).abs().max() < 2e-3 * gp_o.abs().max()
    prb = sdfrt.integrators.create('sdf_prb_reparam', {}, sdf)
    prb.warp_field = integ.warp_field
    with pytest.raises(NotImplementedError):
        sdfrt.render(sdfrt.Scene(prb, scene.sensors()), sensor=0, seed=1, spp=1)


def test_deferred_differentiable_shadow_march_changes_nothing(sdfrt_mod):
    """Adjoint sweep 1 marches shadow segments with the value tracer first and only re-marches differentiably those whose |phi|
    stopped growing somewhere (the others provably end with an invalid warp record, python/shapes.py:195-201,278-283).
    SDFRT_DEFER=0 marches every shadow segment differentiably as the reference does: same films, same gradients, fewer Hessian
    evaluations -- with and without the BSDF-sampling branch."""
    import baseline_cases as B
    sdfrt = sdfrt_mod
    kind, R, AR, W, spp, mult, ns = B.CONFIGS['c2']
    sdf = torch.from_numpy(B.grid(kind, R)).cuda()
    cam = sdfrt.get_regular_cameras(ns, resx=W, resy=W)[0]
    G = torch.from_numpy(np.random.default_rng(2).uniform(-1, 1, (W, W, 3))).float().cuda()

    def run(defer, mis, integrator='sdf_direct_reparam', spp=spp):
        old = os.environ.get('SDFRT_DEFER')
        os.environ['SDFRT_DEFER'] = '1' if defer else '0'
        try:
            ctx = sdfrt.Context(0)
        finally:
            if old is None:
                os.environ.pop('SDFRT_DEFER', None)
            else:
                os.environ['SDFRT_DEFER'] = old
        ctx.set_sdf(sdf)
        ctx.set_sensor(cam.to_world, W, W)
        ctx.set_shading()
        ctx.set_integrator(integrator)
        ctx.set_use_mis(mis)
        ctx.set_warp()
        ctx.set_prb(3, 5)
        ctx.stats(reset=True)
        fa = ctx.adjoint_accumulate(18, spp)
        g = torch.zeros_like(sdf)
        ctx.adjoint_backprop(fa, G, g)
        return fa.clone(), g, ctx.stats(reset=True)

    for mis in (False, True):
        f0, g0, s0 = run(False, mis)
        f1, g1, s1 = run(True, mis)
        assert (f0 - f1).abs().max() < 1e-5 * f0.abs().max()
        assert g0.abs().max() > 0 and (g0 - g1).abs().max() < 1e-4 * g0.abs().max(), (mis, ((g0 - g1).abs().max() / g0.abs().max()).item())
        assert s1['n_evals_hess'] < 0.8 * s0['n_evals_hess'], (s0, s1)
        assert abs(s1['n_scatter_sites'] - s0['n_scatter_sites']) <= 1e-3 * s0['n_scatter_sites'], (s0, s1)
    # sdf_prb_reparam (wavefront form): shadow AND continuation segments of every bounce take the same route -- value march with the
    # approach flag, then only hits (continuation) and flagged misses go through the differentiable marcher (k_prb_flag)
    f0, g0, s0 = run(False, False, 'sdf_prb_reparam', 4)
    f1, g1, s1 = run(True, False, 'sdf_prb_reparam', 4)
    assert (f0 - f1).abs().max() < 2e-5 * f0.abs().max()
    assert g0.abs().max() > 0 and (g0 - g1).norm() < 2e-4 * g0.norm(), ((g0 - g1).norm() / g0.norm()).item()
    assert s1['n_evals_hess'] < 0.8 * s0['n_evals_hess'], (s0, s1)
    assert abs(s1['n_scatter_sites'] - s0['n_scatter_sites']) <= 2e-3 * s0['n_scatter_sites'], (s0, s1)
