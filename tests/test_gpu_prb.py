"""GPU parity tests (-m gpu) of the path replay integrator `sdf_prb_reparam` (csrc/sdfrt_prb.cuh; wavefront form by default,
`SDFRT_PRB_WAVEFRONT=0` runs the same tests on the per-lane walker), called through the C ABI via the Python mirror: golden fixtures (fp64 oracle), the oracle on a
seeded case, world-size independence of the row partition and a size-independent property at a BASELINE-sized film.
Stated fp32 tolerances as in test_gpu_parity.py."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def sdfrt_mod(built):
    import sdfrt
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return sdfrt


def build_scene(sdfrt, kind, R, W, albedo_res=None, max_depth=4, rr_depth=5, mrd=-1, warp=True):
    from oracle import sdf_oracle as O  # data + camera only; the renderer under test is the CUDA path
    osc = O.make_scene(kind, res=R, width=W, height=W, albedo_grid_res=albedo_res)
    sdf = sdfrt.Grid3d(osc.sdf.data.float().numpy(), device=0)
    integrator = sdfrt.integrators.create('sdf_prb_reparam', dict(max_depth=max_depth, rr_depth=rr_depth), sdf=sdf)
    if warp:
        wf = sdfrt.WarpField2D(sdf, weight_strategy=6, edge_eps=0.01)
        wf.clamping_thresh = 0.05
        wf.max_reparam_depth = mrd
        integrator.warp_field = wf
    alb = None if albedo_res is None else osc.albedo_grid.float().cuda().contiguous()
    return sdfrt.Scene(integrator, [sdfrt.Sensor(osc.sensor.to_world, W, W)], albedo_grid=alb), osc


def check_grad(gs, want):
    assert np.abs(gs - want).max() < 2e-2 * np.abs(want).max()
    assert np.linalg.norm(gs - want) < 5e-3 * np.linalg.norm(want)


@pytest.mark.parametrize('name', ['blobs_tex_prb', 'knot_prb_rr', 'blobs_prb_primary'])
def test_golden_image_and_gradient(sdfrt_mod, name):
    from golden import make_golden as MG
    kind, R, W, spp, ar, max_depth, rr_depth, mrd = MG.PRB_CASES[name]
    g = dict(np.load(os.path.join(GOLD, name + '.npz')))
    scene, _ = build_scene(sdfrt_mod, kind, R, W, ar, max_depth, rr_depth, mrd)
    if ar is not None:
        scene.albedo_grid.requires_grad_(True)
    sdf = scene.integrator().sdf
    sdf.data.requires_grad_(True)
    img = sdfrt_mod.render(scene, sensor=0, seed=MG.SEED, spp=spp, seed_grad=MG.SEED_GRAD, spp_grad=spp)
    assert np.abs(img.detach().cpu().numpy() - g['image']).max() < 2e-5
    (img * torch.from_numpy(MG.grad_image(W)).float().cuda()).sum().backward()
    check_grad(sdf.data.grad[..., 0].cpu().numpy(), g['grad_sdf'])
    assert np.abs(sdf.p_grad.cpu().numpy() - g['grad_p']).max() < 2e-2 * max(1.0, np.abs(g['grad_p']).max())
    if ar is not None:
        ga = scene.albedo_grid.grad.cpu().numpy()
        assert np.abs(ga - g['grad_albedo']).max() < 1e-3 * np.abs(g['grad_albedo']).max()
    img_adj = sdf.ctx.develop(scene.integrator().last_film).cpu().numpy()
    assert np.abs(img_adj - g['image_adj']).max() < 2e-5


def test_against_oracle_without_warp_and_separate_calls(sdfrt_mod):
    """No warp field (shading gradients only), through render() / render_backward() instead of the fused view call."""
    from oracle import sdf_oracle as O
    W, spp = 16, 4
    scene, osc = build_scene(sdfrt_mod, 'blobs', 24, W, None, 3, 5, -1, warp=False)
    integ = scene.integrator()
    cfg = O.PrbConfig(max_depth=3, rr_depth=5)
    img = integ.render(scene, sensor=0, seed=3, spp=spp)
    assert np.abs(img.cpu().numpy() - O.render_prb(osc, None, seed=3, spp=spp, cfg=cfg).numpy()).max() < 2e-5
    g_in = torch.from_numpy(np.random.default_rng(5).uniform(-1, 1, size=(W, W, 3)))
    _, g_o, _, _ = O.render_backward_prb(osc, None, g_in, seed=11, spp=spp, cfg=cfg)
    integ.render_backward(scene, None, g_in.float().cuda(), sensor=0, seed=11, spp=spp)
    check_grad(integ.sdf.data.grad[..., 0].cpu().numpy(), g_o.numpy())


def test_dummy_warp_field_is_rejected(sdfrt_mod):
    scene, _ = build_scene(sdfrt_mod, 'sphere', 16, 8, warp=False)
    scene.integrator().warp_field = sdfrt_mod.DummyWarpField(scene.integrator().sdf)
    with pytest.raises(AttributeError):
        scene.integrator().render(scene, sensor=0, seed=1, spp=1)


def test_partition_is_world_size_independent(sdfrt_mod):
    """Two contexts rendering the two halves of the interleaved band partition on one GPU = one context rendering all."""
    scene, _ = build_scene(sdfrt_mod, 'blobs', 32, 24, None, 3)
    integ = scene.integrator()
    ctx = integ._bind(scene, 0)
    g_in = torch.from_numpy(np.random.default_rng(2).uniform(-1, 1, size=(24, 24, 3))).float().cuda()
    film, film_a = ctx.render_view(3, 4, 11, 4)
    g1 = torch.zeros(32, 32, 32, device='cuda')
    ctx.adjoint_backprop(film_a, g_in, g1)
    fsum, fasum, g2 = torch.zeros_like(film), torch.zeros_like(film_a), torch.zeros_like(g1)
    parts = []
    for r in range(2):
        ctx.set_partition(r, 2)
        f, fa = ctx.render_view(3, 4, 11, 4)
        fsum += f
        fasum += fa
        parts.append(r)
        # the adjoint of I = V / W needs the complete accumulators: emulate the all-reduce by
        # running both ranks' sweep 1 first, then each rank's backprop with the summed film
    for r in parts:
        ctx.set_partition(r, 2)
        ctx.render_view(3, 4, 11, 4)            # restores this rank's per-sample L records
        ctx.adjoint_backprop(fasum, g_in, g2)
    ctx.set_partition(0, 1)
    assert (fsum - film).abs().max() < 1e-4 and (fasum - film_a).abs().max() < 1e-4
    assert (g2 - g1).abs().max() < 1e-3 * g1.abs().max()


def test_convex_object_expectation_at_baseline_film_size(sdfrt_mod):
    """Size-independent property at 512x512 (BASELINE film): a convex diffuse object under a constant environment has
    expected radiance rho * Le at every hit, so the image is 1 - 0.5 * coverage; and PRB with max_depth = 1 is the
    complement of the silhouette image bit for bit in expectation (same sampler streams)."""
    W, spp = 512, 16
    scene, _ = build_scene(sdfrt_mod, 'sphere', 64, W, None, 4, warp=False)
    img = scene.integrator().render(scene, sensor=0, seed=1, spp=spp)
    sdf = scene.integrator().sdf
    sil_integ = sdfrt_mod.integrators.create('sdf_silhouette_reparam', sdf=sdf)
    sil = sil_integ.render(sdfrt_mod.Scene(sil_integ, scene.sensors()), sensor=0, seed=1, spp=spp)
    want = 1.0 - 0.5 * sil
    inside = sil[..., 0] > 0.999                 # pixels fully covered by the object
    assert inside.sum() > 10000
    assert abs(float(img[inside].mean()) - 0.5) < 5e-3
    assert float((img - want).abs().mean()) < 0.02
    assert torch.isfinite(img).all()


@pytest.mark.parametrize('env', [{'SDFRT_SITEQ': '0'}, {'SDFRT_SITEQ_CAP': '64'}])
def test_site_queue_changes_nothing(sdfrt_mod, env):
    """The replay appends its B-spline sites to a queue that k_scatter_sites drains through the grouped transposed scatter; sites that do
    not fit are scattered where they arise.  Queue off, and a queue far too small (almost everything takes the overflow path), give the
    gradient of the default configuration."""
    from golden import make_golden as MG
    kind, R, W, spp, ar, max_depth, rr_depth, mrd = MG.PRB_CASES['blobs_tex_prb']

    def grad(extra_env):
        old = {k: os.environ.get(k) for k in extra_env}
        os.environ.update(extra_env)
        try:
            scene, _ = build_scene(sdfrt_mod, kind, R, W, ar, max_depth, rr_depth, mrd)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        sdf = scene.integrator().sdf
        sdf.data.requires_grad_(True)
        img = sdfrt_mod.render(scene, sensor=0, seed=MG.SEED, spp=spp, seed_grad=MG.SEED_GRAD, spp_grad=spp)
        (img * torch.from_numpy(MG.grad_image(W)).float().cuda()).sum().backward()
        return sdf.data.grad[..., 0].cpu().double(), sdf.ctx.stats(reset=True)

    g0, s0 = grad({})
    g1, s1 = grad(env)
    assert g0.abs().max() > 0 and (g0 - g1).norm() < 1e-5 * g0.norm()          # (atomics order only)
    assert s0['n_scatter_sites'] == s1['n_scatter_sites'] > 256
