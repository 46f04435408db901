"""The per-lane math of the CUDA kernels (csrc/sdfrt_core.cuh, sdfrt_lane.cuh) compiled for the host and
compared with the oracle.  CPU only; the same comparisons run through the C ABI on the GPU in
test_gpu_parity.py.  Tolerances: the kernels compute in fp32, the oracle in fp64."""
import numpy as np
import pytest
import torch

import hostmath as HM
from oracle import sdf_oracle as O

INTEG = {'sdf_silhouette_reparam': 0, 'sdf_simple_shading_reparam': 1, 'sdf_direct_reparam': 2}


def make_pair(kind='sphere', R=32, W=24, integ='sdf_direct_reparam', albedo_res=None, warp=None, p=(0, 0, 0)):
    sc = O.make_scene(kind, res=R, width=W, height=W, albedo_grid_res=albedo_res)
    sc.sdf.p = torch.tensor(p, dtype=torch.float64)
    hs = HM.HostScene(sc.sdf.data.numpy(), sc.sensor.to_world, W, W, integrator=INTEG[integ],
                      albedo_grid=None if albedo_res is None else sc.albedo_grid.numpy(), p=p)
    warp = dict(weight_strategy=6, edge_eps=0.01, clamping_thresh=0.05, normalize=True, max_depth=-1, kind=2) | (warp or {})
    if warp['kind'] == 2:
        wf = O.WarpField2D(sc.sdf, weight_strategy=warp['weight_strategy'], edge_eps=warp['edge_eps'])
        wf.clamping_thresh = warp['clamping_thresh']
        wf.normalize_warp_field = warp['normalize']
        wf.max_reparam_depth = warp['max_depth']
    elif warp['kind'] == 3:
        wf = O.WarpFieldConvolution(sc.sdf, n_aux_rays=warp.get('n_aux_rays', 4))
        wf.kappa, wf.power = warp.get('kappa', 1e5), warp.get('power', 3.0)
        wf.max_reparam_depth = warp['max_depth']
        hs.h.warp.n_aux_rays, hs.h.warp.kappa, hs.h.warp.power = wf.n_aux_rays, wf.kappa, wf.power
    elif warp['kind'] == 1:
        wf = O.DummyWarpField(sc.sdf)
    else:
        wf = None
    hs.h.warp.kind = warp['kind']
    hs.h.warp.edge_eps = wf.edge_eps if warp['kind'] == 2 else 0.0
    hs.h.warp.weight_strategy = warp['weight_strategy']
    hs.h.warp.clamping_thresh = warp['clamping_thresh']
    hs.h.warp.normalize_warp_field = int(warp['normalize'])
    hs.h.warp.max_reparam_depth = warp['max_depth']
    return sc, wf, hs


def test_eval_value_grad_hessian():
    sc, _, hs = make_pair('blobs')
    x = np.random.default_rng(0).uniform(-0.1, 1.1, size=(2000, 3))
    v, g, H = sc.sdf._taps(torch.from_numpy(x), 2)
    out = hs.eval_sdf(x, 2)
    hs_scale = H.abs().max().item()
    assert np.abs(out[:, 0] - v.numpy()).max() < 2e-6
    assert np.abs(out[:, 1:4] - g.numpy()).max() < 2e-5
    Hh = np.stack([out[:, 4], out[:, 7], out[:, 8], out[:, 7], out[:, 5], out[:, 9], out[:, 8], out[:, 9], out[:, 6]], -1).reshape(-1, 3, 3)
    assert np.abs(Hh - H.numpy()).max() < 1e-5 * max(1.0, hs_scale)
    assert np.abs(hs.eval_sdf(x, 0)[:, 0] - v.numpy()).max() < 2e-6
    assert np.abs(hs.eval_sdf(x, 1)[:, 1:4] - g.numpy()).max() < 2e-5


def _camera_rays(sc, n, seed):
    rng = np.random.default_rng(seed)
    pos = torch.from_numpy(rng.uniform(-2, sc.sensor.width + 2, size=(n, 2)))
    return sc.sensor.sample_ray(pos, torch.float64)


@pytest.mark.parametrize('kind', ['sphere', 'knot'])
def test_tracer_outputs_per_ray(kind):
    sc, _, hs = make_pair(kind)
    o, d, maxt = _camera_rays(sc, 600, 1)
    act = torch.ones(600, dtype=torch.bool)
    its_t, warp_t, warp_t_d, ww, wwd = sc.sdf.ray_intersect(o, d, maxt, act)
    h = hs.trace(o.numpy(), d.numpy(), maxt.numpy(), differentiable=True)
    fin = torch.isfinite(its_t).numpy()
    assert (np.isfinite(h['its_t']) == fin).mean() > 0.995          # borderline rays may flip in fp32
    both = fin & np.isfinite(h['its_t'])
    assert np.abs(h['its_t'][both] - its_t.numpy()[both]).max() < 5e-5
    wfin = torch.isfinite(warp_t).numpy() & np.isfinite(h['warp_t'])
    assert (np.isfinite(h['warp_t']) == torch.isfinite(warp_t).numpy()).mean() > 0.99
    # warp_t etc. are ratios of sums of 1/denom^3 weights: compare where well conditioned, relative to scale
    rel = np.abs(h['warp_t'][wfin] - warp_t.numpy()[wfin]) / np.maximum(1.0, np.abs(warp_t.numpy()[wfin]))
    assert np.quantile(rel, 0.99) < 1e-3
    dsc = np.maximum(1.0, np.abs(warp_t_d.numpy()[wfin]).max(-1))
    reld = np.abs(h['warp_t_d'][wfin] - warp_t_d.numpy()[wfin]).max(-1) / dsc
    assert np.quantile(reld, 0.98) < 2e-2
    assert np.quantile(np.abs(h['warp_weight'][wfin] - ww.numpy()[wfin]), 0.99) < 1e-3
    # value-only tracer agrees with the differentiable one on its_t
    h0 = hs.trace(o.numpy(), d.numpy(), maxt.numpy(), differentiable=False)
    assert np.abs(h0['its_t'][both] - h['its_t'][both]).max() < 2e-5


CASES = [
    ('sphere', 'sdf_silhouette_reparam', None, None),
    ('sphere', 'sdf_simple_shading_reparam', None, None),
    ('sphere', 'sdf_direct_reparam', None, None),
    ('knot', 'sdf_direct_reparam', None, None),
    ('blobs', 'sdf_direct_reparam', 8, None),                                   # reflectance grid
    ('knot', 'sdf_direct_reparam', None, dict(max_depth=0)),                     # configs.WarpPrimary
    ('knot', 'sdf_direct_reparam', None, dict(normalize=False)),                 # configs.WarpNotNormalized
    ('sphere', 'sdf_direct_reparam', None, dict(kind=1)),                        # DummyWarpField: shading gradients only
    ('knot', 'sdf_silhouette_reparam', None, dict(weight_strategy=4, edge_eps=0.05, clamping_thresh=0.0)),  # WarpField2D ctor defaults
]


@pytest.mark.parametrize('kind,integ,albedo_res,warp', CASES)
def test_image_and_gradient_parity(kind, integ, albedo_res, warp):
    W, spp = 24, 8
    sc, wf, hs = make_pair(kind, 32, W, integ, albedo_res, warp)
    with torch.no_grad():
        img_o = O.render(sc, wf, integ, seed=3, spp=spp).numpy()
    film, _, _ = hs.forward(3, spp, False)
    assert np.abs(hs.develop(film) - img_o).max() < 2e-5
    g_in = np.random.default_rng(5).uniform(-1, 1, size=(W, W, 3))
    img2, g_o, gp_o, ga_o = O.render_backward(sc, wf, integ, torch.from_numpy(g_in), seed=11, spp=spp)
    film, rec, _ = hs.forward(11, spp, True)
    assert np.abs(hs.develop(film) - img2.numpy()).max() < 2e-5
    g_h, ga_h, gp_h, st = hs.backward(11, spp, film, rec, g_in)
    g_o = g_o.numpy()
    scale = np.abs(g_o).max()
    assert scale > 0
    # fp32 kernels vs fp64 oracle.  With clamping_thresh=0 the 1/warp_t factor makes single rays ill-conditioned
    # (the fp32 oracle itself is then 8e-3 away from the fp64 one), hence the looser bound for that case.
    assert np.abs(g_h - g_o).max() < 2e-2 * scale
    assert np.linalg.norm(g_h - g_o) < 5e-3 * np.linalg.norm(g_o)
    cos = (g_h * g_o).sum() / np.linalg.norm(g_h) / np.linalg.norm(g_o)
    assert cos > 0.9999
    assert np.abs(gp_h - gp_o.numpy()).max() < 1e-2 * max(1.0, np.abs(gp_o.numpy()).max())
    if albedo_res is not None:
        ga_o = ga_o.numpy()
        assert np.abs(ga_h - ga_o).max() < 1e-3 * np.abs(ga_o).max()


def test_translated_sdf_p():
    sc, wf, hs = make_pair('sphere', 32, 16, 'sdf_direct_reparam', p=(0.03, -0.02, 0.01))
    with torch.no_grad():
        img_o = O.render(sc, wf, 'sdf_direct_reparam', seed=3, spp=4).numpy()
    film, _, _ = hs.forward(3, 4, False)
    assert np.abs(hs.develop(film) - img_o).max() < 2e-5


def test_adjoint_is_linear_in_grad_image():
    _, _, hs = make_pair('knot', 32, 16, 'sdf_direct_reparam')
    film, rec, _ = hs.forward(5, 4, True)
    rng = np.random.default_rng(0)
    a, b = rng.normal(size=(16, 16, 3)), rng.normal(size=(16, 16, 3))
    ga = hs.backward(5, 4, film, rec, a)[0]
    gb = hs.backward(5, 4, film, rec, b)[0]
    gab = hs.backward(5, 4, film, rec, 2 * a - 3 * b)[0]
    assert np.abs(gab - (2 * ga - 3 * gb)).max() < 1e-4 * max(np.abs(ga).max(), np.abs(gb).max())


@pytest.mark.parametrize('kind,R,W', [('sphere', 64, 48), ('knot', 64, 64), ('blobs', 96, 40)])
def test_empty_space_test_is_conservative(kind, R, W):
    """csrc pixel_provably_clear: no sample ray of a pixel declared clear ever sees an SDF value below its threshold (so
    skipping its march cannot change its_t, and its warp weight is exactly zero)."""
    import ctypes as C
    sc, _, hs = make_pair(kind, R, W)
    for skip_eps in (0.0, 0.01):
        out = np.zeros(3)
        HM.lib().hh_check_empty_space.restype = C.c_int
        v = HM.lib().hh_check_empty_space(C.byref(hs.h), 3, 8, C.c_float(skip_eps), HM.fptr(out))
        assert v == 0
        assert out[0] > 0.1 * out[1], out    # and the test is not vacuous: a good share of the rays is skipped


def test_antithetic_sampling_parity():
    """antithetic_sampling (python/integrators/reparam.py:19,167-178): every lane evaluated twice, the second sample at the
    mirrored film position with a clone of the sampler -- image, gradient and the rank partition of the doubled work list."""
    W, spp = 16, 4
    sc, wf, hs = make_pair('knot', 32, W, 'sdf_direct_reparam')
    hs.h.antithetic = 1
    assert hs.local_lanes(spp) == 2 * hs.wavefront(spp)
    with torch.no_grad():
        img_o = O.render(sc, wf, 'sdf_direct_reparam', seed=3, spp=spp, antithetic=True).numpy()
        img_plain = O.render(sc, wf, 'sdf_direct_reparam', seed=3, spp=spp).numpy()
    film, _, _ = hs.forward(3, spp, False)
    assert np.abs(hs.develop(film) - img_o).max() < 2e-5 and np.abs(img_o - img_plain).max() > 1e-2
    g_in = np.random.default_rng(5).uniform(-1, 1, size=(W, W, 3))
    img2, g_o, _, _ = O.render_backward(sc, wf, 'sdf_direct_reparam', torch.from_numpy(g_in), seed=11, spp=spp, antithetic=True)
    film, rec, _ = hs.forward(11, spp, True)
    assert np.abs(hs.develop(film) - img2.numpy()).max() < 2e-5
    g_h = hs.backward(11, spp, film, rec, g_in)[0]
    g_o = g_o.numpy()
    assert np.abs(g_h - g_o).max() < 2e-2 * np.abs(g_o).max() and np.linalg.norm(g_h - g_o) < 5e-3 * np.linalg.norm(g_o)
    whole = hs.forward(3, spp, False)[0]
    for world in (2, 3):
        parts = sum(hs.forward(3, spp, False, r, world)[0] for r in range(world))
        assert np.abs(parts - whole).max() < 2e-5


def _affine():
    """A general affine sdf_to_world: two rotations, anisotropic scale, shift about the cube centre."""
    def rot(ax, a):
        c, s = np.cos(a), np.sin(a)
        R = np.eye(3)
        i, j = [(1, 2), (0, 2), (0, 1)][ax]
        R[i, i] = R[j, j] = c
        R[i, j], R[j, i] = -s, s
        return R
    M = np.eye(4)
    L = rot(1, 0.4) @ rot(0, -0.3) @ np.diag([1.1, 0.9, 1.2])
    M[:3, :3] = L
    M[:3, 3] = np.array([0.5, 0.5, 0.5]) - L @ np.array([0.5, 0.5, 0.5]) + np.array([0.03, -0.02, 0.04])
    return M


@pytest.mark.parametrize('kind,integ', [('knot', 'sdf_direct_reparam'), ('blobs', 'sdf_silhouette_reparam')])
def test_sdf_to_world_transform_parity(kind, integ):
    """Grid3d(data, transform) / prop sdf_to_world (python/shapes.py:378-448): lookups through to_local, gradient / Hessian
    pulled back with to_local^T, bounding box = AABB of the transformed cube, scatter in grid space."""
    W, spp, M = 16, 4, _affine()
    sc = O.make_scene(kind, res=32, width=W, height=W)
    sc.sdf = O.Grid3d(sc.sdf.data, to_world=M)
    hs = HM.HostScene(sc.sdf.data.numpy(), sc.sensor.to_world, W, W, integrator=INTEG[integ], sdf_to_world=M)
    hs.h.warp.kind, hs.h.warp.edge_eps, hs.h.warp.weight_strategy, hs.h.warp.clamping_thresh = 2, 0.01, 6, 0.05
    hs.h.warp.normalize_warp_field, hs.h.warp.max_reparam_depth = 1, -1
    wf = O.WarpField2D(sc.sdf, weight_strategy=6, edge_eps=0.01)
    wf.clamping_thresh = 0.05
    x = np.random.default_rng(0).uniform(-0.1, 1.1, size=(500, 3))
    v, g, H = sc.sdf._taps(torch.from_numpy(x), 2)
    out = hs.eval_sdf(x, 2)
    assert np.abs(out[:, 0] - v.numpy()).max() < 2e-6 and np.abs(out[:, 1:4] - g.numpy()).max() < 2e-5
    Hh = np.stack([out[:, 4], out[:, 7], out[:, 8], out[:, 7], out[:, 5], out[:, 9], out[:, 8], out[:, 9], out[:, 6]], -1).reshape(-1, 3, 3)
    assert np.abs(Hh - H.numpy()).max() < 1e-5 * max(1.0, H.abs().max().item())
    lo, hi = sc.sdf.bbox()
    assert (hi - lo).min() > 1.1          # the box really changed (rotated cube: AABB larger than 1.1)
    with torch.no_grad():
        img_o = O.render(sc, wf, integ, seed=3, spp=spp).numpy()
    film, _, _ = hs.forward(3, spp, False)
    assert np.abs(hs.develop(film) - img_o).max() < 2e-5
    g_in = np.random.default_rng(5).uniform(-1, 1, size=(W, W, 3))
    img2, g_o, gp_o, _ = O.render_backward(sc, wf, integ, torch.from_numpy(g_in), seed=11, spp=spp)
    film, rec, _ = hs.forward(11, spp, True)
    assert np.abs(hs.develop(film) - img2.numpy()).max() < 2e-5
    g_h, _, gp_h, _ = hs.backward(11, spp, film, rec, g_in)
    g_o = g_o.numpy()
    assert np.abs(g_h - g_o).max() < 2e-2 * np.abs(g_o).max() and np.linalg.norm(g_h - g_o) < 5e-3 * np.linalg.norm(g_o)
    assert np.abs(gp_h - gp_o.numpy()).max() < 1e-2 * max(1.0, np.abs(gp_o.numpy()).max())


@pytest.mark.parametrize('mode', [1, 2], ids=['detach_indirect_si', 'decouple_reparam'])
def test_direct_integrator_shadow_origin_variants(mode):
    """props detach_indirect_si / decouple_reparam (python/integrators/sdf_direct_reparam.py:13-14,44-49): the interaction the
    shadow ray is spawned from.  Same image, different gradient (the variants differ by 20-80 % of ||g||_inf on this scene)."""
    W, spp = 16, 4
    sc, wf, hs = make_pair('knot', 32, W, 'sdf_direct_reparam')
    g_in = np.random.default_rng(5).uniform(-1, 1, size=(W, W, 3))
    g_default = O.render_backward(sc, wf, 'sdf_direct_reparam', torch.from_numpy(g_in), seed=11, spp=spp)[1].numpy()
    sc.detach_indirect_si, sc.decouple_reparam = mode == 1, mode == 2
    hs.h.shadow_origin = mode
    g_o = O.render_backward(sc, wf, 'sdf_direct_reparam', torch.from_numpy(g_in), seed=11, spp=spp)[1].numpy()
    film, rec, _ = hs.forward(11, spp, True)
    g_h = hs.backward(11, spp, film, rec, g_in)[0]
    assert np.abs(g_o - g_default).max() > 0.1 * np.abs(g_default).max()
    assert np.abs(g_h - g_o).max() < 2e-2 * np.abs(g_o).max() and np.linalg.norm(g_h - g_o) < 5e-3 * np.linalg.norm(g_o)


def test_evaluation_counts_match_the_oracle():
    """the device counters are the basis of the algorithmic-bytes figure; they must equal the oracle's count of
    64-tap evaluations for the same seeds (that equality is the proof the kernel did the work).  Primal silhouette render
    without a warp field: value-only marches + refinement, nothing else -- equal up to the handful of rays whose step count
    differs between the fp32 kernel and the fp64 oracle.  (With a WarpField2D the reference's silhouette integrator evaluates the
    value-free warp even in primal mode; the kernels only evaluate it in the adjoint -- ref_semantics.md.)  Direct integrator: the kernel skips the shadow march of samples whose BSDF gate is closed
    (oracle/ref_semantics.md, "deliberate evaluation-count deviations"), so it may only be lower, and not by much."""
    W, spp = 24, 4
    sc, wf, hs = make_pair('knot', 32, W, 'sdf_silhouette_reparam')
    sc.sdf.n_evals = 0
    with torch.no_grad():
        O.render(sc, None, 'sdf_silhouette_reparam', seed=3, spp=spp)
    n_oracle = sc.sdf.n_evals
    _, _, st = hs.forward(3, spp, False)
    n_kernel = int(st[1] + st[2] + st[3])
    assert n_oracle > 10000 and abs(n_kernel - n_oracle) <= 2e-3 * n_oracle, (n_kernel, n_oracle)
    sc, wf, hs = make_pair('knot', 32, W, 'sdf_direct_reparam')
    sc.sdf.n_evals = 0
    with torch.no_grad():
        O.render(sc, wf, 'sdf_direct_reparam', seed=3, spp=spp)
    n_oracle = sc.sdf.n_evals
    _, _, st = hs.forward(3, spp, False)
    n_kernel = int(st[1] + st[2] + st[3])
    # oracle: primary march + refine, 2 lookups per hit (compute_surface_interaction), every shadow ray marched
    assert 0.75 * n_oracle < n_kernel <= n_oracle, (n_kernel, n_oracle)


@pytest.mark.parametrize('kind,albedo_res', [('sphere', None), ('knot', 8)])
def test_use_mis_image_and_gradient_parity(kind, albedo_res):
    """ReparamIntegrator prop `use_mis` (python/integrators/reparam.py:17) in sdf_direct_reparam
    (python/integrators/sdf_direct_reparam.py:76-79,87-107): power-heuristic weight on the emitter sample plus a BSDF-sampled ray
    whose escape sees the environment emitter; its reparameterisation (depth 1) enters through the divergence only."""
    W, spp = 20, 4
    sc, wf, hs = make_pair(kind, 32, W, 'sdf_direct_reparam', albedo_res=albedo_res)
    sc.use_mis = True
    hs.h.use_mis = 1
    g_in = np.random.default_rng(7).uniform(-1, 1, size=(W, W, 3))
    with torch.no_grad():
        img_o = O.render(sc, wf, 'sdf_direct_reparam', seed=5, spp=spp).numpy()
    film, _, _ = hs.forward(5, spp, False)
    assert np.abs(hs.develop(film) - img_o).max() < 2e-5
    # the MIS image differs from the emitter-sampling-only image (same seeds) but estimates the same integral
    sc.use_mis = False
    with torch.no_grad():
        img_nee = O.render(sc, wf, 'sdf_direct_reparam', seed=5, spp=spp).numpy()
    sc.use_mis = True
    assert np.abs(img_o - img_nee).max() > 1e-3
    img_ob, g_o, gp_o, ga_o = O.render_backward(sc, wf, 'sdf_direct_reparam', torch.from_numpy(g_in), seed=11, spp=spp)
    film, rec, _ = hs.forward(11, spp, True)
    assert np.abs(hs.develop(film) - img_ob.numpy()).max() < 2e-5
    g_h, ga_h, gp_h, _ = hs.backward(11, spp, film, rec, g_in)
    g_o = g_o.numpy()
    assert np.abs(g_h - g_o).max() < 2e-2 * np.abs(g_o).max() and np.linalg.norm(g_h - g_o) < 5e-3 * np.linalg.norm(g_o)
    if albedo_res is not None:
        assert np.abs(ga_h - ga_o.numpy()).max() < 1e-3 * np.abs(ga_o.numpy()).max()


def _envmap_bitmap(h=8, w=16, seed=0):
    """Smooth, strictly positive lat-long bitmap with a bright lobe (so importance sampling matters)."""
    y, x = np.meshgrid((np.arange(h) + 0.5) / h, np.arange(w) / w, indexing='ij')
    lobe = np.exp(-((x - 0.3) ** 2 / 0.02 + (y - 0.35) ** 2 / 0.03))
    base = 0.4 + 0.3 * np.sin(2 * np.pi * x)[..., None] * np.array([1.0, 0.5, -0.5]) + 0.2 * y[..., None]
    return (base + 4.0 * lobe[..., None] * np.array([1.0, 0.9, 0.7])).astype(np.float32)


def _rotation(seed=5):
    """A general rotation matrix (to_world of an envmap emitter)."""
    q, r = np.linalg.qr(np.random.default_rng(seed).normal(size=(3, 3)))
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


@pytest.mark.parametrize('rotated', [False, True])
def test_envmap_lookup_sampling_pdf(rotated):
    """`envmap` emitter (oracle/ref_semantics.md S18): lookup, sampled directions and pdf of the kernel math against the oracle,
    without and with a to_world rotation of the emitter."""
    bm = _envmap_bitmap()
    R = _rotation() if rotated else None
    env = O.EnvMap(bm.astype(np.float64), to_world=R)
    _, _, hs = make_pair('sphere', 16, 8)
    hs.set_envmap(bm, to_world=R)
    rng = np.random.default_rng(3)
    d = rng.normal(size=(4000, 3))
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    u = rng.uniform(0, 1, size=(4000, 2))
    got = hs.envmap_query(d, u)
    dd = torch.from_numpy(d.astype(np.float32).astype(np.float64))
    assert np.abs(got['Le'] - env.eval(dd).numpy()).max() < 2e-4          # cell coordinates in fp32: ~1e-5 of a cell
    pdf_o = env.pdf(dd).numpy()      # cell probability / sin(theta): near the (rotated) poles the fp32 direction limits sin(theta)
    assert np.quantile(np.abs(got['pdf'] - pdf_o), 0.995) < 1e-4 * pdf_o.max() and (np.abs(got['pdf'] - pdf_o) < 5e-3 * pdf_o + 1e-6).all()
    # closed-form dLe_c/dd of the kernel math (what the adjoint pushes into the warp sites) against autograd through the oracle's lookup
    dg = dd[:500].clone().requires_grad_(True)
    Le_o = env.eval(dg)
    for c in range(3):
        jac = torch.autograd.grad(Le_o[:, c].sum(), dg, retain_graph=True)[0].numpy()
        scale = max(1.0, np.abs(jac).max())
        # (cells are bilinear: the derivative jumps across cell borders, where fp32 may sit on the other side)
        assert np.quantile(np.abs(got['dLe'][:500, c] - jac).max(-1), 0.98) < 2e-3 * scale
    ds, ps = env.sample(torch.from_numpy(u.astype(np.float32).astype(np.float64)))
    same = np.abs(got['d'] - ds.numpy()).max(-1) < 1e-3                    # a sample on a CDF boundary may pick the neighbouring cell
    assert same.mean() > 0.995
    assert np.abs(got['d_pdf'][same] - ps.numpy()[same]).max() < 1e-3 * ps.numpy().max()


@pytest.mark.parametrize('use_mis,albedo_res,rotated', [(False, None, False), (True, 8, True)])
def test_envmap_image_and_gradient_parity(use_mis, albedo_res, rotated):
    """sdf_direct_reparam lit by an `envmap` emitter: escaped camera rays, the importance-sampled emitter direction (emitter_val /
    ds.pdf) and the MIS branch's escaped BSDF rays all look the bitmap up with the ATTACHED directions
    (python/integrators/sdf_direct_reparam.py:31,62-70,105), so dLe/dd flows into the three warp sites."""
    W, spp = 20, 4
    sc, wf, hs = make_pair('knot', 32, W, 'sdf_direct_reparam', albedo_res=albedo_res)
    bm = _envmap_bitmap()
    R = _rotation() if rotated else None
    sc.envmap = O.EnvMap(bm.astype(np.float64), to_world=R)
    hs.set_envmap(bm, to_world=R)
    sc.use_mis = use_mis
    hs.h.use_mis = int(use_mis)
    g_in = np.random.default_rng(7).uniform(-1, 1, size=(W, W, 3))
    with torch.no_grad():
        img_o = O.render(sc, wf, 'sdf_direct_reparam', seed=5, spp=spp).numpy()
    film, _, _ = hs.forward(5, spp, False)
    err = np.abs(hs.develop(film) - img_o)
    # a lane whose sample falls on a CDF boundary may pick the neighbouring cell in fp32: a few pixels may differ visibly
    assert np.quantile(err, 0.99) < 5e-5 and err.max() < 5e-2, (np.quantile(err, 0.99), err.max())
    img_ob, g_o, gp_o, ga_o = O.render_backward(sc, wf, 'sdf_direct_reparam', torch.from_numpy(g_in), seed=11, spp=spp)
    film, rec, _ = hs.forward(11, spp, True)
    g_h, ga_h, gp_h, _ = hs.backward(11, spp, film, rec, g_in)
    g_o = g_o.numpy()
    assert np.abs(g_h - g_o).max() < 2e-2 * np.abs(g_o).max() and np.linalg.norm(g_h - g_o) < 5e-3 * np.linalg.norm(g_o)
    # the envmap's direction derivative matters: with a constant emitter of the same mean radiance the gradient is different
    sc.envmap = None
    g_c = O.render_backward(sc, wf, 'sdf_direct_reparam', torch.from_numpy(g_in), seed=11, spp=spp)[1].numpy()
    assert np.linalg.norm(g_c - g_o) > 0.05 * np.linalg.norm(g_o)
    if albedo_res is not None:
        assert np.abs(ga_h - ga_o.numpy()).max() < 1e-3 * np.abs(ga_o.numpy()).max()


def test_envmap_direction_derivative():
    """The dLe/dd terms in isolation: with a high-contrast bitmap the lookup's dependence on the reparameterised directions is a few
    percent of the gradient.  The oracle evaluated with detached lookup directions gives the rest; the kernel math must reproduce
    the DIFFERENCE (closed-form dLe/dd through the three warp sites) as well as the total."""
    W, spp = 20, 4
    sc, wf, hs = make_pair('knot', 32, W, 'sdf_direct_reparam')
    bm = np.random.default_rng(1).uniform(0.05, 4.0, size=(48, 96, 3)).astype(np.float32)
    R = _rotation(9)                  # (with a rotated emitter: d/d(d_world) = R d/d(d_local))
    sc.envmap = O.EnvMap(bm.astype(np.float64), to_world=R)
    hs.set_envmap(bm, to_world=R)
    g_in = np.random.default_rng(7).uniform(-1, 1, size=(W, W, 3))
    g_att = O.render_backward(sc, wf, 'sdf_direct_reparam', torch.from_numpy(g_in), seed=11, spp=spp)[1].numpy()
    ev = O.EnvMap.eval
    try:
        O.EnvMap.eval = lambda self, d: ev(self, d.detach())
        g_det = O.render_backward(sc, wf, 'sdf_direct_reparam', torch.from_numpy(g_in), seed=11, spp=spp)[1].numpy()
    finally:
        O.EnvMap.eval = ev
    film, rec, _ = hs.forward(11, spp, True)
    g_h = hs.backward(11, spp, film, rec, g_in)[0]
    term = g_att - g_det
    assert np.linalg.norm(term) > 1e-2 * np.linalg.norm(g_att)
    assert np.linalg.norm(g_h - g_att) < 1e-3 * np.linalg.norm(g_att)
    assert np.linalg.norm((g_h - g_det) - term) < 2e-2 * np.linalg.norm(term)


@pytest.mark.parametrize('integ,kind,mis,albedo_res,shadow_origin,max_depth,n_aux', [
    ('sdf_silhouette_reparam', 'sphere', False, None, 0, -1, 4),
    ('sdf_simple_shading_reparam', 'knot', False, None, 0, -1, 3),
    ('sdf_direct_reparam', 'knot', False, 8, 0, -1, 4),
    ('sdf_direct_reparam', 'sphere', True, None, 0, -1, 2),
    ('sdf_direct_reparam', 'knot', False, None, 2, -1, 4),     # decouple_reparam: shadow origin on the detached interaction
    ('sdf_direct_reparam', 'knot', False, None, 1, -1, 4),     # detach_indirect_si
    ('sdf_direct_reparam', 'sphere', True, None, 0, 0, 4),     # max_reparam_depth = 0: camera ray only -- no auxiliary draws for the others
])
def test_convolution_warp_gradient_parity(integ, kind, mis, albedo_res, shadow_origin, max_depth, n_aux):
    """WarpFieldConvolution (python/warp.py:126-176, python/warp_conv.py; oracle/ref_semantics.md S19): closed-form adjoint of the
    reference's CustomOp `backward` in the kernel math vs torch autograd through the oracle's restatement, same auxiliary samples
    (drawn after the lane's forward draws: BSDF-sampled, shadow, camera ray).  The images are those of a render without warp field."""
    W, spp = 16, 4
    sc, wf, hs = make_pair(kind, 24, W, integ, albedo_res=albedo_res, warp=dict(kind=3, n_aux_rays=n_aux, max_depth=max_depth))
    sc.use_mis = mis
    hs.h.use_mis = int(mis)
    sc.detach_indirect_si, sc.decouple_reparam = shadow_origin == 1, shadow_origin == 2
    hs.h.shadow_origin = shadow_origin
    g_in = np.random.default_rng(7).uniform(-1, 1, size=(W, W, 3))
    img_o, g_o, gp_o, ga_o = O.render_backward(sc, wf, integ, torch.from_numpy(g_in), seed=11, spp=spp)
    with torch.no_grad():
        img_none = O.render(sc, None, integ, seed=11, spp=spp)
    assert (img_o - img_none).abs().max() < 1e-12                  # the primal of the reparameterisation is the identity
    film, rec, _ = hs.forward(11, spp, True)
    assert np.abs(hs.develop(film) - img_o.numpy()).max() < 2e-5
    g_h, ga_h, gp_h, _ = hs.backward(11, spp, film, rec, g_in)
    g_o = g_o.numpy()
    # harmonic weights 1 / (D + B)^3 with B -> 0 at silhouettes: single auxiliary rays can be ill conditioned in fp32
    assert np.linalg.norm(g_h - g_o) < 2e-3 * np.linalg.norm(g_o) and np.abs(g_h - g_o).max() < 2e-2 * np.abs(g_o).max()
    assert np.abs(gp_h - gp_o.numpy()).max() < 2e-3 * np.abs(gp_o.numpy()).max()
    if albedo_res is not None:
        assert np.abs(ga_h - ga_o.numpy()).max() < 1e-3 * np.abs(ga_o.numpy()).max()
    # and it is not the gradient without discontinuity handling
    if integ != 'sdf_silhouette_reparam':      # (without a warp field the silhouette image has no derivative at all)
        g_plain = O.render_backward(sc, None, integ, torch.from_numpy(g_in), seed=11, spp=spp)[1].numpy()
        assert np.linalg.norm(g_plain - g_o) > 0.2 * np.linalg.norm(g_o)


def test_convolution_warp_sampling_functions():
    """von Mises-Fisher sampling as restated (S19): unit directions around +z, E[cos] = coth(kappa) - 1 / kappa, pdf integrates to 1
    under its own samples' importance weights; mi.Frame3f basis orthonormal."""
    u = torch.from_numpy(np.random.default_rng(0).uniform(0, 1, (200000, 2)))
    for kappa in (5.0, 1e3):
        v = O.square_to_von_mises_fisher(u, kappa)
        assert (v.norm(dim=-1) - 1).abs().max() < 1e-9
        assert abs(v[:, 2].mean().item() - (1 / np.tanh(kappa) - 1 / kappa)) < 5e-3 / np.sqrt(kappa)
        c = 1 - 1.0 / kappa                                          # P(cos > c) = (e^k - e^{k c}) / (e^k - e^{-k})
        want = (1 - np.exp(kappa * (c - 1))) / (1 - np.exp(-2 * kappa))
        assert abs((v[:, 2] > c).double().mean().item() - want) < 5e-3
        th = torch.linspace(0, np.pi, 200001, dtype=torch.float64)
        vv = torch.stack([torch.sin(th), torch.zeros_like(th), torch.cos(th)], -1)
        integral = torch.trapezoid(O.square_to_von_mises_fisher_pdf(vv, kappa) * 2 * np.pi * torch.sin(th), th).item()
        assert abs(integral - 1.0) < 1e-4
    n = O.normalize(torch.from_numpy(np.random.default_rng(1).normal(size=(100, 3))))
    s, t = O.coordinate_system(n)
    assert (O.dot(s, t).abs().max() < 1e-12) and (O.dot(s, n).abs().max() < 1e-12) and ((torch.linalg.cross(s, t) - n).abs().max() < 1e-12)


def test_convolution_warp_with_antithetic_sampling():
    """antithetic_sampling + WarpFieldConvolution: the mirrored sample runs on a CLONE of the sampler taken after the jitter draw
    (python/integrators/reparam.py:171-178), so its auxiliary draws continue the same stream positions as the first sample's."""
    W, spp = 12, 2
    sc, wf, hs = make_pair('sphere', 24, W, 'sdf_direct_reparam', warp=dict(kind=3, n_aux_rays=3))
    hs.h.antithetic = 1
    g_in = np.random.default_rng(5).uniform(-1, 1, size=(W, W, 3))
    img_o, g_o, gp_o, _ = O.render_backward(sc, wf, 'sdf_direct_reparam', torch.from_numpy(g_in), seed=11, spp=spp, antithetic=True)
    film, rec, _ = hs.forward(11, spp, True)
    assert np.abs(hs.develop(film) - img_o.numpy()).max() < 2e-5
    g_h, _, gp_h, _ = hs.backward(11, spp, film, rec, g_in)
    g_o = g_o.numpy()
    assert np.linalg.norm(g_h - g_o) < 2e-3 * np.linalg.norm(g_o)
    assert np.abs(gp_h - gp_o.numpy()).max() < 2e-3 * np.abs(gp_o.numpy()).max()


@pytest.mark.parametrize('rows', [1, 2, 4])
def test_row_partition_granularity(rows):
    """sdfrt_set_partition_rows: ranks own interleaved sub-bands of 1 / 2 / 4 film rows and walk their own rows in 8-row tiles; every
    camera sample belongs to exactly one rank (lane counts add up, films and gradients add up to the single-rank result), also when the
    film height is not a multiple of anything (15 + 2 * 2 = 19 rows)."""
    W, spp = 15, 2
    sc, wf, hs = make_pair('sphere', 24, W, 'sdf_direct_reparam')
    g_in = np.random.default_rng(3).uniform(-1, 1, size=(W, W, 3))
    whole, rec, _ = hs.forward(7, spp, True)
    g_whole = hs.backward(7, spp, whole, rec, g_in)[0]
    try:
        HM.lib().hh_set_assign_rows(rows)
        for world in (2, 3, 8):
            assert sum(hs.local_lanes(spp, r, world) for r in range(world)) == hs.wavefront(spp)
            films, g = [], np.zeros_like(g_whole)
            for r in range(world):
                f, rec_r, _ = hs.forward(7, spp, True, r, world)
                films.append(f)
            total = sum(films)
            assert np.abs(total - whole).max() < 2e-5
            for r in range(world):
                _, rec_r, _ = hs.forward(7, spp, True, r, world)
                g += hs.backward(7, spp, total, rec_r, g_in, r, world)[0]
            assert np.abs(g - g_whole).max() < 1e-4 * np.abs(g_whole).max()
    finally:
        HM.lib().hh_set_assign_rows(8)


def test_partition_with_envmap_and_convolution_warp():
    """The row partition with the round-2 additions: an envmap-lit scene rendered by two ranks, and the convolution warp's adjoint
    (auxiliary samples continue each lane's own sampler stream, so they do not depend on which rank owns the lane) over three ranks,
    add up to the single-rank film and gradient."""
    W, spp = 14, 2
    g_in = np.random.default_rng(2).uniform(-1, 1, size=(W, W, 3))
    for warp, env, world in ((dict(kind=2), True, 2), (dict(kind=3, n_aux_rays=3), False, 3)):
        sc, wf, hs = make_pair('sphere', 24, W, 'sdf_direct_reparam', warp=warp)
        if env:
            hs.set_envmap(_envmap_bitmap())
        whole, rec, _ = hs.forward(9, spp, True)
        g_whole = hs.backward(9, spp, whole, rec, g_in)[0]
        g, total = np.zeros_like(g_whole), np.zeros_like(whole)
        recs = []
        for r in range(world):
            f, rec_r, _ = hs.forward(9, spp, True, r, world)
            total += f
            recs.append(rec_r)
        assert np.abs(total - whole).max() < 2e-5
        for r in range(world):
            g += hs.backward(9, spp, total, recs[r], g_in, r, world)[0]
        assert np.abs(g_whole).max() > 0 and np.abs(g - g_whole).max() < 1e-4 * np.abs(g_whole).max()
