"""GPU parity at the BASELINE sizes (-m gpu): the CUDA path through the C ABI against the fp64 oracle on the C2 / C3 grids
(128^3 @ 256x256/16 spp, 256^3 + 256^3 reflectance @ 512x512/64 spp; BASELINE.json configs[1..2]).

The oracle cannot back-propagate a full 17 M-lane wavefront in test time, so the gradient is pinned on a lane subsample: the adjoint
is linear in the per-lane contributions once the film adjoint is fixed (oracle/window_check.py), and `sdfrt_set_lane_window`
restricts adjoint sweep 2 to the same lanes.  Tolerances as everywhere (fp32 kernel vs fp64 oracle): film 1e-4 relative to its
maximum, per-voxel gradients max-abs <= 2e-2 * ||g||_inf and relative L2 <= 5e-3; tracer outputs: tests/baseline_cases.py."""
import numpy as np
import pytest
import torch

import baseline_cases as B

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def sdfrt_mod(built):
    import sdfrt
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return sdfrt


def make_ctx(sdfrt, name, osc, integ='sdf_direct_reparam', warp=True):
    kind, R, AR, W, spp, mult, ns = B.CONFIGS[name]
    ctx = sdfrt.Context(0)
    sdf = torch.from_numpy(B.grid(kind, R)).cuda()
    ctx.set_sdf(sdf)
    alb = None if osc.albedo_grid is None else osc.albedo_grid.float().cuda().contiguous()
    ctx.set_albedo(alb)
    ctx.set_sensor(osc.sensor.to_world, W, W)
    ctx.set_shading()
    ctx.set_integrator(integ)
    if warp:
        ctx.set_warp()                      # configs.Warp: WarpField2D(strategy 6, edge_eps 0.01), clamping 0.05
    else:
        ctx.set_warp(kind=sdfrt._capi.WARP_NONE)
    return ctx, sdf, alb


@pytest.mark.parametrize('name,n_cam,n_shadow', [('c2', 6000, 4000), ('c3', 12000, 8000)])
def test_tracer_outputs_against_fp64_oracle(sdfrt_mod, name, n_cam, n_shadow):
    """SDFBase.ray_intersect (python/shapes.py:115-288) on 10 k / 20 k camera + shadow rays of the C2 / C3 grid: its_t, warp_t,
    warp_t_d, warp_weight, warp_weight_d against the fp64 oracle; ray_intersect_non_diff's its_t (:290-339) on the same rays."""
    osc = B.oracle_scene(name, with_albedo=False)
    ctx, sdf, _ = make_ctx(sdfrt_mod, name, osc)
    o, d, mt, nc = B.rays(osc, n_cam, n_shadow)
    to = lambda a: torch.from_numpy(a).double()
    act = torch.ones(o.shape[0], dtype=torch.bool)
    w = osc.sdf.ray_intersect(to(o), to(d), to(mt), act)
    want = dict(zip(['its_t', 'warp_t', 'warp_t_d', 'warp_weight', 'warp_weight_d'], [x.numpy() for x in w]))
    got = {k: v.cpu().numpy() for k, v in ctx.trace_rays(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(),
                                                          torch.from_numpy(mt).cuda(), differentiable=True).items()}
    for sl, nm in ((slice(0, nc), 'camera'), (slice(nc, None), 'shadow')):
        st = B.compare_tracer_outputs({k: v[sl] for k, v in got.items()}, {k: v[sl] for k, v in want.items()}, f'{name} {nm}')
        print(name, nm, st)
        assert st['n_valid'] > 500
    its_v = ctx.trace_rays(torch.from_numpy(o).cuda(), torch.from_numpy(d).cuda(), torch.from_numpy(mt).cuda(),
                           differentiable=False)['its_t'].cpu().numpy()
    want_v = osc.sdf.ray_intersect_non_diff(to(o), to(d), to(mt), act).numpy()
    assert (np.isfinite(its_v) == np.isfinite(want_v)).mean() > 0.999
    both = np.isfinite(its_v) & np.isfinite(want_v)
    assert np.quantile(np.abs(its_v[both] - want_v[both]), 0.999) < 5e-5


def _check_grad(got, want, what):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert np.abs(want).max() > 0, what
    assert np.abs(got - want).max() < 2e-2 * np.abs(want).max(), (what, np.abs(got - want).max() / np.abs(want).max())
    assert np.linalg.norm(got - want) < 5e-3 * np.linalg.norm(want), (what, np.linalg.norm(got - want) / np.linalg.norm(want))


def test_c2_band_film_and_gradient_of_its_lanes(sdfrt_mod):
    """C2 (128^3, 256x256, 16 spp, sdf_direct_reparam + WarpField2D): the 8 film rows through the middle of the object.
    Film contribution of those 33 k lanes (the CUDA path renders exactly them as rank r of a partition into single bands) and
    their scattered gradient under the completed film of the whole wavefront, against the oracle evaluating the same lanes."""
    from oracle import window_check as WC
    name = 'c2'
    kind, R, AR, W, spp, mult, ns = B.CONFIGS[name]
    osc = B.oracle_scene(name)
    owf = B.oracle_warp(osc)
    ctx, sdf, _ = make_ctx(sdfrt_mod, name, osc)
    Hb, Wb, b = ctx.film_dims()
    nbands = (Hb + 7) // 8
    r = (Hb // 2) // 8
    lo, hi = 8 * r * Wb * spp, min(Hb, 8 * r + 8) * Wb * spp
    G = np.random.default_rng(3).uniform(-1, 1, size=(W, W, 3)).astype(np.float32)
    seed_grad = 13
    # whole wavefront: completed film + lane records; gradient of the window's lanes
    ctx.set_partition(0, 1)
    film_full = ctx.adjoint_accumulate(seed_grad, spp).clone()
    ctx.set_lane_window(lo, hi)
    g_win = torch.zeros_like(sdf)
    ctx.adjoint_backprop(film_full, torch.from_numpy(G).cuda(), g_win)
    ctx.set_lane_window(0, 0)
    g_all = torch.zeros_like(sdf)
    ctx.adjoint_backprop(film_full, torch.from_numpy(G).cuda(), g_all)
    # the window is a strict part of the whole, and windows add up (linearity in the per-lane contributions)
    ctx.set_lane_window(0, lo)
    g_a = torch.zeros_like(sdf)
    ctx.adjoint_backprop(film_full, torch.from_numpy(G).cuda(), g_a)
    ctx.set_lane_window(hi, Hb * Wb * spp)
    g_b = torch.zeros_like(sdf)
    ctx.adjoint_backprop(film_full, torch.from_numpy(G).cuda(), g_b)
    ctx.set_lane_window(0, 0)
    # the band alone (a new sweep: set_partition invalidates the records above)
    ctx.set_partition(r, nbands)
    film_band = ctx.adjoint_accumulate(seed_grad, spp).cpu().numpy()
    ctx.set_partition(0, 1)
    acc, g_o, _ = WC.window_gradient(osc, owf, 'sdf_direct_reparam', film_full.cpu().numpy(), G, seed_grad, spp, (lo, hi))
    acc = acc.numpy()
    assert np.abs(acc[..., 3]).max() > 1.0
    assert np.abs(film_band - acc).max() < 1e-4 * np.abs(acc).max(), np.abs(film_band - acc).max() / np.abs(acc).max()
    _check_grad(g_win.cpu().numpy(), g_o.numpy(), 'c2 band gradient')
    assert (g_all - g_win).abs().max() > 1e-3 * g_all.abs().max()
    s = g_a + g_win + g_b
    assert (s - g_all).abs().max() < 1e-3 * g_all.abs().max()


def test_c3_window_gradient_with_reflectance_grid(sdfrt_mod):
    """C3 (256^3 SDF + 256^3 x 3 reflectance grid, 512x512, 64 spp): 16 k lanes (256 pixels of the film's middle row) -- gradient
    w.r.t. the SDF grid and w.r.t. the reflectance grid under the completed film, against the oracle on the same lanes."""
    from oracle import window_check as WC
    name = 'c3'
    kind, R, AR, W, spp, mult, ns = B.CONFIGS[name]
    osc = B.oracle_scene(name)
    owf = B.oracle_warp(osc)
    ctx, sdf, alb = make_ctx(sdfrt_mod, name, osc)
    lo, hi = WC.centre_window(osc.sensor, spp, 16384)
    G = np.random.default_rng(4).uniform(-1, 1, size=(W, W, 3)).astype(np.float32)
    seed_grad = 13
    film_full = ctx.adjoint_accumulate(seed_grad, spp)
    ctx.set_lane_window(lo, hi)
    g_win, ga_win = torch.zeros_like(sdf), torch.zeros_like(alb)
    ctx.adjoint_backprop(film_full, torch.from_numpy(G).cuda(), g_win, ga_win)
    ctx.set_lane_window(0, 0)
    _, g_o, ga_o = WC.window_gradient(osc, owf, 'sdf_direct_reparam', film_full.cpu().numpy(), G, seed_grad, spp, (lo, hi))
    print('c3 window:', WC.gradient_errors(g_win.cpu().numpy(), g_o.numpy()))
    _check_grad(g_win.cpu().numpy(), g_o.numpy(), 'c3 window sdf gradient')
    ga, gao = ga_win.cpu().numpy(), ga_o.numpy()
    assert np.abs(ga - gao).max() < 1e-3 * np.abs(gao).max()


def test_device_counters_equal_the_oracle_count(sdfrt_mod):
    """algorithmic bytes = 256 B x the device-counted B-spline evaluations, and "the equality with the oracle's count is
    the proof the kernel did the work".  With the empty-space test off (sdfrt_set_empty_space_skip(0): every camera ray marched,
    like the reference) the counters of `k_march` / `k_shade` equal the oracle's number of 64-tap evaluations for the same lanes
    up to the few rays whose fp32 / fp64 step counts differ: (i) a whole small film, (ii) one 8-row band of C2."""
    from oracle import sdf_oracle as O
    sdfrt = sdfrt_mod
    # (i) knot 32^3, 24x24, 4 spp, silhouette integrator without a warp field: value marches + refinement only
    W, spp = 24, 4
    osc = O.make_scene('knot', res=32, width=W, height=W)
    ctx = sdfrt.Context(0)
    ctx.set_sdf(osc.sdf.data.float().cuda().contiguous())
    ctx.set_sensor(osc.sensor.to_world, W, W)
    ctx.set_shading()
    ctx.set_integrator('sdf_silhouette_reparam')
    ctx.set_warp(kind=sdfrt._capi.WARP_NONE)
    osc.sdf.n_evals = 0
    with torch.no_grad():
        O.render(osc, None, 'sdf_silhouette_reparam', seed=3, spp=spp)
    n_oracle = osc.sdf.n_evals
    counts, films = {}, {}
    for skip in (0, 1):
        ctx.set_empty_space_skip(skip)
        ctx.stats(reset=True)
        films[skip] = ctx.render_primal(3, spp).clone()
        st = ctx.stats(reset=True)
        counts[skip] = st['n_evals_value'] + st['n_evals_grad'] + st['n_evals_hess']
    assert n_oracle > 10000 and abs(counts[0] - n_oracle) <= 2e-3 * n_oracle, (counts, n_oracle)
    assert counts[1] <= counts[0]
    assert (films[0] - films[1]).abs().max() < 1e-5 * films[0].abs().max()   # skipping changes no result (up to the order of the atomics)
    # (ii) the middle band of C2 at the adjoint spp, primal silhouette render: oracle restricted to the band's lanes
    name = 'c2'
    kind, R, AR, Wc, sppc, mult, ns = B.CONFIGS[name]
    osc = B.oracle_scene(name)
    ctx, sdf, _ = make_ctx(sdfrt, name, osc, integ='sdf_silhouette_reparam', warp=False)
    Hb, Wb, b = ctx.film_dims()
    nbands, r = (Hb + 7) // 8, (Hb // 2) // 8
    lo, hi = 8 * r * Wb * sppc, (8 * r + 8) * Wb * sppc
    ctx.set_partition(r, nbands)
    ctx.set_empty_space_skip(0)
    ctx.stats(reset=True)
    ctx.render_primal(5, sppc)
    st = ctx.stats(reset=True)
    n_kernel = st['n_evals_value'] + st['n_evals_grad'] + st['n_evals_hess']
    osc.sdf.n_evals = 0
    with torch.no_grad():
        O.render(osc, None, 'sdf_silhouette_reparam', seed=5, spp=sppc, lane_range=(lo, hi), develop=False)
    assert st['n_rays'] == hi - lo
    assert osc.sdf.n_evals > 100000 and abs(n_kernel - osc.sdf.n_evals) <= 2e-3 * osc.sdf.n_evals, (n_kernel, osc.sdf.n_evals)
