"""CPU twins of tests/test_gpu_baseline_sizes.py: the kernel arithmetic compiled for the host (tests/host_harness) against the
fp64 oracle at the C2 size (tracer outputs) and the lane-window gradient check (oracle/window_check.py) that pins the gradient
of a lane subsample under the completed film of the whole wavefront."""
import numpy as np
import torch

import baseline_cases as B
import hostmath as HM
from oracle import sdf_oracle as O
from oracle import window_check as WC


def test_c2_tracer_outputs_host_math_against_fp64_oracle():
    name = 'c2'
    kind, R, AR, W, spp, mult, ns = B.CONFIGS[name]
    osc = B.oracle_scene(name)
    o, d, mt, nc = B.rays(osc, 3000, 2000)
    to = lambda a: torch.from_numpy(a).double()
    w = osc.sdf.ray_intersect(to(o), to(d), to(mt), torch.ones(o.shape[0], dtype=torch.bool))
    want = dict(zip(['its_t', 'warp_t', 'warp_t_d', 'warp_weight', 'warp_weight_d'], [x.numpy() for x in w]))
    hs = HM.HostScene(B.grid(kind, R), osc.sensor.to_world, W, W)
    got = hs.trace(o, d, mt, True)
    for sl, nm in ((slice(0, nc), 'camera'), (slice(nc, None), 'shadow')):
        st = B.compare_tracer_outputs({k: v[sl] for k, v in got.items()}, {k: v[sl] for k, v in want.items()}, nm)
        assert st['n_valid'] > 300, st


def test_window_gradient_host_math_against_oracle():
    """Linearity of the adjoint in the per-lane contributions: gradient of a lane window under the completed film (host twin of
    sdfrt_set_lane_window) = oracle autograd over the same lanes; windows add up to the full gradient."""
    R, W, spp = 48, 40, 4
    sc = O.make_scene('blobs', res=R, width=W, height=W, albedo_grid_res=16)
    wf = O.WarpField2D(sc.sdf, weight_strategy=6, edge_eps=0.01)
    wf.clamping_thresh = 0.05
    hs = HM.HostScene(sc.sdf.data.numpy(), sc.sensor.to_world, W, W, integrator=2, albedo_grid=sc.albedo_grid.numpy())
    G = np.random.default_rng(2).uniform(-1, 1, size=(W, W, 3)).astype(np.float32)
    film, rec, _ = hs.forward(13, spp, True)
    lo, hi = WC.centre_window(sc.sensor, spp, 3 * (W + 4) * spp)
    lo -= (W + 4) * spp // 2                                 # ~3 film rows around the centre, not aligned to rows
    hi -= (W + 4) * spp // 2
    g_win, ga_win, _, _ = hs.backward(13, spp, film, rec, G, lane_window=(lo, hi))
    g_all, ga_all, _, _ = hs.backward(13, spp, film, rec, G)
    acc, g_o, ga_o = WC.window_gradient(sc, wf, 'sdf_direct_reparam', film, G, 13, spp, (lo, hi))
    e = WC.gradient_errors(g_win, g_o.numpy())
    assert e['ref_nonzero_voxels'] > 1000
    assert e['max_abs_over_ref_max'] < 2e-2 and e['rel_l2'] < 5e-3, e
    assert np.abs(ga_win - ga_o.numpy()).max() < 1e-3 * np.abs(ga_o.numpy()).max()
    g_a = hs.backward(13, spp, film, rec, G, lane_window=(0, lo))[0] if lo > 0 else 0
    g_b = hs.backward(13, spp, film, rec, G, lane_window=(hi, hs.wavefront(spp)))[0]
    assert np.abs(g_a + g_win + g_b - g_all).max() < 1e-4 * np.abs(g_all).max()
    assert np.abs(g_all - g_win).max() > 1e-2 * np.abs(g_all).max()
    # full window == render_backward
    g_full_o = O.render_backward(sc, wf, 'sdf_direct_reparam', torch.from_numpy(G).double(), seed=13, spp=spp)[1].numpy()
    _, g_o2, _ = WC.window_gradient(sc, wf, 'sdf_direct_reparam', O.render(sc, wf, 'sdf_direct_reparam', 13, spp, mode_backward=True,
                                                                            develop=False).detach().numpy(), G, 13, spp,
                                    (0, hs.wavefront(spp)))
    assert np.abs(g_o2.numpy() - g_full_o).max() < 1e-9 * max(1.0, np.abs(g_full_o).max())
