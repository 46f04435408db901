"""Inputs for the parity tests at the BASELINE sizes (BASELINE.json configs[1..3]):
the synthetic stand-in grids, the rig's sensor 0, and a seeded set of camera + shadow rays through the object.
Shared by the `-m gpu` tests (CUDA path vs oracle) and the CPU tests (host-compiled kernel math vs oracle)."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# name: (sdf kind, sdf res, albedo res, film W = H, adjoint spp, primal spp multiplier, sensors of the rig)
CONFIGS = {
    'c2': ('blobs', 128, None, 256, 16, 4, 12),    # "dragon no-tex-12, 128^3, 256x256 @16spp"
    'c3': ('blobs', 256, 256, 512, 64, 4, 12),     # "bunny with-tex-12-hq, 256^3 + 256^3 albedo, 512x512 @64spp"
}
_GRIDS = {}


def grid(kind, R):
    """(R, R, R) float32 numpy grid, the same values bench.py renders (synthetic.make_sdf_device, evaluated on the CPU)."""
    if (kind, R) not in _GRIDS:
        import sdfrt
        _GRIDS[(kind, R)] = sdfrt.synthetic.make_sdf_device(kind, R, 'cpu').numpy()
    return _GRIDS[(kind, R)]


def oracle_scene(name, with_albedo=True):
    from oracle import sdf_oracle as O
    kind, R, AR, W, spp, mult, nsens = CONFIGS[name]
    tw = O.look_at(O.regular_camera_origins(nsens)[0], [0.5, 0.5, 0.5], [0, 1, 0])
    sensor = O.Sensor(to_world=tw, width=W, height=W)
    sdf = O.Grid3d(torch.from_numpy(grid(kind, R)).to(torch.float64))
    alb = None
    if AR is not None and with_albedo:
        import sdfrt
        alb = torch.from_numpy(sdfrt.synthetic.make_albedo_grid(AR)).to(torch.float64)
    return O.Scene(sdf=sdf, sensor=sensor, albedo_grid=alb)


def oracle_warp(osc):
    from oracle import sdf_oracle as O
    wf = O.WarpField2D(osc.sdf, weight_strategy=6, edge_eps=0.01)   # configs.Warp (python/configs.py:12-40)
    wf.clamping_thresh = 0.05
    return wf


def rays(osc, n_cam, n_shadow, seed=0):
    """Seeded float32 rays (o, d, maxt): `n_cam` camera rays through film positions over the object's footprint and `n_shadow`
    shadow rays spawned like sdf_direct_reparam does (python/integrators/sdf_direct_reparam.py:38-51) from camera-ray hit points
    found with the oracle's value tracer.  Returned as numpy float32 -- both implementations start from the same bits."""
    from oracle import sdf_oracle as O
    rng = np.random.default_rng(seed)
    S = osc.sensor
    W, H = S.width, S.height
    pos = torch.from_numpy(np.stack([rng.uniform(0.0, W, n_cam + 12 * n_shadow),
                                     rng.uniform(0.0, H, n_cam + 12 * n_shadow)], -1))
    o, d, maxt = S.sample_ray(pos, torch.float64)
    cam = (o[:n_cam], d[:n_cam], maxt[:n_cam])
    o2, d2, m2 = o[n_cam:], d[n_cam:], maxt[n_cam:]
    its = osc.sdf.ray_intersect_non_diff(o2, d2, m2, torch.ones(o2.shape[0], dtype=torch.bool))
    h = torch.nonzero(torch.isfinite(its)).squeeze(-1)[:n_shadow]
    p = o2[h] + its[h][:, None] * O.normalize(d2[h])
    n = O.normalize(osc.sdf.eval_grad(p))
    u = torch.from_numpy(rng.uniform(0, 1, (h.numel(), 2)))
    d_e = O.square_to_uniform_sphere(u)
    flip = (O.dot(d_e, n) < 0)
    d_e = torch.where(flip[:, None], -d_e, d_e)                      # keep the BSDF gate open (cos_o > 0), like the marched ones
    ds_p = p + d_e * osc.env_dist
    mag = (1 + p.abs().max(-1).values) * O.RAY_EPSILON
    o_s = p + (mag * O.dr_sign(O.dot(n, ds_p - p)))[:, None] * n
    dv = ds_p - o_s
    dist = torch.linalg.norm(dv, dim=-1)
    sh = (o_s, dv / dist[:, None], dist * (1 - O.SHADOW_EPSILON))
    f = lambda a, b: np.ascontiguousarray(torch.cat([a, b]).numpy().astype(np.float32))
    return f(cam[0], sh[0]), f(cam[1], sh[1]), f(cam[2], sh[2]), n_cam


def compare_tracer_outputs(got, want, what=''):
    """Stated fp32-vs-fp64 tolerances for SDFBase.ray_intersect's five outputs (python/shapes.py:259-288) at the BASELINE sizes.
    The silhouette weights go as 1/denom^3 (python/shapes.py:74-75), so single grazing rays are ill-conditioned: hit decisions
    and distances are checked on (nearly) all rays, the weight-derived outputs by quantiles and in aggregate."""
    its_g, its_w = got['its_t'], want['its_t']
    agree = np.isfinite(its_g) == np.isfinite(its_w)
    assert agree.mean() > 0.999, (what, agree.mean())
    both = np.isfinite(its_g) & np.isfinite(its_w)
    assert np.quantile(np.abs(its_g[both] - its_w[both]), 0.999) < 5e-5, what
    wt_g, wt_w = got['warp_t'], want['warp_t']
    vagree = np.isfinite(wt_g) == np.isfinite(wt_w)
    assert vagree.mean() > 0.999, (what, vagree.mean())
    v = np.isfinite(wt_g) & np.isfinite(wt_w) & agree
    out = {'n_valid': int(v.sum()), 'n_hit': int(both.sum())}
    rel = lambda a, b: np.abs(a - b) / np.maximum(1.0, np.abs(b))
    out['warp_t_q99'] = float(np.quantile(rel(wt_g[v], wt_w[v]), 0.99))
    out['warp_weight_q99'] = float(np.quantile(np.abs(got['warp_weight'][v] - want['warp_weight'][v]), 0.99))
    a, b = got['warp_t_d'][v], want['warp_t_d'][v]
    e = np.abs(a - b).max(-1) / np.maximum(1.0, np.abs(b).max(-1))
    out['warp_t_d_q95'], out['warp_t_d_q99'] = float(np.quantile(e, 0.95)), float(np.quantile(e, 0.99))
    assert out['warp_t_q99'] < 1e-3 and out['warp_weight_q99'] < 1e-3, (what, out)
    assert out['warp_t_d_q95'] < 5e-3 and out['warp_t_d_q99'] < 5e-2, (what, out)
    # warp_weight / warp_weight_d are only non-trivial where 0 < sum of weights < 1 (python/shapes.py:270-272): rays that pass the
    # surface at a distance.  Checked on that subset, relative to the reference values.
    ww = want['warp_weight']
    s = v & (ww > 0) & (ww < 1) & (got['warp_weight'] > 0) & (got['warp_weight'] < 1)
    out['n_partial_weight'] = int(s.sum())
    if s.sum() >= 20:
        out['warp_weight_rel_q90'] = float(np.quantile(np.abs(got['warp_weight'][s] - ww[s]) / ww[s], 0.90))
        a, b = got['warp_weight_d'][s], want['warp_weight_d'][s]
        out['warp_weight_d_rel_q90'] = float(np.quantile(np.abs(a - b).max(-1) / np.maximum(np.abs(b).max(-1), 1e-6), 0.90))
        assert out['warp_weight_rel_q90'] < 1e-2 and out['warp_weight_d_rel_q90'] < 1e-2, (what, out)
    return out
