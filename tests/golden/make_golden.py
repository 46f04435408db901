"""Generates the golden fixtures in this directory with the CPU oracle (fp64).

The reference ships no golden data and cannot be imported offline (PARITY UNPINNED, see
oracle/sdf_oracle.py), so these vectors pin *our restatement*: they detect drift of the oracle and
give the GPU tests something to compare against that does not require running the oracle.
Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import sdf_oracle as O  # noqa: E402

CASES = {
    # name: (sdf kind, R, W, spp, integrator, albedo_res)
    'sphere_direct': ('sphere', 32, 24, 8, 'sdf_direct_reparam', None),
    'knot_direct': ('knot', 32, 24, 8, 'sdf_direct_reparam', None),
    'blobs_tex_direct': ('blobs', 32, 24, 8, 'sdf_direct_reparam', 8),
    'knot_silhouette': ('knot', 32, 24, 8, 'sdf_silhouette_reparam', None),
    'sphere_simple': ('sphere', 32, 24, 8, 'sdf_simple_shading_reparam', None),
}
# sdf_prb_reparam fixtures: name: (sdf kind, R, W, spp, albedo_res, max_depth, rr_depth, max_reparam_depth)
PRB_CASES = {
    'blobs_tex_prb': ('blobs', 32, 16, 4, 8, 4, 5, -1),      # interreflections + reflectance grid
    'knot_prb_rr': ('knot', 32, 16, 4, None, 6, 1, -1),      # Russian roulette from the second vertex on
    'blobs_prb_primary': ('blobs', 32, 16, 4, None, 3, 5, 0),  # configs.WarpPRBPrimary: max_reparam_depth = 0
}
SEED, SEED_GRAD = 3, 11


def grad_image(W):
    return np.random.default_rng(5).uniform(-1, 1, size=(W, W, 3))


def run_case(name):
    kind, R, W, spp, integ, ar = CASES[name]
    sc = O.make_scene(kind, res=R, width=W, height=W, albedo_grid_res=ar)
    wf = O.WarpField2D(sc.sdf, weight_strategy=6, edge_eps=0.01)
    wf.clamping_thresh = 0.05
    with torch.no_grad():
        img = O.render(sc, wf, integ, seed=SEED, spp=spp)
    img_adj, g, gp, ga = O.render_backward(sc, wf, integ, torch.from_numpy(grad_image(W)), seed=SEED_GRAD, spp=spp)
    # per-ray tracer outputs for the first 256 lanes of the central film rows (fixed rays)
    pos = torch.from_numpy(np.random.default_rng(9).uniform(0, W, size=(256, 2)))
    o, d, maxt = sc.sensor.sample_ray(pos, torch.float64)
    its_t, warp_t, warp_t_d, ww, wwd = sc.sdf.ray_intersect(o, d, maxt, torch.ones(256, dtype=torch.bool))
    out = dict(image=img.numpy().astype(np.float32), image_adj=img_adj.numpy().astype(np.float32),
               grad_sdf=g.numpy().astype(np.float32), grad_p=gp.numpy().astype(np.float32),
               ray_o=o.numpy().astype(np.float32), ray_d=d.numpy().astype(np.float32), ray_maxt=maxt.numpy().astype(np.float32),
               its_t=its_t.numpy().astype(np.float32), warp_t=warp_t.numpy().astype(np.float32),
               warp_t_d=warp_t_d.numpy().astype(np.float32), warp_weight=ww.numpy().astype(np.float32))
    if ga is not None:
        out['grad_albedo'] = ga.numpy().astype(np.float32)
    return out


def run_prb_case(name):
    kind, R, W, spp, ar, max_depth, rr_depth, mrd = PRB_CASES[name]
    sc = O.make_scene(kind, res=R, width=W, height=W, albedo_grid_res=ar)
    wf = O.WarpField2D(sc.sdf, weight_strategy=6, edge_eps=0.01)
    wf.clamping_thresh = 0.05
    wf.max_reparam_depth = mrd
    cfg = O.PrbConfig(max_depth=max_depth, rr_depth=rr_depth)
    img = O.render_prb(sc, wf, seed=SEED, spp=spp, cfg=cfg)
    img_adj, g, gp, ga = O.render_backward_prb(sc, wf, torch.from_numpy(grad_image(W)), seed=SEED_GRAD, spp=spp, cfg=cfg)
    out = dict(image=img.numpy().astype(np.float32), image_adj=img_adj.numpy().astype(np.float32),
               grad_sdf=g.numpy().astype(np.float32), grad_p=gp.numpy().astype(np.float32))
    if ga is not None:
        out['grad_albedo'] = ga.numpy().astype(np.float32)
    return out


def redistance_case():
    """Re-distancing fixture (oracle/redistance_oracle.c): a level-set function with a wobbly sphere-like zero set on an
    anisotropic 28 x 24 x 20 grid and its re-distanced field."""
    from oracle import redistance_oracle as ro
    sys.path.insert(0, os.path.dirname(HERE))
    import test_redistance_oracle as T
    phi = T.wobbly((20, 24, 28), seed=42)
    return dict(phi=phi, dist=ro.redistance(phi))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'redistance':      # only the grid-update fixtures (the render cases take minutes)
        np.savez_compressed(os.path.join(HERE, 'redistance_wobbly.npz'), **redistance_case())
        print('wrote redistance_wobbly')
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'prb':
        for name in PRB_CASES:
            np.savez_compressed(os.path.join(HERE, name + '.npz'), **run_prb_case(name))
            print('wrote', name)
        sys.exit(0)
    np.savez_compressed(os.path.join(HERE, 'redistance_wobbly.npz'), **redistance_case())
    for name in PRB_CASES:
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **run_prb_case(name))
        print('wrote', name)
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **run_case(name))
        print('wrote', name)
