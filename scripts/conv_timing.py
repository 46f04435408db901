"""Time of one view (primal + adjoint sweeps) at BASELINE configs[1] (128^3, 256x256, 16 spp) with the reference's warp configurations:
WarpField2D ("warp"), no warp field ("onlyshadinggrad") and WarpFieldConvolution with 2..32 auxiliary rays (figures/benchmark/benchmark.py:86-92)."""
import json
import sys
import os
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import sdfrt  # noqa: E402
import baseline_cases as B  # noqa: E402

kind, R, AR, W, spp, mult, ns = B.CONFIGS['c2']
sdf = sdfrt.Grid3d(B.grid(kind, R), device=0)
cam = sdfrt.get_regular_cameras(ns, resx=W, resy=W)[0]
integ = sdfrt.integrators.create('sdf_direct_reparam', {}, sdf)
scene = sdfrt.Scene(integ, [cam])
G = torch.from_numpy(np.random.default_rng(2).uniform(-1, 1, (W, W, 3))).float().cuda()
out = {}
wf2 = sdfrt.WarpField2D(sdf, weight_strategy=6, edge_eps=0.01)
wf2.clamping_thresh = 0.05
cases = [('warp', wf2), ('onlyshadinggrad', sdfrt.DummyWarpField(sdf))] + [(f'conv{n}', sdfrt.WarpFieldConvolution(sdf, n_aux_rays=n)) for n in (2, 4, 8, 16, 32)]
for name, wf in cases:
    integ.warp_field = wf
    ts = []
    for it in range(9):
        sdf.data.grad = None
        sdf.data.requires_grad_(True)
        torch.cuda.synchronize()
        t = time.perf_counter()
        img = sdfrt.render(scene, sensor=0, seed=it, spp=spp * mult, seed_grad=100 + it, spp_grad=spp)
        (img * G).sum().backward()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    out[name] = {'ms_per_view': float(np.median(ts[5:])), 'grad_norm': float(sdf.data.grad.norm())}
    print(name, out[name], flush=True)
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'conv_timing.json'), 'w'), indent=1)
