"""sdf_prb_reparam at BASELINE configs[4] (256^3, 512x512, 128 spp, max_depth 3): time of the two sweeps and of the replay, with and
without the scatter (grad_sdf = NULL: walker only), per launch role."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import sdfrt  # noqa: E402
import baseline_cases as B  # noqa: E402

R, W, spp = 256, 512, 128
sdf = torch.from_numpy(B.grid('blobs', R)).cuda()
alb = torch.from_numpy(sdfrt.synthetic.make_albedo_grid(R)).float().cuda()
cam = sdfrt.get_regular_cameras(12, resx=W, resy=W)[0]
G = torch.from_numpy(np.random.default_rng(2).uniform(-1, 1, (W, W, 3))).float().cuda()
ctx = sdfrt.Context(0)
ctx.set_sdf(sdf)
ctx.set_albedo(alb)
ctx.set_sensor(cam.to_world, W, W)
ctx.set_shading()
ctx.set_integrator('sdf_prb_reparam')
ctx.set_warp()
ctx.set_prb(3, 5)
out = {}


def timed(fn, n=3):
    ts = []
    for _ in range(n + 1):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t) * 1e3)
    return float(np.median(ts[1:]))


film = {}
out['primal_ms'] = timed(lambda: ctx.render_primal(1, spp))
out['adjoint_sweep1_ms'] = timed(lambda: film.__setitem__('a', ctx.adjoint_accumulate(2, spp)))
ctx.stats(reset=True)
g = torch.zeros_like(sdf)
ga = torch.zeros_like(alb)
out['backprop_ms'] = timed(lambda: ctx.adjoint_backprop(film['a'], G, g, ga))
out['stats_4_backprops'] = {k: int(v) for k, v in ctx.stats(reset=True).items()}
print(json.dumps(out), flush=True)
gp = torch.zeros(3, device='cuda')
out['backprop_walker_only_ms'] = timed(lambda: ctx.adjoint_backprop(film['a'], G, None, None, gp))
print(json.dumps(out, indent=1))
json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'prb_timing.json'), 'w'), indent=1)
