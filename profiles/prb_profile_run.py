"""One PRB view (primal + re-render + replay/backprop) on a 128^3 grid, 256x256 film, 16 spp -- the command profiled under ncu
for profiles/r1_prb_per_lane.md:
  ncu --set full --clock-control none --import-source on -k regex:k_prb -c 3 -o gpurun_out/prb python profiles/prb_profile_run.py
r1_prb_per_lane.md was captured with the per-lane walker (now SDFRT_PRB_WAVEFRONT=0).  For the wavefront form (the default) the
kernels of interest are k_prb_wf_emit, the k_march launches of the bounces and k_prb_backward:  -k "regex:k_prb|k_march" -c 40.
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sdfrt  # noqa: E402
from sdfrt import synthetic as syn  # noqa: E402

dev = torch.device('cuda', 0)
ctx = sdfrt.Context(0)
sdf = syn.make_sdf_device('blobs', 128, dev)
cam = sdfrt.get_regular_cameras(12, resx=256, resy=256)[0]
ctx.set_sdf(sdf)
ctx.set_albedo(None)
ctx.set_sensor(cam.to_world, 256, 256)
ctx.set_shading()
ctx.set_integrator('sdf_prb_reparam')
ctx.set_prb(max_depth=3, rr_depth=5)
ctx.set_warp()
film, film_a = ctx.render_view(1, 16, 14, 16)
g = torch.zeros_like(sdf)
ctx.adjoint_backprop(film_a, torch.ones((256, 256, 3), device=dev), g)
torch.cuda.synchronize()
print('stats', ctx.stats(), 'grad max', float(g.abs().max()))
