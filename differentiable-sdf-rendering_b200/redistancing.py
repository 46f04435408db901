"""python/redistancing.py: `redistance(phi, method='fastsweep')`.

The reference hands the grid to the native `fastsweep` module; here it goes to `sdfrt_redistance` (csrc/sdfrt_grid.cuh), the
block-asynchronous solver of the same discrete problem.  No CPU fallback: a CUDA tensor (or an array that is moved to the
GPU) comes back as a tensor of the same shape on the GPU."""
import numpy as np
import torch

from . import engine

_shared = {}


def shared_context(device=0):
    """One sdfrt context per device for the stateless grid operations (redistance, Adam step, regulariser)."""
    index = device if isinstance(device, int) else (torch.device(device).index or 0)
    if index not in _shared:
        _shared[index] = engine.Context(index)
    return _shared[index]


def redistance(phi, method='fastsweep', lower_bound=None, context=None):
    """python/redistancing.py:4-13.  `lower_bound` (optional) fuses the `dr.maximum(sdf, bbox_sdf)` of
    python/variables.py:185-187 into the call.  `method='fmm'` (scikit-fmm) is not provided."""
    if method != 'fastsweep':
        if method == 'fmm':
            raise NotImplementedError("the scikit-fmm back end is not provided; use method='fastsweep'")
        raise ValueError("Invalid re-distancing method")
    if not torch.is_tensor(phi):
        phi = torch.from_numpy(np.ascontiguousarray(np.asarray(phi, dtype=np.float32)))
    ctx = context if context is not None else shared_context(phi.device.index or 0)
    phi = phi.to(ctx.device, torch.float32).contiguous()
    if lower_bound is not None:
        lower_bound = torch.as_tensor(lower_bound, dtype=torch.float32).to(phi.device).contiguous()
    # synchronous on purpose: the solver's error word (barrier time-out / iteration cap) is checked before the grid is handed
    # to the optimiser -- an aborted solve would leave unreached nodes at +-1e30
    out, _ = ctx.redistance(phi, lower_bound, return_iters=True)
    return out
