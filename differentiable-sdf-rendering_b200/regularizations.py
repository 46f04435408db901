"""python/regularizations.py.  `eval_discrete_laplacian_reg` runs `sdfrt_laplacian_reg` (csrc/sdfrt_grid.cuh) and is
differentiable through torch autograd; `reference_laplacian_reg` is the same sum written with plain torch ops (the fp32/fp64
reference the tests compare the kernel with -- not used by the product path)."""
import torch

from ._capi import SdfrtError


class _LaplacianReg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data):
        from . import redistancing
        sd = redistancing.shared_context(data.device.index or 0)
        ctx.save_for_backward(data)
        return sd.laplacian_reg(data.detach().contiguous(), 1.0, None, True).to(torch.float32)

    @staticmethod
    def backward(ctx, grad_out):
        from . import redistancing
        (data,) = ctx.saved_tensors
        sd = redistancing.shared_context(data.device.index or 0)
        g = torch.zeros_like(data, memory_format=torch.contiguous_format)
        sd.laplacian_reg(data.detach().contiguous(), 1.0, g, False)
        return g * grad_out


def eval_discrete_laplacian_reg(data, _=None):
    """Sum over voxels of (c - mean of the 6 axis neighbours)^2 with clamped neighbour indices (python/regularizations.py:5-25).
    data: (Z, Y, X[, 1]) float32 CUDA tensor."""
    from . import redistancing
    if not (torch.is_tensor(data) and data.dtype == torch.float32 and data.device == redistancing.shared_context(data.device.index or 0).device):
        raise SdfrtError("eval_discrete_laplacian_reg expects a float32 tensor on the rendering device (no CPU path)")
    return _LaplacianReg.apply(data)


def reference_laplacian_reg(data, _=None):
    """The same quantity with torch ops on any device / dtype (test reference)."""
    d = data[..., 0] if data.dim() == 4 else data

    def shifted(t, dim, step):
        idx = (torch.arange(t.shape[dim], device=t.device) + step).clamp(0, t.shape[dim] - 1)
        return t.index_select(dim, idx)

    nb = sum(shifted(d, dim, s) for dim in range(3) for s in (-1, 1))
    return ((d - nb / 6) ** 2).sum()
