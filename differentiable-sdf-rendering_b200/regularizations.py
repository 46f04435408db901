"""python/regularizations.py on torch tensors."""
import torch


def eval_discrete_laplacian_reg(data, _=None):
    """Sum over voxels of (c - mean of the 6 axis neighbours)^2 with clamped neighbour indices (python/regularizations.py:5-25)."""
    d = data[..., 0] if data.dim() == 4 else data

    def shifted(t, dim, step):
        idx = (torch.arange(t.shape[dim], device=t.device) + step).clamp(0, t.shape[dim] - 1)
        return t.index_select(dim, idx)

    nb = sum(shifted(d, dim, s) for dim in range(3) for s in (-1, 1))
    return ((d - nb / 6) ** 2).sum()
