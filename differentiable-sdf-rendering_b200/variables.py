"""Host-side pieces of python/variables.py that sit next to the hot path (minus redistancing):
LR schedule, gradient clamp, bbox constraint, EMA of the parameters, x2 upsampling through the cubic lookup."""
import numpy as np
import torch

from . import synthetic


def simple_lr_decay(initial_lrate, decay, i):
    """python/variables.py:28-36."""
    lr = initial_lrate / (1 + decay * i)
    if i > 480:
        lr = lr / 2
    if i > 500:
        lr = lr / 2
    return lr


def adaptive_sdf_lr(initial_lr, res, i, decay_rate=0.02):
    """python/variables.py:179-183: scale by 32 / resolution."""
    return 32 / res * simple_lr_decay(initial_lr, decay_rate, i)


def validate_gradient(grad, r=1e-1):
    """python/variables.py:193-199: clamp to [-r, r], NaN -> 0 (in place)."""
    grad.nan_to_num_(nan=0.0).clamp_(-r, r)
    return grad


def bbox_constraint(sdf, box=None):
    """python/variables.py:161-166,185-187: sdf = max(sdf, BoxSDF(0, 0.49, smoothing 0.01) sampled on the grid)."""
    d = sdf[..., 0] if sdf.dim() == 4 else sdf
    if box is None:
        if not (d.shape[0] == d.shape[1] == d.shape[2]):
            raise ValueError("cubic grid expected")
        box = torch.from_numpy(synthetic.box_sdf_grid(d.shape[0]).astype(np.float32)).to(d.device)
    out = torch.maximum(d, box)
    return out[..., None] if sdf.dim() == 4 else out


class ParameterMean:
    """python/variables.py:121-132: exponential moving average of a parameter tensor."""

    def __init__(self, beta):
        self.beta, self.mean = beta, None

    def update(self, value):
        v = value.detach()
        if self.beta is None:
            return
        if self.mean is None or self.mean.shape != v.shape:
            self.mean = v.clone()
        else:
            self.mean = self.beta * self.mean + (1 - self.beta) * v


def upsample_sdf(sdf_object):
    """python/variables.py:18-23: evaluate the cubic interpolant at the cell centres of the x2 grid (runs on the GPU
    through sdfrt_eval_sdf).  Returns a (2Z, 2Y, 2X, 1) tensor on the SDF's device."""
    z, y, x = sdf_object.shape[:3]
    dev = sdf_object.ctx.device
    cz = (torch.arange(2 * z, device=dev, dtype=torch.float32) + 0.5) / (2 * z)
    cy = (torch.arange(2 * y, device=dev, dtype=torch.float32) + 0.5) / (2 * y)
    cx = (torch.arange(2 * x, device=dev, dtype=torch.float32) + 0.5) / (2 * x)
    Z, Y, X = torch.meshgrid(cz, cy, cx, indexing='ij')
    pts = torch.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], -1).contiguous()
    return sdf_object.eval(pts).reshape(2 * z, 2 * y, 2 * x, 1)
