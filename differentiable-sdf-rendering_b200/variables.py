"""python/variables.py: the optimised variables and their per-iteration validation.

Free functions (LR schedule, gradient clamp, bbox constraint, EMA, x2 upsampling) plus the `Variable` / `VolumeVariable` /
`SdfVariable` classes with the reference's interface.  `opt` is an `optimizers.Adam` (key -> CUDA tensor).  The heavy steps
run in libsdfrt.so: cubic x2 upsampling (`sdfrt_eval_sdf`), max(sdf, bbox_sdf) + re-distancing (`sdfrt_redistance`)."""
import os

import numpy as np
import torch

from . import redistancing, synthetic, volio


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


def upsample_grid(data):
    """python/variables.py:25-26 (`dr.upsample` of a Texture3f by 2 per axis): trilinear resampling at the cell centres of
    the x2 grid with clamped borders.  data: (Z, Y, X, C) tensor."""
    t = data.permute(3, 0, 1, 2)[None]
    up = torch.nn.functional.interpolate(t, scale_factor=2, mode='trilinear', align_corners=False)
    return up[0].permute(1, 2, 3, 0).contiguous()


class Variable:
    """python/variables.py:39-78."""

    def __init__(self, k, beta=None, regularizer_weight=0.0, regularizer=None, lr=None):
        self.k = k
        self.mean = None
        self.beta = beta
        self.regularizer_weight = regularizer_weight
        self.regularizer = regularizer
        self.lr = None  # (sic) python/variables.py:48

    def initialize(self, opt):
        return

    def save(self, opt, output_dir, suffix):
        return

    def restore(self, opt, output_dir, suffix):
        return

    def validate_gradient(self, opt, i):
        return

    def validate(self, opt, i):
        return

    def update_mean(self, opt, i):
        return

    def load_mean(self, opt):
        if self.mean is not None:
            opt[self.k] = self.mean

    def eval_regularizer(self, opt, sdf_object, i):
        if self.regularizer is not None:
            return self.regularizer_weight * self.regularizer(opt[self.k])
        return 0.0


class VolumeVariable(Variable):
    """python/variables.py:81-132."""

    def __init__(self, k, shape, init_value=0.5, upsample_iter=(64, 128), device=0, **kwargs):
        super().__init__(k, **kwargs)
        self.shape = np.array(shape)
        self.init_value = init_value
        self.upsample_iter = upsample_iter
        self.device = redistancing.shared_context(device).device   # the rendering device (cuda:<device>)
        if self.upsample_iter is not None:
            self.upsample_iter = list(self.upsample_iter)
            for i in range(3):
                self.shape[i] = self.shape[i] // 2 ** len(self.upsample_iter)

    def initialize(self, opt):
        opt[self.k] = torch.full(tuple(int(v) for v in self.shape), float(self.init_value), dtype=torch.float32, device=self.device)
        if self.lr is not None:
            opt.set_learning_rate({self.k: self.lr})

    def get_variable_path(self, output_dir, suffix, suffix2=''):
        suffix_str = f'{suffix:04d}' if isinstance(suffix, int) else suffix
        return os.path.join(output_dir, f'{self.k.replace(".", "-")}-{suffix_str}{suffix2}.vol')

    def save(self, opt, output_dir, suffix):
        volio.write_vol(self.get_variable_path(output_dir, suffix), opt[self.k].detach().cpu().numpy())

    def restore(self, opt, output_dir, suffix):
        loaded = volio.read_vol(self.get_variable_path(output_dir, suffix))[0]
        if self.k in opt and opt[self.k].dim() == 3 and loaded.ndim == 4 and loaded.shape[-1] == 1:
            loaded = loaded[..., 0]
        opt[self.k] = torch.from_numpy(np.ascontiguousarray(loaded)).to(self.device)

    def value_clamp(self):
        k = self.k
        if k.endswith('reflectance.volume.data') or k.endswith('base_color.volume.data'):
            return (1e-5, 1.0)
        if k.endswith('roughness.volume.data'):
            return (0.1, 0.8)
        return None

    def validate(self, opt, i):
        k = self.k
        if self.upsample_iter is not None and i in self.upsample_iter:
            opt[k] = upsample_grid(opt[k].detach())
        c = self.value_clamp()
        if c is not None:
            with torch.no_grad():
                opt[k].clamp_(c[0], c[1])

    def update_mean(self, opt, i):
        if self.beta is None:
            return
        v = opt[self.k].detach()
        if self.mean is None or v.shape != self.mean.shape:
            self.mean = v.clone()
        else:
            self.mean = self.beta * self.mean + (1 - self.beta) * v


class SdfVariable(VolumeVariable):
    """python/variables.py:135-205.  `sdf_init_fn(shape)` returns a (Z, Y, X) array or tensor."""

    def __init__(self, k, resolution, sdf_init_fn=None, adaptive_learning_rate=True, **kwargs):
        super().__init__(k, shape=(resolution,) * 3, **kwargs)
        from .shapes import create_sphere_sdf
        self.adaptive_learning_rate = adaptive_learning_rate
        self.bbox_constraint = True
        self.sdf_init_fn = sdf_init_fn if sdf_init_fn is not None else (lambda shape: create_sphere_sdf(shape, redistance=True))
        if self.bbox_constraint:
            self.update_box_sdf(self.shape)
        self.lr_decay_rate = 0.02

    def initialize(self, opt):
        self.initial_lr = opt.lr.get(self.k, opt.lr_default)
        self.initial_shape = self.shape
        init = self.sdf_init_fn(tuple(int(v) for v in self.shape))
        init = torch.as_tensor(init, dtype=torch.float32).to(self.device)
        opt[self.k] = init[..., None] if init.dim() == 3 else init
        if self.lr is not None:
            opt.set_learning_rate({self.k: self.lr})

    def get_variable_path(self, output_dir, suffix, suffix2=''):
        k = self.k.replace("SamplingIntegrator.", "")
        suffix_str = f'{suffix:04d}' if isinstance(suffix, int) else suffix
        return os.path.join(output_dir, f'{k.replace(".", "-")}-{suffix_str}{suffix2}.vol')

    def update_box_sdf(self, res):
        """python/variables.py:161-166: BoxSDF(0, 0.49, smoothing 0.01) at linspace(-0.5, 0.5, res) per axis."""
        res = tuple(int(v) for v in res[:3])
        if not res[0] == res[1] == res[2]:
            raise ValueError("cubic grid expected")
        self.bbox_sdf = torch.from_numpy(synthetic.box_sdf_grid(res[0]).astype(np.float32)).to(self.device)[..., None]

    def validate(self, opt, i):
        k = self.k
        if self.upsample_iter is not None and i in self.upsample_iter:
            from .shapes import Grid3d
            sdf = upsample_sdf(Grid3d(opt[k].detach(), context=redistancing.shared_context(self.device.index)))
            self.shape = np.array(sdf.shape)
            if self.bbox_constraint:
                self.update_box_sdf(self.shape)
        else:
            self.shape = np.array(opt[k].shape)
            sdf = opt[k].detach()
        if self.adaptive_learning_rate and i is not None:
            lr_scale = 32 / self.shape[0]
            opt.set_learning_rate({k: lr_scale * simple_lr_decay(self.initial_lr, self.lr_decay_rate, i)})
        lower = None
        if self.bbox_constraint:
            assert tuple(sdf.shape) == tuple(self.bbox_sdf.shape)
            lower = self.bbox_sdf                        # max(sdf, bbox_sdf) is fused into the re-distancing call
        opt[k] = redistancing.redistance(sdf, lower_bound=lower)

    def validate_gradient(self, opt, i):
        g = opt[self.k].grad
        if g is not None:
            validate_gradient(g)

    def eval_regularizer(self, opt, sdf_object, i):
        if self.regularizer is not None and self.regularizer_weight > 0.0:
            return self.regularizer_weight * self.regularizer(opt[self.k], sdf_object)
        return 0.0
