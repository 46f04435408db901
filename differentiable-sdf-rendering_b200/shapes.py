"""Host-side mirror of the reference's SDF shape interface (python/shapes.py).

`Grid3d` owns the SDF voxel grid as a CUDA tensor and forwards every lookup / trace to the
CUDA kernels through the C ABI; attribute names, defaults and return conventions follow
python/shapes.py:25-41,115-288,375-483 so that callers (configs.get_warpfield, variables.upsample_sdf,
the integrators) find what they expect.
"""
import numpy as np
import torch

from . import engine


class Ray3f:
    """Minimal stand-in for mi.Ray3f: o, d (N,3) tensors and maxt (N,) or None."""

    def __init__(self, o, d, maxt=None):
        self.o, self.d, self.maxt = o, d, maxt


class SDFBase:
    """Tracer knobs of python/shapes.py:28-41 (same names, same defaults)."""

    def __init__(self):
        self.step_scale = 1.0
        self.trace_eps = 1e-6
        self.refine_intersection = True
        self.use_extra_weight = True
        self.extra_thresh = 0.05
        self.sil_weight_offset = 0.05
        self.sil_weight_epsilon = 1e-6
        self.weight_power = 3
        self.use_weight_sum_weight = True
        self.use_weight_ad = False  # python/shapes.py:41; the AD variant of the weights is not provided

    def tracer_params(self):
        if self.use_weight_ad:
            raise NotImplementedError("use_weight_ad=True (python/shapes.py:105-111) is not supported")
        return dict(step_scale=self.step_scale, trace_eps=self.trace_eps,
                    refine_intersection=int(self.refine_intersection), use_extra_weight=int(self.use_extra_weight),
                    extra_thresh=self.extra_thresh, sil_weight_offset=self.sil_weight_offset,
                    sil_weight_epsilon=self.sil_weight_epsilon, weight_power=int(self.weight_power),
                    use_weight_sum_weight=int(self.use_weight_sum_weight))


def atleast_4d(t):
    """python/util.py:219-223."""
    return t if t.dim() == 4 else t[..., None]


class Grid3d(SDFBase):
    """python/shapes.py:375-483.  `data`: (Z, Y, X[, 1]) array/tensor; `transform`: affine to_world (4x4) or None."""

    def __init__(self, data, transform=None, device=0, context=None):
        super().__init__()
        if isinstance(data, str):  # python/shapes.py:384-386: load a .vol grid and re-distance it
            from . import redistancing, volio
            data = redistancing.redistance(volio.read_vol(data)[0][..., 0])
        self.has_transform = transform is not None          # python/shapes.py:381
        self.to_world = np.eye(4) if transform is None else np.asarray(getattr(transform, 'matrix', transform), dtype=np.float64).reshape(4, 4)
        self.to_local = np.linalg.inv(self.to_world)
        self.ctx = context if context is not None else engine.Context(device)
        dev = self.ctx.device
        t = torch.as_tensor(np.asarray(data) if not torch.is_tensor(data) else data, dtype=torch.float32)
        self._data = atleast_4d(t).to(dev).contiguous().clone()
        self.p = [0.0, 0.0, 0.0]  # sdf.p (python/shapes.py:389)
        self.update()

    # -- parameter exposure (python/shapes.py:469-483) --
    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, value):
        self._data = atleast_4d(torch.as_tensor(value, dtype=torch.float32)).to(self.ctx.device).contiguous()
        self.update()

    @property
    def shape(self):
        return tuple(self._data.shape)

    def traverse(self, callback):
        callback.put_parameter("sdf.data", self._data, True)
        callback.put_parameter("sdf.p", self.p, True)

    def update(self):
        self.ctx.set_sdf(self._data, self.p)
        self.ctx.set_sdf_transform(self.to_world if self.has_transform else None)
        self.ctx.set_tracer(**self.tracer_params())

    def parameters_changed(self, keys=None):
        self.update()

    def bbox(self, expand=True):
        """python/shapes.py:394-403,416-418: AABB of the transformed unit cube, inflated by 0.05."""
        delta = 0.05 if expand else 0.0
        c = np.array([[x, y, z, 1.0] for x in (0, 1) for y in (0, 1) for z in (0, 1)]) @ self.to_world.T
        return tuple(c[:, :3].min(0) - delta), tuple(c[:, :3].max(0) + delta)

    # -- lookups (python/shapes.py:420-450) --
    def _pts(self, x):
        return torch.as_tensor(x, dtype=torch.float32, device=self.ctx.device).reshape(-1, 3).contiguous()

    def eval(self, x, detached=False):
        return self.ctx.eval_sdf(self._pts(x), 0)[:, 0]

    def eval_and_grad(self, x, detached=False):
        o = self.ctx.eval_sdf(self._pts(x), 1)
        return o[:, 0], o[:, 1:4]

    def eval_grad(self, x, detached=False):
        return self.ctx.eval_sdf(self._pts(x), 1)[:, 1:4]

    def eval_all(self, x, detach_w=False):
        o = self.ctx.eval_sdf(self._pts(x), 2)
        v, g = o[:, 0], o[:, 1:4]
        H = torch.stack([o[:, 4], o[:, 7], o[:, 8], o[:, 7], o[:, 5], o[:, 9], o[:, 8], o[:, 9], o[:, 6]], -1).reshape(-1, 3, 3)
        return v, v, g, g, H

    # -- tracers (python/shapes.py:115-345) --
    def ray_intersect(self, ray, warp=None, active=True, extra_outputs=None):
        """Returns (its_t, warp_t, warp_t_d, warp_weight, warp_weight_d) like python/shapes.py:288."""
        self.ctx.set_tracer(**self.tracer_params())
        o, d = self._pts(ray.o), self._pts(ray.d)
        maxt = None if ray.maxt is None else torch.as_tensor(ray.maxt, dtype=torch.float32, device=o.device).contiguous()
        diff = warp is not None and not getattr(warp, 'is_dummy', False)
        out = self.ctx.trace_rays(o, d, maxt, differentiable=diff)
        if extra_outputs is not None:
            extra_outputs['i'] = out['n_steps']
        if not diff:  # python/shapes.py:339
            z = torch.zeros_like(out['its_t'])
            return out['its_t'], z, torch.zeros_like(out['warp_t_d']), None, None
        return out['its_t'], out['warp_t'], out['warp_t_d'], out['warp_weight'], out['warp_weight_d']

    def ray_intersect_non_diff(self, ray, active=True):
        return self.ray_intersect(ray, warp=None)


class BoxSDF(SDFBase):
    """Smooth box SDF, python/shapes.py:540-554 (used as bbox constraint, python/variables.py:161-166)."""

    def __init__(self, p, extents, smoothing=0.01):
        super().__init__()
        self.p, self.extents, self.smoothing = p, extents, smoothing

    def eval(self, x, detached=False):
        x = torch.as_tensor(x)
        q = (x - torch.as_tensor(self.p, dtype=x.dtype, device=x.device)).abs() - torch.as_tensor(self.extents, dtype=x.dtype, device=x.device)
        return torch.linalg.norm(q.clamp_min(0.0), dim=-1) + q.max(-1).values.clamp_max(0.0) - self.smoothing


def create_sphere_sdf(res, center=(0.5, 0.5, 0.5), radius=0.3, redistance=True):
    """python/shapes.py:557-581 without noise.  Like the reference it finishes with `redistancing.redistance` (:580) and returns
    the re-distanced grid (a CUDA tensor: the pass runs on the GPU); `redistance=False` is the explicit opt-out that returns the
    analytic numpy grid (CPU-only tests)."""
    z, y, x = np.meshgrid(np.linspace(0, 1, res[0]), np.linspace(0, 1, res[1]), np.linspace(0, 1, res[2]), indexing='ij')
    pts = np.stack([x.ravel(), y.ravel(), z.ravel()], axis=1)
    sdf = np.reshape(np.linalg.norm(pts - np.asarray(center), axis=-1) - radius, res).astype(np.float32)
    if redistance:
        from . import redistancing
        return redistancing.redistance(sdf)
    return sdf


def create_block_sdf(resolution, center=(0.5, 0.5, 0.5)):
    """python/shapes.py:584-590: a +-1 occupancy block turned into an SDF by the redistancing pass (CUDA tensor)."""
    from . import redistancing
    r2 = resolution // 2
    signed_dist = np.ones([resolution] * 3, dtype=np.float32)
    signed_dist[r2 - r2 // 6:r2 + r2 // 6, r2 - r2 // 6:r2 + r2 // 6, r2 - r2 // 2:r2 + r2 // 2] = -1
    return redistancing.redistance(signed_dist)
