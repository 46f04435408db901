"""`mi.ad.Adam` as the optimisation loop uses it (python/shape_opt.py:53,102; python/variables.py:89-94,179-183):
a key -> tensor map with per-key learning rates whose `step()` runs `sdfrt_adam_step` (one fused pass over parameter,
gradient and both moments on the GPU).

Restated semantics of Mitsuba 3's Adam (third-party, unpinned -- see oracle/ref_semantics.md S11):
    t_k += 1;  lr_t = lr_k * sqrt(1 - beta_2^t) / (1 - beta_1^t)
    m = beta_1 m + (1 - beta_1) g;  v = beta_2 v + (1 - beta_2) g^2;  p = p - lr_t * m / (sqrt(v) + epsilon)
State and step count of a key are reset when the shape of its tensor changes (the x2 upsampling steps).
Parameters are torch leaf tensors whose `.grad` the rendering backward pass fills."""
import torch

from . import redistancing


class Adam:
    def __init__(self, lr, beta_1=0.9, beta_2=0.999, epsilon=1e-8, mask_updates=False, params=None):
        if mask_updates:
            raise NotImplementedError("mask_updates=True is not supported (the reference runs with False, python/configs.py:26)")
        self.lr_default = lr
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self.variables, self.lr, self.state, self.t = {}, {}, {}, {}
        for k, v in (params or {}).items():
            self[k] = v

    # -- mapping protocol (opt[k], k in opt, opt.keys()/items()) --
    def __contains__(self, k):
        return k in self.variables

    def __getitem__(self, k):
        return self.variables[k]

    def __setitem__(self, k, value):
        value = value.detach().to(torch.float32).contiguous().clone().requires_grad_(True)
        if k not in self.variables or self.variables[k].shape != value.shape:
            self.reset(k)
        self.variables[k] = value
        self.lr.setdefault(k, self.lr_default)

    def keys(self):
        return self.variables.keys()

    def items(self):
        return self.variables.items()

    def reset(self, k):
        self.state.pop(k, None)
        self.t[k] = 0

    def set_learning_rate(self, lr):
        """float: all keys; dict: per key (python/variables.py:183)."""
        if isinstance(lr, dict):
            self.lr.update(lr)
        else:
            self.lr_default = lr
            for k in self.lr:
                self.lr[k] = lr

    def step(self, grad_clamp=None, value_clamp=None):
        """One Adam step on every key that has a gradient.  `grad_clamp` {key: r} fuses SdfVariable.validate_gradient
        (NaN -> 0, clamp to [-r, r], python/variables.py:193-199), `value_clamp` {key: (lo, hi)} the clamps of
        VolumeVariable.validate (python/variables.py:114-117)."""
        for k, p in self.variables.items():
            g = p.grad
            if g is None:
                continue
            ctx = redistancing.shared_context(p.device.index or 0)
            if k not in self.state or self.state[k][0].shape != p.shape:
                self.reset(k)
                self.state[k] = (torch.zeros_like(p), torch.zeros_like(p))
            self.t[k] += 1
            m, v = self.state[k]
            lo, hi = (value_clamp or {}).get(k, (1.0, 0.0))
            with torch.no_grad():
                ctx.adam_step(p, g if g.is_contiguous() else g.contiguous(), m, v, self.lr[k], self.t[k], self.beta_1,
                              self.beta_2, self.epsilon, (grad_clamp or {}).get(k, 0.0), lo, hi)

    def zero_grad(self):
        for p in self.variables.values():
            p.grad = None
