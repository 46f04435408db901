"""CPU restatements (numpy, float64 unless noted) of the small per-iteration grid operations next to the re-distancing:
the Adam step with the reference's gradient / value validation and the discrete Laplacian regulariser.
TEST INFRASTRUCTURE ONLY (tests/ compare the CUDA kernels of csrc/sdfrt_grid.cuh with these); PARITY UNPINNED for the Adam
part: `mi.ad.Adam` lives in Mitsuba (not installable offline, version unpinned); its update rule is restated from the
published source of Mitsuba 3.x (`mitsuba/python/ad/optimizers.py`), see oracle/ref_semantics.md S11."""
import numpy as np


def validate_gradient(g, r=1e-1):
    """SdfVariable.validate_gradient, python/variables.py:193-199: NaN -> 0, clamp to [-r, r]."""
    g = np.asarray(g, dtype=np.float64)
    return np.where(np.isnan(g), 0.0, np.clip(g, -r, r))


def adam_step(p, g, m, v, lr, t, beta_1=0.9, beta_2=0.999, epsilon=1e-8, grad_clamp=0.0, lo=1.0, hi=0.0):
    """One `mi.ad.Adam.step()` for one tensor (t = step count after the increment, i.e. 1 for the first step).
    Returns (p, g, m, v) as float64 arrays."""
    p, m, v = (np.asarray(a, dtype=np.float64) for a in (p, m, v))
    g = validate_gradient(g, grad_clamp) if grad_clamp > 0 else np.asarray(g, dtype=np.float64)
    lr_t = lr * np.sqrt(1 - beta_2 ** t) / (1 - beta_1 ** t)
    m = beta_1 * m + (1 - beta_1) * g
    v = beta_2 * v + (1 - beta_2) * g * g
    p = p - lr_t * m / (np.sqrt(v) + epsilon)
    if lo <= hi:
        p = np.clip(p, lo, hi)        # VolumeVariable.validate, python/variables.py:114-117
    return p, g, m, v


def _neighbour_sum(d):
    out = np.zeros_like(d)
    for axis in range(3):
        n = d.shape[axis]
        for step in (-1, 1):
            idx = np.clip(np.arange(n) + step, 0, n - 1)
            out += np.take(d, idx, axis=axis)
    return out


def laplacian_reg(data):
    """regularizations.eval_discrete_laplacian_reg, python/regularizations.py:5-25 -> (value, d value / d data), float64."""
    d = np.asarray(data, dtype=np.float64)
    d = d[..., 0] if d.ndim == 4 else d
    r = d - _neighbour_sum(d) / 6
    # clamped neighbour access is symmetric (node a reads node b exactly as often as b reads a), so the transpose of the
    # neighbour sum is the neighbour sum itself
    grad = 2 * r - _neighbour_sum(r) / 3
    return float((r * r).sum()), grad
