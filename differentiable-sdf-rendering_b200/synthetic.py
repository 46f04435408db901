"""Deterministic synthetic stand-ins for the reference's scenes (which are a separate download,
README.md:66-68): analytic SDFs sampled like create_sphere_sdf (python/shapes.py:557-560) and clamped by
the bbox constraint of python/variables.py:161-166,185-187; a seeded reflectance grid."""
import numpy as np


def _grid_points(res):
    """Nodes at linspace(0,1,R) like create_sphere_sdf (python/shapes.py:557-560)."""
    lin = np.linspace(0, 1, res)
    z, y, x = np.meshgrid(lin, lin, lin, indexing='ij')
    return x, y, z


def box_sdf_grid(res, extents=0.49, smoothing=0.01):
    """variables.update_box_sdf (python/variables.py:161-166) with BoxSDF (python/shapes.py:548-550)."""
    lin = np.linspace(-0.5, 0.5, res)
    z, y, x = np.meshgrid(lin, lin, lin, indexing='ij')
    q = np.stack([np.abs(x), np.abs(y), np.abs(z)], -1) - extents
    return np.linalg.norm(np.maximum(q, 0.0), axis=-1) + np.minimum(q.max(-1), 0.0) - smoothing


def make_sdf(kind, res):
    """'sphere' (python/shapes.py:557: c=.5, r=.3), 'knot' (torus-knot tube), 'blobs' (smooth union)."""
    x, y, z = _grid_points(res)
    if kind == 'sphere':
        sdf = np.sqrt((x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.5) ** 2) - 0.3
    elif kind == 'knot':
        # (2,3) torus knot sampled as a polyline; tube radius 0.045; conservative distance
        s = np.linspace(0, 2 * np.pi, 400, endpoint=False)
        R, r = 0.24, 0.10
        cx = 0.5 + (R + r * np.cos(3 * s)) * np.cos(2 * s)
        cy = 0.5 + r * np.sin(3 * s) * 1.3
        cz = 0.5 + (R + r * np.cos(3 * s)) * np.sin(2 * s)
        P = np.stack([cx, cy, cz], -1)
        pts = np.stack([x.ravel(), y.ravel(), z.ravel()], -1)
        best = np.full(pts.shape[0], np.inf)
        for k in range(0, P.shape[0], 8):
            dd = np.linalg.norm(pts[:, None, :] - P[None, k:k + 8, :], axis=-1).min(1)
            best = np.minimum(best, dd)
        sdf = (best - 0.045).reshape(x.shape)
    elif kind == 'blobs':
        rng = np.random.default_rng(0)
        c = rng.uniform(0.25, 0.75, size=(24, 3))
        rad = rng.uniform(0.05, 0.13, size=24)
        k = 0.03
        acc = np.zeros_like(x)
        for ci, ri in zip(c, rad):
            dd = np.sqrt((x - ci[0]) ** 2 + (y - ci[1]) ** 2 + (z - ci[2]) ** 2) - ri
            acc += np.exp(-dd / k)
        sdf = -k * np.log(acc)
    else:
        raise ValueError(kind)
    sdf = np.maximum(sdf, box_sdf_grid(res))  # bbox constraint, python/variables.py:185-187
    return sdf.astype(np.float32)


def make_albedo_grid(res, seed=1):
    """(res, res, res, 3) reflectance in [0.2, 0.8]."""
    rng = np.random.default_rng(seed)
    return rng.uniform(0.2, 0.8, size=(res,) * 3 + (3,)).astype(np.float32)


def make_sdf_device(kind, res, device):
    """Same shapes as make_sdf, evaluated with torch on `device` (fast path for large grids; values agree
    with make_sdf to float32 round-off).  Only 'sphere' and 'blobs' are provided."""
    import torch
    lin = torch.linspace(0, 1, res, dtype=torch.float64, device=device)
    z, y, x = torch.meshgrid(lin, lin, lin, indexing='ij')
    if kind == 'sphere':
        sdf = torch.sqrt((x - 0.5) ** 2 + (y - 0.5) ** 2 + (z - 0.5) ** 2) - 0.3
    elif kind == 'blobs':
        rng = np.random.default_rng(0)
        c = rng.uniform(0.25, 0.75, size=(24, 3))
        rad = rng.uniform(0.05, 0.13, size=24)
        k = 0.03
        acc = torch.zeros_like(x)
        for ci, ri in zip(c, rad):
            dd = torch.sqrt((x - ci[0]) ** 2 + (y - ci[1]) ** 2 + (z - ci[2]) ** 2) - ri
            acc += torch.exp(-dd / k)
        sdf = -k * torch.log(acc)
    else:
        raise ValueError(kind)
    linb = torch.linspace(-0.5, 0.5, res, dtype=torch.float64, device=device)
    bz, by, bx = torch.meshgrid(linb, linb, linb, indexing='ij')
    q = torch.stack([bx.abs(), by.abs(), bz.abs()], -1) - 0.49
    box = torch.linalg.norm(q.clamp_min(0.0), dim=-1) + q.max(-1).values.clamp_max(0.0) - 0.01
    return torch.maximum(sdf, box).float().contiguous()
