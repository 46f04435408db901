"""TEST / BENCH INFRASTRUCTURE (checker only, never on the product path).

Gradient of a lane subsample at full-size configurations.  The oracle (oracle/sdf_oracle.py, torch autograd, fp64) is far too
slow to back-propagate a 17 M-lane wavefront, but the adjoint is LINEAR in the per-lane contributions once the film adjoint is
fixed:  dL/dtheta = sum_lanes < F, d splat(lane)/dtheta >,  F = (dL/dV, dL/dW) of I = V / W on the completed accumulators
(python/integrators/reparam.py:183-190, HDRFilm::develop).  So for a window of wavefront lanes [lo, hi) the oracle evaluates
only those lanes (`render(..., lane_range=...)`, RNG streams keyed by the global lane index), contracts their accumulator with
F built from the completed film, and differentiates -- which is exactly what `sdfrt_adjoint_backprop` scatters for the same
lanes under `sdfrt_set_lane_window`."""
import numpy as np
import torch

from . import sdf_oracle as O


def film_adjoint(film, grad_image, border):
    """(dL/dV_rgb, dL/dW) per film pixel for I = V / W on the crop window, zero on the border.  film: (Hb, Wb, 4) float64."""
    Hb, Wb, _ = film.shape
    F = torch.zeros_like(film)
    b = border
    V = film[b:Hb - b, b:Wb - b, :3]
    W = film[b:Hb - b, b:Wb - b, 3]
    ok = W != 0
    iw = torch.where(ok, 1.0 / torch.where(ok, W, torch.ones_like(W)), torch.zeros_like(W))
    F[b:Hb - b, b:Wb - b, :3] = grad_image * iw[..., None]
    F[b:Hb - b, b:Wb - b, 3] = -(grad_image * V).sum(-1) * iw * iw
    return F


def centre_window(sensor, spp, n_lanes):
    """`n_lanes` consecutive wavefront lanes (whole pixels) centred on the middle of the film's middle row."""
    b = sensor.border
    Wb, Hb = sensor.width + 2 * b, sensor.height + 2 * b
    npx = max(1, n_lanes // spp)
    row = Hb // 2
    x0 = max(0, Wb // 2 - npx // 2)
    lo = (row * Wb + x0) * spp
    return lo, lo + npx * spp


def window_gradient(scene, warp_field, integrator, film_full, grad_image, seed_grad, spp, lane_range):
    """Returns (acc_window (Hb, Wb, 4), dL/d(sdf.data) (Z, Y, X), dL/d(albedo) or None) for the lanes of `lane_range`, with the
    film adjoint taken from `film_full` (the completed accumulator of the same re-render, any array-like of shape (Hb, Wb, 4))."""
    sdf = scene.sdf
    dt = sdf.dtype
    film_full = torch.as_tensor(np.asarray(film_full), dtype=dt)
    grad_image = torch.as_tensor(np.asarray(grad_image), dtype=dt)
    F = film_adjoint(film_full, grad_image, scene.sensor.border)
    data0, alb0 = sdf.data, scene.albedo_grid
    sdf.data = data0.detach().clone().requires_grad_(True)
    if alb0 is not None:
        scene.albedo_grid = alb0.detach().clone().requires_grad_(True)
    try:
        acc = O.render(scene, warp_field, integrator, seed_grad, spp, mode_backward=True, lane_range=lane_range, develop=False)
        wrt = [sdf.data] + ([scene.albedo_grid] if alb0 is not None else [])
        grads = torch.autograd.grad((acc * F).sum(), wrt, allow_unused=True)
        g = grads[0] if grads[0] is not None else torch.zeros_like(sdf.data)
        ga = None
        if alb0 is not None:
            ga = grads[1] if grads[1] is not None else torch.zeros_like(scene.albedo_grid)
        return acc.detach(), g.detach(), ga
    finally:
        sdf.data, scene.albedo_grid = data0, alb0


def gradient_errors(got, want):
    """The error figures reported next to the metric ("grad MSE vs ref"): MSE over all voxels, max-abs and L2 relative to the oracle."""
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    want = np.asarray(want, dtype=np.float64).reshape(-1)
    d = got - want
    return {"mse": float((d * d).mean()), "ref_mean_sq": float((want * want).mean()),
            "max_abs_over_ref_max": float(np.abs(d).max() / max(np.abs(want).max(), 1e-300)),
            "rel_l2": float(np.linalg.norm(d) / max(np.linalg.norm(want), 1e-300)),
            "ref_nonzero_voxels": int((want != 0).sum())}
