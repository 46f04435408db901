"""Thin object wrapper over the C ABI (libsdfrt.so).  PyTorch is used for device memory and
streams only; all rendering work happens in the hand-written CUDA kernels behind the C ABI."""
import ctypes as C

import numpy as np
import torch

from . import _capi
from ._capi import (INTEGRATOR_DIRECT, INTEGRATOR_PRB, INTEGRATOR_SILHOUETTE, INTEGRATOR_SIMPLE_SHADING, WARP_2D, WARP_DUMMY, WARP_NONE,
                    SdfrtError, check)

INTEGRATOR_IDS = {'sdf_silhouette_reparam': INTEGRATOR_SILHOUETTE,
                  'sdf_simple_shading_reparam': INTEGRATOR_SIMPLE_SHADING,
                  'sdf_direct_reparam': INTEGRATOR_DIRECT,
                  'sdf_prb_reparam': INTEGRATOR_PRB}


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _f32c(t, device):
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.device == device):
        raise SdfrtError("expected a contiguous float32 CUDA tensor on " + str(device))
    return t


class Context:
    """One sdfrt_ctx per (process, device)."""

    def __init__(self, device=0):
        self.lib = _capi.load()
        if not torch.cuda.is_available():
            raise SdfrtError("no CUDA device: libsdfrt has no CPU fallback")
        self.device = torch.device('cuda', device if isinstance(device, int) else device.index or 0)
        st = C.c_int(0)
        self.ctx = self.lib.sdfrt_create(self.device.index, C.byref(st))
        if not self.ctx:
            raise SdfrtError(f"sdfrt_create failed ({st.value}): need an sm_100 GPU")
        self._keep = {}
        self.sensor = _capi.SensorParams()
        self.lib.sdfrt_sensor_defaults(C.byref(self.sensor))

    def __del__(self):
        try:
            if getattr(self, 'ctx', None):
                self.lib.sdfrt_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # ---- scene state ------------------------------------------------------------------------
    def set_sdf(self, data, p=(0.0, 0.0, 0.0)):
        """data: (Z, Y, X[, 1]) float32 CUDA tensor (borrowed)."""
        _f32c(data, self.device)
        rz, ry, rx = data.shape[:3]
        pv = (C.c_float * 3)(*[float(v) for v in p])
        check(self.ctx, self.lib.sdfrt_set_sdf(self.ctx, _ptr(data), rx, ry, rz, pv))
        self._keep['sdf'] = data
        self.sdf_shape = (rz, ry, rx)

    def set_sdf_transform(self, to_world=None):
        """Grid3d's `transform` / integrator prop `sdf_to_world`: 4x4 affine (row-major) or None for the identity."""
        if to_world is None:
            check(self.ctx, self.lib.sdfrt_set_sdf_transform(self.ctx, None))
        else:
            m = (C.c_float * 16)(*[float(v) for v in np.asarray(to_world, dtype=np.float64).reshape(16)])
            check(self.ctx, self.lib.sdfrt_set_sdf_transform(self.ctx, m))

    def set_albedo(self, grid):
        if grid is None:
            check(self.ctx, self.lib.sdfrt_set_albedo(self.ctx, None, 0, 0, 0))
        else:
            _f32c(grid, self.device)
            rz, ry, rx, ch = grid.shape
            if ch != 3:
                raise SdfrtError("albedo grid must be (Z, Y, X, 3)")
            check(self.ctx, self.lib.sdfrt_set_albedo(self.ctx, _ptr(grid), rx, ry, rz))
        self._keep['albedo'] = grid

    def set_tracer(self, **kw):
        p = _capi.TracerParams()
        self.lib.sdfrt_tracer_defaults(C.byref(p))
        for k, v in kw.items():
            if not hasattr(p, k):
                raise SdfrtError(f"unknown tracer attribute {k}")
            setattr(p, k, v)
        check(self.ctx, self.lib.sdfrt_set_tracer(self.ctx, C.byref(p)))

    def set_warp(self, kind=WARP_2D, edge_eps=0.01, weight_strategy=6, clamping_thresh=0.05, normalize_warp_field=True,
                 max_reparam_depth=-1, n_aux_rays=16, kappa=1e5, power=3.0):
        p = _capi.WarpParams(kind, edge_eps, weight_strategy, clamping_thresh, int(normalize_warp_field), max_reparam_depth,
                             int(n_aux_rays), kappa, power)
        check(self.ctx, self.lib.sdfrt_set_warp(self.ctx, C.byref(p)))

    def set_sensor(self, to_world, width, height, fov=39.0, near_clip=1e-2, far_clip=1e4, filter_radius=2.0,
                   filter_stddev=0.5, sample_border=True):
        s = self.sensor
        s.to_world[:] = [float(v) for v in np.asarray(to_world, dtype=np.float64).reshape(-1)]
        s.fov_x_deg, s.near_clip, s.far_clip = fov, near_clip, far_clip
        s.width, s.height = int(width), int(height)
        s.filter_radius, s.filter_stddev, s.sample_border = filter_radius, filter_stddev, int(sample_border)
        check(self.ctx, self.lib.sdfrt_set_sensor(self.ctx, C.byref(s)))

    def set_shading(self, env_radiance=(1.0, 1.0, 1.0), env_dist=2.0 * 3.0 ** 0.5, albedo=(0.5, 0.5, 0.5), hide_emitters=False):
        p = _capi.ShadingParams()
        p.env_radiance[:] = list(env_radiance)
        p.env_dist = env_dist
        p.albedo[:] = list(albedo)
        p.hide_emitters = int(hide_emitters)
        check(self.ctx, self.lib.sdfrt_set_shading(self.ctx, C.byref(p)))

    def set_integrator(self, name_or_id):
        i = INTEGRATOR_IDS[name_or_id] if isinstance(name_or_id, str) else int(name_or_id)
        check(self.ctx, self.lib.sdfrt_set_integrator(self.ctx, i))

    def set_antithetic(self, enable):
        check(self.ctx, self.lib.sdfrt_set_antithetic(self.ctx, int(bool(enable))))

    def set_use_mis(self, enable):
        check(self.ctx, self.lib.sdfrt_set_use_mis(self.ctx, int(bool(enable))))

    def set_envmap(self, rgb=None, scale=1.0, to_world=None):
        """`envmap` emitter from an (H, W, 3) lat-long bitmap (numpy / CPU tensor; row 0 = +y pole), or None for the constant emitter
        of set_shading.  to_world: 4x4 (or 3x3) rotation of the emitter, None = identity.  sdf_direct_reparam only.  Copies the
        bitmap; synchronises the device."""
        if to_world is None:
            check(self.ctx, self.lib.sdfrt_set_envmap_transform(self.ctx, None))
        else:
            m = np.asarray(to_world, dtype=np.float64)
            m4 = np.eye(4)
            m4[:3, :3] = m[:3, :3]
            m4 = np.ascontiguousarray(m4, dtype=np.float32)
            check(self.ctx, self.lib.sdfrt_set_envmap_transform(self.ctx, m4.ctypes.data_as(C.c_void_p)))
        if rgb is None:
            check(self.ctx, self.lib.sdfrt_set_envmap_host(self.ctx, None, 0, 0, 1.0))
            return
        a = np.ascontiguousarray(rgb.detach().cpu().numpy() if torch.is_tensor(rgb) else rgb, dtype=np.float32)
        if a.ndim != 3 or a.shape[2] != 3:
            raise ValueError("envmap must have shape (H, W, 3)")
        check(self.ctx, self.lib.sdfrt_set_envmap_host(self.ctx, a.ctypes.data_as(C.c_void_p), a.shape[1], a.shape[0], float(scale)))

    def set_shadow_origin(self, mode):
        check(self.ctx, self.lib.sdfrt_set_shadow_origin(self.ctx, int(mode)))

    def set_prb(self, max_depth=4, rr_depth=5):
        p = _capi.PrbParams(int(max_depth), int(rr_depth))
        check(self.ctx, self.lib.sdfrt_set_prb(self.ctx, C.byref(p)))

    def set_partition(self, rank, world, rows=8):
        """Rank `rank` of `world` renders the sub-bands rank, rank + world, ... of `rows` film rows (1, 2, 4 or 8)."""
        check(self.ctx, self.lib.sdfrt_set_partition_rows(self.ctx, rank, world, rows))

    def set_lane_window(self, lane_lo=0, lane_hi=0):
        """Validation hook: adjoint_backprop only back-propagates wavefront lanes [lane_lo, lane_hi); (0, 0) = all."""
        check(self.ctx, self.lib.sdfrt_set_lane_window(self.ctx, int(lane_lo), int(lane_hi)))

    def set_empty_space_skip(self, enable=True):
        """enable=False marches every camera ray (reference behaviour); results are identical, only the work counters change."""
        check(self.ctx, self.lib.sdfrt_set_empty_space_skip(self.ctx, int(bool(enable))))

    def set_view_slot(self, slot):
        """Select the adjoint-state slot (view in flight) used by the following sweep / backprop / token calls."""
        check(self.ctx, self.lib.sdfrt_set_view_slot(self.ctx, int(slot)))

    def num_view_slots(self):
        return int(self.lib.sdfrt_num_view_slots())

    def adjoint_token(self):
        """Identifies the adjoint sweep 1 whose lane records the context holds (0: none valid)."""
        return int(self.lib.sdfrt_adjoint_token(self.ctx))

    # ---- rendering ----------------------------------------------------------------------------
    def film_dims(self):
        h, w, b = C.c_int(), C.c_int(), C.c_int()
        check(self.ctx, self.lib.sdfrt_film_dims(self.ctx, C.byref(h), C.byref(w), C.byref(b)))
        return h.value, w.value, b.value

    def wavefront_size(self, spp):
        return int(self.lib.sdfrt_wavefront_size(self.ctx, spp))

    def new_film(self):
        h, w, _ = self.film_dims()
        return torch.empty((h, w, 4), dtype=torch.float32, device=self.device)

    def render_primal(self, seed, spp, film=None):
        film = self.new_film() if film is None else film
        check(self.ctx, self.lib.sdfrt_render_primal(self.ctx, seed & 0xFFFFFFFF, spp, _ptr(film), _stream(self.device)))
        return film

    def develop(self, film, image=None):
        if image is None:
            image = torch.empty((self.sensor.height, self.sensor.width, 3), dtype=torch.float32, device=self.device)
        check(self.ctx, self.lib.sdfrt_develop(self.ctx, _ptr(film), _ptr(image), _stream(self.device)))
        return image

    def adjoint_accumulate(self, seed_grad, spp_grad, film=None):
        film = self.new_film() if film is None else film
        check(self.ctx, self.lib.sdfrt_adjoint_accumulate(self.ctx, seed_grad & 0xFFFFFFFF, spp_grad, _ptr(film),
                                                          _stream(self.device)))
        return film

    def render_view(self, seed, spp, seed_grad, spp_grad, film_primal=None, film_adjoint=None):
        """Primal render + adjoint sweep 1 in one pipelined call (they are independent); returns both film accumulators."""
        film_primal = self.new_film() if film_primal is None else film_primal
        film_adjoint = self.new_film() if film_adjoint is None else film_adjoint
        check(self.ctx, self.lib.sdfrt_render_view(self.ctx, seed & 0xFFFFFFFF, spp, _ptr(film_primal), seed_grad & 0xFFFFFFFF, spp_grad,
                                                   _ptr(film_adjoint), _stream(self.device)))
        return film_primal, film_adjoint

    def adjoint_backprop(self, film, grad_image, grad_sdf, grad_albedo=None, grad_p=None):
        _f32c(grad_image, self.device)
        if grad_sdf is not None:
            _f32c(grad_sdf, self.device)
        if grad_albedo is not None:
            _f32c(grad_albedo, self.device)
        check(self.ctx, self.lib.sdfrt_adjoint_backprop(self.ctx, _ptr(film), _ptr(grad_image), _ptr(grad_sdf),
                                                        _ptr(grad_albedo), _ptr(grad_p), _stream(self.device)))

    def stats(self, reset=True):
        s = _capi.Stats()
        check(self.ctx, self.lib.sdfrt_get_stats(self.ctx, C.byref(s), int(reset), _stream(self.device)))
        return s.as_dict()

    KERNEL_CLASSES = ('gen_primary', 'march_primary', 'shade', 'march_shadow', 'shadow_splat', 'film_adjoint', 'backward', 'develop')

    def set_profiling(self, enable=True):
        check(self.ctx, self.lib.sdfrt_set_profiling(self.ctx, int(enable)))

    def set_overlap(self, enable=True):
        check(self.ctx, self.lib.sdfrt_set_overlap(self.ctx, int(enable)))

    def timing(self, union=False):
        """{class: (milliseconds, launches)} of the launches since the last call (CUDA events on the launching stream); with
        union=True {class: (summed ms, launches, ms of the union of the launches' intervals)}."""
        ms = (C.c_float * 8)()
        n = (C.c_int * 8)()
        if not union:
            check(self.ctx, self.lib.sdfrt_get_timing(self.ctx, ms, n))
            return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.KERNEL_CLASSES)}
        un = (C.c_float * 9)()
        check(self.ctx, self.lib.sdfrt_get_timing_ex(self.ctx, ms, n, un))
        out = {k: (float(ms[i]), int(n[i]), float(un[i])) for i, k in enumerate(self.KERNEL_CLASSES)}
        out['march_all'] = (float(ms[1] + ms[3]), int(n[1] + n[3]), float(un[8]))
        return out

    # ---- test hooks -----------------------------------------------------------------------------
    def trace_rays(self, o, d, maxt=None, differentiable=True):
        _f32c(o, self.device)
        _f32c(d, self.device)
        n = o.shape[0]
        f = dict(dtype=torch.float32, device=self.device)
        out = dict(its_t=torch.empty(n, **f), warp_t=torch.empty(n, **f), warp_t_d=torch.empty(n, 3, **f),
                   warp_weight=torch.empty(n, **f), warp_weight_d=torch.empty(n, 3, **f),
                   n_steps=torch.empty(n, dtype=torch.int32, device=self.device))
        ro = _capi.RayOut(*[out[k].data_ptr() for k in ('its_t', 'warp_t', 'warp_t_d', 'warp_weight', 'warp_weight_d', 'n_steps')])
        check(self.ctx, self.lib.sdfrt_trace_rays(self.ctx, _ptr(o), _ptr(d), _ptr(maxt), n, int(differentiable),
                                                  C.byref(ro), _stream(self.device)))
        return out

    def eval_sdf(self, x, order=0):
        _f32c(x, self.device)
        n = x.shape[0]
        out = torch.empty((n, {0: 1, 1: 4, 2: 10}[order]), dtype=torch.float32, device=self.device)
        check(self.ctx, self.lib.sdfrt_eval_sdf(self.ctx, _ptr(x), n, order, _ptr(out), _stream(self.device)))
        return out

    # ---- per-iteration grid update (sdfrt_grid.cuh) -------------------------------------------------------------
    def redistance(self, phi, lower_bound=None, return_iters=False):
        """redistancing.redistance(max(phi, lower_bound)) (python/redistancing.py:4-13, python/variables.py:185-189)."""
        shape = phi.shape
        grid = phi.reshape(shape[:3]) if phi.dim() == 4 else phi
        _f32c(grid, self.device)
        if lower_bound is not None:
            lower_bound = lower_bound.reshape(grid.shape)
            _f32c(lower_bound, self.device)
        out = torch.empty_like(grid)
        rz, ry, rx = grid.shape
        iters = C.c_int(0)
        check(self.ctx, self.lib.sdfrt_redistance(self.ctx, _ptr(grid), _ptr(lower_bound) if lower_bound is not None else None,
                                                  _ptr(out), rx, ry, rz, C.byref(iters) if return_iters else None,
                                                  _stream(self.device)))
        out = out.reshape(shape)
        return (out, iters.value) if return_iters else out

    def redistance_status(self):
        """Waits for the last asynchronous redistance and raises SdfrtError if the solver aborted."""
        check(self.ctx, self.lib.sdfrt_redistance_status(self.ctx))

    def adam_step(self, param, grad, m, v, lr, t, beta1=0.9, beta2=0.999, eps=1e-8, grad_clamp=0.0, lo=1.0, hi=0.0):
        """In-place mi.ad.Adam.step on one tensor with the reference's gradient / value validation fused in."""
        for a in (param, grad, m, v):
            _f32c(a, self.device)
        check(self.ctx, self.lib.sdfrt_adam_step(self.ctx, _ptr(param), _ptr(grad), _ptr(m), _ptr(v), param.numel(), lr, beta1,
                                                 beta2, eps, int(t), grad_clamp, lo, hi, _stream(self.device)))

    def laplacian_reg(self, data, weight=1.0, grad=None, return_loss=True):
        """regularizations.eval_discrete_laplacian_reg (python/regularizations.py:5-25): returns weight * R as a 0-d float64
        tensor (if return_loss) and adds weight * dR/d(data) to `grad` (if given)."""
        grid = data.reshape(data.shape[:3]) if data.dim() == 4 else data
        _f32c(grid, self.device)
        rz, ry, rx = grid.shape
        loss = torch.zeros((), dtype=torch.float64, device=self.device) if return_loss else None
        if grad is not None:
            _f32c(grad, self.device)
        check(self.ctx, self.lib.sdfrt_laplacian_reg(self.ctx, _ptr(grid), rx, ry, rz, weight,
                                                     _ptr(loss) if loss is not None else None,
                                                     _ptr(grad) if grad is not None else None, _stream(self.device)))
        return loss

    def render_view_host(self, sdf_host, seed, spp, seed_grad, spp_grad, grad_image_host=None, image_host=None,
                         grad_sdf_host=None, albedo_host=None, grad_albedo_host=None):
        """End-to-end call with HOST (numpy / pinned torch CPU) buffers; with `albedo_host` the reflectance grid is a second
        uploaded parameter and `grad_albedo_host` (if given) receives its gradient."""
        def hp(a):
            if a is None:
                return None
            return C.c_void_p(a.data_ptr() if torch.is_tensor(a) else a.ctypes.data)
        rz, ry, rx = sdf_host.shape[:3]
        if image_host is None:
            image_host = torch.empty((self.sensor.height, self.sensor.width, 3), dtype=torch.float32).pin_memory()
        if grad_sdf_host is None:
            grad_sdf_host = torch.empty((rz, ry, rx), dtype=torch.float32).pin_memory()
        az, ay, ax = (albedo_host.shape[:3] if albedo_host is not None else (0, 0, 0))
        check(self.ctx, self.lib.sdfrt_render_view_host_ex(self.ctx, hp(sdf_host), rx, ry, rz, hp(albedo_host), ax, ay, az,
                                                           seed & 0xFFFFFFFF, spp, seed_grad & 0xFFFFFFFF, spp_grad,
                                                           hp(grad_image_host), hp(image_host), hp(grad_sdf_host), hp(grad_albedo_host)))
        return image_host, grad_sdf_host
