"""ctypes binding of tests/host_harness (the per-lane kernel math compiled for the host).

Test infrastructure only: lets `-m "not gpu"` tests compare the CUDA path's arithmetic with the
oracle on a machine without a GPU.  The product (libsdfrt.so) never uses it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'host_harness', 'harness.cpp')
LIB = os.path.join(HERE, 'host_harness', 'libsdfrt_hostmath.so')


class TracerParams(C.Structure):
    _fields_ = [('step_scale', C.c_float), ('trace_eps', C.c_float), ('refine_intersection', C.c_int32),
                ('use_extra_weight', C.c_int32), ('extra_thresh', C.c_float), ('sil_weight_offset', C.c_float),
                ('sil_weight_epsilon', C.c_float), ('weight_power', C.c_int32), ('use_weight_sum_weight', C.c_int32),
                ('max_steps', C.c_int32)]


class WarpParams(C.Structure):
    _fields_ = [('kind', C.c_int32), ('edge_eps', C.c_float), ('weight_strategy', C.c_int32),
                ('clamping_thresh', C.c_float), ('normalize_warp_field', C.c_int32), ('max_reparam_depth', C.c_int32),
                ('n_aux_rays', C.c_int32), ('kappa', C.c_float), ('power', C.c_float)]


class SensorParams(C.Structure):
    _fields_ = [('to_world', C.c_float * 16), ('fov_x_deg', C.c_float), ('near_clip', C.c_float), ('far_clip', C.c_float),
                ('width', C.c_int32), ('height', C.c_int32), ('filter_radius', C.c_float), ('filter_stddev', C.c_float),
                ('sample_border', C.c_int32)]


class ShadingParams(C.Structure):
    _fields_ = [('env_radiance', C.c_float * 3), ('env_dist', C.c_float), ('albedo', C.c_float * 3),
                ('hide_emitters', C.c_int32)]


class PrbParams(C.Structure):
    _fields_ = [('max_depth', C.c_int32), ('rr_depth', C.c_int32)]


class HHScene(C.Structure):
    _fields_ = [('sdf', C.c_void_p), ('rx', C.c_int), ('ry', C.c_int), ('rz', C.c_int), ('p', C.c_float * 3),
                ('albedo', C.c_void_p), ('ax', C.c_int), ('ay', C.c_int), ('az', C.c_int),
                ('tracer', TracerParams), ('warp', WarpParams), ('sensor', SensorParams), ('shading', ShadingParams),
                ('integrator', C.c_int), ('prb', PrbParams), ('antithetic', C.c_int), ('has_transform', C.c_int),
                ('to_world', C.c_float * 16), ('shadow_origin', C.c_int), ('use_mis', C.c_int), ('envmap', C.c_void_p), ('env_has_rot', C.c_int), ('env_R', C.c_float * 9)]


def build():
    srcs = [SRC] + [os.path.join(HERE, '..', 'differentiable-sdf-rendering_b200', 'csrc', f)
                    for f in ('sdfrt_core.cuh', 'sdfrt_lane.cuh', 'sdfrt_prb.cuh', 'sdfrt_params.h', 'sdfrt_grid.cuh')]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-x', 'c++', SRC, '-o', LIB])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def fptr(a):
    return a.ctypes.data_as(C.c_void_p)


class HostScene:
    """Owns the numpy buffers referenced by an HHScene."""

    def __init__(self, sdf, sensor_to_world, width, height, integrator=2, albedo_grid=None, p=(0, 0, 0), sdf_to_world=None):
        self.h = HHScene()
        lib().hh_defaults(C.byref(self.h))
        self.sdf = np.ascontiguousarray(sdf, dtype=np.float32)
        rz, ry, rx = self.sdf.shape
        self.h.sdf = self.sdf.ctypes.data
        self.h.rx, self.h.ry, self.h.rz = rx, ry, rz
        self.h.p[:] = list(p)
        self.albedo = None
        if albedo_grid is not None:
            self.albedo = np.ascontiguousarray(albedo_grid, dtype=np.float32)
            self.h.albedo = self.albedo.ctypes.data
            self.h.az, self.h.ay, self.h.ax = self.albedo.shape[:3]
        tw = np.asarray(sensor_to_world, dtype=np.float32).reshape(-1)
        self.h.sensor.to_world[:] = list(tw)
        self.h.sensor.width, self.h.sensor.height = width, height
        self.h.integrator = integrator
        if sdf_to_world is not None:
            self.h.has_transform = 1
            self.h.to_world[:] = [float(v) for v in np.asarray(sdf_to_world, dtype=np.float64).reshape(16)]
        self.border = 2
        self._env = None

    def set_envmap(self, rgb, scale=1.0, to_world=None):
        """`envmap` emitter from an (H, W, 3) bitmap (None: back to the constant emitter); to_world: 3x3 rotation or None."""
        self.h.env_has_rot = 0 if to_world is None else 1
        self._env_R = None if to_world is None else np.ascontiguousarray(np.asarray(to_world, dtype=np.float64)[:3, :3], dtype=np.float32)
        if to_world is not None:
            self.h.env_R[:] = [float(v) for v in self._env_R.reshape(-1)]
        L = lib()
        L.hh_envmap_create.restype = C.c_void_p
        if self._env is not None:
            L.hh_envmap_destroy(C.c_void_p(self._env))
            self._env = None
        if rgb is not None:
            a = np.ascontiguousarray(rgb, dtype=np.float32)
            self._env = L.hh_envmap_create(fptr(a), a.shape[1], a.shape[0], C.c_float(scale))
            if not self._env:
                raise ValueError('invalid envmap')
        self.h.envmap = self._env

    def envmap_query(self, d, u):
        d = np.ascontiguousarray(d, dtype=np.float32)
        u = np.ascontiguousarray(u, dtype=np.float32)
        out = np.zeros((d.shape[0], 17), np.float32)
        R = getattr(self, '_env_R', None)
        lib().hh_envmap_query(C.c_void_p(self._env), None if R is None else fptr(R), fptr(d), fptr(u), d.shape[0], fptr(out))
        return dict(Le=out[:, :3], pdf=out[:, 3], d=out[:, 4:7], d_pdf=out[:, 7], dLe=out[:, 8:17].reshape(-1, 3, 3))

    @property
    def film_shape(self):
        b = self.border if self.h.sensor.sample_border else 0
        return (self.h.sensor.height + 2 * b, self.h.sensor.width + 2 * b, 4)

    def wavefront(self, spp):
        fs = self.film_shape
        return fs[0] * fs[1] * spp

    def eval_sdf(self, x, order):
        x = np.ascontiguousarray(x, dtype=np.float32)
        n = x.shape[0]
        out = np.zeros((n, {0: 1, 1: 4, 2: 10}[order]), dtype=np.float32)
        lib().hh_eval_sdf(C.byref(self.h), fptr(x), n, order, fptr(out))
        return out

    def trace(self, o, d, maxt=None, differentiable=True):
        o = np.ascontiguousarray(o, dtype=np.float32)
        d = np.ascontiguousarray(d, dtype=np.float32)
        n = o.shape[0]
        its_t = np.zeros(n, np.float32); warp_t = np.zeros(n, np.float32); ww = np.zeros(n, np.float32)
        wtd = np.zeros((n, 3), np.float32); wwd = np.zeros((n, 3), np.float32); ns = np.zeros(n, np.int32)
        mt = None if maxt is None else np.ascontiguousarray(maxt, dtype=np.float32)
        lib().hh_trace(C.byref(self.h), fptr(o), fptr(d), None if mt is None else fptr(mt), n, int(differentiable),
                       fptr(its_t), fptr(warp_t), fptr(wtd), fptr(ww), fptr(wwd), fptr(ns))
        return dict(its_t=its_t, warp_t=warp_t, warp_t_d=wtd, warp_weight=ww, warp_weight_d=wwd, n_steps=ns)

    def local_lanes(self, spp, rank=0, world=1):
        lib().hh_local_lanes.restype = C.c_ulonglong
        return int(lib().hh_local_lanes(C.byref(self.h), spp, rank, world))

    def forward(self, seed, spp, diff, rank=0, world=1):
        film = np.zeros(self.film_shape, np.float32)
        n = self.local_lanes(spp, rank, world)
        rec = np.zeros((lib().hh_rec_count(), max(n, 1)), np.float32) if diff else np.zeros((1, 1), np.float32)
        stats = np.zeros(8, np.uint64)
        lib().hh_forward(C.byref(self.h), C.c_uint(seed), spp, int(diff), fptr(film), fptr(rec), fptr(stats), rank, world)
        return film, rec, stats

    def develop(self, film):
        img = np.zeros((self.h.sensor.height, self.h.sensor.width, 3), np.float32)
        lib().hh_develop(C.byref(self.h), fptr(film), fptr(img))
        return img

    def backward(self, seed, spp, film, rec, grad_image, rank=0, world=1, lane_window=None):
        lo, hi = lane_window if lane_window is not None else (0, 0)
        lib().hh_set_lane_window(C.c_ulonglong(lo), C.c_ulonglong(hi))
        grad_image = np.ascontiguousarray(grad_image, dtype=np.float32)
        g = np.zeros_like(self.sdf)
        ga = None if self.albedo is None else np.zeros_like(self.albedo)
        gp = np.zeros(3, np.float32)
        stats = np.zeros(8, np.uint64)
        lib().hh_backward(C.byref(self.h), C.c_uint(seed), spp, fptr(film), fptr(rec), fptr(grad_image), fptr(g),
                          None if ga is None else fptr(ga), fptr(gp), fptr(stats), rank, world)
        return g, ga, gp, stats

    # ---- sdf_prb_reparam (integrator 3) ----
    def prb_forward(self, seed, spp, rank=0, world=1):
        film = np.zeros(self.film_shape, np.float32)
        n = self.local_lanes(spp, rank, world)
        L = np.zeros((3, max(n, 1)), np.float32)
        lib().hh_prb_forward(C.byref(self.h), C.c_uint(seed), spp, fptr(film), fptr(L), rank, world)
        return film, L

    def prb_backward(self, seed, spp, film, L, grad_image, rank=0, world=1):
        grad_image = np.ascontiguousarray(grad_image, dtype=np.float32)
        g = np.zeros_like(self.sdf)
        ga = None if self.albedo is None else np.zeros_like(self.albedo)
        gp = np.zeros(3, np.float32)
        stats = np.zeros(8, np.uint64)
        lib().hh_prb_backward(C.byref(self.h), C.c_uint(seed), spp, fptr(film), fptr(L), fptr(grad_image), fptr(g),
                              None if ga is None else fptr(ga), fptr(gp), fptr(stats), rank, world)
        return g, ga, gp, stats

    # wavefront form (records + stop hooks), emulated on the host
    def prb_wf_forward(self, seed, spp, diff, rank=0, world=1):
        film = np.zeros(self.film_shape, np.float32)
        n = self.local_lanes(spp, rank, world)
        L = np.zeros((3, max(n, 1)), np.float32)
        recs = np.zeros((self.h.prb.max_depth * 18, max(n, 1)), np.float32)
        lib().hh_prb_wf_forward(C.byref(self.h), C.c_uint(seed), spp, int(diff), fptr(film), fptr(L), fptr(recs), rank, world)
        return film, L, recs

    def prb_wf_backward(self, seed, spp, film, L, recs, grad_image, rank=0, world=1):
        grad_image = np.ascontiguousarray(grad_image, dtype=np.float32)
        g = np.zeros_like(self.sdf)
        ga = None if self.albedo is None else np.zeros_like(self.albedo)
        gp = np.zeros(3, np.float32)
        lib().hh_prb_wf_backward(C.byref(self.h), C.c_uint(seed), spp, fptr(film), fptr(L), fptr(recs), fptr(grad_image), fptr(g),
                                 None if ga is None else fptr(ga), fptr(gp), rank, world)
        return g, ga, gp
