"""ctypes driver of oracle/cpu_port.cpp (compiled, all-cores CPU build of the per-lane restatement).
Bench infrastructure only (bench.py `cpu_baseline` and `--impl reference`)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, 'cpu_port.cpp')
LIB = os.path.join(HERE, 'libsdfrt_cpu_port.so')
FLAGS = ['-O3', '-march=x86-64-v3', '-fopenmp', '-std=c++17', '-fPIC', '-shared', '-x', 'c++']

sys.path.insert(0, os.path.join(ROOT, 'tests'))
from hostmath import PrbParams, SensorParams, ShadingParams, TracerParams, WarpParams  # noqa: E402  (same public structs)


class CPScene(C.Structure):
    _fields_ = [('sdf', C.c_void_p), ('rx', C.c_int), ('ry', C.c_int), ('rz', C.c_int),
                ('albedo', C.c_void_p), ('ax', C.c_int), ('ay', C.c_int), ('az', C.c_int),
                ('tracer', TracerParams), ('warp', WarpParams), ('sensor', SensorParams), ('shading', ShadingParams),
                ('integrator', C.c_int), ('prb', PrbParams)]


def build():
    deps = [SRC] + [os.path.join(ROOT, 'differentiable-sdf-rendering_b200', 'csrc', f)
                    for f in ('sdfrt_core.cuh', 'sdfrt_lane.cuh', 'sdfrt_prb.cuh', 'sdfrt_params.h')]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    subprocess.check_call(['g++'] + FLAGS + [SRC, '-o', LIB])
    return LIB


_NATIVE = {}


def build_native():
    """`-march=native` build for the host that runs the timing.  The portable x86-64-v3 build above is
    what travels with the repository snapshot; this one is compiled on the timing host itself (g++ is part of the image) into
    oracle/_native/ and falls back to the portable build if the compiler is missing.  Returns (path, flags description)."""
    if 'lib' in _NATIVE:
        return _NATIVE['lib']
    portable = build()
    out_dir = os.path.join(HERE, '_native')
    out = os.path.join(out_dir, 'libsdfrt_cpu_port_native.so')
    res = (portable, '-O3 -march=x86-64-v3 -fopenmp')
    try:
        os.makedirs(out_dir, exist_ok=True)
        flags = [f if f != '-march=x86-64-v3' else '-march=native' for f in FLAGS]
        subprocess.check_call(['g++'] + flags + [SRC, '-o', out], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        C.CDLL(out)
        res = (out, '-O3 -march=native -fopenmp')
    except Exception:
        pass
    _NATIVE['lib'] = res
    return res


def time_sample(sdf, albedo, to_world, W, H, seed, spp_primal, spp_adjoint, n_primal, n_adjoint, prb=None, threads=None):
    """Times `n_primal` primal + `n_adjoint` adjoint lanes strided over the whole wavefronts on `threads` OpenMP threads (default:
    the affinity mask -- set explicitly because launchers such as torchrun export OMP_NUM_THREADS=1).
    Returns dict(seconds_primal, seconds_adjoint, rays, threads, evals, hits, build)."""
    path, how = build_native()
    lib = C.CDLL(path)
    lib.cp_set_threads(int(threads) if threads else len(os.sched_getaffinity(0)))
    sc = CPScene()
    lib.cp_defaults(C.byref(sc))
    sdf = np.ascontiguousarray(sdf, dtype=np.float32)
    sc.sdf = sdf.ctypes.data
    sc.rz, sc.ry, sc.rx = sdf.shape
    if albedo is not None:
        albedo = np.ascontiguousarray(albedo, dtype=np.float32)
        sc.albedo = albedo.ctypes.data
        sc.az, sc.ay, sc.ax = albedo.shape[:3]
    sc.sensor.to_world[:] = [float(v) for v in np.asarray(to_world, dtype=np.float64).reshape(-1)]
    sc.sensor.width, sc.sensor.height = W, H
    if prb is not None:  # sdf_prb_reparam: dict(max_depth, rr_depth)
        sc.integrator = 3
        sc.prb.max_depth, sc.prb.rr_depth = prb['max_depth'], prb['rr_depth']
    grad = np.zeros_like(sdf)
    out = np.zeros(8, dtype=np.float64)
    lib.cp_time_sample(C.byref(sc), C.c_uint(seed), spp_primal, spp_adjoint, C.c_longlong(n_primal), C.c_longlong(n_adjoint),
                       grad.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return dict(seconds_primal=out[0], seconds_adjoint=out[1], rays=n_primal + n_adjoint, threads=lib.cp_threads(),
                evals=out[2] + out[3], hits=out[4], build=how, grad_abs_max=float(np.abs(grad).max()))
