"""ctypes driver of oracle/redistance_oracle.c (CPU restatement of `redistancing.redistance`, python/redistancing.py:4-13).
TEST INFRASTRUCTURE ONLY -- see the header of the C file (parity unpinned: fastsweep is not installable offline)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'redistance_oracle.c')
LIB = os.path.join(HERE, 'libredistance_oracle.so')
FLAGS = ['-O2', '-ffp-contract=off', '-std=c99', '-fPIC', '-shared']

_lib = None


def build():
    if not (os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC)):
        subprocess.check_call(['gcc'] + FLAGS + [SRC, '-o', LIB, '-lm'])
    return LIB


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.rd_redistance.restype = C.c_int
        _lib.rd_redistance.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lib.rd_check.restype = None
        _lib.rd_check.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return _lib


def redistance(phi, return_cycles=False):
    """phi: (rz, ry, rx) float array -> re-distanced float32 array of the same shape."""
    phi = np.ascontiguousarray(np.asarray(phi, dtype=np.float32).reshape(np.asarray(phi).shape[:3]))
    out = np.empty_like(phi)
    rz, ry, rx = phi.shape
    cycles = _load().rd_redistance(phi.ctypes.data, out.ctypes.data, rx, ry, rz)
    return (out, cycles) if return_cycles else out


def check(phi, dist):
    """Size-independent properties of a re-distanced grid: dict(eikonal, frozen, sign_mismatches)."""
    phi = np.ascontiguousarray(phi, dtype=np.float32)
    dist = np.ascontiguousarray(dist, dtype=np.float32)
    rz, ry, rx = phi.shape
    res = np.zeros(3, dtype=np.float64)
    _load().rd_check(phi.ctypes.data, dist.ctypes.data, rx, ry, rz, res.ctypes.data)
    return dict(eikonal=float(res[0]), frozen=float(res[1]), sign_mismatches=int(res[2]))
