"""Re-distancing, CPU side: the oracle (oracle/redistance_oracle.c) against analytic distance fields and its own
fixed-point properties, and the block-asynchronous update order of the CUDA solver -- emulated sequentially on the host from the
kernel's shared node arithmetic (tests/host_harness) -- against the oracle's sweeps."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hostmath  # noqa: E402
from oracle import redistance_oracle as ro  # noqa: E402


def centres(r):
    return (np.arange(r) + 0.5) / r


def sphere(shape, c=(0.5, 0.5, 0.5), radius=0.3):
    z, y, x = np.meshgrid(centres(shape[0]), centres(shape[1]), centres(shape[2]), indexing='ij')
    return (np.sqrt((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) - radius).astype(np.float32)


def wobbly(shape, seed=0, noise=0.02):
    """a non-distance level-set function with a sphere-like zero set"""
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(centres(shape[0]), centres(shape[1]), centres(shape[2]), indexing='ij')
    phi = sphere(shape, (0.5, 0.45, 0.55)) * (1.5 + np.sin(9 * x) * np.cos(7 * y)) + noise * rng.standard_normal(shape)
    return phi.astype(np.float32)


def tiles(phi, lower_bound=None):
    lib = C.CDLL(hostmath.build())
    phi = np.ascontiguousarray(phi, np.float32)
    out = np.empty_like(phi)
    rz, ry, rx = phi.shape
    visits = C.c_longlong(0)
    lb = None if lower_bound is None else np.ascontiguousarray(lower_bound, np.float32)
    lists = lib.hh_redistance_tiles(phi.ctypes.data_as(C.c_void_p), None if lb is None else lb.ctypes.data_as(C.c_void_p),
                                    out.ctypes.data_as(C.c_void_p), rx, ry, rz, C.byref(visits))
    return out, lists, visits.value


@pytest.mark.parametrize('res', [24, 48])
def test_oracle_recovers_a_sphere(res):
    exact = sphere((res,) * 3)
    for scale in (1.0, 3.0, 0.25):                        # the zero set is what matters, not the magnitudes
        d = ro.redistance(exact * scale)
        assert np.abs(d - exact).max() < 1.5 / res        # first-order scheme: error of the order of one cell
        assert np.abs(d - exact)[np.abs(exact) < 0.5 / res].max() < 0.1 / res   # sub-voxel accurate at the frozen nodes
        chk = ro.check(exact * scale, d)
        assert chk['sign_mismatches'] == 0 and chk['frozen'] == 0.0 and chk['eikonal'] < 1e-7


def test_oracle_fixed_point_and_idempotence():
    phi = wobbly((20, 28, 36))
    d, cycles = ro.redistance(phi, return_cycles=True)
    assert cycles >= 2
    chk = ro.check(phi, d)
    assert chk['sign_mismatches'] == 0 and chk['frozen'] == 0.0 and chk['eikonal'] < 1e-7
    d2 = ro.redistance(d)                                  # re-distancing a distance field moves it by less than a cell
    assert np.abs(d2 - d).max() < 1.0 / 20
    # anisotropic spacing: gradient magnitude (central differences, away from the surface / borders / the medial axis) ~ 1
    gz, gy, gx = np.gradient(d.astype(np.float64), 1 / 20, 1 / 28, 1 / 36)
    gm = np.sqrt(gx ** 2 + gy ** 2 + gz ** 2)[2:-2, 2:-2, 2:-2]
    assert abs(np.median(gm) - 1.0) < 0.05


def test_oracle_edge_cases():
    # no interface at all: nothing is frozen, every node stays "unknown" (kept at +-1e30 like an unreached node)
    d = ro.redistance(np.full((8, 8, 8), 0.7, np.float32))
    assert (d >= 1e29).all()
    # exact zeros are frozen at zero; their neighbours are one cell away
    phi = np.ones((8, 8, 8), np.float32)
    phi[4, 4, 4] = 0.0
    d = ro.redistance(phi)
    assert d[4, 4, 4] == 0.0 and abs(d[4, 4, 5] - 1 / 8) < 1e-7 and abs(d[3, 4, 4] - 1 / 8) < 1e-7
    # +-1 occupancy block (create_block_sdf, python/shapes.py:584-590): interface half way between the cells
    occ = np.ones((16, 16, 16), np.float32)
    occ[6:10, 6:10, 4:12] = -1
    d = ro.redistance(occ)
    assert abs(d[8, 8, 3] - 0.5 / 16) < 1e-6 and abs(d[8, 8, 4] + 0.5 / 16) < 1e-6 and (np.sign(d) == np.sign(occ)).all()


@pytest.mark.parametrize('shape', [(16, 16, 16), (20, 13, 9), (40, 40, 40), (33, 24, 17)])
def test_tile_order_reaches_the_oracle_fixed_point(shape):
    phi = wobbly(shape, seed=shape[0])
    want = ro.redistance(phi)
    got, lists, visits = tiles(phi)
    assert lists >= 2 and visits > 0
    assert np.abs(got - want).max() <= 1e-6               # same fixed point up to the rounding of the update order
    chk = ro.check(phi, got)
    assert chk['sign_mismatches'] == 0 and chk['frozen'] == 0.0 and chk['eikonal'] < 1e-6


def test_tile_order_with_lower_bound():
    phi = wobbly((24, 24, 24), seed=5)
    lb = (sphere((24, 24, 24), radius=0.45) * -1.0).astype(np.float32)     # keeps the shape inside a ball of radius 0.45
    want = ro.redistance(np.maximum(phi, lb))
    got, _, _ = tiles(phi, lb)
    assert np.abs(got - want).max() <= 1e-6


def test_golden_fixture():
    """tests/golden/redistance_wobbly.npz (written by tests/golden/make_golden.py): the oracle reproduces it bit for bit, the
    tile order within the stated tolerance."""
    g = dict(np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'redistance_wobbly.npz')))
    assert np.array_equal(ro.redistance(g['phi']), g['dist'])
    assert np.abs(tiles(g['phi'])[0] - g['dist']).max() <= 1e-6
