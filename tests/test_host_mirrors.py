"""Host-side mirrors next to the hot path: .vol I/O, losses, regulariser, variable helpers.  CPU only."""
import struct

import numpy as np
import torch

import pytest

import sdfrt
from sdfrt import losses, regularizations, variables, volio


def test_vol_round_trip_and_header(tmp_path):
    rng = np.random.default_rng(0)
    a = rng.normal(size=(3, 4, 5, 2)).astype(np.float32)
    p = str(tmp_path / 'x.vol')
    volio.write_vol(p, a)
    raw = open(p, 'rb').read()
    assert raw[:4] == b'VOL\x03'
    assert struct.unpack('<iiiii', raw[4:24]) == (1, 5, 4, 3, 2)                 # encoding, xres, yres, zres, channels
    assert struct.unpack('<6f', raw[24:48]) == (0.0, 0.0, 0.0, 1.0, 1.0, 1.0)
    assert len(raw) == 48 + a.size * 4
    # x fastest: element (z, y, x, c) at ((z*yres + y)*xres + x)*channels + c
    off = 48 + 4 * (((2 * 4 + 1) * 5 + 3) * 2 + 1)
    assert struct.unpack('<f', raw[off:off + 4])[0] == a[2, 1, 3, 1]
    b, lo, hi = volio.read_vol(p)
    assert np.array_equal(a, b) and lo == (0.0, 0.0, 0.0) and hi == (1.0, 1.0, 1.0)
    volio.write_vol(p, a[..., 0])                                                   # 3-D input -> 1 channel
    assert volio.read_vol(p)[0].shape == (3, 4, 5, 1)
    assert volio.variable_path('out', 'SamplingIntegrator.sdf.data', 64).replace('\\\\', '/') == 'out/sdf-data-0064.vol'
    assert volio.variable_path('out', 'main-bsdf.reflectance.volume.data', 'final').endswith('main-bsdf-reflectance-volume-data-final.vol')


def test_losses_against_numpy():
    rng = np.random.default_rng(1)
    a, b = rng.uniform(size=(8, 8, 3)), rng.uniform(size=(8, 8, 3))
    ta, tb = torch.from_numpy(a), torch.from_numpy(b)
    assert abs(losses.l1(ta, tb).item() - np.abs(a - b).mean()) < 1e-12
    assert abs(losses.l2(ta, tb).item() - ((a - b) ** 2).mean()) < 1e-12
    d = losses.downsample(ta).numpy()
    ref = np.zeros_like(a)
    for y in range(8):
        for x in range(8):
            y1, x1 = min(y + 1, 7), min(x + 1, 7)
            ref[y, x] = 0.25 * (a[y, x] + a[y, x1] + a[y1, x] + a[y1, x1])
    assert np.allclose(d, ref)
    ms = losses.multiscale_l1(ta, tb).item()
    x, y, acc = a, b, 0.0
    for lvl in range(4):
        acc += np.abs(x - y).mean()
        x, y = losses.downsample(torch.from_numpy(x)).numpy(), losses.downsample(torch.from_numpy(y)).numpy()
    assert abs(ms - acc / 4) < 1e-12


def test_laplacian_regulariser():
    rng = np.random.default_rng(2)
    d = rng.normal(size=(5, 5, 5))
    ref = 0.0
    for z in range(5):
        for y in range(5):
            for x in range(5):
                nb = sum(d[min(max(z + dz, 0), 4), min(max(y + dy, 0), 4), min(max(x + dx, 0), 4)]
                         for dz, dy, dx in ((-1, 0, 0), (1, 0, 0), (0, -1, 0), (0, 1, 0), (0, 0, -1), (0, 0, 1)))
                ref += (d[z, y, x] - nb / 6) ** 2
    assert abs(regularizations.reference_laplacian_reg(torch.from_numpy(d)[..., None]).item() - ref) < 1e-10
    lin = torch.arange(5, dtype=torch.float64)[None, None, :].expand(5, 5, 5)
    # a linear field has zero Laplacian away from the clamped faces
    assert regularizations.reference_laplacian_reg(lin.contiguous()).item() < 2 * 25 * (1 / 6) ** 2 + 1e-12
    with pytest.raises(sdfrt.SdfrtError):   # the product entry point has no CPU path
        regularizations.eval_discrete_laplacian_reg(lin.contiguous().float())


def test_variable_helpers():
    assert variables.simple_lr_decay(4e-2, 0.02, 0) == 4e-2
    assert abs(variables.simple_lr_decay(4e-2, 0.02, 490) - 4e-2 / (1 + 0.02 * 490) / 2) < 1e-15
    assert abs(variables.simple_lr_decay(4e-2, 0.02, 510) - 4e-2 / (1 + 0.02 * 510) / 4) < 1e-15
    assert abs(variables.adaptive_sdf_lr(4e-2, 64, 0) - 2e-2) < 1e-15
    g = torch.tensor([float('nan'), 5.0, -3.0, 0.05])
    assert torch.equal(variables.validate_gradient(g), torch.tensor([0.0, 0.1, -0.1, 0.05]))
    sdf = torch.full((8, 8, 8, 1), -1.0)
    con = variables.bbox_constraint(sdf)
    assert con.shape == sdf.shape and con[0, 0, 0, 0] > 0 and con[4, 4, 4, 0] < 0     # outside the 0.49 box the SDF is pushed positive
    m = variables.ParameterMean(0.95)
    m.update(torch.ones(3))
    m.update(torch.zeros(3))
    assert torch.allclose(m.mean, torch.full((3,), 0.95))
