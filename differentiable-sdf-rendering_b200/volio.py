"""Mitsuba `.vol` grid files: the on-disk format the reference's callers read and write
(`mi.VolumeGrid(...).write`, python/variables.py:96-108,156-159; `sdf-data-XXXX.vol`, python/util.py:202-216).

Layout (little endian): 'V','O','L', u8 version = 3, i32 encoding = 1 (float32), i32 xres, yres, zres, i32 channels,
6 x f32 bounding box (min xyz, max xyz), then float32 data with x fastest: ((z*yres + y)*xres + x)*channels + c.
"""
import os
import struct

import numpy as np


def write_vol(path, data, bbox_min=(0.0, 0.0, 0.0), bbox_max=(1.0, 1.0, 1.0)):
    """data: (Z, Y, X) or (Z, Y, X, C) array (the tensor layout of `sdf.data`, python/shapes.py:388)."""
    a = np.asarray(data, dtype=np.float32)
    if a.ndim == 3:
        a = a[..., None]
    if a.ndim != 4:
        raise ValueError("expected a (Z, Y, X[, C]) array")
    z, y, x, c = a.shape
    with open(path, 'wb') as f:
        f.write(b'VOL' + struct.pack('<B', 3))
        f.write(struct.pack('<iiiii', 1, x, y, z, c))
        f.write(struct.pack('<6f', *bbox_min, *bbox_max))
        f.write(np.ascontiguousarray(a).astype('<f4').tobytes())


def read_vol(path):
    """Returns (data (Z, Y, X, C) float32, bbox_min, bbox_max)."""
    with open(path, 'rb') as f:
        head = f.read(48)
        if len(head) < 48 or head[:3] != b'VOL':
            raise ValueError(f"{path}: not a Mitsuba volume file")
        version = head[3]
        enc, x, y, z, c = struct.unpack('<iiiii', head[4:24])
        if version != 3 or enc != 1:
            raise ValueError(f"{path}: unsupported version/encoding {version}/{enc} (only float32 v3)")
        bbox = struct.unpack('<6f', head[24:48])
        n = x * y * z * c
        a = np.frombuffer(f.read(4 * n), dtype='<f4')
        if a.size != n:
            raise ValueError(f"{path}: truncated data")
    return a.reshape(z, y, x, c).astype(np.float32), bbox[:3], bbox[3:]


def variable_path(output_dir, key, suffix, suffix2=''):
    """python/variables.py:92-94 / :153-156: `SamplingIntegrator.` prefix dropped for the SDF, dots -> dashes."""
    k = key.replace("SamplingIntegrator.", "")
    suffix_str = f'{suffix:04d}' if isinstance(suffix, int) else suffix
    return os.path.join(output_dir, f'{k.replace(".", "-")}-{suffix_str}{suffix2}.vol')
