"""`util` of the reference (python/util.py): the helpers its drivers import, on the sdfrt mirror's sensors and on the array-backed
bitmaps of the compat `mitsuba` module."""
import json

import mitsuba as mi
import numpy as np

from sdfrt.scene import get_regular_cameras, set_sensor_res  # noqa: F401
from sdfrt.shapes import atleast_4d  # noqa: F401


def get_regular_cameras_top(n_sensors, angle_shift=0.0, resx=128, resy=128, radius=2.0):
    return get_regular_cameras(n_sensors, angle_shift, resx, resy, radius, height_scale=1.3)


def get_file_sensors(fn, resx, resy, indices=None):
    raise NotImplementedError("sensors stored in Mitsuba XML scenes are not available; use get_regular_cameras configs")


def resize_img(img, target_res, smooth=False):
    """python/util.py:14-23 (box / gaussian resampling of a Bitmap; here area resampling for both)."""
    assert isinstance(img, mi.Bitmap)
    source_res = img.size()
    if target_res[0] == source_res[0] and target_res[1] == source_res[1]:
        return img
    return img.resample([target_res[1], target_res[0]])


def render_turntable(scene, output_dir, **kw):
    return None      # presentation only (python/util.py:26-60)


def dump_metadata(config, opt_config, extra=None, fn='test.json'):
    """python/util.py:152-186: configs + loss curve as JSON."""
    def conv(obj):
        if hasattr(obj, 'name'):
            return obj.name
        if callable(obj):
            return getattr(obj, '__name__', str(obj))
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        return obj
    d = {'config': {k: conv(v) for k, v in dict(config.__dict__).items()},
         'opt_config': {k: conv(v) for k, v in dict(opt_config.__dict__).items() if k not in ('sensors', 'sensors_reordered', 'variables')}}
    if extra is not None:
        d.update(extra)
    with open(fn, 'wt') as f:
        json.dump(d, f, indent=4, default=str)
