"""`mitsuba` as the reference's drivers use it (python/optimize.py, python/shape_opt.py, python/opt_configs.py, python/configs.py):
scene loading from the registry of sdfrt.compat, parameter traversal, the Adam optimiser, `mi.render`, bitmaps on disk."""
import os

import numpy as np
import torch

import sdfrt
from sdfrt import compat as _compat

_variant = 'cuda_ad_rgb'


class LogLevel:
    Warn = 2


def variants():
    return ['cuda_ad_rgb']            # the one back end: hand-written sm_100a kernels behind libsdfrt (no LLVM/CPU variant)


def set_variant(v):
    global _variant
    if v != 'cuda_ad_rgb':
        raise RuntimeError(f"variant {v!r} is not available: libsdfrt has no CPU back end")
    _variant = v


def variant():
    return _variant


def set_log_level(level):
    return None


class Thread:
    @staticmethod
    def wait_for_tasks():
        return None


# ---- small value types ------------------------------------------------------------------------------------------------
class ScalarPoint2i(np.ndarray):
    """2-vector of ints with numpy arithmetic (`target_res // 2**k`, `init_res * 2**k`, indexing)."""

    def __new__(cls, *a):
        v = np.asarray(a[0] if len(a) == 1 else a, dtype=np.int64).reshape(-1)
        return v.view(cls)


ScalarVector2i = ScalarPoint2i


class Float:
    """The scalar accumulator of the optimisation loop: `loss = mi.Float(0.0); loss += view_loss; loss[0]`
    (python/shape_opt.py:76,87,92,96-101).  Dr.Jit hands back a Python float from `loss[0]`; so does this (the backward pass of
    every term has already run when it is added, python/shape_opt.py:83,91)."""

    def __init__(self, v=0.0):
        self.v = float(v)

    def __iadd__(self, x):
        self.v += float(x.detach().sum()) if torch.is_tensor(x) else float(x[0] if isinstance(x, Float) else x)
        return self

    def __getitem__(self, i):
        return self.v


def TensorXf(x, shape=None):
    """mi.TensorXf(bitmap | array): a float32 tensor on the rendering device."""
    if isinstance(x, Bitmap):
        x = x.array
    t = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x, dtype=torch.float32)
    if shape is not None:
        t = t.reshape(tuple(int(v) for v in shape))
    return t.to(sdfrt.redistancing.shared_context(0).device)


# ---- bitmaps: arrays on disk (the reference writes .exr / .png through OpenEXR / libpng; here a file of the same NAME holds the
# ---- array in numpy's format, which is all the drivers need: they write, copy and re-load these files) ---------------------
class Bitmap:
    def __init__(self, src):
        if isinstance(src, Bitmap):
            self.array = src.array.copy()
        elif isinstance(src, (str, os.PathLike)):
            with open(src, 'rb') as f:
                self.array = np.load(f)
        else:
            self.array = np.asarray(src.detach().cpu() if torch.is_tensor(src) else src, dtype=np.float32)

    def size(self):
        return ScalarPoint2i(self.array.shape[1], self.array.shape[0])

    def resample(self, res, rfilter=None):
        """Bitmap.resample([height, width], filter): area (box) resampling -- the drivers only shrink by powers of two."""
        t = torch.from_numpy(self.array).permute(2, 0, 1)[None]
        out = torch.nn.functional.interpolate(t, size=(int(res[0]), int(res[1])), mode='area')
        return Bitmap(out[0].permute(1, 2, 0).contiguous().numpy())


class _Util:
    @staticmethod
    def write_bitmap(fn, data, write_async=True):
        arr = data.array if isinstance(data, Bitmap) else np.asarray(data.detach().cpu() if torch.is_tensor(data) else data, dtype=np.float32)
        os.makedirs(os.path.dirname(os.path.abspath(fn)), exist_ok=True)
        with open(fn, 'wb') as f:
            np.save(f, arr)


util = _Util()


# ---- scenes ------------------------------------------------------------------------------------------------------------
class _Shape:
    def __init__(self, sid):
        self._id = sid

    def id(self):
        return self._id


def _shapes(scene):
    return [_Shape('dummy_sdf_shape')]          # the placeholder shape of dummysdf.xml (python/integrators/reparam.py:56-80)


def load_file(path, **kwargs):
    """mi.load_file('<SCENE_DIR>/<name>/<name>.xml', integrator=..., resx=..., resy=..., sdf_filename=..., shape_file=...)."""
    name = os.path.splitext(os.path.basename(path))[0]
    if name not in _compat.SCENES:
        raise FileNotFoundError(f"scene {name!r} is not registered (sdfrt.compat.register_scene); Mitsuba XML scenes are not loaded")
    target, optimisable = _compat.SCENES[name]
    is_target = kwargs.get('sdf_filename', None) == '' and 'shape_file' not in kwargs
    scene = (target if is_target else optimisable)(**kwargs)
    if not hasattr(scene, 'shapes'):
        scene.shapes = lambda: _shapes(scene)
    return scene


def load_dict(d):
    raise NotImplementedError("mi.load_dict: only scenes registered with sdfrt.compat.register_scene are available")


# ---- parameters / optimiser -------------------------------------------------------------------------------------------
SDF_KEY = 'SamplingIntegrator.sdf.data'
REFLECTANCE_KEY = 'main-bsdf.reflectance.volume.data'


class SceneParameters(dict):
    """mi.traverse(scene): key -> tensor, `keep`, `update(opt)` (python/shape_opt.py:49-51,105)."""

    def __init__(self, scene):
        super().__init__()
        self.scene = scene
        self[SDF_KEY] = scene.integrator().sdf.data
        if getattr(scene, 'albedo_grid', None) is not None:
            self[REFLECTANCE_KEY] = scene.albedo_grid

    def keep(self, keys):
        for k in list(self.keys()):
            if k not in keys:
                del self[k]

    def update(self, values=None):
        if values is not None:
            for k in list(self.keys()):
                if k in values:
                    self[k] = values[k]
        integrator = self.scene.integrator()
        if SDF_KEY in self:
            integrator.sdf.data = self[SDF_KEY]          # Grid3d.parameters_changed (python/shapes.py:477-479)
            integrator.parameters_changed()
        if REFLECTANCE_KEY in self:
            self.scene.albedo_grid = self[REFLECTANCE_KEY]


def traverse(obj):
    if hasattr(obj, 'integrator'):
        return SceneParameters(obj)
    if hasattr(obj, 'film'):                             # a sensor: `params['film.size'] = res; params.update()` (python/util.py:146-150)
        return _SensorParameters(obj)
    raise TypeError(f"cannot traverse {type(obj).__name__}")


class _SensorParameters(dict):
    def __init__(self, sensor):
        super().__init__()
        self.sensor = sensor
        self['film.size'] = ScalarPoint2i(sensor.film().width, sensor.film().height)

    def update(self):
        sdfrt.set_sensor_res(self.sensor, self['film.size'])


class _Ad:
    class Adam(sdfrt.optimizers.Adam):
        """mi.ad.Adam(lr=, params=, mask_updates=) (python/shape_opt.py:53)."""

        def __init__(self, lr, params=None, mask_updates=False, **kw):
            super().__init__(lr=lr, mask_updates=mask_updates, params=dict(params) if params is not None else None, **kw)


ad = _Ad()


def render(scene, params=None, sensor=0, seed=0, spp=0, seed_grad=0, spp_grad=0):
    """mi.render: `sensor` may be a sensor object of the scene config (python/shape_opt.py:78-80) or an index."""
    if not isinstance(sensor, int):
        sensors = scene.sensors()
        if sensor not in sensors:
            sensors.append(sensor)
        sensor = sensors.index(sensor)
    return sdfrt.render(scene, params=params, sensor=sensor, seed=seed, spp=spp, seed_grad=seed_grad, spp_grad=spp_grad)
