"""sdfrt -- B200-native hot path of rgl-epfl/differentiable-sdf-rendering.

Python surface mirroring the reference plugins: shapes.Grid3d, warp.WarpField2D / DummyWarpField,
integrators.sdf_{direct,silhouette,simple_shading}_reparam and a `render()` with the mi.render contract.
All rendering runs in hand-written sm_100a CUDA kernels behind the C ABI of libsdfrt.so.
"""
from . import (_capi, dist, engine, integrators, losses, optimizers, redistancing, regularizations, scene, shape_opt, shapes,  # noqa: F401
               synthetic, variables, volio, warp)  # noqa: F401
from ._capi import SdfrtError  # noqa: F401
from .engine import Context  # noqa: F401
from .render import render  # noqa: F401
from .scene import Scene, Sensor, get_regular_cameras, look_at, set_sensor_res  # noqa: F401
from .shapes import BoxSDF, Grid3d, Ray3f, create_block_sdf, create_sphere_sdf  # noqa: F401
from .warp import DummyWarpField, WarpField2D, WarpFieldConvolution  # noqa: F401
