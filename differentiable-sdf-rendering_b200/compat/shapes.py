"""`shapes` of the reference (python/shapes.py): the names its drivers import, provided by the sdfrt mirror."""
from sdfrt.shapes import BoxSDF, Grid3d, Ray3f, SDFBase, atleast_4d, create_block_sdf, create_sphere_sdf  # noqa: F401
