"""python/integrators/sdf_simple_shading_reparam.py: the plugin lives in sdfrt.integrators (kernels behind the C ABI)."""
from sdfrt.integrators.sdf_simple_shading_reparam import *  # noqa: F401,F403
