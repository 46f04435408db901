"""`warp` of the reference (python/warp.py): the parameter-holder classes of the sdfrt mirror under the reference's names."""
from sdfrt.warp import DummyWarpField, WarpField2D, WarpFieldConvolution  # noqa: F401
