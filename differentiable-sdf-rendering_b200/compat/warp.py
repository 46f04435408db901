"""`warp` of the reference (python/warp.py): WarpField2D / DummyWarpField from the sdfrt mirror; the Bangaru-2020 baseline
(WarpFieldConvolution, python/warp.py:131-176) is not provided and fails on construction."""
from sdfrt.warp import DummyWarpField, WarpField2D  # noqa: F401


class WarpFieldConvolution:
    def __init__(self, *a, **kw):
        raise NotImplementedError("WarpFieldConvolution (the Bangaru et al. 2020 comparison baseline) is not provided")
