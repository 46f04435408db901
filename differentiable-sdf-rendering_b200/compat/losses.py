"""`losses` of the reference (python/losses.py)."""
from sdfrt.losses import downsample, l1, l2, mape, multiscale, multiscale_l1  # noqa: F401
