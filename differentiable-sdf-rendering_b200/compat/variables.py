"""`variables` of the reference (python/variables.py), provided by the sdfrt mirror (same class names and constructor arguments)."""
from sdfrt.variables import SdfVariable, Variable, VolumeVariable, upsample_grid, upsample_sdf  # noqa: F401
