"""Stand-in used only when matplotlib is not installed: python/shape_opt.py:107-115 plots the loss curve in a `finally:` block."""
