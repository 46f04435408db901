"""`redistancing` of the reference (python/redistancing.py)."""
from sdfrt.redistancing import redistance  # noqa: F401
