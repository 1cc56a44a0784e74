"""Compatibility surface: Pyfhel 2.3.1-style objects and the reference's FLPyfhelin module."""
from .pyfhel_shim import PyCtxt, Pyfhel, PyPtxt  # noqa: F401
