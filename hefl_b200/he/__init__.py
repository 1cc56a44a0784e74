"""HE scheme layer: packed CKKS (product surface) and BFV-fractional (compat surface)."""
from .context import CKKSContext, CtBatch, RelinKey  # noqa: F401
