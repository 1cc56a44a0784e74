"""Multi-GPU layer: symmetric peer memory, fused ciphertext all-reduce, baselines."""
from .allreduce import (CollectiveTransport, FusedTransport, LoopbackTransport, Transport,  # noqa: F401
                        make_transport)
from .symm import SymmetricBuffer, probe  # noqa: F401
