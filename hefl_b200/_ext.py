"""Loader for the in-tree native library (``hefl_b200/_native.so``).

``ops()`` returns ``torch.ops.hefl`` after loading the library. If the library is missing
it is built on the spot (nvcc cross-compiles without a GPU). On a GPU box a missing or
unloadable library is a hard error: there is no PyTorch fallback for the kernels.
"""
from __future__ import annotations

import os
import threading
from pathlib import Path

import torch

_LIB = Path(__file__).resolve().parent / "_native.so"
_lock = threading.Lock()
_loaded = False


def native_path() -> Path:
    return _LIB


def load(build_if_missing: bool = True) -> None:
    global _loaded
    if _loaded:
        return
    with _lock:
        if _loaded:
            return
        if not _LIB.exists():
            if not build_if_missing or os.environ.get("HEFL_NO_BUILD") == "1":
                raise RuntimeError(f"native library {_LIB} is missing; run `python -m hefl_b200._build`")
            from . import _build

            _build.build(verbose=True)
        torch.ops.load_library(str(_LIB))
        _loaded = True


def ops():
    load()
    return torch.ops.hefl


def is_loaded() -> bool:
    return _loaded
