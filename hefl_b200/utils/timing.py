"""Device-side stage timing.

The reference brackets six stages with ``time.time()`` and prints them
(FLPyfhelin.py:203/223, :235/238, :243/247, :264/266, :306/326, :369/388). Here every stage
is bracketed by CUDA events on the launching stream (wall clock on CPU), reported per rank
and reduced to the max over ranks; NVTX ranges carry the reference's stage names.
"""
from __future__ import annotations

import contextlib
import time
from typing import Dict, List, Optional

import torch

REFERENCE_STAGE_NAMES = {
    "train": "local training (model.fit, FLPyfhelin.py:193)",
    "encrypt": "Time to encrypt weights (FLPyfhelin.py:224)",
    "export": "Time to export weights to pickle (FLPyfhelin.py:239)",
    "import": "Time to import (FLPyfhelin.py:327)",
    "aggregate": "Time to aggregate (FLPyfhelin.py:389)",
    "decrypt": "Time to decrypt (FLPyfhelin.py:267)",
}


@contextlib.contextmanager
def nvtx_range(name: str):
    if torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield


class DeviceTimer:
    """Records (start, end) event pairs per stage; resolve() after a synchronize."""

    def __init__(self, device: torch.device):
        self.cuda = device.type == "cuda"
        self.pending: List[tuple] = []
        self.done: Dict[str, float] = {}

    @contextlib.contextmanager
    def stage(self, name: str):
        with nvtx_range(name):
            if self.cuda:
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                yield
                e.record()
                self.pending.append((name, s, e))
            else:
                t0 = time.perf_counter()
                yield
                self.done[name] = self.done.get(name, 0.0) + (time.perf_counter() - t0) * 1e3

    def resolve(self) -> Dict[str, float]:
        """Milliseconds per stage (summed over repeated entries). Synchronizes."""
        if self.cuda and self.pending:
            torch.cuda.synchronize()
            for name, s, e in self.pending:
                self.done[name] = self.done.get(name, 0.0) + s.elapsed_time(e)
            self.pending.clear()
        out, self.done = self.done, {}
        return out


class StageTimes:
    """Max-over-ranks reduction of a stage->ms dict."""

    @staticmethod
    def max_over_ranks(times: Dict[str, float], device: torch.device, group=None) -> Dict[str, float]:
        import torch.distributed as dist

        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return dict(times)
        keys = sorted(times)
        t = torch.tensor([times[k] for k in keys], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return {k: float(v) for k, v in zip(keys, t.tolist())}
