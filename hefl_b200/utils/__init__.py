"""Timing, clock sampling, structured logging, NVTX ranges."""
from .timing import DeviceTimer, StageTimes, nvtx_range  # noqa: F401
from .clocks import ClockSampler  # noqa: F401
from .jsonl import JsonlLogger  # noqa: F401
