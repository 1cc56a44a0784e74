"""Structured per-round records as JSON lines (the reference only prints)."""
from __future__ import annotations

import json
import time
from typing import Any, Dict, Optional


class JsonlLogger:
    def __init__(self, path: Optional[str], rank: int = 0, echo: bool = False):
        self.path = path if rank == 0 else None
        self.echo = echo and rank == 0

    def write(self, record: Dict[str, Any]) -> None:
        record = dict(record, ts=time.time())
        if self.echo:
            print(json.dumps(record))
        if self.path:
            with open(self.path, "a") as f:
                f.write(json.dumps(record) + "\n")
