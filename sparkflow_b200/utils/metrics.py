"""Structured per-iteration metrics (JSON lines). The reference only prints the partition loss when
``verbose`` is set (HogwildSparkModel.py:94-98); this adds a machine-readable sink with throughput."""
from __future__ import annotations

import json
import os
import threading
import time
from typing import Any, Dict, Optional

_LOCK = threading.Lock()


class MetricsLogger:
    def __init__(self, path: Optional[str] = None):
        self.path = path or os.environ.get("SPARKFLOW_METRICS")
        self.t0 = time.perf_counter()

    @property
    def enabled(self) -> bool:
        return bool(self.path)

    def log(self, **record: Any) -> None:
        if not self.path:
            return
        record.setdefault("t", round(time.perf_counter() - self.t0, 6))
        line = json.dumps(record)
        with _LOCK:
            with open(self.path, "a") as fh:
                fh.write(line + "\n")


def read_metrics(path: str):
    with open(path) as fh:
        return [json.loads(l) for l in fh if l.strip()]
