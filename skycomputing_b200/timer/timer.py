"""Timers.

``DistributedTimer`` keeps the reference's file format (``timestamp: <time.time()>`` lines in
``<root>/dist_timer.txt``, scaelum/timer/timer.py:10-29) for log compatibility.  The hot path does
not use it: per-stage forward/backward times come from ``DeviceTimer`` (CUDA events on the compute
stream, no host synchronisation until the numbers are read).
"""
from __future__ import annotations

import os
import os.path as osp
import time
from typing import Dict, List

import torch


class DistributedTimer:
    def __init__(self, root: str = "/tmp"):
        self.root = root
        self.file_path = osp.join(root, "dist_timer.txt")

    def clean_prev_file(self) -> None:
        if osp.exists(self.file_path):
            os.remove(self.file_path)

    def add_timestamp(self) -> None:
        os.makedirs(self.root, exist_ok=True)
        with open(self.file_path, "a") as f:
            f.write("timestamp: {}\n".format(time.time()))

    def get_prev_interval(self) -> float:
        with open(self.file_path, "r") as f:
            lines = f.readlines()
        return float(lines[-1].split(":")[-1]) - float(lines[-2].split(":")[-1])


class DeviceTimer:
    """Named intervals measured with CUDA events (falls back to wall clock on CPU)."""

    def __init__(self, enabled: bool = True):
        self.enabled = enabled
        self._open: Dict[str, object] = {}
        self._pairs: Dict[str, List[tuple]] = {}
        self._cuda = torch.cuda.is_available()

    def start(self, name: str) -> None:
        if not self.enabled:
            return
        if self._cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._open[name] = ev
        else:
            self._open[name] = time.perf_counter()

    def stop(self, name: str) -> None:
        if not self.enabled or name not in self._open:
            return
        s = self._open.pop(name)
        if self._cuda:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
        else:
            e = time.perf_counter()
        self._pairs.setdefault(name, []).append((s, e))

    def collect(self, reset: bool = True) -> Dict[str, float]:
        """Total seconds per name (synchronises once)."""
        out: Dict[str, float] = {}
        if self._cuda and self._pairs:
            torch.cuda.synchronize()
        for name, pairs in self._pairs.items():
            if self._cuda:
                out[name] = sum(s.elapsed_time(e) for s, e in pairs) * 1e-3
            else:
                out[name] = sum(e - s for s, e in pairs)
        if reset:
            self._pairs = {}
        return out
