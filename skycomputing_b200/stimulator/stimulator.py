"""Synthetic heterogeneity multipliers.

Same numbers as scaelum/stimulator/stimulator.py:4-24 (numpy default_rng seeds 22 / 32 / 32), but
produced by the C++ core's bit-exact PCG64/SeedSequence re-implementation (csrc/alloc/
stimulator.cc); falls back to numpy if the native module is unavailable.  ``STIMULATE=1`` makes the
DeviceBenchmarker apply them (the reference's import path for this is broken).
"""
from __future__ import annotations

import numpy as np

try:
    from .. import _core
except Exception:  # pragma: no cover
    _core = None


class Stimulator:
    def __init__(self, worker_num: int, mem_seed: int = 22, net_seed: int = 32, comp_seed: int = 32):
        self.worker_num = worker_num
        n = worker_num + 1
        if _core is not None:
            s = _core.Stimulator(worker_num, mem_seed, net_seed, comp_seed)
            self.m_slowdown = np.asarray(s.m_slowdown)
            self.n_slowdown = np.asarray(s.n_slowdown)
            self.c_slowdown = np.asarray(s.c_slowdown)
        else:  # pragma: no cover
            self.m_slowdown = 2 * np.random.default_rng(seed=mem_seed).random((n,)) + 1
            self.n_slowdown = np.random.default_rng(seed=net_seed).random((n,)) + 1
            self.c_slowdown = np.random.default_rng(seed=comp_seed).random((n,)) + 1

    def memory_slowdown(self, worker_id: int) -> float:
        return float(self.m_slowdown[worker_id])

    def compute_slowdown(self, worker_id: int) -> float:
        return float(self.c_slowdown[worker_id])

    def network_stimulate(self, worker_id: int) -> float:
        return float(self.n_slowdown[worker_id])
