"""Schedule planner: picks micro-batch size, micro-batch count and virtual stages for a pipeline.

The allocator answers *which layers go where*; this module answers *how the step is scheduled* on
that allocation.  It is a closed-form model, validated against device timelines of 8-GPU runs
(`tools/predict_schedule.py`, profiles/bench_history.md: within 1.5 % of the measured step):

    plain 1F1B, weight gradients deferred:   (P - 1) F  +  m T  +  (P - 1) B
    looped / breadth-first, v chunks per GPU: m T  +  (P - 1) (F + B) / v  +  m (v - 1) X

with, per micro-batch on the slowest stage, F = forward, B = input-gradient chain, T = steady-state
period (F + B + the part of the weight gradients that does not hide on the side stream), X = exposed
cost of one extra boundary crossing (forward + backward).  Smaller micro-batches shrink fill / drain
but run less efficiently; that is captured by an efficiency curve ``eff(sequences)`` (time per
sequence relative to the reference micro-batch size).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence


@dataclass
class StageCosts:
    """Per-micro-batch times (seconds) of the slowest stage at ``ref_sequences`` sequences."""
    forward: float
    backward: float
    period: float
    ref_sequences: int = 32
    crossing: float = 40e-6          # exposed boundary cost of one extra (fwd + bwd) crossing
    # time-per-sequence multiplier vs ref_sequences; measured on B200 / BERT-large:
    # 16 sequences cost 1.25-1.36x per sequence, 8 sequences ~2.2x (profiles/bench_history.md)
    efficiency: Dict[int, float] = field(default_factory=lambda: {64: 0.97, 32: 1.0, 16: 1.3, 8: 2.2})

    def scale(self, sequences: int) -> float:
        pts = sorted(self.efficiency.items())
        if sequences <= pts[0][0]:
            e = pts[0][1]
        elif sequences >= pts[-1][0]:
            e = pts[-1][1]
        else:
            for (s0, e0), (s1, e1) in zip(pts[:-1], pts[1:]):
                if s0 <= sequences <= s1:
                    w = (sequences - s0) / (s1 - s0)
                    e = e0 + w * (e1 - e0)
                    break
        return e * sequences / self.ref_sequences


@dataclass
class Plan:
    micro_batch: int
    micro_batches: int
    virtual_stages: int
    schedule: str
    step_seconds: float


class SchedulePlanner:
    def __init__(self, num_stages: int, costs: StageCosts, blocks_per_stage: Optional[int] = None):
        """``blocks_per_stage`` bounds the number of chunks a GPU's span can be cut into (one
        transformer block per chunk at most when cuts must stay fusable)."""
        self.P = num_stages
        self.costs = costs
        self.max_v = max(1, blocks_per_stage or 1)

    def step_time(self, micro_batch: int, micro_batches: int, virtual_stages: int = 1) -> float:
        c, P, m, v = self.costs, self.P, micro_batches, virtual_stages
        k = c.scale(micro_batch)
        F, B, T = c.forward * k, c.backward * k, c.period * k
        if P == 1:
            return m * T
        if v <= 1:
            return (P - 1) * F + m * T + (P - 1) * B
        stall = max(0, P - m) * (F + B) / v          # a rank waits for the ring when m < P
        return m * T + (P - 1) * (F + B) / v + m * (v - 1) * c.crossing + stall

    def candidates(self, global_batch: int, micro_batch_sizes: Sequence[int] = (8, 16, 32, 64),
                   allow_looped: bool = True) -> List[Plan]:
        out = []
        for mb in micro_batch_sizes:
            if mb <= 0 or global_batch % mb:
                continue
            m = global_batch // mb
            vs = range(1, self.max_v + 1) if (allow_looped and self.P > 1) else (1,)
            for v in vs:
                if self.max_v % v:
                    continue                           # chunks of equal size only
                sched = "looped" if v > 1 else ("1f1b" if m > 1 else "sequential")
                out.append(Plan(mb, m, v, sched, self.step_time(mb, m, v)))
        return sorted(out, key=lambda p: p.step_seconds)

    def best(self, global_batch: int, **kw) -> Plan:
        plans = self.candidates(global_batch, **kw)
        if not plans:
            raise ValueError("no micro-batch size divides the global batch")
        return plans[0]


def costs_from_single_gpu_step(step_seconds: float, num_stages: int, sequences: int = 32,
                               forward_share: float = 0.335, backward_share: float = 0.60,
                               crossing: float = 40e-6) -> StageCosts:
    """Derive the per-stage costs of a balanced P-stage pipeline from a measured single-GPU step of
    the same per-GPU batch (the shares are the measured F : B(dgrad) : rest split of a BERT-large
    block on B200: 0.53 / 0.94 / 1.57 ms per 3 blocks)."""
    per_stage = step_seconds / num_stages
    return StageCosts(forward=per_stage * forward_share, backward=per_stage * backward_share,
                      period=per_stage, ref_sequences=sequences, crossing=crossing)
