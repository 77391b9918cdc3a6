"""Looped ("breadth-first") pipeline engine: v virtual stages per rank.

A plain 1F1B step costs ``m x T + (P - 1) x (F + B)``: the fill / drain term is a whole stage's
forward + backward per pipeline hop and dominates at 8 GPUs (profiles/bench_history.md: 9.4 of
21.2 ms).  Here every rank owns v NON-adjacent chunks of the model (virtual stage k = chunk k // P
of rank k % P, ``Allocator.looped_allocate``), the ranks form a ring, and a micro-batch visits
every rank v times.  Each rank runs, per step,

    for c in chunks:            for j in micro-batches:  forward (c, j)
    for c in reversed(chunks):  for j in micro-batches:  backward(c, j)

(the breadth-first order of Lamy-Poirier, "Breadth-first pipeline parallelism", 2022).  A hop now
costs 1/v of a rank's work, so the step is ``m x T + (P - 1) x (F + B) / v`` as long as
``m >= P``; activation memory grows to all micro-batches of all chunks, which 180 GB of HBM absorb
easily for the models this framework targets.  Weight gradients use the same deferral as the
plain engine.

Transport in this version is ``torch.distributed`` point-to-point (gloo on CPU, NCCL on GPUs):
both ends of a link post their transfers in the same (chunk, micro-batch) order, so plain FIFO
matching is enough and no tags are needed.  The fused NVLink boundary (one slot per (chunk,
micro-batch), ring channels incl. the wrap-around link) and CUDA-graph capture are the plain
engine's and are not wired in here yet - see DESIGN.md §7.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

from .comm import TorchDistComm


def looped_order(num_chunks: int, micro_batches: int) -> List[tuple]:
    """[('F' | 'B', chunk, micro_batch)] for one rank."""
    order = [("F", c, j) for c in range(num_chunks) for j in range(micro_batches)]
    order += [("B", c, j) for c in reversed(range(num_chunks)) for j in range(micro_batches)]
    return order


class LoopedPipelineEngine:
    def __init__(self, stages: Sequence, virtual_indices: Sequence[int], num_ranks: int,
                 ring: Sequence[int], device: torch.device, optimizer,
                 loss_fn: Optional[Callable] = None, micro_batches: int = 1, group=None,
                 advance_rng: bool = True):
        """``stages[i]`` is the ModuleWrapper of virtual stage ``virtual_indices[i]``; ``ring[p]`` is
        the process rank at ring position p (virtual stage k lives at position k % num_ranks)."""
        self.stages = list(stages)
        self.vidx = list(virtual_indices)
        self.P = num_ranks
        self.v = len(self.stages)
        self.total = self.P * self.v
        assert self.P >= 2, "a looped pipeline needs at least two ranks"
        assert self.vidx == sorted(self.vidx) and len(set(k % self.P for k in self.vidx)) == 1
        self.pos = self.vidx[0] % self.P
        assert self.vidx == [c * self.P + self.pos for c in range(self.v)], self.vidx
        self.ring = list(ring)
        self.device = device
        self.optimizer = optimizer
        self.loss_fn = loss_fn
        self.m = micro_batches
        self.group = group
        self.prev_rank = self.ring[(self.pos - 1) % self.P]
        self.next_rank = self.ring[(self.pos + 1) % self.P]
        self.is_first = self.vidx[0] == 0                 # owns the stage that reads the data
        self.is_last = self.vidx[-1] == self.total - 1    # owns the stage that computes the loss
        self.comm = TorchDistComm(device, group=group)
        self._loss_acc: Optional[torch.Tensor] = None
        self._advance_rng = advance_rng
        self._pending: list = []
        self._order = looped_order(self.v, self.m)
        for st in self.stages:
            st.engine_managed_backward = True
        # attributes the Runner / bench read on any engine
        self.schedule = "looped"
        self.fused = None
        self.in_fused = self.out_fused = False
        self.launches_per_step = 0
        self._graph = None
        self._defer_wgrad = False
        self._setup_done = False

    # ------------------------------------------------------------------ helpers
    def _native_active(self) -> bool:
        from ..models.bert_layers import get_backend
        from ..ops import native as nat

        return self.device.type == "cuda" and get_backend() != "torch" and nat.available()

    @staticmethod
    def _diff_flags(tensors) -> List[bool]:
        n = len(tensors)
        return [torch.is_tensor(t) and t.is_floating_point() and (i < n - 1 or n == 1)
                for i, t in enumerate(tensors)]

    def _prep_inputs(self, tensors):
        for t, d in zip(tensors, self._diff_flags(tensors)):
            if d:
                t.requires_grad_(True)
        return tuple(tensors)

    def _flush_wgrads(self) -> None:
        if self._defer_wgrad:
            from ..ops.functions import flush_wgrads

            flush_wgrads()

    # ------------------------------------------------------------------ one optimisation step
    def train_step(self, inputs: Optional[Sequence[torch.Tensor]] = None,
                   labels: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        if not self._setup_done:
            self._loss_acc = torch.zeros((), dtype=torch.float32, device=self.device)
            if self._native_active():
                from ..ops.functions import set_wgrad_deferral

                self._defer_wgrad = True
                set_wgrad_deferral(True)
            self._setup_done = True
        if self._advance_rng and self._native_active():
            from ..models.bert_layers import advance_rng

            advance_rng()
        self._loss_acc.zero_()
        chunks_in = [t.chunk(self.m, dim=0) for t in inputs] if self.is_first else None
        label_chunks = labels.chunk(self.m, dim=0) if (self.is_last and labels is not None) else None
        saved = {}
        for kind, c, j in self._order:
            k = self.vidx[c]
            st = self.stages[c]
            st.microbatch = c * self.m + j
            st.in_channel = st.out_channel = None
            first_stage, last_stage = k == 0, k == self.total - 1
            if kind == "F":
                if first_stage:
                    args = tuple(t[j] for t in chunks_in)
                else:
                    got, reqs = self.comm.recv(self.prev_rank, f"fwd{c}")
                    self.comm.wait(reqs)
                    args = self._prep_inputs(got)
                outs = st(*args)
                loss = None
                if last_stage:
                    loss = self.loss_fn(outs[0], label_chunks[j]) / self.m
                    self._loss_acc += loss.detach().float()
                else:
                    self._pending += self.comm.send(
                        [o.detach() if torch.is_tensor(o) else None for o in outs],
                        self.next_rank, f"fwd{c}")
                saved[(c, j)] = (args, outs, loss)
            else:
                args, outs, loss = saved.pop((c, j))
                self._flush_wgrads()
                st.begin_backward()
                if last_stage:
                    loss.backward()
                else:
                    metas = [(o.dtype, tuple(o.shape)) if d else None
                             for o, d in zip(outs, self._diff_flags(outs))]
                    grads, reqs = self.comm.recv(self.next_rank, f"bwd{c}", metas=metas)
                    self.comm.wait(reqs)
                    ts, gs = [], []
                    for o, g in zip(outs, grads):
                        if g is not None and torch.is_tensor(o) and o.requires_grad:
                            ts.append(o)
                            gs.append(g.to(o.dtype))
                    torch.autograd.backward(ts, gs)
                st.end_backward()
                if not first_stage:
                    in_grads = [(a.grad if a.grad is not None else torch.zeros_like(a)) if d else None
                                for a, d in zip(args, self._diff_flags(args))]
                    self._pending += self.comm.send(in_grads, self.prev_rank, f"bwd{c}",
                                                    with_meta=False)
        self._flush_wgrads()
        self.optimizer.step()
        self.comm.wait(self._pending)
        self._pending = []
        return self._loss_acc if self.is_last else None

    # ------------------------------------------------------------------ plain forward (eval)
    @torch.no_grad()
    def forward_only(self, inputs: Optional[Sequence[torch.Tensor]] = None):
        outs = None
        for c, st in enumerate(self.stages):
            k = self.vidx[c]
            if k == 0:
                args = tuple(inputs)
            else:
                args, reqs = self.comm.recv(self.prev_rank, f"eval{c}")
                self.comm.wait(reqs)
            st.in_channel = st.out_channel = None
            outs = st(*args)
            if k != self.total - 1:
                self.comm.wait(self.comm.send(list(outs), self.next_rank, f"eval{c}"))
                outs = None
        return outs

    def close(self) -> None:
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        if self._defer_wgrad:
            from ..ops.functions import set_wgrad_deferral

            set_wgrad_deferral(False)
            self._defer_wgrad = False

    def trace(self) -> list:
        return []

    @staticmethod
    def barrier(group=None) -> None:
        if dist.is_available() and dist.is_initialized():
            dist.barrier(group=group)
