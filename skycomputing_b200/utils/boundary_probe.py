"""Exposed stage-boundary communication time, measured (BASELINE.json names it as part of the
headline metric; the reference's hop is scaelum/builder/module_wrapper.py:148-175).

Both boundary kernels are timed twice on the ranks that own them, CUDA-graph replays of
back-to-back launches, both directions running at the same time like in the pipeline:

* forward link  (rank 0 -> 1): the stage's closing FFN2 GEMM with bias + dropout + residual +
  LayerNorm in its tcgen05 epilogue, storing y into the NEXT stage's HBM + panel flags (for
  micro-batches too small for that kernel: the standalone LayerNorm producer), vs local stores;
* backward link (rank 1 -> 0): the QKV dgrad GEMM whose epilogue stores the input gradient into
  the PREVIOUS stage's HBM + per-tile flags, vs local stores.

``exposed`` = peer - local is what one crossing adds to the critical path; the roofline of a fused
path is the slower of its compute with local stores and its bytes over NVLink 5 (900 GB/s per
direction), and ``roofline_fraction`` = roofline / measured.  Collective over the process group.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

NVLINK_BYTES_PER_S = 900e9


def _time_us(fn, iters: int = 20, warmup: int = 3) -> float:
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warmup):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def probe(tokens: int, hidden: int = 1024, intermediate: int = 4096, group=None) -> Optional[dict]:
    """Returns (on every rank) {"forward": {...}, "backward": {...}} measured on ranks 0 / 1 of a
    >= 2-rank group, or None for a single rank."""
    from ..ops import native as nat
    from ..ops.functions import _ln_fwd
    from ..parallel.p2p import FusedBoundaryManager

    world = dist.get_world_size(group)
    if world < 2:
        return None
    rank = dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    M, H = tokens, hidden
    mgr = FusedBoundaryManager(rank, world, list(range(world)), 1, M, H, M, dev, group=group)
    nbytes = M * H * 2
    floor_us = nbytes / NVLINK_BYTES_PER_S * 1e6
    row = None
    torch.manual_seed(rank)
    if rank == 0:
        ch = mgr.next
        g, b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
        if nat.gemm_ln_supported(M, H):
            inter = torch.randn(M, intermediate, device=dev).bfloat16()
            w2 = (torch.randn(H, intermediate, device=dev) * 0.02).bfloat16()
            res = torch.randn(M, H, device=dev).bfloat16()
            bias = torch.zeros(H, device=dev)
            kern = "FFN2 GEMM + bias + residual + LayerNorm epilogue (tcgen05)"
            local = lambda: nat.gemm_ln(inter, w2, g, b, bias=bias, residual=res)  # noqa: E731
            peer = lambda: nat.gemm_ln(inter, w2, g, b, bias=bias, residual=res,   # noqa: E731
                                       y_ptr=ch.peer_act_ptr(0), y_ld=H,
                                       signal_flags=ch.peer_act_flags_ptr(0))
        else:
            z = torch.randn(M, H, device=dev).bfloat16()
            kern = "closing LayerNorm"
            local = lambda: _ln_fwd(z, g, b, 1e-12)                                # noqa: E731
            peer = lambda: _ln_fwd(z, g, b, 1e-12, ch.peer_act_ptr(0),             # noqa: E731
                                   ch.peer_act_flags_ptr(0))
        name = "forward"
    elif rank == 1:
        ch = mgr.prev
        dqkv = torch.randn(M, 3 * H, device=dev).bfloat16()
        w = (torch.randn(3 * H, H, device=dev) * 0.02).bfloat16()
        aux = torch.randn(M, H, device=dev).bfloat16()
        kern = "QKV dgrad GEMM + residual-path gradient (tcgen05)"
        local = lambda: nat.gemm(dqkv, w, b_mn=True, aux=aux, add_aux=True)        # noqa: E731
        peer = lambda: nat.gemm(dqkv, w, b_mn=True, aux=aux, add_aux=True,         # noqa: E731
                                out_ptr=ch.peer_grad_ptr(0), out_ld=ch.grad_ld,
                                signal_flags=ch.peer_grad_flags_ptr(0))
        name = "backward"
    if rank in (0, 1):
        t_local = _time_us(local)
        torch.cuda.synchronize()
    dist.barrier(group=group)
    if rank in (0, 1):
        t_peer = _time_us(peer)      # ranks 0 and 1 run their peer variants at the same time
        roof = max(t_local, floor_us)
        row = dict(direction=name, kernel=kern, tokens=M, bytes=nbytes,
                   nvlink_floor_us=round(floor_us, 2), kernel_local_us=round(t_local, 2),
                   kernel_peer_us=round(t_peer, 2), exposed_us=round(max(t_peer - t_local, 0.0), 2),
                   roofline_fraction=round(roof / t_peer, 3), flag_errors=int(mgr.error_code()))
    gathered = [None] * world
    dist.all_gather_object(gathered, row, group=group)
    mgr.close()
    out = {r["direction"]: r for r in gathered if r}
    return out
