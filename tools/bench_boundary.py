"""Fused stage-boundary paths against their roofline (2 ranks: torchrun --nproc-per-node 2).

Forward link  (rank 0 -> rank 1): the LayerNorm that closes a transformer block stores its output
rows into the NEXT stage's HBM over NVLink and bumps per-panel flags.
Backward link (rank 1 -> rank 0): the QKV dgrad tcgen05 GEMM stores its output tiles from the
epilogue into the PREVIOUS stage's HBM and signals per tile.

For both: time of the producing kernel with local stores vs with peer stores + flags (CUDA-graph
replays of 20 launches, both directions running at the same time like in the pipeline), the NVLink
floor (bytes / 900 GB/s) and the achieved fraction of the roofline max(local compute, NVLink floor);
next to it what the un-fused baseline costs: producer kernel + NCCL send/recv of the same tensor.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skycomputing_b200.ops import native as nat  # noqa: E402
from skycomputing_b200.ops.functions import _ln_fwd  # noqa: E402
from skycomputing_b200.parallel.p2p import FusedBoundaryManager  # noqa: E402
from tools.triage_gemm import timeit  # noqa: E402

NVLINK_BPS = 900e9  # one direction, per GPU (NVLink 5)


def nccl_p2p_us(t: torch.Tensor, rank: int, iters: int = 20) -> float:
    """rank 0 -> rank 1 transfer of `t` with torch.distributed (device-timed on the receiver side)."""
    for _ in range(3):
        (dist.send(t, 1) if rank == 0 else dist.recv(t, 0))
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        (dist.send(t, 1) if rank == 0 else dist.recv(t, 0))
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=2, device_id=dev)
    rows_out = []
    H = 1024
    for M in (4096, 2048):
        mgr = FusedBoundaryManager(rank, 2, [0, 1], 1, M, H, M, dev)
        nbytes = M * H * 2
        floor_us = nbytes / NVLINK_BPS * 1e6
        torch.manual_seed(rank)
        if rank == 0:
            ch = mgr.next
            z = torch.randn(M, H, device=dev).bfloat16()
            g, b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
            t_local = timeit(lambda: _ln_fwd(z, g, b, 1e-12)) * 1e3
            t_peer = timeit(lambda: _ln_fwd(z, g, b, 1e-12, ch.peer_act_ptr(0),
                                            ch.peer_act_flags_ptr(0))) * 1e3
            t_noflag = timeit(lambda: _ln_fwd(z, g, b, 1e-12, ch.peer_act_ptr(0), 0)) * 1e3
            name = "forward: LayerNorm -> next stage's HBM"
        else:
            ch = mgr.prev
            dqkv = torch.randn(M, 3 * H, device=dev).bfloat16()
            w = (torch.randn(3 * H, H, device=dev) * 0.02).bfloat16()
            aux = torch.randn(M, H, device=dev).bfloat16()
            t_local = timeit(lambda: nat.gemm(dqkv, w, b_mn=True, aux=aux, add_aux=True)) * 1e3
            t_peer = timeit(lambda: nat.gemm(dqkv, w, b_mn=True, aux=aux, add_aux=True,
                                             out_ptr=ch.peer_grad_ptr(0), out_ld=ch.grad_ld,
                                             signal_flags=ch.peer_grad_flags_ptr(0))) * 1e3
            t_noflag = timeit(lambda: nat.gemm(dqkv, w, b_mn=True, aux=aux, add_aux=True,
                                               out_ptr=ch.peer_grad_ptr(0), out_ld=ch.grad_ld)) * 1e3
            name = "backward: QKV-dgrad GEMM epilogue -> previous stage's HBM"
        torch.cuda.synchronize()
        dist.barrier()
        buf = torch.empty(M, H, device=dev, dtype=torch.bfloat16)
        t_nccl = nccl_p2p_us(buf, rank)
        roof = max(t_local, floor_us)
        row = dict(path=name, tokens=M, bytes=nbytes, nvlink_floor_us=round(floor_us, 2),
                   kernel_local_us=round(t_local, 2), kernel_peer_us=round(t_peer, 2), kernel_peer_noflags_us=round(t_noflag, 2),
                   exposed_us=round(t_peer - t_local, 2), roofline_fraction=round(roof / t_peer, 3),
                   nccl_p2p_us=round(t_nccl, 2),
                   unfused_baseline_us=round(t_local + t_nccl, 2),
                   flag_errors=int(mgr.error_code()))
        gathered = [None, None]
        dist.all_gather_object(gathered, row)
        if rank == 0:
            rows_out += gathered
            for r in gathered:
                print(json.dumps(r), flush=True)
        mgr.close()
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(rows_out, open("gpurun_out/bench_boundary.json", "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
