"""NVLink traffic of the two fused stage-boundary kernels, counted by ncu (run under torchrun with
2 ranks; every rank re-executes itself under its own ncu):

    rank 0: FFN2 GEMM + LayerNorm epilogue storing y into rank 1's HBM   (forward link)
    rank 1: QKV dgrad GEMM storing the input gradient into rank 0's HBM  (backward link)

Counters: nvltx__bytes (all / user data / protocol) and nvlrx__bytes of ONE launch next to the
algorithmic payload (tokens x hidden x 2 B).  CSVs land in gpurun_out/nvlink_rank{0,1}.csv.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rank = int(os.environ.get("RANK", "0"))
if os.environ.get("SKY_NCU_CHILD") != "1":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    os.environ["SKY_NCU_CHILD"] = "1"
    kern = "regex:gemm_ln" if rank == 0 else "regex:gemm_tcgen05"
    os.execvp("ncu", ["ncu", "--metrics",
                      "nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvltx__bytes_data_protocol.sum,"
                      "nvlrx__bytes.sum,gpu__time_duration.sum",
                      "--clock-control", "none", "-k", kern, "-s", "2", "-c", "1", "--csv",
                      "--log-file", os.path.join(ROOT, "gpurun_out", f"nvlink_rank{rank}.csv"),
                      sys.executable, os.path.abspath(__file__)])

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

sys.path.insert(0, ROOT)
from skycomputing_b200.ops import native as nat  # noqa: E402
from skycomputing_b200.parallel.p2p import FusedBoundaryManager  # noqa: E402

torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", rank=rank, world_size=2, device_id=dev)
M, H, I = 4096, 1024, 4096
mgr = FusedBoundaryManager(rank, 2, [0, 1], 1, M, H, M, dev)
torch.manual_seed(rank)
if rank == 0:
    ch = mgr.next
    inter = torch.randn(M, I, device=dev).bfloat16()
    w2 = (torch.randn(H, I, device=dev) * 0.02).bfloat16()
    res = torch.randn(M, H, device=dev).bfloat16()
    g, b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    for _ in range(4):
        nat.gemm_ln(inter, w2, g, b, bias=b, residual=res, y_ptr=ch.peer_act_ptr(0), y_ld=H,
                    signal_flags=ch.peer_act_flags_ptr(0))
else:
    ch = mgr.prev
    dqkv = torch.randn(M, 3 * H, device=dev).bfloat16()
    w = (torch.randn(3 * H, H, device=dev) * 0.02).bfloat16()
    aux = torch.randn(M, H, device=dev).bfloat16()
    for _ in range(4):
        nat.gemm(dqkv, w, b_mn=True, aux=aux, add_aux=True, out_ptr=ch.peer_grad_ptr(0),
                 out_ld=ch.grad_ld, signal_flags=ch.peer_grad_flags_ptr(0))
torch.cuda.synchronize()
dist.barrier()
print(f"rank {rank} done; algorithmic payload {M * H * 2} bytes", flush=True)
mgr.close()
dist.destroy_process_group()
