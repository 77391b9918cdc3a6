"""A few launches of our GEMM and of cuBLAS on the same shapes for an `ncu --set full` comparison."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skycomputing_b200.ops import native as nat  # noqa: E402

torch.manual_seed(0)
for (m, n, k) in [(2048, 1024, 4096), (4096, 3072, 1024), (2048, 1024, 1024)]:
    a = torch.randn(m, k, device="cuda").bfloat16()
    b = torch.randn(n, k, device="cuda").bfloat16()
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    for _ in range(3):
        nat.gemm(a, b, out=out, block_n=256, pair=0, stream_k=0)
        nat.gemm(a, b, out=out, block_n=128, pair=0, stream_k=0)
        nat.gemm(a, b, out=out, block_n=256, pair=1, stream_k=0)
        torch.matmul(a, b.t(), out=out)
    torch.cuda.synchronize()
print("done")
