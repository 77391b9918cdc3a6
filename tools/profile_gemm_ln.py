"""Launches the GEMM + LayerNorm epilogue kernel (FFN2 shape, BERT-large, 4096 tokens) a few
times: the target of  ncu --set full --clock-control none --import-source on -k regex:gemm_ln -s 3 -c 1"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skycomputing_b200.ops import native as nat  # noqa: E402

M, N, K = 4096, 1024, 4096
torch.manual_seed(0)
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
res = torch.randn(M, N, device="cuda").bfloat16()
bias = torch.randn(N, device="cuda")
g, b = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
rng = nat.RngState(1)
for _ in range(6):
    nat.gemm_ln(a, w, g, b, bias=bias, residual=res, dropout_p=0.1, rng=rng, rng_stream=3)
torch.cuda.synchronize()
print("done")
