"""Run one GEMM configuration a few times (for `ncu -k regex:gemm_tcgen05`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skycomputing_b200.ops import native as nat  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "ffn1"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cfgs = {
    "qkv": (M, 3072, 1024, {}),
    "attn_out": (M, 1024, 1024, {}),
    "ffn1": (M, 4096, 1024, {"act": nat.ACT_GELU}),
    "ffn2": (M, 1024, 4096, {}),
}
if which == "wgrad":
    dy = torch.randn(M, 4096, device="cuda").bfloat16()
    x = torch.randn(M, 1024, device="cuda").bfloat16()
    dw = torch.zeros(4096, 1024, device="cuda")
    for _ in range(8):
        nat.gemm(dy, x, a_mn=True, b_mn=True, out=dw, accumulate=True)
elif which == "dgrad":
    dy = torch.randn(M, 4096, device="cuda").bfloat16()
    w = torch.randn(4096, 1024, device="cuda").bfloat16()
    for _ in range(8):
        nat.gemm(dy, w, b_mn=True)
else:
    m, n, k, kw = cfgs[which]
    a = torch.randn(m, k, device="cuda").bfloat16()
    b = torch.randn(n, k, device="cuda").bfloat16()
    bias = torch.zeros(n, device="cuda")
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    out2 = torch.empty(m, n, device="cuda", dtype=torch.bfloat16) if "act" in kw else None
    for _ in range(8):
        nat.gemm(a, b, out=out, bias=bias, out2=out2, **kw)
torch.cuda.synchronize()
