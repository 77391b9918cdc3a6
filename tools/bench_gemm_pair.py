"""cta_group::2 (CTA pair) vs single-CTA tcgen05 GEMM vs cuBLAS: correctness + time on the BERT-large
shapes at 4096 and 2048 tokens, forward / dgrad / wgrad operand layouts."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skycomputing_b200.ops import native as nat  # noqa: E402
from tools.bench_gemm import timeit  # noqa: E402


def main():
    res = []
    torch.manual_seed(0)
    shapes = []
    for M in (4096, 2048):
        shapes += [("qkv", M, 3072, 1024), ("attn_out", M, 1024, 1024), ("ffn1", M, 4096, 1024),
                   ("ffn2", M, 1024, 4096)]
    shapes.append(("big", 8192, 8192, 8192))
    shapes.append(("odd", 1000, 520, 328))
    for name, m, n, k in shapes:
        a = torch.randn(m, k, device="cuda").bfloat16()
        b = torch.randn(n, k, device="cuda").bfloat16()
        bias = torch.randn(n, device="cuda")
        ref = (a.float() @ b.float().t() + bias)
        dy = torch.randn(m, n, device="cuda").bfloat16()
        ref_dx = dy.float() @ b.float()
        ref_dw = dy.float().t() @ a.float()
        scale = ref.abs().max().item()
        for pair in (0, 1):
            for bn in (128, 256):
                tag = f"pair{pair}_bn{bn}"
                out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
                nat.gemm(a, b, out=out, bias=bias, block_n=bn, pair=pair)
                err = (out.float() - ref).abs().max().item() / scale
                t = timeit(lambda: nat.gemm(a, b, out=out, bias=bias, block_n=bn, pair=pair))
                res.append(dict(name=name, M=m, impl="fwd_" + tag, ms=t, tflops=2 * m * n * k / t / 1e9, err=err))
                dx = torch.empty(m, k, device="cuda", dtype=torch.bfloat16)
                nat.gemm(dy, b, b_mn=True, out=dx, block_n=bn, pair=pair)
                err = (dx.float() - ref_dx).abs().max().item() / ref_dx.abs().max().item()
                t = timeit(lambda: nat.gemm(dy, b, b_mn=True, out=dx, block_n=bn, pair=pair))
                res.append(dict(name=name, M=m, impl="dgrad_" + tag, ms=t, tflops=2 * m * n * k / t / 1e9, err=err))
                if m % 8 == 0:
                    dw = torch.zeros(n, k, device="cuda", dtype=torch.float32)
                    nat.gemm(dy, a, a_mn=True, b_mn=True, out=dw, accumulate=True, block_n=bn, pair=pair)
                    err = (dw - ref_dw).abs().max().item() / ref_dw.abs().max().item()
                    t = timeit(lambda: nat.gemm(dy, a, a_mn=True, b_mn=True, out=dw, accumulate=True,
                                                block_n=bn, pair=pair))
                    res.append(dict(name=name, M=m, impl="wgrad_" + tag, ms=t, tflops=2 * m * n * k / t / 1e9, err=err))
        out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        t = timeit(lambda: torch.matmul(a, b.t(), out=out))
        res.append(dict(name=name, M=m, impl="cublas_fwd", ms=t, tflops=2 * m * n * k / t / 1e9, err=0))
    for r in res:
        print(json.dumps(r))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_gemm_pair.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
