"""Micro-benchmark of the tcgen05 GEMM on BERT-large shapes vs torch.matmul (cuBLAS)."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skycomputing_b200.ops import native as nat  # noqa: E402


def timeit(fn, iters=20, warmup=5):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


def main():
    res = []
    M = 4096
    shapes = [("qkv", M, 3072, 1024, {}), ("attn_out", M, 1024, 1024, {}),
              ("ffn1_gelu", M, 4096, 1024, {"act": nat.ACT_GELU}), ("ffn2", M, 1024, 4096, {}),
              ("big", 8192, 8192, 8192, {})]
    for name, m, n, k, kw in shapes:
        a = torch.randn(m, k, device="cuda").bfloat16()
        b = torch.randn(n, k, device="cuda").bfloat16()
        bias = torch.zeros(n, device="cuda")
        out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        for bn in (128, 256):
            t = timeit(lambda: nat.gemm(a, b, out=out, bias=bias, block_n=bn, **kw))
            res.append(dict(name=name, impl=f"tcgen05_bn{bn}", ms=t, tflops=2 * m * n * k / t / 1e9))
        t = timeit(lambda: torch.matmul(a, b.t(), out=out))
        res.append(dict(name=name, impl="cublas", ms=t, tflops=2 * m * n * k / t / 1e9))
        # dgrad / wgrad layouts
        dy = torch.randn(m, n, device="cuda").bfloat16()
        dx = torch.empty(m, k, device="cuda", dtype=torch.bfloat16)
        t = timeit(lambda: nat.gemm(dy, b, b_mn=True, out=dx))
        res.append(dict(name=name, impl="tcgen05_dgrad", ms=t, tflops=2 * m * n * k / t / 1e9))
        dw = torch.zeros(n, k, device="cuda", dtype=torch.float32)
        t = timeit(lambda: nat.gemm(dy, a, a_mn=True, b_mn=True, out=dw, accumulate=True))
        res.append(dict(name=name, impl="tcgen05_wgrad", ms=t, tflops=2 * m * n * k / t / 1e9))
    for r in res:
        print(json.dumps(r))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_gemm.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
