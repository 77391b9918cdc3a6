"""GEMM + LayerNorm epilogue (one tcgen05 kernel, cluster-wide row statistics) against the two
kernels it replaces (GEMM with bias/dropout/residual epilogue, then layernorm_fwd) and against
cuBLAS + torch.layer_norm, on the BERT-large K4 / K6 shapes.  One JSON line per shape;
gpurun_out/bench_gemm_ln.json collects them."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skycomputing_b200.ops import native as nat  # noqa: E402
from tools.triage_gemm import timeit  # noqa: E402


def main():
    torch.manual_seed(0)
    rows = []
    rng = nat.RngState(1)
    for name, M, N, K in (("attn_out (K4) 4096 tok", 4096, 1024, 1024),
                          ("ffn2 (K6) 4096 tok", 4096, 1024, 4096),
                          ("attn_out (K4) 2048 tok", 2048, 1024, 1024),
                          ("ffn2 (K6) 2048 tok", 2048, 1024, 4096)):
        a = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
        res = torch.randn(M, N, device="cuda").bfloat16()
        bias = torch.randn(N, device="cuda")
        g, b = torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
        y = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

        def two_kernels():
            z = nat.gemm(a, w, bias=bias, aux=res, add_aux=True, dropout_p=0.1, rng=rng, rng_stream=3)
            nat.layernorm_fwd(z, g, b, y=y)

        def fused(block_n=0):
            nat.gemm_ln(a, w, g, b, bias=bias, residual=res, dropout_p=0.1, rng=rng, rng_stream=3,
                        block_n=block_n)

        def cublas():
            z = torch.nn.functional.linear(a, w) + res
            torch.nn.functional.layer_norm(z, (N,), g.bfloat16(), b.bfloat16())

        def gemm_only():
            nat.gemm(a, w, bias=bias, aux=res, add_aux=True, dropout_p=0.1, rng=rng, rng_stream=3)

        t2 = timeit(two_kernels) * 1e3
        tf = timeit(fused) * 1e3
        tf128 = timeit(lambda: fused(128)) * 1e3
        tf256 = timeit(lambda: fused(256)) * 1e3
        tc = timeit(cublas) * 1e3
        tg = timeit(gemm_only) * 1e3
        flops = 2.0 * M * N * K
        row = dict(shape=name, M=M, N=N, K=K, gemm_then_layernorm_us=round(t2, 2),
                   gemm_alone_us=round(tg, 2), fused_us=round(tf, 2), fused_bn128_us=round(tf128, 2),
                   fused_bn256_us=round(tf256, 2), cublas_plus_torch_ln_us=round(tc, 2),
                   speedup_vs_two_kernels=round(t2 / tf, 3),
                   fused_tflops=round(flops / tf / 1e6, 1),
                   block_n_auto=nat.ext().gemm_ln_block_n(M, N))
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/bench_gemm_ln.json", "w"), indent=1)


if __name__ == "__main__":
    main()
