"""Stream-K / CTA-pair / classic schedules of the tcgen05 GEMM vs cuBLAS: max error against an fp32
reference and CUDA-graph-replayed time, BERT-large shapes at 4096 / 2048 / 1024 tokens."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skycomputing_b200.ops import native as nat  # noqa: E402
from tools.triage_gemm import timeit  # noqa: E402

CONFIGS = [
    ("classic_auto", dict(pair=0, stream_k=0)),
    ("classic_256", dict(pair=0, stream_k=0, block_n=256)),
    ("classic_128", dict(pair=0, stream_k=0, block_n=128)),
    ("pair_256", dict(pair=1, stream_k=0, block_n=256)),
    ("quad_256", dict(pair=2, stream_k=0, block_n=256)),
    ("quad_128", dict(pair=2, stream_k=0, block_n=128)),
    ("quad_sk_128", dict(pair=2, stream_k=1, block_n=128)),
    ("auto", dict()),
]


def main():
    res = []
    torch.manual_seed(0)
    shapes = []
    for M in (4096, 2048, 1024):
        shapes += [("qkv", M, 3072, 1024), ("attn_out", M, 1024, 1024), ("ffn1", M, 4096, 1024),
                   ("ffn2", M, 1024, 4096)]
    shapes.append(("odd", 1000, 520, 1096))
    for name, m, n, k in shapes:
        a = torch.randn(m, k, device="cuda").bfloat16()
        b = torch.randn(n, k, device="cuda").bfloat16()
        bias = torch.randn(n, device="cuda")
        ref = (a.float() @ b.float().t() + bias)
        dy = torch.randn(m, n, device="cuda").bfloat16()
        ref_dx = dy.float() @ b.float()
        ref_dw = dy.float().t() @ a.float()
        for tag, kw in CONFIGS:
            out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
            for _ in range(2):  # twice: the second launch sees the counters the first one left
                out.zero_()
                nat.gemm(a, b, out=out, bias=bias, **kw)
            err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
            t = timeit(lambda: nat.gemm(a, b, out=out, bias=bias, **kw))
            res.append(dict(name=name, M=m, kind="fwd", impl=tag, us=t * 1e3, err=err))
            dx = torch.empty(m, k, device="cuda", dtype=torch.bfloat16)
            nat.gemm(dy, b, b_mn=True, out=dx, **kw)
            err = (dx.float() - ref_dx).abs().max().item() / ref_dx.abs().max().item()
            t = timeit(lambda: nat.gemm(dy, b, b_mn=True, out=dx, **kw))
            res.append(dict(name=name, M=m, kind="dgrad", impl=tag, us=t * 1e3, err=err))
            if m % 8 == 0:
                dw = torch.zeros(n, k, device="cuda", dtype=torch.float32)
                nat.gemm(dy, a, a_mn=True, b_mn=True, out=dw, accumulate=True, **kw)
                err = (dw - ref_dw).abs().max().item() / ref_dw.abs().max().item()
                t = timeit(lambda: nat.gemm(dy, a, a_mn=True, b_mn=True, out=dw, accumulate=True, **kw))
                res.append(dict(name=name, M=m, kind="wgrad", impl=tag, us=t * 1e3, err=err))
        out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
        t = timeit(lambda: torch.matmul(a, b.t(), out=out))
        res.append(dict(name=name, M=m, kind="fwd", impl="cublas", us=t * 1e3, err=0))
        dx = torch.empty(m, k, device="cuda", dtype=torch.bfloat16)
        t = timeit(lambda: torch.matmul(dy, b, out=dx))
        res.append(dict(name=name, M=m, kind="dgrad", impl="cublas", us=t * 1e3, err=0))
        dwb = torch.empty(n, k, device="cuda", dtype=torch.bfloat16)
        t = timeit(lambda: torch.matmul(dy.t(), a, out=dwb))
        res.append(dict(name=name, M=m, kind="wgrad", impl="cublas", us=t * 1e3, err=0))
        print(name, m, "done", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_gemm_streamk.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
