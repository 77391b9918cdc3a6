"""Where does a GEMM's time go?  Runs each BERT shape with the epilogue progressively disabled
(debug=1: accumulator discarded, debug=2: epilogue math but no global stores) and with different
CTA counts, so main-loop (TMA + tcgen05) speed can be separated from epilogue cost."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skycomputing_b200.ops import native as nat  # noqa: E402


_SIDE = None


def timeit(fn, iters=20, warmup=5):
    """Time `fn` as `iters` back-to-back launches replayed from a CUDA graph: host launch cost
    (tensor-map encode + pybind, ~14 us per call) would otherwise hide every kernel shorter than
    that."""
    global _SIDE
    if _SIDE is None:
        _SIDE = torch.cuda.Stream()
    side = _SIDE
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        # warm up ON the capture stream: per-stream resources (the stream-K workspace) cannot be
        # created while the stream is capturing
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
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * iters)


def main():
    out_rows = []
    for M in (4096, 2048):
        for name, n, k, kw in [("qkv", 3072, 1024, {}), ("attn_out", 1024, 1024, {}),
                               ("ffn1_gelu", 4096, 1024, {"act": nat.ACT_GELU}),
                               ("ffn2", 1024, 4096, {})]:
            a = torch.randn(M, k, device="cuda").bfloat16()
            b = torch.randn(n, k, device="cuda").bfloat16()
            bias = torch.zeros(n, device="cuda")
            out = torch.empty(M, n, device="cuda", dtype=torch.bfloat16)
            out2 = torch.empty_like(out) if "act" in kw else None
            row = dict(M=M, name=name)
            for bn in (128, 256):
                for dbg in (0, 2, 1):
                    t = timeit(lambda: nat.gemm(a, b, out=out, bias=bias, out2=out2, block_n=bn,
                                                debug=dbg, **kw))
                    row[f"bn{bn}_dbg{dbg}_us"] = round(t * 1e3, 1)
            t = timeit(lambda: torch.matmul(a, b.t(), out=out))
            row["cublas_us"] = round(t * 1e3, 1)
            row["ideal_us_at_1460TF"] = round(2 * M * n * k / 1460e12 * 1e6, 1)
            out_rows.append(row)
            print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out_rows, open("gpurun_out/triage_gemm.json", "w"), indent=1)


if __name__ == "__main__":
    main()
