"""Main-loop-only (debug=1) vs full time, single CTA vs CTA pair, plus the fixed cost of a launch
with a single k-block (K=64), back-to-back launches with a warm L2."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skycomputing_b200.ops import native as nat  # noqa: E402
from tools.triage_gemm import timeit  # noqa: E402


def main():
    rows = []
    for M in (4096, 2048):
        for name, n, k in [("qkv", 3072, 1024), ("attn_out", 1024, 1024), ("ffn1", 4096, 1024),
                           ("ffn2", 1024, 4096)]:
            a = torch.randn(M, k, device="cuda").bfloat16()
            b = torch.randn(n, k, device="cuda").bfloat16()
            bias = torch.zeros(n, device="cuda")
            out = torch.empty(M, n, device="cuda", dtype=torch.bfloat16)
            row = dict(M=M, name=name)
            for pair in (0, 1):
                for bn in (128, 256):
                    for dbg in (0, 1):
                        t = timeit(lambda: nat.gemm(a, b, out=out, bias=bias, block_n=bn, pair=pair, debug=dbg))
                        row[f"p{pair}_bn{bn}_{'full' if dbg == 0 else 'main'}"] = round(t * 1e3, 1)
                    t = timeit(lambda: nat.gemm(a[:, :64], b[:, :64], out=out, bias=bias, block_n=bn,
                                                pair=pair, debug=1))
                    row[f"p{pair}_bn{bn}_k64main"] = round(t * 1e3, 1)
            t = timeit(lambda: torch.matmul(a, b.t(), out=out))
            row["cublas"] = round(t * 1e3, 1)
            rows.append(row)
            print(json.dumps(row), flush=True)
    json.dump(rows, open("gpurun_out/triage_gemm_pair.json", "w"), indent=1)


if __name__ == "__main__":
    main()
