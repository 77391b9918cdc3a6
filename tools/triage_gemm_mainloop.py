"""Which side of the main loop is slow?  debug=1: full main loop, epilogue dropped; debug=3: MMAs
only (no TMA loads, stale smem); debug=4: TMA loads only (no MMAs).  CUDA-graph timed."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skycomputing_b200.ops import native as nat  # noqa: E402
from tools.triage_gemm import timeit  # noqa: E402

rows = []
for (name, m, n, k) in [("attn_out", 2048, 1024, 1024), ("ffn2", 2048, 1024, 4096), ("ffn2", 4096, 1024, 4096),
                        ("qkv", 4096, 3072, 1024), ("big", 8192, 8192, 8192), ("deepK", 128, 256, 65536)]:
    a = torch.randn(m, k, device="cuda").bfloat16()
    b = torch.randn(n, k, device="cuda").bfloat16()
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    row = dict(name=name, M=m, N=n, K=k)
    for bn in (128, 256):
        for dbg, tag in ((0, "full"), (1, "main"), (3, "mma_only"), (4, "load_only")):
            t = timeit(lambda: nat.gemm(a, b, out=out, block_n=bn, pair=0, stream_k=0, debug=dbg), iters=10)
            row[f"bn{bn}_{tag}"] = round(t * 1e3, 1)
    rows.append(row)
    print(json.dumps(row), flush=True)
json.dump(rows, open("gpurun_out/triage_mainloop.json", "w"), indent=1)
