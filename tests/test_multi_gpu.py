"""Distributed GPU tier: needs >= 2 B200s, skipped otherwise.  Loss trajectories of the
same model / seeds / data must agree between one GPU, a 2-stage pipeline over NCCL p2p and a
2-stage pipeline over the fused NVLink boundary inside one CUDA graph (tools/check_pipeline.py)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world: int, *extra) -> dict:
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "check_pipeline.py"), *extra]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("CHECK ")][-1]
    return json.loads(line[len("CHECK "):])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_stage_pipeline_matches_single_gpu_over_both_boundaries():
    single = _run(1)
    nccl = _run(2, "--boundary", "nccl")
    fused = _run(2, "--boundary", "fused")
    assert fused["fused_any"] and fused["graph_all"] and fused["err_any"] == 0
    assert not nccl["fused_any"]
    assert len(single["losses"]) == len(nccl["losses"]) == len(fused["losses"]) == 6
    for a, b, c in zip(single["losses"], nccl["losses"], fused["losses"]):
        assert b == pytest.approx(a, rel=2e-2, abs=2e-2)
        assert c == pytest.approx(a, rel=2e-2, abs=2e-2)
