#!/bin/bash
# compute-sanitizer over the kernel numerics tests (the in-kernel producer / consumer
# flag protocol and the mbarrier pipelines are the race surface of this code base).
#   tools/run_sanitizer.sh [memcheck|racecheck|synccheck|initcheck] [pytest -k expression]
# Runs on ONE GPU (e.g. through `gpurun --timeout 1800 -- tools/run_sanitizer.sh racecheck gemm`);
# reports land in gpurun_out/sanitizer_<tool>.log.  The sanitizer slows kernels 10-100x, so the
# flag-wait timeouts (4 s) can fire under racecheck on the cross-GPU tests: those are skipped here.
set -euo pipefail
TOOL=${1:-memcheck}
EXPR=${2:-"gemm or layernorm or attention or colsum or softmax"}
mkdir -p gpurun_out
exec compute-sanitizer --tool "$TOOL" --print-limit 50 --error-exitcode 7 \
  --log-file "gpurun_out/sanitizer_${TOOL}.log" \
  python -m pytest tests/test_kernels_gpu.py -x -q -k "($EXPR) and not handshake" -p no:cacheprovider
