#!/bin/bash
# First GPU validation of the looped pipeline (needs >= 2 GPUs; run through
#   gpurun --gpus 2 --timeout 900 -- tools/validate_looped.sh 2
# ).  Step 1 compares loss trajectories (same seeds per global layer index, dropout off):
#   plain fused pipeline  ==  looped over NCCL p2p  ==  looped over the fused ring boundary.
# Step 2 times bench.py with and without virtual stages.  Results land in gpurun_out/looped/.
set -u
N=${1:-2}
V=${2:-2}
mkdir -p gpurun_out/looped
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" \
          --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
run 29531 tools/check_pipeline.py --boundary fused --layers $((N * V)) --micro-batches "$N" \
    --tag plain 2>&1 | grep CHECK | tee gpurun_out/looped/check_plain.json
run 29532 tools/check_pipeline.py --virtual-stages "$V" --layers $((N * V)) --micro-batches "$N" \
    --tag looped_p2p 2>&1 | grep CHECK | tee gpurun_out/looped/check_looped_p2p.json
SKY_LOOPED_FUSED=1 run 29533 tools/check_pipeline.py --virtual-stages "$V" --layers $((N * V)) \
    --micro-batches "$N" --tag looped_fused 2>&1 | grep "CHECK\|Error\|error" | tee gpurun_out/looped/check_looped_fused.json
for v in 1 "$V"; do
  for fused in 0 1; do
    [ "$v" = 1 ] && [ "$fused" = 1 ] && continue
    SKY_LOOPED_FUSED=$fused run 29534 bench.py --gpus "$N" --steps 10 --warmup 6 --micro-batch 32 \
        --virtual-stages "$v" 2>/dev/null | tail -n 1 > "gpurun_out/looped/bench_v${v}_fused${fused}.json"
    grep -o "ms_per_step\": [0-9.]*" "gpurun_out/looped/bench_v${v}_fused${fused}.json" | head -1 | sed "s/^/v=$v fused=$fused /"
  done
done
