#!/bin/bash
# GPU validation of the looped pipeline (needs >= 2 GPUs; run through
#   gpurun --gpus N --timeout 900 -- tools/validate_looped.sh N V
# ).  Step 1 compares loss trajectories (same seeds per global layer index, dropout off):
#   plain pipeline over NCCL  ==  plain fused  ==  looped over NCCL p2p  ==  looped fused ring.
# Step 2 times bench.py: plain 1F1B vs the planner's looped choice.  Results: gpurun_out/looped/.
set -u
N=${1:-2}
V=${2:-2}
OUT=gpurun_out/looped
mkdir -p $OUT
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$N" \
          --master-addr 127.0.0.1 --master-port "$1" "${@:2}"; }
L=$((N * V))
run 29530 tools/check_pipeline.py --boundary nccl --layers $L --micro-batches "$N" \
    --tag plain_nccl 2>&1 | grep "CHECK\|Error\|error" | tee $OUT/check_n${N}_plain_nccl.json
run 29531 tools/check_pipeline.py --boundary fused --layers $L --micro-batches "$N" \
    --tag plain_fused 2>&1 | grep "CHECK\|Error\|error" | tee $OUT/check_n${N}_plain_fused.json
SKY_LOOPED_FUSED=0 run 29532 tools/check_pipeline.py --virtual-stages "$V" --layers $L \
    --micro-batches "$N" --tag looped_nccl 2>&1 | grep "CHECK\|Error\|error" | tee $OUT/check_n${N}_looped_nccl.json
run 29533 tools/check_pipeline.py --virtual-stages "$V" --layers $L \
    --micro-batches "$N" --tag looped_fused 2>&1 | grep "CHECK\|Error\|error" | tee $OUT/check_n${N}_looped_fused.json
python - <<PY
import json
runs = {}
for tag in ("plain_nccl", "plain_fused", "looped_nccl", "looped_fused"):
    try:
        line = [l for l in open("$OUT/check_n${N}_%s.json" % tag) if l.startswith("CHECK")][-1]
        runs[tag] = json.loads(line[6:])
    except Exception as e:
        runs[tag] = {"error": repr(e)}
ref = runs["plain_nccl"].get("losses")
verdict = {t: dict(losses=r.get("losses"), graph_all=r.get("graph_all"), err_any=r.get("err_any"),
                   bit_equal_to_plain_nccl=(r.get("losses") == ref and ref is not None),
                   max_abs_diff=(max(abs(a - b) for a, b in zip(r["losses"], ref))
                                 if r.get("losses") and ref and len(ref) == len(r["losses"]) else None))
           for t, r in runs.items()}
json.dump(verdict, open("$OUT/verdict_n${N}.json", "w"), indent=1)
print("VERDICT", json.dumps(verdict))
PY
for v in 1 0; do
    run 29534 bench.py --gpus "$N" --steps 10 --warmup 6 --virtual-stages "$v" 2>$OUT/bench_n${N}_v${v}.err \
        | tail -n 1 > "$OUT/bench_n${N}_v${v}.json"
    python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_n${N}_v${v}.json"))
    print("BENCH N=$N v=$v", d["ms_per_step"], d["value"], d["config"]["plan"], d["config"]["schedule"],
          "graph", d["config"]["cuda_graph"], "err", d["flag_wait_errors"], "loss", d["final_loss"])
except Exception as e:
    print("BENCH N=$N v=$v failed", repr(e)); print(open("$OUT/bench_n${N}_v${v}.err").read()[-3000:])
PY
done
