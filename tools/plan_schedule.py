"""Print the schedule planner's ranking of (micro-batch size, count, virtual stages) plans.

    python tools/plan_schedule.py --gpus 8 --single-gpu-ms 11.72 --per-gpu-batch 32 --blocks 24
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skycomputing_b200 as sky  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, nargs="+", default=[2, 4, 8])
    ap.add_argument("--single-gpu-ms", type=float, default=11.72)
    ap.add_argument("--per-gpu-batch", type=int, default=32)
    ap.add_argument("--blocks", type=int, default=24)
    ap.add_argument("--top", type=int, default=5)
    a = ap.parse_args()
    for n in a.gpus:
        costs = sky.costs_from_single_gpu_step(a.single_gpu_ms * 1e-3, n, sequences=a.per_gpu_batch)
        planner = sky.SchedulePlanner(n, costs, blocks_per_stage=max(1, a.blocks // n))
        gb = a.per_gpu_batch * n
        print(f"== {n} GPUs, global batch {gb}")
        for p in planner.candidates(gb)[: a.top]:
            eff = gb / p.step_seconds / (n * a.per_gpu_batch / (a.single_gpu_ms * 1e-3))
            print(f"   {p.schedule:10s} micro-batch {p.micro_batch:3d} x {p.micro_batches:3d}  "
                  f"v={p.virtual_stages:2d}  {p.step_seconds * 1e3:6.2f} ms  "
                  f"{gb / p.step_seconds:8.0f} seq/s  scaling efficiency {eff:.2f}")


if __name__ == "__main__":
    main()
