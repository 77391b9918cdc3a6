"""Step-time model of the pipeline schedules, fed by a measured device timeline.

Reads the per-rank traces written by `SKY_TRACE=1 bench.py` (profiles/trace_n8_*/), extracts the
steady-state forward / input-gradient / weight-gradient time of one micro-batch on one stage, and
predicts the step for

  plain 1F1B with deferred weight gradients:  fill (P-1) F  +  m (F + B + W*)  +  drain (P-1) B
  looped / breadth-first with v chunks:       m (F + B + W*)  +  (P-1) (F + B) / v

(W* = the part of the weight-gradient work that does not hide behind the dgrad chain on the side
stream).  Prints the measured span next to the predictions so the model can be checked.
"""
import argparse
import glob
import json
import os
import statistics


def phases(path):
    r = json.load(open(path))
    off = r["clock_offset_ns"]
    open_, out = {}, []
    for (kind, j, edge), ns in r["events"]:
        t = (ns - off) / 1e3
        if edge == "begin":
            open_[(kind, j)] = t
        elif (kind, j) in open_:
            out.append((kind, j, open_.pop((kind, j)), t))
    return r["stage"], out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default="profiles/trace_n8_mb32")
    ap.add_argument("--n", type=int, default=8)
    args = ap.parse_args()
    stages = dict(phases(p) for p in sorted(glob.glob(os.path.join(args.dir, f"n{args.n}_rank*.json"))))
    P = len(stages)
    last = stages[P - 1]
    f = [t1 - t0 for k, j, t0, t1 in last if k == "F"]
    b = [t1 - t0 for k, j, t0, t1 in last if k == "B"]
    w = [t1 - t0 for k, j, t0, t1 in last if k == "W"]
    m = len(b)
    # the last stage never waits in steady state: its F / B phases are pure compute (skip the first
    # F, which includes the pipeline fill)
    F = statistics.median(f[1:]) / 1e3
    B = statistics.median(b) / 1e3
    W = statistics.median(w) / 1e3 if w else 0.0
    t_all = [t for s in stages.values() for k, j, t0, t1 in s if k != "STEP" for t in (t0, t1)]
    span = (max(t_all) - min(t_all)) / 1e3
    cyc = [b2[2] - b1[2] for b1, b2 in zip([x for x in last if x[0] == "B"][:-1],
                                           [x for x in last if x[0] == "B"][1:])]
    T = statistics.median(cyc) / 1e3 if cyc else F + B      # steady-state period of the last stage
    print(f"P={P} m={m}  F={F:.3f} ms  B(dgrad)={B:.3f} ms  W(side stream)={W:.3f} ms  period T={T:.3f} ms")
    print(f"measured step span           {span:6.2f} ms")
    plain = (P - 1) * F + m * T + (P - 1) * B
    print(f"plain 1F1B model             {plain:6.2f} ms  = (P-1) F + m T + (P-1) B")
    for v in (2, 3, 6):
        looped = m * T + (P - 1) * (F + B) / v
        print(f"looped, v={v} chunks per GPU    {looped:6.2f} ms  = m T + (P-1)(F+B)/v")


if __name__ == "__main__":
    main()
