"""Draw the device timeline written by `SKY_TRACE=1 bench.py` (gpurun_out/trace/n{N}_rank*.json).

One row per pipeline stage, one column per `--res` microseconds: F / B = forward / backward of a
micro-batch (lower-case letter = the same phase still running), W = weight-gradient flush on the
side stream (drawn on its own row), O = optimizer, '.' = idle.  Also prints per-stage busy times.
"""
import argparse
import glob
import json
import os


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default="gpurun_out/trace")
    ap.add_argument("--n", type=int, required=True)
    ap.add_argument("--res", type=float, default=250.0, help="microseconds per character")
    args = ap.parse_args()
    ranks = []
    for p in sorted(glob.glob(os.path.join(args.dir, f"n{args.n}_rank*.json"))):
        ranks.append(json.load(open(p)))
    ranks.sort(key=lambda r: r["stage"])
    spans = {}
    t_min, t_max = None, None
    for r in ranks:
        off = r["clock_offset_ns"]
        open_ = {}
        out = []
        for (kind, j, edge), ns in r["events"]:
            t = (ns - off) / 1e3
            if edge == "begin":
                open_[(kind, j)] = t
            else:
                t0 = open_.pop((kind, j), None)
                if t0 is not None:
                    out.append((kind, j, t0, t))
            if kind != "STEP":
                t_min = t if t_min is None else min(t_min, t)
                t_max = t if t_max is None else max(t_max, t)
        spans[r["stage"]] = out
    width = int((t_max - t_min) / args.res) + 1
    print(f"step span {(t_max - t_min) / 1e3:.2f} ms, {args.res:.0f} us / char")
    for s in sorted(spans):
        main_row = ["."] * width
        w_row = [" "] * width
        busy = {"F": 0.0, "B": 0.0, "W": 0.0, "OPT": 0.0}
        for kind, j, t0, t1 in spans[s]:
            busy[kind] = busy.get(kind, 0.0) + (t1 - t0)
            a, b = int((t0 - t_min) / args.res), int((t1 - t_min) / args.res)
            row = w_row if kind == "W" else main_row
            ch = {"F": "F", "B": "B", "W": "W", "OPT": "O"}[kind]
            for k in range(a, min(b + 1, width)):
                row[k] = ch if k == a else ch.lower()
        print(f"s{s} " + "".join(main_row))
        if any(c != " " for c in w_row):
            print("   " + "".join(w_row))
        n = max(1, sum(1 for k in spans[s] if k[0] == "F"))
        print(f"    F {busy['F'] / n / 1e3:.3f} ms/mb  B {busy['B'] / n / 1e3:.3f} ms/mb  "
              f"W total {busy['W'] / 1e3:.3f} ms  OPT {busy['OPT'] / 1e3:.3f} ms   (F/B include flag waits)")


if __name__ == "__main__":
    main()
