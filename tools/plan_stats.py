#!/usr/bin/env python
"""What the planner asks of csrc/tile.cu, pass by pass: rounds, records by kind, FP64 instructions per amplitude as the
kernel executes them, bank-conflict-free rounds, TMA copies per tile -- and the HBM / FP64 floors that follow.

    python tools/plan_stats.py [--workload layers|qft|qaoa] [--qubits 30] [--depth 20] [--times detail.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from qaqarot_b200 import planner as P, workloads as W          # noqa: E402
from qaqarot_b200.lowering import lower                        # noqa: E402

SM, CLK, DFMA_PER_CLK = 148, 1.965e9, 64


pass_stats = P.describe_tile_pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="layers")
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--times", default="")
    ap.add_argument("--max-cost", type=float, default=0.0)
    args = ap.parse_args()
    n = args.qubits
    circ = {"layers": lambda: W.random_layers(n, args.depth), "qft": lambda: W.qft(n),
            "qaoa": lambda: W.qaoa_maxcut(n)[0]}[args.workload]()
    cfg = P.PlanConfig(**({"max_cost": args.max_cost} if args.max_cost else {}))
    prog = P.plan(lower(list(circ.ops), n).prims, n, cfg, free_initial_layout=True)
    times = None
    if args.times:
        with open(args.times) as f:
            times = json.load(f)["ms"][-1]
    hbm_ms = 32.0 * 2 ** n / 6554.9e9 * 1e3
    tot = dict(fp64=0.0, ms=0.0)
    all_stats = []
    k = 0
    for i, o in enumerate(prog._ops):
        if o[0] != 16:
            continue
        st = pass_stats(prog._words, o[5], n, prog._reals)
        all_stats.append(st)
        fp_ms = st["fp64"] * 2 ** n / (SM * DFMA_PER_CLK * CLK) * 1e3
        ms = times[i + (len(times) - len(prog._ops))] if times else float("nan")
        tot["fp64"] += st["fp64"]
        tot["ms"] += ms
        print(f"pass {k:2d}: rounds {st['rounds']} records {st['records']:3d} (layer {st['layers']:2d} rot {st['rotations']:2d}/{st['rot3']} "
              f"tab {st['tabs']:2d} fan {st['fans']:2d} other {st['other']:2d}) fp64/amp {st['fp64']:6.1f} = {fp_ms:5.2f} ms "
              f"(hbm {hbm_ms:.2f}) runs {st['runs']} conflict-rounds {st['conflict_rounds']}  measured {ms:6.2f} ms  tile {st['tile'][3:]}")
        k += 1
    # round-2 fit of the measured pass times at 30 qubits: 2.8 ms + 1.3 ms per round + 0.017 ms per DFMA/amp +
    # 0.17 ms per record, scaled with the state size
    model = sum(2.8 + 1.3 * s["rounds"] + 0.017 * s["fp64"] + 0.17 * s["records"] for s in all_stats) * 2.0 ** (n - 30)
    print(f"cost model (round-2 fit): {model:.1f} ms for {sum(s['rounds'] for s in all_stats)} rounds, "
          f"{sum(s['records'] for s in all_stats)} records")
    print(f"total fp64/amp {tot['fp64']:.0f} = {tot['fp64'] * 2 ** n / (SM * DFMA_PER_CLK * CLK) * 1e3:.1f} ms at 64 DFMA/clk/SM, "
          f"{k} passes = {k * hbm_ms:.1f} ms of HBM time, measured {tot['ms']:.1f} ms")


if __name__ == "__main__":
    main()
