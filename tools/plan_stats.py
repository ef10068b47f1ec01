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


def pass_stats(words, woff, n, reals):
    h0 = words[woff]
    a, r, n_rounds = h0 & 0xFF, (h0 >> 16) & 0xFF, (h0 >> 24) & 0xFF
    tile = [((words[woff + 1 + i // 8]) >> (8 * (i % 8))) & 0xFF for i in range(a)]
    runs, prev = 0, None
    for q in tile[3:]:
        runs += prev is None or q != prev + 1
        prev = q
    pos = woff + (words[woff + 3] >> 32)
    per = 1 << r
    out = dict(a=a, rounds=n_rounds, runs=runs, records=0, layers=0, rotations=0, rot3=0, tabs=0, fans=0, other=0, fp64=0.0,
               conflict_rounds=0, tile=tile)
    for _ in range(n_rounds):
        ng, nw = words[pos] & 0xFFFF, words[pos] >> 16
        order = [((words[pos + 1 + i // 8]) >> (8 * (i % 8))) & 0xFF for i in range(a)]
        lanes = order[r:r + 3]
        classes = {(b % 3) for b in lanes if b < 6}
        if len(classes) < 3:
            out["conflict_rounds"] += 1
        for g in range(ng):
            g0 = words[pos + 3 + 3 * g]
            tk, j, dk, rm = g0 & 0xFF, (g0 >> 8) & 0xFF, (g0 >> 16) & 0xFF, (g0 >> 48) & 0xFFFF
            out["records"] += 1
            f = 0.0
            if tk == P.T_LAYER:
                k = bin(rm & 15).count("1")
                out["layers"] += 1
                out["rotations"] += k
                roff = words[pos + 5 + 3 * g] & 0xFFFFFFFF
                for b in range(4):                # 4 or 6 DFMA per pair (third shear only when y != 0)
                    if (rm >> b) & 1:
                        three = reals[roff + 4 * b + 2] != 0.0
                        out["rot3"] += three
                        f += 3.0 if three else 2.0
            elif tk in (P.T_MATC, P.T_MATR):
                f += 8.0 if tk == P.T_MATC else 4.0
                out["other"] += 1
            elif tk == P.T_X:
                out["other"] += 1
            if dk == P.D_TAB:
                out["tabs"] += 1
                f += 4.0
            elif dk in (P.D_FAN, P.D_FANT, P.D_W, P.D_FIX):
                out["fans"] += 1
                frac = 0.5 if j < r else 1.0
                f += frac * (8.0 if dk == P.D_FAN else 4.0) + 8.0 / per
            elif dk == P.D_PHASE:
                f += 4.0 / (1 << bin(rm).count("1"))
            out["fp64"] += f
        pos += nw
    return out


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
    k = 0
    for i, o in enumerate(prog._ops):
        if o[0] != 16:
            continue
        st = pass_stats(prog._words, o[5], n, prog._reals)
        fp_ms = st["fp64"] * 2 ** n / (SM * DFMA_PER_CLK * CLK) * 1e3
        ms = times[i] if times else float("nan")
        tot["fp64"] += st["fp64"]
        tot["ms"] += ms
        print(f"pass {k:2d}: rounds {st['rounds']} records {st['records']:3d} (layer {st['layers']:2d} rot {st['rotations']:2d}/{st['rot3']} "
              f"tab {st['tabs']:2d} fan {st['fans']:2d} other {st['other']:2d}) fp64/amp {st['fp64']:6.1f} = {fp_ms:5.2f} ms "
              f"(hbm {hbm_ms:.2f}) runs {st['runs']} conflict-rounds {st['conflict_rounds']}  measured {ms:6.2f} ms  tile {st['tile'][3:]}")
        k += 1
    print(f"total fp64/amp {tot['fp64']:.0f} = {tot['fp64'] * 2 ** n / (SM * DFMA_PER_CLK * CLK) * 1e3:.1f} ms at 64 DFMA/clk/SM, "
          f"{k} passes = {k * hbm_ms:.1f} ms of HBM time, measured {tot['ms']:.1f} ms")


if __name__ == "__main__":
    main()
