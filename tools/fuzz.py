#!/usr/bin/env python
"""Randomised parity fuzzing (not part of the test suite): many random circuits through the planner and either the
numpy descriptor interpreter (default, CPU) or the CUDA library (--gpu), against the oracle.

    python tools/fuzz.py --cases 300                 # CPU, small tile configurations
    python tools/fuzz.py --gpu --cases 150           # on a B200
"""
import argparse
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)

import cases as C                                    # noqa: E402
from oracle import sv_oracle as O                     # noqa: E402
from qaqarot_b200 import B200Backend, Circuit        # noqa: E402
from qaqarot_b200.planner import PlanConfig          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--row-bits", type=int, default=4, help="--gpu: PlanConfig.c (low index bits in every tile)")
    ap.add_argument("--shots", action="store_true", help="fuzz measurement: mid-circuit measure / reset, seeded shot Counters")
    args = ap.parse_args()
    rng = np.random.default_rng(args.seed)
    all_tiles = tuple(range(1, 20))
    worst = 0.0
    for k in range(args.cases):
        if args.gpu:
            n = int(rng.integers(9, 19))
            a = int(rng.integers(9, min(n, 13) + 1))
            cfg = PlanConfig(a=a, c=args.row_bits, max_cost=float(rng.choice([20, 150, 300])), direct_io=bool(rng.integers(2)),
                             direct_out=bool(rng.integers(2)))
            be = B200Backend(plan_config=cfg)
        else:
            from oracle.program_interp import OracleState
            n = int(rng.integers(2, 12))
            a = int(rng.integers(1, n + 1))
            r = int(rng.integers(1, min(a, 4) + 1))
            c = int(rng.integers(0, min(a, 4) + 1))
            cfg = PlanConfig(a=a, r=r, c=c, min_qubits=int(rng.integers(0, 4)), kernel_tiles=all_tiles,
                             max_cost=float(rng.choice([10, 150, 300])), max_rounds=int(rng.integers(1, 7)))
            be = B200Backend(state_factory=lambda nn, d: OracleState(nn, d), plan_config=cfg)
        ops = C.random_ops(rng, n, int(rng.integers(1, 200)), True)
        if args.shots:
            import random
            ops = ops[:60]
            for _ in range(int(rng.integers(0, 3))):          # mid-circuit measurements / resets
                pos = int(rng.integers(0, len(ops) + 1))
                ops.insert(pos, C._op("reset" if rng.integers(3) == 0 else "m", int(rng.integers(n))))
            order = [int(q) for q in rng.permutation(n)[: int(rng.integers(1, n + 1))]]
            ops.append(C._op("m", tuple(order) if len(order) > 1 else order[0]))
            circ = C.build(Circuit, {"n": n, "ops": ops})
            shots = int(rng.integers(1, 40))
            seed = int(rng.integers(1 << 30))
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                random.seed(seed)
                ref = O.OracleBackend().run(circ.ops, n, shots=shots)
                random.seed(seed)
                got = circ.run(backend=be, shots=shots)
            be.close()
            if got != ref:
                print(f"SHOT MISMATCH case {k}: n={n} cfg={cfg} shots={shots} seed={seed}\n  got {dict(got)}\n  ref {dict(ref)}")
                return 1
            continue
        circ = C.build(Circuit, {"n": n, "ops": ops})
        init = None
        if rng.integers(2):
            init = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
            init /= np.linalg.norm(init)
        kw = {} if init is None else {"initial": init}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = O.OracleBackend().run(circ.ops, n, **kw)
            got = circ.run(backend=be, **kw)
        be.close()
        err = float(np.abs(got - ref).max())
        worst = max(worst, err)
        if err > 1e-12:
            print(f"MISMATCH case {k}: n={n} cfg={cfg} ops={len(ops)} init={init is not None} err={err}")
            return 1
    print(f"{args.cases} cases ok, worst max-abs error {worst:.2e}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
