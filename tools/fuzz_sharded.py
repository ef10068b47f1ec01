#!/usr/bin/env python
"""Randomised parity fuzzing of the sharded path on CPU (not part of the test suite): world_size ranks under gloo,
the numpy program interpreter as every rank's device, random circuits and random small tile configurations
against the oracle.

    python tools/fuzz_sharded.py --world 4 --cases 40 [--seed 0]
"""
import argparse
import os
import socket
import sys
import traceback
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, cases, seed, errors):
    try:
        for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
            if p not in sys.path:
                sys.path.insert(0, p)
        import random
        import torch.distributed as dist
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
        import cases as C
        from oracle import sv_oracle as O
        from oracle.program_interp import OracleState
        from qaqarot_b200 import Circuit, ShardedB200Backend
        from qaqarot_b200.planner import PlanConfig

        g = world.bit_length() - 1
        rng = np.random.default_rng(seed)                   # identical on every rank
        worst = 0.0
        for k in range(cases):
            n = int(rng.integers(g + 3, 11))
            nl = n - g
            a = int(rng.integers(2, nl + 1))
            r = int(rng.integers(1, min(a, 4) + 1))
            c = int(rng.integers(0, min(a - 1, 4) + 1))
            cfg = PlanConfig(a=a, r=r, c=c, min_qubits=0, kernel_tiles=tuple(range(1, 20)),
                             max_cost=float(rng.choice([10, 150, 300])), max_rounds=int(rng.integers(1, 7)))
            be = ShardedB200Backend(plan_config=cfg, host_state_factory=lambda nl_, off: OracleState(nl_, 0, off))
            ops = C.random_ops(rng, n, int(rng.integers(1, 120)), True)
            shots = bool(rng.integers(3) == 0)
            if shots:
                order = [int(q) for q in rng.permutation(n)[: int(rng.integers(1, n + 1))]]
                ops = ops[:60] + [C._op("m", tuple(order) if len(order) > 1 else order[0])]
            circ = C.build(Circuit, {"n": n, "ops": ops})
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                if shots:
                    random.seed(k)
                    want = O.OracleBackend().run(circ.ops, n, shots=50)
                    random.seed(k)
                    got = circ.run(backend=be, shots=50)
                    assert got == want, f"case {k}: shot counters differ (n={n}, cfg={cfg})"
                else:
                    want = O.OracleBackend().run(circ.ops, n)
                    got = circ.run(backend=be)
                    err = float(np.abs(got - want).max())
                    worst = max(worst, err)
                    assert err <= 1e-12, f"case {k}: max-abs error {err:.2e} (n={n}, cfg={cfg})"
        if rank == 0:
            print(f"world {world}: {cases} cases ok, worst max-abs error {worst:.2e}")
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        errors.put((rank, traceback.format_exc()))
        raise


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=2)
    ap.add_argument("--cases", type=int, default=40)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    errors = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, args.world, port, args.cases, args.seed, errors))
             for r in range(args.world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(1800)
    msgs = []
    while not errors.empty():
        msgs.append(errors.get())
    for p in procs:
        if p.is_alive():
            p.kill()
            msgs.append((-1, "worker timed out"))
    if msgs:
        print("\n".join(f"rank {r}:\n{m}" for r, m in msgs[:2]))
        sys.exit(1)


if __name__ == "__main__":
    main()
