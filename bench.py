#!/usr/bin/env python
"""bench.py -- gates/sec and HBM GB/s of the B200 statevector path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--qubits n] [--workload qft|layers]

A *step* is one run of the workload circuit on a device-resident state: reset to |0..0>, then the whole op
program (BASELINE configs[1]: the 30-qubit QFT of basis state 123456789 with gate fusion, at N = 1).  The state
(16 GiB at 30 qubits) is far larger than the 126 MB L2, so no flush is needed between steps.

Printed JSON line (rank 0):
  value        gates/s with the state resident in HBM, timed with CUDA events on the state's stream over the
               K steps (max over ranks), blueqat-level gates after slice expansion and before fusion;
  e2e          the same metric through ``Circuit.run(backend="b200")`` with host buffers: planning, program
               upload and the device->host copy of the statevector (into page-locked memory) inside the timed
               region;
  roofline     the dominant kernel (the fused tile pass) against measured HBM copy bandwidth
               (MEASURED_PEAKS.json): algorithmic bytes per launch = 32 * 2^n_local (one read + one write of
               every amplitude), duration from per-op CUDA events recorded during the timed steps;
  cpu_baseline the numpy oracle (a port of blueqat's NumPyBackend) on this box's host, on a bounded sample of
               the same gate list at the same qubit count.

``--impl reference`` times only that CPU path (rank 0; other ranks exit) and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np


# ---------------------------------------------------------------------------------------------------
def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for t, line in self.rows:
            if not (t0 <= t <= t1 + 0.2):
                continue
            parts = [p.strip() for p in line.split(",")]
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                 parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(name: str, n: int, depth: int):
    from qaqarot_b200 import workloads as W
    if name == "qft":
        return W.qft(n), f"{n}-qubit QFT of basis state 123456789, complex128, gate fusion on"
    if name == "layers":
        return W.random_layers(n, depth), f"{n}-qubit random H/RZ/CX layers depth {depth}, complex128, gate fusion on"
    raise SystemExit(f"unknown workload {name}")


# ---------------------------------------------------------------------------------------------------
# CPU path (numpy port of the reference), on a bounded sample of the workload's gate list
def cpu_sample(circ, n: int, budget_s: float):
    """Every `stride`-th op of the gate list, stride chosen so the predicted cost fits the budget.
    Per-op cost model (seconds at 26 qubits on one core, scaled by 2^(n-26)): full-state 2x2 passes ~0.9,
    controlled-diagonal quarter passes ~0.065."""
    scale = 2.0 ** (n - 26)
    cost = {"h": 0.95, "x": 0.85, "swap": 0.85 * 3, "cphase": 0.065, "rz": 0.5, "cx": 0.6}
    ops = list(circ.ops)
    per_op = [cost.get(o.lowername, 0.9) * scale for o in ops]
    stride = 1
    while stride < len(ops) and sum(per_op[stride // 2::stride]) > budget_s:
        stride += 1
    picked = ops[stride // 2::stride]
    names = {}
    for o in picked:
        names[o.lowername] = names.get(o.lowername, 0) + 1
    desc = (f"every {stride}th op of the {len(ops)}-op gate list at n={n} "
            f"({', '.join(f'{v} {k}' for k, v in sorted(names.items()))}) on a numpy port of NumPyBackend")
    return picked, desc


def host_mem_available_gib() -> float:
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 2 ** 20
    except Exception:
        pass
    return 0.0


def time_cpu_path(circ, n: int, budget_s: float, steps: int = 1, warmup: int = 0):
    """Returns (gates/s, sample description, n actually used)."""
    from oracle import sv_oracle as O
    from qaqarot_b200.lowering import lower
    n_cpu = n
    need_gib = 2.2 * 16 * 2 ** n / 2 ** 30
    while n_cpu > 20 and 2.2 * 16 * 2 ** n_cpu / 2 ** 30 > 0.8 * host_mem_available_gib():
        n_cpu -= 1
    if n_cpu != n:
        raise SystemExit(f"host RAM too small for a {n}-qubit CPU sample (needs {need_gib:.0f} GiB)")
    picked, desc = cpu_sample(circ, n, budget_s)
    gates = lower(picked, n).n_gates
    ctx = O.Ctx(n)
    ctx.prepare(None)
    rng = np.random.default_rng(0)
    # a dense non-trivial state so that every pass does real arithmetic (|0..0> would be mostly zeros)
    blk = 1 << 20
    for off in range(0, 1 << n, blk):
        m = min(blk, (1 << n) - off)
        ctx.qubits[off:off + m] = (rng.standard_normal(m) + 1j * rng.standard_normal(m)) * (2.0 ** (-(n + 1) / 2))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        for op in picked:
            O.run_gate(ctx, op)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = float(np.mean(times))
    return gates / sec, desc, sec


# ---------------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    circ, wl = build_workload(args.workload, args.qubits, args.depth)
    budget = max(4.0, 170.0 / max(1, args.steps + args.warmup))
    gps, desc, sec = time_cpu_path(circ, args.qubits, budget, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "gates/sec", "value": gps, "unit": "gates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": {"workload": wl, "n_qubits": args.qubits},
        "cpu_baseline": {"value": gps, "unit": "gates/s", "cores": 1, "kind": "port", "sample": desc},
        "e2e": {"value": gps, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    from qaqarot_b200 import B200Backend
    from qaqarot_b200 import _lib
    from qaqarot_b200.device import PinnedBuffer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        from qaqarot_b200 import bench_sharded
        return bench_sharded.run(args)

    n = args.qubits
    circ, wl = build_workload(args.workload, n, args.depth)
    cfg = None
    if args.tile_bits:
        from qaqarot_b200.planner import PlanConfig
        cfg = PlanConfig(a=args.tile_bits)
    be = B200Backend(device=0, fusion=not args.no_fusion, plan_config=cfg)
    comp = be._compile(circ.ops, n, 0)
    prog = comp.prefix
    st = be._state(n)
    st.set_profiling(True)

    def step():
        st.set_basis(0)
        st.apply(prog)

    for _ in range(args.warmup):
        step()
    st.sync()
    launches0 = st.launch_count()
    sampler = ClockSampler(0)
    time.sleep(0.3)
    t0 = time.perf_counter()
    op_ms = []
    st.event_record(0)
    for _ in range(args.steps):
        step()
        ms, kinds = st.last_op_times()
        op_ms.append((ms.copy(), kinds.copy()))
    st.event_record(1)
    st.sync()
    t1 = time.perf_counter()
    total_ms = st.event_elapsed_ms(0, 1)
    clocks = sampler.stop(t0, t1)
    launches = st.launch_count() - launches0
    st.set_profiling(False)
    sec_per_step = total_ms / 1e3 / args.steps
    gates = comp.n_gates
    value = gates / sec_per_step

    # roofline of the dominant kernel
    peak, peak_src = measured_peak_gbs()
    pass_bytes = 32.0 * 2 ** n
    by_kind = {}
    for ms, kinds in op_ms:
        for m, k in zip(ms, kinds):
            d = by_kind.setdefault(int(k), [0, 0.0])
            d[0] += 1
            d[1] += float(m)
    dom = max(by_kind, key=lambda k: by_kind[k][1])
    cnt, tot = by_kind[dom]
    kind_names = {1: "k_mat", 2: "k_diag/k_phase", 3: "k_x", 4: "k_swapbits", 16: "k_tile", 17: "k_scale"}
    avg_ms = tot / cnt
    dom_bytes = [b for b, o in zip(prog.op_bytes, prog._ops) if o[0] == dom]
    alg_bytes = float(np.mean(dom_bytes)) if dom_bytes else pass_bytes
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as f:
                traffic = json.load(f).get(kind_names.get(dom, ""), {}).get(str(n))
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kind_names.get(dom, str(dom)), "launches_timed": cnt,
                "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "share_of_step": tot / total_ms}
    hbm_gbs = prog.pass_bytes / sec_per_step / 1e9
    if args.detail:
        with open(args.detail, "w") as f:
            json.dump({"kinds": [int(k) for k in op_ms[-1][1]], "ms": [[float(x) for x in ms] for ms, _ in op_ms],
                       "bytes": prog.op_bytes}, f)

    # end to end through the plugin API with host buffers
    pinned = PinnedBuffer(1 << n)
    be.result_buffer = pinned.array
    e2e_steps = max(1, min(args.steps, 3))
    circ.run(backend=be)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        be._compiled = None                      # plan again: planning is part of the user's call
        out = circ.run(backend=be)
    e2e_sec = (time.perf_counter() - t0) / e2e_steps
    assert out.shape == (1 << n,)
    ops_arr, n_ops, words, reals = prog.frozen()
    h2d = int(n_ops * 40 + words.nbytes + reals.nbytes)
    e2e = {"value": gates / e2e_sec, "unit": "gates/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": int(16 * 2 ** n), "ms_per_step": e2e_sec * 1e3, "steps": e2e_steps,
           "note": "Circuit.run(backend='b200') -> statevector ndarray in page-locked host memory"}
    check = {"norm2": st.norm2()}
    be.close()
    pinned.close()

    cpu = None
    if not args.no_cpu:
        try:
            gps, desc, _ = time_cpu_path(circ, n, args.cpu_budget)
            cpu = {"value": gps, "unit": "gates/s", "cores": 1, "kind": "port", "sample": desc}
        except SystemExit as e:
            cpu = {"value": None, "unit": "gates/s", "cores": 1, "kind": "port", "sample": str(e)}

    line = {
        "metric": "gates/sec", "value": value, "unit": "gates/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": {"workload": wl, "n_qubits": n, "gates": gates, "passes": prog.n_passes,
                   "state_bytes": int(16 * 2 ** n), "l2": "state larger than L2 (no flush needed)"},
        "hbm_gbs": hbm_gbs, "hbm_frac_of_measured": hbm_gbs / peak,
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        "check": check,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="qft", choices=["qft", "layers"])
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--no-fusion", action="store_true")
    ap.add_argument("--tile-bits", type=int, default=0, help="override the planner's tile size (log2 amplitudes)")
    ap.add_argument("--detail", default="", help="write per-op device times of the timed steps to this JSON file")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for cpu_baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
