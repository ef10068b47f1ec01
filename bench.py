#!/usr/bin/env python
"""bench.py -- gates/sec and HBM GB/s of the B200 statevector path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload layers|qft|qaoa] [--qubits n]

A *step* is one run of the workload circuit on a device-resident state: the whole op program, starting from
|0...0> (synthesised by the first pass).  N = 1 headline workload = the one BASELINE.json's metric is quoted on:
the 30-qubit random H/RZ/CX layer circuit, depth 20 (`config.workload`); the 30-qubit QFT (configs[1]) and the
28-qubit QAOA MaxCut p=4 run(hamiltonian=...) (configs[2]) are measured in the same run and reported under
`extra`.  Under torchrun (N > 1) the workload is the same generator at 33 (N = 2, 4) or 36 qubits (N = 8), depth 8,
sharded on the top log2 N index bits, preceded by a sharded parity check against the oracle (`check`).
The state (16 GiB at 30 qubits) is far larger than the 126 MB L2, so no flush is needed between steps.

Printed JSON line (rank 0):
  value        gates/s with the state resident in HBM, CUDA events on the state's stream over the K steps (max over
               ranks), blueqat-level gates after slice expansion and before fusion;
  e2e          the same metric through ``Circuit.run(backend="b200")`` with host buffers: planning, program upload
               and the device->host copy of the statevector (into page-locked memory) inside the timed region;
  roofline     the dominant kernel (the fused tile pass) against BOTH bounds that apply to it: `hbm` -- algorithmic
               bytes per launch (32 * 2^n_local, 16 * 2^n_local for the pass that synthesises |0...0>) over the average
               launch time from per-op CUDA events of the timed steps, against the measured copy bandwidth
               (MEASURED_PEAKS.json) -- and `fp64` -- the FP64 instructions per amplitude the pass descriptors ask of
               the kernel (planner.describe_tile_pass) against the DFMA rate measured by tools/fp64_probe.cu
               (profiles/r01_fp64_probe.json).  `bound` names the one whose floor time is larger for this workload, and
               the top-level achieved / peak / frac are that bound's; `traffic` is a stored ncu figure, labelled so;
  cpu_baseline the numpy oracle (a port of blueqat's NumPyBackend) on this box's host, on a bounded sample of
               the same gate list.

``--impl reference`` times only that CPU path (rank 0; other ranks exit) and prints the same line shape with the
same `config`.
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
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
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
    if name == "qaoa":
        return W.qaoa_maxcut(n)[0], f"{n}-qubit QAOA MaxCut on C_n(1, n/2), p=4, complex128, gate fusion on"
    raise SystemExit(f"unknown workload {name}")


# ---------------------------------------------------------------------------------------------------
def workload_config(wname: str, n: int, depth: int, world: int) -> dict:
    """The `config` dict, identical in both arms (b200 and reference) for the same command line."""
    circ, wl = build_workload(wname, n, depth)
    from qaqarot_b200.lowering import lower
    gates = lower(list(circ.ops), n).n_gates
    cfg = {"workload": wl if world == 1 else wl + f", sharded on the top {world.bit_length() - 1} index bits",
           "n_qubits": n, "gates": gates, "state_bytes": int(16 * 2 ** n),
           "l2": "state (shard) larger than L2: no flush needed between steps"}
    if world > 1:
        cfg["scaling_note"] = ("BASELINE.json names a different register per GPU count (30 q at 1, 33 q at 2 and 4, 36 q "
                               "at 8), and gates/s falls with 2^n by construction: compare amp_updates_per_s "
                               "(= gates x 2^n / s) or hbm_gbs across N, not value")
    return cfg


# CPU path (numpy port of the reference), on a bounded sample of the workload's gate list
# seconds per op at 26 qubits on one core of the numpy port, by gate kind (scaled by 2^(n - 26)); used only to size
# the sample -- the reported numbers are measured
CPU_COST_26 = {"h": 0.95, "x": 0.85, "swap": 0.85, "cphase": 0.065, "rz": 0.5, "cx": 0.6, "rx": 0.95}


def host_mem_available_gib() -> float:
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 2 ** 20
    except Exception:
        pass
    return 0.0


def cpu_plan(wname: str, n: int, depth: int, budget_s: float):
    """Choose the register size the CPU sample runs at and the ops it times: at least ONE op of every gate kind of the
    workload, more (the same fraction of every kind, evenly spaced) when the budget allows.  A register whose sample
    does not fit the budget or the host memory is timed at the largest size that does, on the same generator, and
    scaled by 2^(n - n_cpu): every numpy gate pass is linear in the state size."""
    n_cpu = min(n, 30)
    while n_cpu > 16 and 2.2 * 16 * 2 ** n_cpu / 2 ** 30 > 0.8 * host_mem_available_gib():
        n_cpu -= 1
    while True:
        circ = build_workload(wname, n_cpu, depth)[0]
        by_kind = {}
        for i, o in enumerate(circ.ops):
            by_kind.setdefault(o.lowername, []).append(i)
        scale = 2.0 ** (n_cpu - 26)
        one_each = sum(CPU_COST_26.get(k, 0.9) * scale for k in by_kind)
        if one_each <= budget_s or n_cpu <= 16:
            break
        n_cpu -= 1
    total = sum(CPU_COST_26.get(k, 0.9) * scale * len(v) for k, v in by_kind.items())
    frac = min(1.0, budget_s / total)
    picked = []
    for k, idxs in by_kind.items():
        m = max(1, int(len(idxs) * frac))
        step = len(idxs) / m
        picked += [idxs[int((j + 0.5) * step)] for j in range(m)]
    picked.sort()
    return n_cpu, circ, [circ.ops[i] for i in picked]


def time_cpu_path(wname: str, n: int, depth: int, budget_s: float, steps: int = 1, warmup: int = 0):
    """Returns (gates/s at n qubits, sample description, CPU seconds of one step's sample)."""
    from oracle import sv_oracle as O
    from qaqarot_b200.lowering import lower
    n_cpu, circ_cpu, picked = cpu_plan(wname, n, depth, budget_s)
    circ = build_workload(wname, n, depth)[0]
    ctx = O.Ctx(n_cpu)
    ctx.prepare(None)
    rng = np.random.default_rng(0)
    # a dense non-trivial state so that every pass does real arithmetic (|0..0> would be mostly zeros)
    blk = 1 << 20
    for off in range(0, 1 << n_cpu, blk):
        m = min(blk, (1 << n_cpu) - off)
        ctx.qubits[off:off + m] = (rng.standard_normal(m) + 1j * rng.standard_normal(m)) * (2.0 ** (-(n_cpu + 1) / 2))
    kinds = {}
    for o in circ.ops:
        kinds[o.lowername] = kinds.get(o.lowername, 0) + 1
    all_gates = lower(list(circ.ops), n).n_gates
    est, sampled = [], []
    for it in range(warmup + steps):
        per_kind = {}
        for op in picked:
            t0 = time.perf_counter()
            O.run_gate(ctx, op)
            per_kind.setdefault(op.lowername, []).append(time.perf_counter() - t0)
        missing = sorted(set(kinds) - set(per_kind))
        if missing:
            raise SystemExit(f"cpu sample holds no op of kind(s) {missing}")          # never a cross-kind guess
        if it >= warmup:
            est.append(sum(cnt * float(np.mean(per_kind[k])) for k, cnt in kinds.items()) * 2.0 ** (n - n_cpu))
            sampled.append(sum(sum(v) for v in per_kind.values()))
    sec_full, sampled_sec = float(np.mean(est)), float(np.mean(sampled))
    names = {}
    for o in picked:
        names[o.lowername] = names.get(o.lowername, 0) + 1
    desc = (f"{len(picked)} of the {len(circ_cpu.ops)} ops of the gate list at n={n_cpu} "
            f"({', '.join(f'{v} {k}' for k, v in sorted(names.items()))}: at least one of every kind), on a dense random "
            f"state, numpy port of NumPyBackend, single thread (host has {os.cpu_count()} logical cores; the reference is "
            f"single-threaded numpy); per-kind mean op times extrapolated to the full {len(circ.ops)}-op list "
            f"({sampled_sec:.1f} s of CPU work per step sampled, {sec_full:.0f} s extrapolated)")
    if n_cpu != n:
        desc += (f"; measured on the same generator at {n_cpu} qubits and multiplied by 2^{n - n_cpu}: every numpy gate "
                 f"pass is linear in the state size")
    return all_gates / sec_full, desc, sampled_sec


# ---------------------------------------------------------------------------------------------------
def run_reference(args):
    """The CPU path on this arm's config.  Under torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    wname, n, depth = sharded_workload(args, world) if world > 1 else (args.workload, args.qubits, args.depth)
    budget = max(4.0, 150.0 / max(1, args.steps + args.warmup))
    gps, desc, sec = time_cpu_path(wname, n, depth, budget, steps=args.steps, warmup=args.warmup)
    cpu = {"value": gps, "unit": "gates/s", "cores": 1, "host_logical_cores": os.cpu_count(), "kind": "port", "sample": desc}
    line = {
        "impl": "reference", "metric": "gates/sec", "value": gps, "unit": "gates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": workload_config(wname, n, depth, world), "cpu_baseline": cpu,
        "e2e": {"value": gps, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


KIND_NAMES = {1: "k_mat", 2: "k_diag/k_phase", 3: "k_x", 4: "k_swapbits", 5: "k_involution", 16: "k_tile", 17: "k_scale"}
OP_TILE, OP_INIT = 16, 18


def fp64_peak_tflops():
    """DFMA rate measured on this pool's B200 by tools/fp64_probe.cu (2 flops per DFMA)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01_fp64_probe.json")) as f:
            return float(json.load(f)["dfma_tflops"]), "measured (tools/fp64_probe.cu, profiles/r01_fp64_probe.json)"
    except Exception:
        return 148 * 64 * 2 * 1.965e9 / 1e12, "nominal (148 SMs x 64 DFMA/clk x 1.965 GHz)"


def tile_roofline(prog, n_local: int, op_ms, total_ms: float, from_zero: bool, stored_traffic=None) -> dict:
    """Both rooflines of the fused tile pass over the timed launches.  op_ms: [(ms per op, kinds)] per step, as
    b200_last_op_times returns them (an OP_INIT entry first when the program starts from a basis state)."""
    from qaqarot_b200 import planner as P
    peak, peak_src = measured_peak_gbs()
    fpeak, fpeak_src = fp64_peak_tflops()
    by_kind = {}
    for ms, kinds in op_ms:
        for m, k in zip(ms, kinds):
            d = by_kind.setdefault(int(k), [0, 0.0])
            d[0] += 1
            d[1] += float(m)
    dom = max(by_kind, key=lambda k: by_kind[k][1])
    cnt, tot = by_kind[dom]
    avg_ms = tot / cnt
    # algorithmic bytes and FP64 instructions of the dominant kind's launches in ONE step
    nbytes, fp64_amp, first_tile = [], [], True
    for o, b in zip(prog._ops, prog.op_bytes):
        if o[0] != dom:
            continue
        if dom == OP_TILE:
            d = P.describe_tile_pass(prog._words, o[5], n_local, prog._reals)
            fp64_amp.append(d["fp64"])
            if first_tile and from_zero:
                b = b / 2             # the pass that synthesises |0...0> writes the state and reads nothing
            first_tile = False
        nbytes.append(float(b))
    alg_bytes = float(np.mean(nbytes)) if nbytes else 32.0 * 2 ** n_local
    gbs = alg_bytes / (avg_ms * 1e-3) / 1e9
    hbm = {"achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "algorithmic_bytes_per_launch": alg_bytes,
           "floor_ms_per_launch": alg_bytes / (peak * 1e9) * 1e3, "peak_source": peak_src}
    out = {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak}
    if fp64_amp:
        flops = 2.0 * float(np.mean(fp64_amp)) * 2 ** n_local          # per launch, a DFMA counted as 2
        tfl = flops / (avg_ms * 1e-3) / 1e12
        fp64 = {"dfma_per_amp_per_launch": float(np.mean(fp64_amp)), "dfma_per_amp_per_step": float(np.sum(fp64_amp)),
                "achieved": tfl, "peak": fpeak, "unit": "TFLOP/s", "frac": tfl / fpeak,
                "floor_ms_per_launch": flops / (fpeak * 1e12) * 1e3, "peak_source": fpeak_src}
        out["fp64"] = fp64
        if fp64["floor_ms_per_launch"] > hbm["floor_ms_per_launch"]:
            out.update({"bound": "fp64", "achieved": tfl, "peak": fpeak, "unit": "TFLOP/s", "frac": tfl / fpeak})
    out["hbm"] = hbm
    if isinstance(stored_traffic, dict):      # per launch: read + write passes, and the write-only pass that starts from |0...0>
        n_wo = 1 if (from_zero and dom == OP_TILE and nbytes) else 0
        stored_traffic = ((len(nbytes) - n_wo) * stored_traffic["read_write"] + n_wo * stored_traffic["write_only"]) / max(len(nbytes), 1)
    out.update({"traffic": stored_traffic,
                "traffic_source": "stored ncu figure (profiles/traffic.json), not measured in this run" if stored_traffic else None,
                "kernel": KIND_NAMES.get(dom, str(dom)), "launches_timed": cnt, "avg_launch_ms": avg_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src if out["bound"] == "hbm" else fpeak_src,
                "share_of_step": tot / total_ms})
    return out


def stored_traffic(kernel: str, n: int):
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(tpath) as f:
            return json.load(f).get(kernel, {}).get(str(n))
    except Exception:
        return None


def measure_resident(be, circ, n: int, steps: int, warmup: int, clock_index=None, detail: str = ""):
    """K steps of circ's op program on the device-resident state, each from |0...0>.  CUDA events on the state's stream."""
    comp = be._compile(circ.ops, n, 0, from_zero=True)
    prog = comp.prefix
    st = be._state(n)
    st.set_profiling(True)
    for _ in range(warmup):
        st.apply(prog, init_basis=0)             # prepare |0..0> + the whole program (B200_OP_INIT folded in)
    st.sync()
    launches0 = st.launch_count()
    sampler = ClockSampler(clock_index) if clock_index is not None else None
    if sampler:
        time.sleep(0.3)
    t0 = time.perf_counter()
    op_ms = []
    st.event_record(0)
    for _ in range(steps):
        st.apply(prog, init_basis=0)
        ms, kinds = st.last_op_times()
        op_ms.append((ms.copy(), kinds.copy()))
    st.event_record(1)
    st.sync()
    t1 = time.perf_counter()
    total_ms = st.event_elapsed_ms(0, 1)
    clocks = sampler.stop(t0, t1) if sampler else None
    launches = st.launch_count() - launches0
    st.set_profiling(False)
    if detail:
        with open(detail, "w") as f:
            json.dump({"kinds": [int(k) for k in op_ms[-1][1]], "ms": [[float(x) for x in ms] for ms, _ in op_ms],
                       "bytes": prog.op_bytes}, f)
    roof = tile_roofline(prog, n, op_ms, total_ms, True, stored_traffic("k_tile", n))
    return {"comp": comp, "prog": prog, "state": st, "sec_per_step": total_ms / 1e3 / steps, "launches": int(launches),
            "clocks": clocks, "roofline": roof}


def run_b200(args):
    from qaqarot_b200 import B200Backend
    from qaqarot_b200.device import PinnedBuffer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        return run_sharded(args, world)

    n = args.qubits
    circ, wl = build_workload(args.workload, n, args.depth)
    cfg = None
    if args.tile_bits or args.no_shears or args.row_bits != 4 or args.no_bank_pairs:
        from qaqarot_b200.planner import PlanConfig
        cfg = PlanConfig(a=args.tile_bits or 12, r=args.reg_bits, c=args.row_bits, shears=not args.no_shears,
                         bank_pairs=not args.no_bank_pairs, **({"max_cost": args.max_cost} if args.max_cost else {}))
    be = B200Backend(device=0, fusion=not args.no_fusion, plan_config=cfg)
    res = measure_resident(be, circ, n, args.steps, args.warmup, clock_index=0, detail=args.detail)
    comp, prog, st, sec_per_step = res["comp"], res["prog"], res["state"], res["sec_per_step"]
    gates = comp.n_gates
    value = gates / sec_per_step
    peak, _ = measured_peak_gbs()
    moved = prog.pass_bytes - 16.0 * 2 ** n          # the first pass writes only
    hbm_gbs = moved / sec_per_step / 1e9

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
           "note": "Circuit.run(backend='b200') -> statevector ndarray in page-locked host memory; planning and program "
                   "upload inside the timed region"}
    check = {"norm2": st.norm2()}
    be.result_buffer = None
    pinned.close()

    extra = None
    if args.extras:
        extra = run_extras(be, args)
    be.close()

    cpu = None
    if not args.no_cpu:
        try:
            gps, desc, _ = time_cpu_path(args.workload, n, args.depth, args.cpu_budget)
            cpu = {"value": gps, "unit": "gates/s", "cores": 1, "host_logical_cores": os.cpu_count(), "kind": "port",
                   "sample": desc}
        except SystemExit as e:
            cpu = {"value": None, "unit": "gates/s", "cores": 1, "kind": "port", "sample": str(e)}

    line = {
        "metric": "gates/sec", "value": value, "unit": "gates/s", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": workload_config(args.workload, n, args.depth, 1),
        "plan": {"passes": prog.n_passes, "fusion": not args.no_fusion},
        "amp_updates_per_s": gates * 2.0 ** n / sec_per_step,
        "hbm_gbs": hbm_gbs, "hbm_frac_of_measured": hbm_gbs / peak,
        "roofline": res["roofline"], "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": res["launches"],
        "clocks": res["clocks"], "check": check,
    }
    if extra is not None:
        line["extra"] = extra
    print(json.dumps(line), flush=True)


def run_extras(be, args) -> dict:
    """BASELINE configs[1] and [2] next to the headline: the 30-qubit QFT (device resident) and the 28-qubit QAOA MaxCut
    p=4 energy through run(hamiltonian=...)."""
    from qaqarot_b200 import workloads as W
    out = {}
    steps, warmup = max(1, min(args.steps, 5)), 3
    n = 30
    circ, wl = build_workload("qft", n, 0)
    r = measure_resident(be, circ, n, steps, warmup)
    roof = r["roofline"]
    out["qft30"] = {"workload": wl, "ms_per_step": r["sec_per_step"] * 1e3, "value": r["comp"].n_gates / r["sec_per_step"],
                    "unit": "gates/s", "passes": r["prog"].n_passes, "gpu_launches": r["launches"],
                    "roofline": {k: roof[k] for k in ("bound", "achieved", "peak", "unit", "frac", "avg_launch_ms", "hbm", "fp64")
                                 if k in roof},
                    "norm2": r["state"].norm2()}
    n = 28
    circ, ham, edges = W.qaoa_maxcut(n)
    r = measure_resident(be, circ, n, steps, warmup)
    roof = r["roofline"]
    # the objective loop (vqe.py:67-83): a NEW angle vector every call, nothing but the energy returns.  First with the
    # speculative launch from the planner model off (every call plans, then launches pass by pass), then with it on (three
    # calls until the model exists, untimed like any warm-up).
    loops = {}
    for replay in (False, True):
        be.param_replay = replay
        be._models.clear()
        be._last_key = None
        for seed in (900, 901, 902):
            c2, h2, _ = W.qaoa_maxcut(n, seed=seed)
            c2.run(backend=be, hamiltonian=h2)
        t0 = time.perf_counter()
        for k in range(max(steps, 3)):
            c2, h2, _ = W.qaoa_maxcut(n, seed=910 + k)
            energy = c2.run(backend=be, hamiltonian=h2)
        loops[replay] = (time.perf_counter() - t0) / max(steps, 3) * 1e3
    e2e_ms = loops[True]
    replay_stats = dict(be.replay_stats)
    energy = circ.run(backend=be, hamiltonian=ham)
    out["qaoa28_hamiltonian"] = {
        "workload": f"{n}-qubit QAOA MaxCut on C_n(1, n/2), p=4: run(hamiltonian=sum of {len(edges)} Z.Z terms)",
        "ms_per_step": r["sec_per_step"] * 1e3, "value": r["comp"].n_gates / r["sec_per_step"], "unit": "gates/s",
        "passes": r["prog"].n_passes, "gpu_launches": r["launches"],
        "roofline": {k: roof[k] for k in ("bound", "achieved", "peak", "unit", "frac", "avg_launch_ms", "hbm", "fp64") if k in roof},
        "e2e_ms_per_energy": e2e_ms, "e2e_ms_per_energy_without_replay": loops[False], "replay": replay_stats,
        "energy": energy,
        "note": "ms_per_step: the circuit on the device-resident state (plan with two-shear rotations); "
                "e2e_ms_per_energy: Circuit.run(backend='b200', hamiltonian=h) -> float with a NEW angle vector every "
                "call (ansatz construction included): launched at once from the planner model (exact three-shear "
                "rotations), planned for real while the device runs, compared; without_replay: planned first, launched "
                "pass by pass"}
    return out


def sharded_workload(args, world: int):
    """BASELINE configs[3] / [4]: 33-qubit random layers on 2 and 4 GPUs, 36 qubits on 8 (depth 8)."""
    if args.workload != "auto":
        return args.workload, args.qubits, args.depth
    return "layers", (36 if world >= 8 else 33), 8


def sharded_parity_check(be, rank: int, n: int = 20) -> dict:
    """Before anything is timed: a 20-qubit circuit (random layers + QFT) sharded over ALL ranks of this run --
    statevector, 200 seeded shots and an X/Y/Z Hamiltonian -- against the numpy oracle on rank 0.  A wrong
    permutation keeps the norm; it does not keep the amplitudes."""
    import random
    import warnings
    import torch
    import torch.distributed as dist
    from qaqarot_b200 import workloads as W, X, Y, Z
    circ = W.random_layers(n, 6, seed=99)
    for op in W.qft(n, 0).ops:
        circ.ops.append(op)
    got = circ.run(backend=be)
    exchanges = be._compiled.prefix.n_exchanges
    ham = Z[0] * Z[n - 1] + 0.5 * Z[n - 2] * Z[n - 1] - 1.5 * X[n - 1] * Y[2] + 0.25 * X[n - 3] + 2.0
    terms = [(1.0, [("Z", 0), ("Z", n - 1)]), (0.5, [("Z", n - 2), ("Z", n - 1)]), (-1.5, [("X", n - 1), ("Y", 2)]),
             (0.25, [("X", n - 3)]), (2.0, [])]
    energy = circ.run(backend=be, hamiltonian=ham)
    circ_m = circ.copy(copy_backends=False).m[:]
    random.seed(21)
    counts = circ_m.run(backend=be, shots=200)
    res = {"n_qubits": n, "exchanges": int(exchanges), "max_abs_err": None, "counters_equal": None, "energy_err": None}
    ok = torch.ones(1, device=f"cuda:{torch.cuda.current_device()}")
    if rank == 0:
        from oracle import sv_oracle as O
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = O.OracleBackend().run(circ.ops, n)
            random.seed(21)
            want = O.OracleBackend().run(circ_m.ops, n, shots=200)
        res["max_abs_err"] = float(np.abs(got - ref).max())
        res["counters_equal"] = bool(counts == want)
        res["energy_err"] = float(abs(energy - O.expectation(terms, ref)))
        res["tolerance"] = {"max_abs_err": 1e-12, "energy_err": 1e-10}
        if not (res["max_abs_err"] <= 1e-12 and res["counters_equal"] and res["energy_err"] <= 1e-10 and exchanges > 0):
            ok.zero_()
    dist.broadcast(ok, 0)
    res["ok"] = bool(ok.item() == 1)
    return res


def run_sharded(args, world: int):
    import random
    import torch
    import torch.distributed as dist
    from qaqarot_b200 import ShardedB200Backend
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # stdout carries exactly one JSON line: everything else written to file descriptor 1 from here on (NCCL prints its
    # version banner with printf) lands on stderr, where NCCL's own log (INFO: communicator, ranks, transports) goes too
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    os.environ["NCCL_DEBUG"] = os.environ.get("B200_NCCL_DEBUG", "INFO")
    os.environ.pop("NCCL_DEBUG_FILE", None)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    wname, n, depth = sharded_workload(args, world)
    circ, wl = build_workload(wname, n, depth)
    cfg = None
    if args.tile_bits:
        from qaqarot_b200.planner import PlanConfig
        cfg = PlanConfig(a=args.tile_bits)
    be = ShardedB200Backend(device=local, plan_config=cfg)
    parity = sharded_parity_check(be, rank)
    if not parity["ok"]:
        if rank == 0:
            os.write(json_fd, (json.dumps({"error": "sharded parity check failed", "check": parity}) + "\n").encode())
        dist.barrier()
        dist.destroy_process_group()
        raise SystemExit(1)
    comp = be._compile(circ.ops, n, 0, from_zero=True)
    prog = comp.prefix
    st = be._state(n)
    st.use_p2p = not args.nccl_exchange
    n_local = st.n_local
    st.set_profiling(True)

    def step():
        st.apply(prog, init_basis=0)

    def fence():
        st.sync()
        torch.cuda.synchronize(local)
        dist.barrier()
        torch.cuda.synchronize(local)

    for _ in range(args.warmup):
        step()
    fence()
    launches0, ex_ms0, ex_bytes0 = st.launch_count(), st.exchange_ms, st.exchange_bytes
    st.profile = []
    sampler = ClockSampler(local) if rank == 0 else None
    time.sleep(0.3)
    t0 = time.perf_counter()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        step()
    fence()
    stop.record()
    stop.synchronize()
    t1 = time.perf_counter()
    ms = torch.tensor([start.elapsed_time(stop)], device=f"cuda:{local}", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    clocks = sampler.stop(t0, t1) if sampler else None
    launches = st.launch_count() - launches0
    profile = st.profile
    st.set_profiling(False)
    sec_per_step = total_ms / 1e3 / args.steps
    gates = comp.n_gates
    ex = torch.tensor([st.exchange_ms - ex_ms0], device=f"cuda:{local}", dtype=torch.float64)
    dist.all_reduce(ex, op=dist.ReduceOp.MAX)
    ex_ms_step = float(ex.item()) / args.steps
    ex_bytes_step = (st.exchange_bytes - ex_bytes0) / args.steps
    norm2 = st.norm2()

    # dominant kernel on this rank
    roofline = tile_roofline(prog, n_local, [(mss, kinds) for kinds, mss in profile], total_ms, True)
    roofline["per_gpu"] = True
    exchange = {"count_per_step": prog.n_exchanges, "ms_per_step": ex_ms_step,
                "bytes_sent_per_gpu_per_step": ex_bytes_step,
                "nvlink_gbs_per_direction": (ex_bytes_step / (ex_ms_step * 1e-3) / 1e9) if ex_ms_step > 0 else None,
                "share_of_step": ex_ms_step / (sec_per_step * 1e3),
                "note": (st.exchange_note if hasattr(st, "exchange_note") else
                         "exchanged in place by k_peer_swap over CUDA-IPC peer memory (NVLink), fences included"
                         if st.use_p2p else
                         "over NCCL send/recv through a 256 MiB bounce buffer, copy-back included")}

    # end to end through the plugin API: plan, upload, run, sample 1000 shots, Counter back on the host
    shots = 1000
    circ_m = circ.copy(copy_backends=False).m[:]
    random.seed(7)
    be._compiled = None
    fence()
    t0 = time.perf_counter()
    counts = circ_m.run(backend=be, shots=shots)
    fence()
    e2e_sec = time.perf_counter() - t0
    ops_arr, n_ops, words, reals = prog.frozen()
    e2e = {"value": gates / e2e_sec, "unit": "gates/s", "h2d_bytes_per_step": int(n_ops * 40 + words.nbytes + reals.nbytes
                                                                                    + shots * n * 8),
           "d2h_bytes_per_step": int(n * shots * 16), "ms_per_step": e2e_sec * 1e3, "steps": 1,
           "note": f"Circuit.m[:].run(backend=ShardedB200Backend, shots={shots}) -> Counter; planning, program upload, "
                   f"sharded sampling descent included; {len(counts)} distinct outcomes"}
    be.close()
    if rank == 0:
        moved = prog.pass_bytes - 16.0 * 2 ** n_local * prog.n_exchanges - 16.0 * 2 ** n_local
        line = {
            "metric": "gates/sec", "value": gates / sec_per_step, "unit": "gates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
            "config": workload_config(wname, n, depth, world),
            "plan": {"n_local": n_local, "passes": prog.n_passes - prog.n_exchanges, "exchanges": prog.n_exchanges,
                     "state_bytes_per_gpu": int(16 * 2 ** n_local)},
            "amp_updates_per_s": gates * 2.0 ** n / sec_per_step,
            "hbm_gbs": world * moved / sec_per_step / 1e9, "hbm_gbs_per_gpu": moved / sec_per_step / 1e9,
            "roofline": roofline, "exchange": exchange, "cpu_baseline": None, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks, "check": dict(parity, norm2=norm2),
        }
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "qft", "layers", "qaoa"])
    ap.add_argument("--qubits", type=int, default=30)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--no-fusion", action="store_true")
    ap.add_argument("--tile-bits", type=int, default=0, help="override the planner's tile size (log2 amplitudes)")
    ap.add_argument("--reg-bits", type=int, default=4, help="with --tile-bits: register bits per round (3 only with 11)")
    ap.add_argument("--max-cost", type=float, default=0.0, help="with --tile-bits: planner's DFMA budget per amplitude per pass")
    ap.add_argument("--row-bits", type=int, default=4, help="planner experiment: low index bits every tile must contain (2^c * 16 B rows)")
    ap.add_argument("--no-bank-pairs", action="store_true", help="planner experiment: PlanConfig.bank_pairs = False")
    ap.add_argument("--no-shears", action="store_true", help="planner experiment: keep 1-qubit gates as general 2x2 matrices")
    ap.add_argument("--detail", default="", help="write per-op device times of the timed steps to this JSON file")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="N = 1: skip the QFT-30 / QAOA-28 lines under `extra`")
    ap.add_argument("--nccl-exchange", action="store_true", help="sharded runs: exchange through NCCL send/recv instead of the peer-to-peer kernel")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for cpu_baseline")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    args.extras = world == 1 and args.workload == "auto" and not args.no_extras and args.impl == "b200"
    if world == 1 and args.workload == "auto":
        args.workload = "layers"           # BASELINE.json metric: "30q random circuit 1 GPU"
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
