"""bench.py's CPU-side legs (no GPU): the reference arm prints one well-formed JSON line, the sample keeps the gate mix."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--qubits", "14",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "gates/sec" and d["unit"] == "gates/s" and d["value"] > 0
    assert d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["host_logical_cores"] == os.cpu_count() and d["config"]["n_qubits"] == 14
    assert d["e2e"] == {"value": d["value"], "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_cpu_sample_holds_every_gate_kind_at_every_size():
    sys.path.insert(0, ROOT)
    import bench
    from qaqarot_b200 import workloads as W
    for wname, n, depth, budget in (("qft", 30, 0, 20.0), ("layers", 30, 20, 20.0), ("layers", 33, 8, 4.0),
                                    ("layers", 36, 8, 4.0), ("qaoa", 28, 0, 10.0)):
        n_cpu, circ, picked = bench.cpu_plan(wname, n, depth, budget)
        want = {o.lowername for o in bench.build_workload(wname, n, depth)[0].ops}
        assert {o.lowername for o in picked} == want, (wname, n)
        assert 16 <= n_cpu <= min(n, 30)
        cost = sum(bench.CPU_COST_26.get(o.lowername, 0.9) * 2.0 ** (n_cpu - 26) for o in picked)
        assert cost <= 1.5 * budget, (wname, n, n_cpu, cost)
    n_cpu, circ, picked = bench.cpu_plan("qft", 12, 0, 1e9)
    assert n_cpu == 12 and len(picked) == len(W.qft(12).ops)


def test_both_arms_print_the_same_config():
    sys.path.insert(0, ROOT)
    import bench
    a = bench.workload_config("layers", 30, 20, 1)
    assert a == bench.workload_config("layers", 30, 20, 1) and a["gates"] == 1490 and "30-qubit random" in a["workload"]
    assert "sharded" in bench.workload_config("layers", 33, 8, 4)["workload"]
