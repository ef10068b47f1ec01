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
    assert d["e2e"] == {"value": d["value"], "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_cpu_sample_is_stratified():
    sys.path.insert(0, ROOT)
    import bench
    from qaqarot_b200 import workloads as W
    picked, desc = bench.cpu_sample(W.qft(30), 30, 40.0)
    kinds = {o.lowername for o in picked}
    assert "h" in kinds and "cphase" in kinds and len(picked) < 40
    picked, _ = bench.cpu_sample(W.qft(12), 12, 1e9)
    assert len(picked) == len(W.qft(12).ops)
