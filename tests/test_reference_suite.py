"""The reference's OWN test-suite (/root/reference/tests, 1 500 cases) with "b200" registered through
BlueqatGlobalSetting.register_backend and made the default backend: the drop-in claim, checked by blueqat's tests.

* build container (`-m "not gpu"`): blueqat from /root/reference, the device is the numpy interpreter of the op program
  -- this checks the host logic (plugin, lowering, planner), not the kernels;
* GPU box (`-m gpu`): blueqat and its tests from baseline/_ref (the unmodified reference, installed there by
  `__graft_entry__.build()`; git-ignored), the device is libb200sv.so on cuda:0."""
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def _run_suite(ref_root: str, tests: str, device: str):
    plugin_dir = os.path.join(ROOT, "tests", "ref_plugin")
    env = dict(os.environ, B200_REF_ROOT=ref_root, B200_REF_SUITE_DEVICE=device)
    env["PYTHONPATH"] = os.pathsep.join([ref_root, plugin_dir, ROOT, env.get("PYTHONPATH", "")])
    with tempfile.TemporaryDirectory() as cwd:
        res = subprocess.run(
            [sys.executable, "-m", "pytest", "-p", "no:cacheprovider", "-p", "b200_ref_plugin", tests,
             f"--ignore={os.path.join(tests, 'test_ibmq_backend.py')}", "--add-backend", "b200", "-q", "-x"],
            cwd=cwd, env=env, capture_output=True, text=True, timeout=1500)
    tail = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else res.stderr[-500:]
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    m = re.search(r"(\d+) passed", tail)
    assert m and int(m.group(1)) >= 1400, tail
    return tail


@pytest.mark.reference
def test_reference_suite_passes_on_backend_b200():
    _run_suite("/root/reference", "/root/reference/tests", "interpreter")


@pytest.mark.gpu
def test_reference_suite_passes_on_the_cuda_device():
    if not (os.path.isdir(os.path.join(REF, "blueqat")) and os.path.isdir(os.path.join(REF, "_tests"))):
        pytest.skip("baseline/_ref is not installed (run __graft_entry__.build() in the build container)")
    tail = _run_suite(REF, os.path.join(REF, "_tests"), "cuda")
    print("reference suite on cuda:0:", tail)
