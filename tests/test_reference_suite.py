"""The reference's OWN test-suite (/root/reference/tests, 1 500 cases) with "b200" registered through
BlueqatGlobalSetting.register_backend and made the default backend: the drop-in claim, checked by blueqat's tests.
Build container only (the reference does not travel to the GPU box)."""
import os
import re
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.reference
def test_reference_suite_passes_on_backend_b200():
    plugin_dir = os.path.join(ROOT, "tests", "ref_plugin")
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join(["/root/reference", plugin_dir, ROOT, env.get("PYTHONPATH", "")])
    with tempfile.TemporaryDirectory() as cwd:
        res = subprocess.run(
            [sys.executable, "-m", "pytest", "-p", "no:cacheprovider", "-p", "b200_ref_plugin", "/root/reference/tests",
             "--ignore=/root/reference/tests/test_ibmq_backend.py", "--add-backend", "b200", "-q", "-x"],
            cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
    tail = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else res.stderr[-500:]
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    m = re.search(r"(\d+) passed", tail)
    assert m and int(m.group(1)) >= 1400, tail
