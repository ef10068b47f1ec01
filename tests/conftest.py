"""pytest configuration: `gpu` marker, import paths, shared golden-fixture loaders."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, GOLDEN):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE = "/root/reference"          # present in the build container only; never on the GPU box


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def _cuda_devices() -> int:
    try:
        from qaqarot_b200 import _lib
        return _lib.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    if any("gpu" in item.keywords for item in items) and _cuda_devices() == 0:
        # (the library itself never falls back: test_backend_requires_the_cuda_library checks that on the GPU box)
        no_gpu = pytest.mark.skip(reason="no CUDA device (or libb200sv.so not built)")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(no_gpu)
    if os.path.isdir(os.path.join(REFERENCE, "blueqat")):
        return
    skip = pytest.mark.skip(reason="/root/reference not present")
    for item in items:
        if "reference" in item.keywords:
            item.add_marker(skip)


class Golden:
    def __init__(self):
        with open(os.path.join(GOLDEN, "cases.json")) as f:
            self.cases = {c["name"]: c for c in json.load(f)}
        with open(os.path.join(GOLDEN, "golden.json")) as f:
            self.meta = json.load(f)
        self.arrays = np.load(os.path.join(GOLDEN, "golden.npz"))

    def sv(self, name):
        return self.arrays["sv/" + name]


_GOLDEN = None


def golden() -> Golden:
    global _GOLDEN
    if _GOLDEN is None:
        _GOLDEN = Golden()
    return _GOLDEN


def case_names(kind=None):
    g = golden()
    out = []
    for name, c in g.cases.items():
        returns = c["run"]["returns"] or ("statevector" if c["run"]["shots"] is None else "shots")
        if kind is None or returns == kind:
            out.append(name)
    return out


@pytest.fixture(scope="session")
def gold():
    return golden()
