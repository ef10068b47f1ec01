"""The C-ABI library loads and exports every symbol include/b200sv.h declares (no compute calls, no GPU)."""
import ctypes
import os
import re

from qaqarot_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200sv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    _lib.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in b200sv.h but not exported"


def test_python_binding_covers_the_header():
    assert header_symbols() == _lib.EXPORTS


def test_op_struct_layout_matches_header():
    # struct b200_op: 4 x int32, 3 x uint64 = 40 bytes, no padding surprises
    assert ctypes.sizeof(_lib.Op) == 40
    assert _lib.Op.cmask.offset == 16 and _lib.Op.woff.offset == 24 and _lib.Op.roff.offset == 32


def test_errors_are_reported_without_a_gpu():
    lib = _lib.load()
    assert lib.b200_version() == _lib.ABI_VERSION
    assert lib.b200_sync(None) != 0
    assert b"null state" in lib.b200_last_error()
