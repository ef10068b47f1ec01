"""ctypes binding of libb200sv.so (the C ABI declared in include/b200sv.h) and its build recipe.

There is no CPU fallback: if the library is missing or fails to load, ``load()`` raises ``B200LibraryError``
and every backend entry point propagates it.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from typing import List, Optional

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libb200sv.so")
ABI_VERSION = 104            # b200_version() of the library this planner's TILE descriptors are written for
SOURCES = ["b200sv.cu", "sample.cu", "tile.cu", "permute.cu"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


class B200LibraryError(RuntimeError):
    """libb200sv.so is not built / cannot be loaded / no CUDA device."""


class B200Error(RuntimeError):
    """A C-ABI call returned a non-zero status."""


class Op(C.Structure):
    """struct b200_op (include/b200sv.h)."""
    _fields_ = [("kind", C.c_int32), ("target", C.c_int32), ("aux", C.c_int32), ("wlen", C.c_int32),
                ("cmask", C.c_uint64), ("woff", C.c_uint64), ("roff", C.c_uint64)]


OP_MAT, OP_DIAG, OP_X, OP_SWAP, OP_PERMUTE, OP_TILE, OP_SCALE = 1, 2, 3, 4, 5, 16, 17
OP_INIT = 18         # state := |cmask>; free when the next op is a TILE pass
APPLY_ASYNC, APPLY_CONTINUE = 1, 2     # flags of b200_apply_program_ex
OP_EXCHANGE = 32      # host-level: exchange index bit `target` (top local) with global bit `aux`; never sent to the C ABI

_lib: Optional[C.CDLL] = None

_P = C.POINTER
_SIGNATURES = {
    "b200_last_error": (C.c_char_p, []),
    "b200_device_count": (C.c_int, [_P(C.c_int)]),
    "b200_version": (C.c_int, []),
    "b200_state_create": (C.c_int, [C.c_int, C.c_int, _P(C.c_void_p)]),
    "b200_state_wrap": (C.c_int, [C.c_int, C.c_int, C.c_void_p, _P(C.c_void_p)]),
    "b200_state_destroy": (C.c_int, [C.c_void_p]),
    "b200_state_n_qubits": (C.c_int, [C.c_void_p]),
    "b200_state_device_ptr": (C.c_void_p, [C.c_void_p]),
    "b200_state_set_basis": (C.c_int, [C.c_void_p, C.c_uint64]),
    "b200_state_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "b200_state_download": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "b200_state_copy": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200_apply_program": (C.c_int, [C.c_void_p, _P(Op), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "b200_apply_program_ex": (C.c_int, [C.c_void_p, _P(Op), C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]),
    "b200_last_program_ms": (C.c_int, [C.c_void_p, _P(C.c_double)]),
    "b200_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "b200_last_op_times": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, _P(C.c_int)]),
    "b200_launch_count": (C.c_int, [C.c_void_p, _P(C.c_uint64)]),
    "b200_prob_zero": (C.c_int, [C.c_void_p, C.c_int, _P(C.c_double)]),
    "b200_collapse": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int]),
    "b200_norm2": (C.c_int, [C.c_void_p, _P(C.c_double)]),
    "b200_scale": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "b200_first_above": (C.c_int, [C.c_void_p, C.c_double, _P(C.c_uint64), _P(C.c_double), _P(C.c_double)]),
    "b200_first_above_wire": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p, C.c_int, _P(C.c_uint64)]),
    "b200_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "b200_expect_group": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                    _P(C.c_double), _P(C.c_double)]),
    "b200_cond_probs": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "b200_marginal_low": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "b200_state_set_index_offset": (C.c_int, [C.c_void_p, C.c_uint64]),
    "b200_sync": (C.c_int, [C.c_void_p]),
    "b200_state_ipc_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "b200_ipc_open": (C.c_int, [C.c_int, C.c_void_p, _P(C.c_void_p)]),
    "b200_ipc_close": (C.c_int, [C.c_void_p]),
    "b200_peer_swap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint64, C.c_uint64]),
    "b200_peer_swap_bits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64,
                                      C.c_uint64, C.c_int]),
    "b200_event_record": (C.c_int, [C.c_void_p, C.c_int]),
    "b200_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_int, C.c_int, _P(C.c_double)]),
    "b200_host_alloc": (C.c_int, [C.c_size_t, _P(C.c_void_p)]),
    "b200_host_free": (C.c_int, [C.c_void_p]),
}
EXPORTS: List[str] = sorted(_SIGNATURES)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into qaqarot_b200/libb200sv.so (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "b200sv.h"))
    if not force and os.path.exists(LIB_PATH) and all(
            os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise B200LibraryError("nvcc not found: cannot build libb200sv.so")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB_PATH] + srcs
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise B200LibraryError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB_PATH


def load() -> C.CDLL:
    """Load the shared library (once) and declare every exported signature."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200LibraryError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(backend 'b200' has no CPU fallback)")
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:                                    # pragma: no cover - depends on the machine
        raise B200LibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    if lib.b200_version() != ABI_VERSION:
        raise B200LibraryError(
            f"{LIB_PATH} reports version {lib.b200_version()}, this package needs {ABI_VERSION}: the library is stale, "
            "rebuild it with `python -c 'import __graft_entry__ as g; g.build()'`")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().b200_last_error()
        raise B200Error(msg.decode() if msg else f"b200sv call failed with status {status}")


def device_count() -> int:
    n = C.c_int(0)
    lib = load()
    if lib.b200_device_count(C.byref(n)) != 0:
        return 0
    return n.value
