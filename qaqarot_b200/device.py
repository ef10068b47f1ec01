"""DeviceState: a 2^n complex128 statevector resident in B200 HBM, driven through the C ABI."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .program import Program


class DeviceState:
    """Owns (or wraps) one device buffer.  All methods are synchronous; errors raise ``B200Error``."""

    def __init__(self, n_qubits: int, device: int = 0, wrap_ptr: Optional[int] = None):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.n_qubits = n_qubits
        self.device = device
        if wrap_ptr is None:
            _lib.check(self._lib.b200_state_create(device, n_qubits, C.byref(self._h)))
        else:
            _lib.check(self._lib.b200_state_wrap(device, n_qubits, C.c_void_p(wrap_ptr), C.byref(self._h)))

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.b200_state_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def dim(self) -> int:
        return 1 << self.n_qubits

    @property
    def device_ptr(self) -> int:
        return self._lib.b200_state_device_ptr(self._h) or 0

    # -- contents ---------------------------------------------------------------------------------
    def set_basis(self, index: int = 0) -> None:
        _lib.check(self._lib.b200_state_set_basis(self._h, index))

    def upload(self, vec: np.ndarray, offset: int = 0) -> None:
        vec = np.ascontiguousarray(vec, dtype=np.complex128)
        _lib.check(self._lib.b200_state_upload(self._h, vec.ctypes.data_as(C.c_void_p), offset, vec.size))

    def download(self, offset: int = 0, count: Optional[int] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
        count = self.dim - offset if count is None else count
        if out is None:
            out = np.empty(count, dtype=np.complex128)
        assert out.dtype == np.complex128 and out.size == count and out.flags.c_contiguous
        _lib.check(self._lib.b200_state_download(self._h, out.ctypes.data_as(C.c_void_p), offset, count))
        return out

    def copy_from(self, other: "DeviceState") -> None:
        _lib.check(self._lib.b200_state_copy(self._h, other._h))

    def clone(self) -> "DeviceState":
        c = DeviceState(self.n_qubits, self.device)
        c.copy_from(self)
        return c

    # -- evolution --------------------------------------------------------------------------------
    def apply(self, program: Program, init_basis: Optional[int] = None) -> None:
        """`init_basis`: start from |init_basis> (B200_OP_INIT): replaces a set_basis() sweep before the program."""
        if len(program) == 0:
            if init_basis is not None:
                self.set_basis(init_basis & (self.dim - 1))
            return
        ops, n_ops, words, reals = program.frozen() if init_basis is None else program.frozen_with_init(init_basis)
        _lib.check(self._lib.b200_apply_program(
            self._h, ops, n_ops, words.ctypes.data_as(C.c_void_p), words.size,
            reals.ctypes.data_as(C.c_void_p), reals.size))

    def apply_piece(self, program: Program, op_from: int, op_to: int, words_from: int, reals_from: int,
                    init_basis: Optional[int] = None, first: bool = True) -> None:
        """Launch ops [op_from, op_to) of a program that is still being planned, without waiting for the device
        (b200_apply_program_ex).  Only the pool tails the piece can refer to -- words[words_from:], reals[reals_from:] --
        are converted; the library reads them at launch time (a pass travels as kernel parameters)."""
        ops_py = program._ops[op_from:op_to]
        n = len(ops_py) + (1 if init_basis is not None else 0)
        if n == 0:
            return
        ops = (_lib.Op * n)()
        k = 0
        if init_basis is not None:
            ops[0].kind, ops[0].cmask = _lib.OP_INIT, int(init_basis)
            k = 1
        for kind, target, aux, wlen, cmask, woff, roff in ops_py:
            o = ops[k]
            o.kind, o.target, o.aux, o.wlen, o.cmask, o.woff, o.roff = kind, target, aux, wlen, cmask, woff, roff
            k += 1
        words = np.asarray(program._words[words_from:], dtype=np.uint64)
        reals = np.asarray(program._reals[reals_from:], dtype=np.float64)
        # base pointers of the full pools: offsets inside the ops and descriptors are absolute
        wbase = C.c_void_p(words.ctypes.data - 8 * words_from) if words.size else None
        rbase = C.c_void_p(reals.ctypes.data - 8 * reals_from) if reals.size else None
        flags = _lib.APPLY_ASYNC | (0 if first else _lib.APPLY_CONTINUE)
        _lib.check(self._lib.b200_apply_program_ex(self._h, ops, n, wbase, len(program._words), rbase,
                                                   len(program._reals), flags))

    def last_program_ms(self) -> float:
        ms = C.c_double()
        _lib.check(self._lib.b200_last_program_ms(self._h, C.byref(ms)))
        return ms.value

    def set_profiling(self, on: bool) -> None:
        _lib.check(self._lib.b200_set_profiling(self._h, int(on)))

    def last_op_times(self) -> Tuple[np.ndarray, np.ndarray]:
        n = C.c_int()
        _lib.check(self._lib.b200_last_op_times(self._h, None, None, 0, C.byref(n)))
        ms = np.zeros(n.value, dtype=np.float32)
        kinds = np.zeros(n.value, dtype=np.int32)
        if n.value:
            _lib.check(self._lib.b200_last_op_times(self._h, ms.ctypes.data_as(C.c_void_p),
                                                    kinds.ctypes.data_as(C.c_void_p), n.value, C.byref(n)))
        return ms, kinds

    def launch_count(self) -> int:
        v = C.c_uint64()
        _lib.check(self._lib.b200_launch_count(self._h, C.byref(v)))
        return v.value

    # -- measurement ------------------------------------------------------------------------------
    def prob_zero(self, target: int) -> float:
        p = C.c_double()
        _lib.check(self._lib.b200_prob_zero(self._h, target, C.byref(p)))
        return p.value

    def collapse(self, target: int, outcome: int, p: float, reset: bool = False) -> None:
        _lib.check(self._lib.b200_collapse(self._h, target, outcome, p, int(reset)))

    def norm2(self) -> float:
        p = C.c_double()
        _lib.check(self._lib.b200_norm2(self._h, C.byref(p)))
        return p.value

    def scale(self, z: complex) -> None:
        z = complex(z)
        _lib.check(self._lib.b200_scale(self._h, z.real, z.imag))

    def first_above(self, eps: float) -> Optional[Tuple[int, complex]]:
        idx, re, im = C.c_uint64(), C.c_double(), C.c_double()
        _lib.check(self._lib.b200_first_above(self._h, eps, C.byref(idx), C.byref(re), C.byref(im)))
        if idx.value == 0xFFFFFFFFFFFFFFFF:
            return None
        return idx.value, complex(re.value, im.value)

    def first_above_wire(self, eps: float, wire_of_bit: Sequence[int]) -> Optional[int]:
        """Smallest wire-order index of an amplitude of this shard above eps (bit p of the full index is wire
        wire_of_bit[p]); None when there is none."""
        wob = np.ascontiguousarray(wire_of_bit, dtype=np.int32)
        key = C.c_uint64()
        _lib.check(self._lib.b200_first_above_wire(self._h, eps, wob.ctypes.data_as(C.c_void_p), int(wob.size), C.byref(key)))
        return None if key.value == 0xFFFFFFFFFFFFFFFF else int(key.value)

    def sample(self, order: Sequence[int], uniforms: np.ndarray) -> np.ndarray:
        """uniforms: (n_shots, m) float64 -> (n_shots,) uint64, bit j set when qubit order[j] read 1."""
        order = np.ascontiguousarray(order, dtype=np.int32)
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(-1, max(1, order.size))
        n_shots = uniforms.shape[0]
        out = np.zeros(n_shots, dtype=np.uint64)
        _lib.check(self._lib.b200_sample(self._h, order.ctypes.data_as(C.c_void_p), order.size,
                                         uniforms.ctypes.data_as(C.c_void_p), n_shots,
                                         out.ctypes.data_as(C.c_void_p)))
        return out

    def expect_group(self, xmask: int, zymasks: Sequence[int], coeffs: Sequence[complex]) -> complex:
        zy = np.ascontiguousarray(zymasks, dtype=np.uint64)
        cre = np.ascontiguousarray([complex(c).real for c in coeffs], dtype=np.float64)
        cim = np.ascontiguousarray([complex(c).imag for c in coeffs], dtype=np.float64)
        re, im = C.c_double(), C.c_double()
        _lib.check(self._lib.b200_expect_group(self._h, xmask, zy.size, zy.ctypes.data_as(C.c_void_p),
                                               cre.ctypes.data_as(C.c_void_p), cim.ctypes.data_as(C.c_void_p),
                                               C.byref(re), C.byref(im)))
        return complex(re.value, im.value)

    def sync(self) -> None:
        _lib.check(self._lib.b200_sync(self._h))

    # -- shards (qaqarot_b200/sharded.py) -----------------------------------------------------------
    def set_index_offset(self, offset: int) -> None:
        _lib.check(self._lib.b200_state_set_index_offset(self._h, offset))

    def ipc_export(self) -> bytes:
        buf = (C.c_ubyte * 64)()
        _lib.check(self._lib.b200_state_ipc_export(self._h, buf))
        return bytes(buf)

    def ipc_open(self, handle: bytes) -> int:
        """Map a partner process's state (its ipc_export()) into this process; returns the device pointer."""
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        _lib.check(self._lib.b200_ipc_open(self.device, buf, C.byref(p)))
        return p.value

    def ipc_close(self, peer_ptr: int) -> None:
        _lib.check(self._lib.b200_ipc_close(C.c_void_p(peer_ptr)))

    def peer_swap_bits(self, peer_ptr: int, lbits: Sequence[int], mine_pat: int, peer_pat: int, start: int, count: int,
                       sync: bool = False) -> None:
        """One partner's share of a k-bit remap (include/b200sv.h: b200_peer_swap_bits)."""
        lb = np.ascontiguousarray(lbits, dtype=np.int32)
        _lib.check(self._lib.b200_peer_swap_bits(self._h, C.c_void_p(peer_ptr), lb.ctypes.data_as(C.c_void_p), int(lb.size),
                                                 mine_pat, peer_pat, start, count, int(sync)))

    def peer_swap(self, peer_ptr: int, lbit: int, my_bit: int, start: int, count: int) -> None:
        """Pairs [start, start + count): amp[i | lbit = 1 - my_bit] <-> peer[i | lbit = my_bit] (synchronous)."""
        _lib.check(self._lib.b200_peer_swap(self._h, C.c_void_p(peer_ptr), lbit, my_bit, start, count))

    MARGINAL_LOW_BITS = (8, 13)        # b200_marginal_low accepts this many low index bits

    def marginal_low(self, k: int) -> np.ndarray:
        """table[v] = sum |amp[i]|^2 over i mod 2^k = v, one contiguous read of the state (8 <= k <= 13)."""
        out = np.zeros(1 << k, dtype=np.float64)
        _lib.check(self._lib.b200_marginal_low(self._h, k, out.ctypes.data_as(C.c_void_p)))
        return out

    def cond_probs(self, fixed_mask: int, qubit: int, group_bits: np.ndarray) -> np.ndarray:
        """(n_groups, 2): sum |amp|^2 of every group with bit `qubit` = 0 / 1 (qubit < 0: whole group, 0)."""
        gb = np.ascontiguousarray(group_bits, dtype=np.uint64)
        out = np.zeros((gb.size, 2), dtype=np.float64)
        _lib.check(self._lib.b200_cond_probs(self._h, fixed_mask, qubit, gb.ctypes.data_as(C.c_void_p), gb.size,
                                             out.ctypes.data_as(C.c_void_p)))
        return out

    def marginal(self, qubits: Sequence[int]) -> np.ndarray:
        """Probabilities of the 2^k outcomes of the distinct `qubits` (bit j of the index = qubits[j]): one
        reduction over the state, nothing but 2^k doubles comes back (samplers.py)."""
        return marginal_from_cond_probs(self, qubits)

    # -- timing (bench.py / profiling) ------------------------------------------------------------
    def event_record(self, slot: int) -> None:
        _lib.check(self._lib.b200_event_record(self._h, slot))

    def event_elapsed_ms(self, slot_a: int, slot_b: int) -> float:
        ms = C.c_double()
        _lib.check(self._lib.b200_event_elapsed_ms(self._h, slot_a, slot_b, C.byref(ms)))
        return ms.value


def marginal_from_cond_probs(st, qubits: Sequence[int]) -> np.ndarray:
    """Marginal distribution through `st.cond_probs` (any state object with that method): every outcome is one
    group of the reduction."""
    qubits = [int(q) for q in qubits]
    assert len(set(qubits)) == len(qubits) and len(qubits) <= 24
    idx = np.arange(1 << len(qubits), dtype=np.uint64)
    gb = np.zeros_like(idx)
    for j, q in enumerate(qubits):
        gb |= ((idx >> np.uint64(j)) & np.uint64(1)) << np.uint64(q)
    return st.cond_probs(sum(1 << q for q in qubits), -1, gb)[:, 0].copy()


class PinnedBuffer:
    """Page-locked host memory exposed as a complex128 ndarray (destination of fast downloads)."""

    def __init__(self, count: int):
        self._lib = _lib.load()
        self._p = C.c_void_p()
        self.count = count
        _lib.check(self._lib.b200_host_alloc(16 * max(1, count), C.byref(self._p)))
        buf = (C.c_double * (2 * count)).from_address(self._p.value)
        self.array = np.frombuffer(buf, dtype=np.complex128, count=count)

    def close(self) -> None:
        if getattr(self, "_p", None) is not None and self._p.value:
            self.array = None
            self._lib.b200_host_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
