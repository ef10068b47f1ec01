"""Device op program: the arrays handed to ``b200_apply_program`` (include/b200sv.h)."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from . import _lib
from .lowering import Prim


class Program:
    """ops (struct b200_op[]), words (uint64 descriptor pool), reals (float64 payload pool)."""

    def __init__(self):
        self._ops: List[tuple] = []
        self._words: List[int] = []
        self._reals: List[float] = []
        self.n_prims = 0          # primitives folded into this program
        self.n_passes = 0         # state passes (one per op)
        self.pass_bytes = 0.0     # algorithmic bytes moved, summed over passes (set by the planner)
        self._frozen = None
        self._bytes: List[float] = []   # algorithmic bytes per op (read + write of the amplitudes it touches)
        self.n_exchanges = 0            # OP_EXCHANGE ops (run by the host layer of a sharded state over NVLink)
        self.initial_pos: List[int] = []   # wire -> index bit the program assumes / leaves (planner.plan)
        self.final_pos: List[int] = []

    # -- building -------------------------------------------------------------------------------
    def add_reals(self, values: Sequence[complex]) -> int:
        off = len(self._reals)
        if len(values) > 8:                  # (tables: one conversion instead of a Python loop)
            self._reals.extend(np.asarray(values, dtype=np.complex128).view(np.float64).tolist())
        else:
            for v in values:
                v = complex(v)
                self._reals += [v.real, v.imag]
        return off

    def add_words(self, values: Sequence[int]) -> int:
        off = len(self._words)
        self._words += [int(v) & 0xFFFFFFFFFFFFFFFF for v in values]
        return off

    def add_op(self, kind: int, target: int = 0, aux: int = 0, cmask: int = 0, roff: int = 0, woff: int = 0,
               wlen: int = 0, nbytes: float = 0.0) -> None:
        self._ops.append((kind, target, aux, wlen, cmask, woff, roff))
        self._bytes.append(float(nbytes))
        self.n_passes += 1
        self.pass_bytes += float(nbytes)
        self._frozen = None

    @property
    def op_bytes(self) -> List[float]:
        return self._bytes

    def add_prim(self, p: Prim, n_qubits: int = 0) -> None:
        """One primitive as its own single-gate pass."""
        cmask = 0
        for c in p.controls:
            cmask |= 1 << c
        full = 32.0 * 2.0 ** n_qubits / (1 << len(p.controls))
        if p.kind == "mat":
            self.add_op(_lib.OP_MAT, p.target, cmask=cmask, roff=self.add_reals(p.data), nbytes=full)
        elif p.kind == "diag":
            touched = full / 2 if complex(p.data[0]) == 1.0 else full
            self.add_op(_lib.OP_DIAG, p.target, cmask=cmask, roff=self.add_reals(p.data), nbytes=touched)
        elif p.kind == "x":
            self.add_op(_lib.OP_X, p.target, cmask=cmask, nbytes=full)
        elif p.kind == "swap":
            self.add_op(_lib.OP_SWAP, p.target, aux=p.aux, nbytes=full / 2)
        else:
            raise ValueError(f"primitive {p.kind} is not a unitary pass")
        self.n_prims += 1

    # -- consumption ----------------------------------------------------------------------------
    def frozen_with_init(self, index: int):
        """The same program preceded by B200_OP_INIT (state := |index>), for runs that start from a basis state."""
        key = int(index)
        cache = self.__dict__.setdefault("_frozen_init", {})
        if key not in cache or cache[key][0] is not self.frozen():
            base = self.frozen()
            ops0, n_ops, words, reals = base
            ops = (_lib.Op * (n_ops + 1))()
            ops[0].kind, ops[0].cmask = _lib.OP_INIT, key
            for i in range(n_ops):
                for f, _ in _lib.Op._fields_:
                    setattr(ops[i + 1], f, getattr(ops0[i], f))
            cache[key] = (base, (ops, n_ops + 1, words, reals))
        return cache[key][1]

    def frozen(self):
        if self._frozen is None:
            ops = (_lib.Op * max(1, len(self._ops)))()
            for i, (kind, target, aux, wlen, cmask, woff, roff) in enumerate(self._ops):
                o = ops[i]
                o.kind, o.target, o.aux, o.wlen, o.cmask, o.woff, o.roff = kind, target, aux, wlen, cmask, woff, roff
            words = np.asarray(self._words, dtype=np.uint64)
            reals = np.asarray(self._reals, dtype=np.float64)
            self._frozen = (ops, len(self._ops), words, reals)
        return self._frozen

    def __len__(self):
        return len(self._ops)

    def segments(self):
        """Split at the host-level EXCHANGE ops: yields sub-programs (sharing this program's payload pools) and
        (top local bit, global bit) tuples, in order."""
        out, cur = [], None
        for op, nbytes in zip(self._ops, self._bytes):
            if op[0] == _lib.OP_EXCHANGE:
                if cur is not None:
                    out.append(cur)
                    cur = None
                out.append((op[1], op[2]))
                continue
            if cur is None:
                cur = Program()
                cur._words, cur._reals = self._words, self._reals
            cur._ops.append(op)
            cur._bytes.append(nbytes)
            cur.n_passes += 1
        if cur is not None:
            out.append(cur)
        return out


def unfused_program(prims: Sequence[Prim], n_qubits: int = 0) -> Program:
    prog = Program()
    for p in prims:
        prog.add_prim(p, n_qubits)
    return prog
