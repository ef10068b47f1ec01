"""Lower circuit operations to the primitive instruction set the device executes.

Input: any sequence of operations quacking like blueqat ``Operation`` (``lowername``, ``targets``,
``target_iter``, ``control_target_iter``, typed attributes, optional ``fallback``) -- blueqat's own objects or
``qaqarot_b200.circuit.Op``.  Output: a flat list of ``Prim`` records, one per (operation, target) after slice
expansion, in program order:

    mat      2x2 complex matrix on `target`, applied where all `controls` are 1
    diag     diag(d0, d1) on `target`, applied where all `controls` are 1
    x        pair swap on `target` under `controls` (X / CX / CCX as permutations)
    swap     exchange of two qubits
    measure  projective measurement of `target` (carries the key/duplicated bookkeeping of its op)
    reset    measurement followed by a move to |0>

Matrix entries are built with the same scalar expressions as the reference's kernels (cited per gate) so the
only differences from NumPyBackend are the rounding of the fused arithmetic, far inside the 1e-12 budget.
Anything without a native rule goes through ``op.fallback(n_qubits)`` recursively, exactly like
``run_single_gate`` (numpy_backend.py:548-560); an operation with neither raises the reference's ValueError.
"""
from __future__ import annotations

import cmath
import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np


class Prim:
    __slots__ = ("kind", "target", "controls", "data", "aux", "op_index", "group")

    def __init__(self, kind: str, target: int, controls: Tuple[int, ...] = (), data=None, aux: int = -1,
                 op_index: int = -1, group=None):
        self.kind = kind
        self.target = target
        self.controls = tuple(controls)
        self.data = data            # mat: 4 complex (row major); diag: 2 complex
        self.aux = aux              # swap: second qubit
        self.op_index = op_index    # index of the originating op in the circuit's op list
        self.group = group          # measure: (key, duplicated, op serial)

    @property
    def qubits(self) -> Tuple[int, ...]:
        q = (self.target,) + self.controls
        return q + (self.aux,) if self.aux >= 0 else q

    @property
    def unitary(self) -> bool:
        return self.kind not in ("measure", "reset")

    def __repr__(self):
        c = f" c{list(self.controls)}" if self.controls else ""
        return f"<{self.kind} t{self.target}{c}>"


_SQRT2_INV = 1.0 / math.sqrt(2)         # gate_h: numpy_backend.py:130


def _mat(a00, a01, a10, a11) -> np.ndarray:
    return np.array([a00, a01, a10, a11], dtype=complex)


def _u_entries(theta, phi, lam, gamma):
    """gate_u / gate_cu entry construction (numpy_backend.py:275-280, 405-410)."""
    gp = cmath.exp(1j * gamma)
    a00 = math.cos(theta * 0.5) * gp
    a11 = a00 * cmath.exp(1j * (phi + lam))
    a01 = a10 = math.sin(theta * 0.5) * gp
    a01 *= -cmath.exp(1j * lam)
    a10 *= cmath.exp(1j * phi)
    return _mat(a00, a01, a10, a11)


def _rx_entries(theta):                  # numpy_backend.py:144-146, 301-303
    h = theta * 0.5
    return _mat(math.cos(h), -1j * math.sin(h), -1j * math.sin(h), math.cos(h))


def _ry_entries(theta):                  # :165-168, 324-327
    h = theta * 0.5
    return _mat(math.cos(h), -math.sin(h), math.sin(h), math.cos(h))


def _rz_entries(theta):                  # :186-188, 347-349
    h = theta * 0.5
    return np.array([complex(math.cos(h), -math.sin(h)), complex(math.cos(h), math.sin(h))])


def _phase_entries(theta):               # :201-202, 364-365
    return np.array([1.0, complex(math.cos(theta), math.sin(theta))])


# name -> (kind, payload builder(op)); arity decides how targets are expanded
_ONE: Dict[str, Tuple[str, Callable]] = {
    "x": ("x", lambda op: None),
    "y": ("mat", lambda op: _mat(0.0, -1.0j, 1.0j, 0.0)),
    "z": ("diag", lambda op: np.array([1.0, -1.0], dtype=complex)),
    "h": ("mat", lambda op: _mat(_SQRT2_INV, _SQRT2_INV, _SQRT2_INV, -_SQRT2_INV)),
    "s": ("diag", lambda op: np.array([1.0, 1.0j])),
    "t": ("diag", lambda op: np.array([1.0, complex(_SQRT2_INV, _SQRT2_INV)])),     # :215-216
    "phase": ("diag", lambda op: _phase_entries(op.theta)),
    "rz": ("diag", lambda op: _rz_entries(op.theta)),
    "rx": ("mat", lambda op: _rx_entries(op.theta)),
    "ry": ("mat", lambda op: _ry_entries(op.theta)),
    "u": ("mat", lambda op: _u_entries(op.theta, op.phi, op.lam, op.gamma)),
    "mat1": ("mat", lambda op: np.asarray(op.mat, dtype=complex).reshape(4).copy()),
}
_TWO: Dict[str, Tuple[str, Callable]] = {
    "cx": ("x", lambda op: None),
    "cz": ("diag", lambda op: np.array([1.0, -1.0], dtype=complex)),
    "cphase": ("diag", lambda op: _phase_entries(op.theta)),
    "crz": ("diag", lambda op: _rz_entries(op.theta)),
    "crx": ("mat", lambda op: _rx_entries(op.theta)),
    "cry": ("mat", lambda op: _ry_entries(op.theta)),
    "cu": ("mat", lambda op: _u_entries(op.theta, op.phi, op.lam, op.gamma)),
}


class Lowered:
    """Result of lowering: the primitive list plus bookkeeping the backend needs."""

    def __init__(self):
        self.prims: List[Prim] = []
        self.n_gates = 0              # blueqat-level single-target gates after slice expansion
        self.first_nonunitary = None  # index into prims


def lower(ops: Sequence, n_qubits: int, start_index: int = 0) -> Lowered:
    out = Lowered()
    serial = [0]

    def check(q: int) -> int:
        if not 0 <= q < n_qubits:
            raise ValueError(f"qubit index {q} out of range for {n_qubits} qubits")
        return q

    def emit(op, op_index: int, top: bool) -> None:
        name = op.lowername
        if name in _ONE:
            kind, build = _ONE[name]
            data = build(op)
            for t in op.target_iter(n_qubits):
                out.prims.append(Prim(kind, check(t), (), data, op_index=op_index))
                out.n_gates += top
        elif name in _TWO:
            kind, build = _TWO[name]
            data = build(op)
            for c, t in op.control_target_iter(n_qubits):
                out.prims.append(Prim(kind, check(t), (check(c),), data, op_index=op_index))
                out.n_gates += top
        elif name == "swap":
            for a, b in op.control_target_iter(n_qubits):
                out.prims.append(Prim("swap", check(a), (), None, aux=check(b), op_index=op_index))
                out.n_gates += top
        elif name == "ccz":
            c1, c2, t = op.targets
            out.prims.append(Prim("diag", check(t), (check(c1), check(c2)), np.array([1.0, -1.0], dtype=complex),
                                  op_index=op_index))
            out.n_gates += top
        elif name == "ccx":
            # the reference computes H.CCZ.H (numpy_backend.py:385-392); the permutation agrees to ~1e-16
            c1, c2, t = op.targets
            out.prims.append(Prim("x", check(t), (check(c1), check(c2)), None, op_index=op_index))
            out.n_gates += top
        elif name in ("measure", "reset"):
            serial[0] += 1
            grp = (getattr(op, "key", None), getattr(op, "duplicated", None), serial[0])
            targets = [check(t) for t in op.target_iter(n_qubits)]
            if name == "measure" and not targets:
                # an empty measurement still performs the key bookkeeping (numpy_backend.py:467-476)
                out.prims.append(Prim("measure", -1, (), None, op_index=op_index, group=grp))
            for t in targets:
                out.prims.append(Prim(name, t, (), None, op_index=op_index, group=grp))
        elif hasattr(op, "fallback"):
            before = len(out.prims)
            for g in op.fallback(n_qubits):
                emit(g, op_index, False)
            if top:
                # count the expanded applications of the original gate, not its decomposition
                out.n_gates += _applications(op, n_qubits, len(out.prims) - before)
        else:
            raise ValueError(f"Unknown operation {name}.")

    for i, op in enumerate(ops):
        emit(op, start_index + i, True)
    for i, p in enumerate(out.prims):
        if not p.unitary:
            out.first_nonunitary = i
            break
    return out


def _applications(op, n_qubits: int, n_prims: int) -> int:
    """How many single-target gates an operation stands for after slice expansion."""
    try:
        arity = getattr(op, "n_qargs", 1)
        if arity == 1:
            return len(list(op.target_iter(n_qubits)))
        if arity == 2:
            return len(list(op.control_target_iter(n_qubits)))
        return 1
    except Exception:
        return 1 if n_prims else 0
