"""The fusion planner's normal form as a circuit.

``Circuit.run(backend="b200_fuse")`` returns a new circuit, like blueqat's transpiler backends
(``1q_compaction``, backends/onequbitgate_transpiler.py:12-51): chains of 1-qubit gates merged into one ``mat1``
(also across the controls and phases they commute with, which the reference pass cannot do), every diagonal gate as
``phase`` / ``cphase`` grouped per target, a carried X / Y pushed through the phases it meets, CX . RZ . CX folded
into phases.  The plan of a circuit thereby becomes inspectable and testable with any backend: the returned circuit
has the same statevector (including the global phase) as the original.  Measurements and resets keep their place.
"""
from __future__ import annotations

import cmath
import math
from typing import List, Sequence

import numpy as np

from .lowering import lower
from .planner import normalize


def _cu_params(m):
    """(theta, phi, lam, gamma) of blueqat's U gate (numpy_backend.py:405-410) for a 2x2 unitary, row major."""
    a00, a01, a10, a11 = (complex(x) for x in m)
    theta = 2.0 * math.atan2(abs(a10), abs(a00))
    gamma = cmath.phase(a00) if abs(a00) > 1e-15 else 0.0
    phi = cmath.phase(a10) - gamma if abs(a10) > 1e-15 else cmath.phase(a11) - gamma
    lam = cmath.phase(-a01) - gamma if abs(a01) > 1e-15 else 0.0
    gp = cmath.exp(1j * gamma)
    rec = [math.cos(theta / 2) * gp, -cmath.exp(1j * lam) * math.sin(theta / 2) * gp,
           cmath.exp(1j * phi) * math.sin(theta / 2) * gp, cmath.exp(1j * (phi + lam)) * math.cos(theta / 2) * gp]
    if max(abs(r - x) for r, x in zip(rec, (a00, a01, a10, a11))) > 1e-12:
        raise ValueError("controlled 2x2 is not unitary: no blueqat gate expresses it")
    return theta, phi, lam, gamma


def _unit_angle(w: complex) -> float:
    if abs(abs(w) - 1.0) > 1e-12:
        raise ValueError("phase factor is not of unit modulus: no blueqat gate expresses it")
    return cmath.phase(w)


def _emit(c, it) -> None:
    if it.kind == "mat":
        if not it.ctrls:
            c.mat1(np.asarray(it.m, dtype=complex).reshape(2, 2))[it.t]
        elif len(it.ctrls) == 1:
            c.cu(*_cu_params(it.m))[it.ctrls[0], it.t]
        else:
            raise ValueError("multi-controlled 2x2: no blueqat gate expresses it")
    elif it.kind == "x":
        if not it.ctrls:
            c.x[it.t]
        elif len(it.ctrls) == 1:
            c.cx[it.ctrls[0], it.t]
        elif len(it.ctrls) == 2:
            c.ccx[it.ctrls[0], it.ctrls[1], it.t]
        else:
            raise ValueError("X with more than two controls")
    elif it.kind == "swap":
        c.swap[it.t, it.aux]
    elif it.kind == "phase":
        if len(it.bits) == 0:                        # the segment's global factor: blueqat has no such gate
            c.mat1(np.array([[it.w, 0], [0, it.w]], dtype=complex))[0]
        elif len(it.bits) == 1:
            c.phase(_unit_angle(it.w))[it.bits[0]]
        elif len(it.bits) == 2:
            c.cphase(_unit_angle(it.w))[it.bits[0], it.bits[1]]
        elif len(it.bits) == 3 and abs(it.w + 1.0) < 1e-15:
            c.ccz[it.bits[0], it.bits[1], it.bits[2]]
        else:
            raise ValueError("phase on more than two qubits")
    elif it.kind == "fan":
        if it.w != 1.0:
            c.phase(_unit_angle(it.w))[it.t]
        for q, w in it.entries.items():
            c.cphase(_unit_angle(w))[q, it.t]
    else:                                            # pragma: no cover - not produced for an unsharded segment
        raise ValueError(f"unexpected planner item {it.kind}")


def fuse_ops(ops: Sequence, n_qubits: int, circuit_cls=None):
    """The operations `ops` on `n_qubits` qubits -> an equivalent circuit in the planner's normal form."""
    if circuit_cls is None:
        circuit_cls = _circuit_class(ops)
    out = circuit_cls(n_qubits)
    low = lower(ops, n_qubits)
    seg: List = []
    done_ops = set()

    def flush():
        if seg:
            norm = normalize(seg, n_qubits, relabel_swaps=False)
            for it in norm.items:
                _emit(out, it)
            seg.clear()

    for p in low.prims:
        if p.unitary:
            seg.append(p)
            continue
        flush()
        if p.op_index not in done_ops:               # a measurement / reset op lowers to several primitives
            done_ops.add(p.op_index)
            out.ops.append(ops[p.op_index])
    flush()
    return out


def _circuit_class(ops):
    if ops and type(ops[0]).__module__.split(".")[0] == "blueqat":
        from blueqat import Circuit
        return Circuit
    from .circuit import Circuit
    return Circuit


class FuseTranspiler:
    """Backend object for ``Circuit.run(backend="b200_fuse")``: returns the fused circuit instead of running it."""

    def run(self, gates, n_qubits, *args, **kwargs):
        return fuse_ops(list(gates), n_qubits)

    def make_cache(self, gates, n_qubits):
        return None

    def copy(self):
        return FuseTranspiler()
