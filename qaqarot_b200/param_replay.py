"""Parametrised re-launch for the objective loop (vqe.py:67-83, utils.py:44-62).

The optimiser of a VQE / QAOA run calls ``get_circuit(params) -> run`` hundreds of times, and nothing but the angles of
the rotation and phase gates changes between the calls.  The op program of such a circuit has the same descriptor words
every time; only the float64 payload moves -- and, when every rotation runs as three exact shears
(``PlanConfig.exact_rotations``), it moves in a very regular way:

* every table, fan and phase entry is  E0 * exp(i * sum_g a_g (u_g - u0_g))  with exponents a_g that are multiples of
  1/4 (products of the phases exp(+-i theta/2), exp(i theta) ... the gates contribute), and
* every rotation slot (z, x, y) of a LAYER record is  (-tan(t/2), sin t, -tan(t/2))  with  t = t0 + sum_g s_g (u_g - u0_g),

where u_g are the distinct angle values of the call (gates that carried the same angle in two consecutive calls form
one group g: the 2p angles of a QAOA ansatz, not its 300 gates).

``ParamModel`` finds a_g and s_g by PROBING the planner -- one plan per group with that group's angle moved by DELTA --
and checks the result against a further plan at random angles.  It is a model of the planner, not a second planner,
so it is only ever used to SPECULATE: ``B200Backend`` launches the program with the model's payload at once, plans the
circuit for real while the device is already running (the plan costs about as much host time as the circuit costs
device time), and compares.  Equal (<= TOL = 1e-13): the result stands and the iteration cost the device time plus ~1 ms.
Different (a branch of the planner went the other way for these angles): the circuit runs again with the real plan and
the model is dropped.  The answer never depends on the model being right.

The model holds around the angles it was built at (every merged rotation within |t| <= pi/2; beyond, the planner turns
the rotation around and the slot is no longer affine): ``evaluate`` returns None outside, the backend plans first as it
always did, and after RECENTRE_AFTER such calls in a row it builds a new model around the angles the loop has reached.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .program import Program

DELTA = 1.0e-3            # probe step: exponents up to ~3000 / (group) stay below a phase wrap
MAX_GROUPS = 32           # one probe plan per group
RECENTRE_AFTER = 3        # calls in a row the model had to decline before the backend builds a new one around the current angles
MAX_RECENTRES = 4         # ... at most so many times per gate structure (a loop that keeps jumping gets no model)
TOL = 1.0e-13             # predicted against planned payload: 1.3e-14 at most over QAOA-28 p=4 angle sweeps (the planner's
                          # own products of exp() round as much), so an order of magnitude of head-room; a payload entry off
                          # by TOL moves an amplitude by at most TOL per record, well inside the 1e-12 parity bar


class NotModelable(Exception):
    """The circuit's payload is not of the form above (or its structure moves with the angles)."""


def _is_angle(p) -> bool:
    return isinstance(p, (float, np.floating)) and not isinstance(p, bool)


def structure_key(gates: Sequence, n_qubits: int) -> Tuple:
    """Everything about a gate list except the values of its float parameters."""
    key: List = [n_qubits]
    for op in gates:
        params = tuple(getattr(op, "params", ()))
        sig = tuple("f" if _is_angle(p) else repr(p) for p in params)
        key.append((type(op).__name__, getattr(op, "lowername", None), repr(op.targets), sig,
                    repr(getattr(op, "options", None)), repr(getattr(op, "key", None)),
                    repr(getattr(op, "duplicated", None))))
    return tuple(key)


def angle_values(gates: Sequence) -> np.ndarray:
    return np.array([float(p) for op in gates for p in getattr(op, "params", ()) if _is_angle(p)], dtype=np.float64)


def _clone(op, params: tuple):
    if hasattr(op, "spec"):                                   # host mirror (qaqarot_b200.circuit.Op)
        new = type(op)(op.spec, op.targets, params, dict(op.options) if op.options else None)
        new.key, new.duplicated = op.key, op.duplicated
        return new
    create = getattr(type(op), "create", None)                # blueqat.gate.Operation.create(targets, params, options)
    if create is None:
        raise NotModelable(f"cannot rebuild a {type(op).__name__} with other parameters")
    return create(op.targets, tuple(params), getattr(op, "options", None))


def with_angles(gates: Sequence, values: np.ndarray) -> List:
    out, k = [], 0
    for op in gates:
        params = tuple(getattr(op, "params", ()))
        if not any(_is_angle(p) for p in params):
            out.append(op)
            continue
        new = []
        for p in params:
            if _is_angle(p):
                new.append(float(values[k]))
                k += 1
            else:
                new.append(p)
        out.append(_clone(op, tuple(new)))
    assert k == len(values)
    return out


def _layer_payload_offsets(prog: Program) -> List[int]:
    """Offsets (into reals) of the (z, x, y, 0) slots of every LAYER record of every TILE pass."""
    from .planner import T_LAYER
    w = prog._words
    out: List[int] = []
    for kind, _t, _aux, _wlen, _cmask, woff, _roff in prog._ops:
        if kind != _lib.OP_TILE:
            continue
        n_rounds = (w[woff] >> 24) & 0xFF
        pos = woff + (w[woff + 3] >> 32)
        for _ in range(n_rounds):
            ng, nw = w[pos] & 0xFFFF, w[pos] >> 16
            for g in range(ng):
                if w[pos + 3 + 3 * g] & 0xFF == T_LAYER:
                    roff = w[pos + 5 + 3 * g] & 0xFFFFFFFF
                    out += [roff + 4 * k for k in range(4)]
            pos += nw
    return out


class ParamModel:
    def __init__(self, gates: Sequence, prev_values: np.ndarray, plan_fn: Callable[[Sequence], Program],
                 rng: Optional[np.random.Generator] = None):
        """`gates`: the gate list of the current call; `prev_values`: the angles the same structure had in the call
        before (gates whose angles were equal BOTH times share a group); `plan_fn(gates) -> Program` plans the unitary
        prefix exactly as the backend will (exact rotations)."""
        vals = angle_values(gates)
        if vals.size == 0 or vals.size != prev_values.size:
            raise NotModelable("no angles")
        groups: Dict[Tuple[float, float], List[int]] = {}
        for i, (a, b) in enumerate(zip(prev_values, vals)):
            groups.setdefault((float(a), float(b)), []).append(i)
        if len(groups) > MAX_GROUPS:
            raise NotModelable(f"{len(groups)} independent angles")
        self.declined = 0          # consecutive calls evaluate() returned None for (kept by the backend)
        self.generation = 0        # how many models of this structure came before this one
        self.members = [np.array(ix, dtype=np.int64) for ix in groups.values()]
        self.n_values = vals.size
        self.u0 = np.array([vals[ix[0]] for ix in self.members])
        G = len(self.members)
        base = plan_fn(gates)
        self.ops, self.words, self.bytes = list(base._ops), list(base._words), list(base._bytes)
        self.n_prims, self.final_pos, self.initial_pos = base.n_prims, list(base.final_pos), list(base.initial_pos)
        r0 = np.array(base._reals, dtype=np.float64)
        if r0.size % 2:
            raise NotModelable("odd payload")
        shear = np.array(sorted(set(_layer_payload_offsets(base))), dtype=np.int64)
        is_shear = np.zeros(r0.size, dtype=bool)
        for o in shear:
            is_shear[o:o + 4] = True
        cplx = np.nonzero(~is_shear[0::2])[0] * 2               # offsets of the (re, im) pairs outside the layer payloads
        E0 = r0[cplx] + 1j * r0[cplx + 1]
        t0 = -2.0 * np.arctan(r0[shear]) if shear.size else np.zeros(0)
        self._check_shear(r0, shear, t0)
        A = np.zeros((cplx.size, G))
        S = np.zeros((shear.size, G))
        for g in range(G):
            v = vals.copy()
            v[self.members[g]] += DELTA
            r = self._payload_of(plan_fn(with_angles(gates, v)), r0.size)
            E = r[cplx] + 1j * r[cplx + 1]
            live = np.abs(E0) > 0.0
            if np.any(np.abs(np.abs(E) - np.abs(E0)) > TOL) or np.any(np.abs(E[~live]) > 0.0):
                raise NotModelable("an entry changes its modulus with an angle")
            ang = np.zeros(cplx.size)
            ang[live] = np.angle(E[live] / E0[live])
            a = np.round(ang / DELTA * 4.0) / 4.0
            if np.any(np.abs(a * DELTA - ang) > 1e-9):
                raise NotModelable("a phase is not exp(i a theta) with a in Z/4")
            A[:, g] = a
            if shear.size:
                t = -2.0 * np.arctan(r[shear])
                self._check_shear(r, shear, t)
                s = np.round((t - t0) / DELTA * 4.0) / 4.0
                if np.any(np.abs(s * DELTA - (t - t0)) > 1e-9):
                    raise NotModelable("a rotation angle is not affine in the gate angles")
                S[:, g] = s
        self.r0, self.cplx, self.E0, self.A = r0, cplx, E0, A
        self.shear, self.t0, self.S = shear, t0, S
        keep = np.nonzero(np.any(A != 0.0, axis=1))[0]            # only the entries that move at all are evaluated
        self.cplx_live, self.E0_live, self.A_live = cplx[keep], E0[keep], A[keep]
        keep = np.nonzero(np.any(S != 0.0, axis=1))[0]
        self.shear_live, self.t0_live, self.S_live = shear[keep], t0[keep], S[keep]
        # validation at angles the probes have not seen
        rng = rng or np.random.default_rng(12345)
        v = vals.copy()
        for ix in self.members:
            v[ix] += rng.uniform(-0.3, 0.3)
        got = self.evaluate(v)
        if got is None:
            raise NotModelable("validation point outside the model's range")
        want = self._payload_of(plan_fn(with_angles(gates, v)), r0.size)
        if np.abs(got - want).max() > TOL:
            raise NotModelable("the model does not reproduce the planner at the validation point")

    # -- helpers ----------------------------------------------------------------------------------
    def _payload_of(self, prog: Program, size: int) -> np.ndarray:
        if list(prog._words) != self.words or list(prog._ops) != self.ops or len(prog._reals) != size:
            raise NotModelable("the plan's structure moves with the angles")
        return np.array(prog._reals, dtype=np.float64)

    @staticmethod
    def _check_shear(r: np.ndarray, shear: np.ndarray, t: np.ndarray) -> None:
        if not shear.size:
            return
        z, x, y, pad = r[shear], r[shear + 1], r[shear + 2], r[shear + 3]
        if (np.any(np.abs(x - np.sin(t)) > TOL) or np.any(np.abs(y - z) > TOL) or np.any(pad != 0.0)
                or np.any(np.abs(t) > math.pi / 2 + 1e-9)):
            raise NotModelable("a rotation slot is not three exact shears")

    # -- use --------------------------------------------------------------------------------------
    def evaluate(self, values: np.ndarray) -> Optional[np.ndarray]:
        """Payload for the angle vector `values` (as angle_values returns it), or None when the angles do not fit the
        model (gates of one group no longer share a value, a rotation left |t| <= pi/2)."""
        if values.size != self.n_values:
            return None
        u = np.empty(len(self.members))
        for g, ix in enumerate(self.members):
            x = values[ix]
            if x.size > 1 and np.ptp(x) != 0.0:
                return None
            u[g] = x[0]
        du = u - self.u0
        r = self.r0.copy()
        if self.cplx_live.size:
            E = self.E0_live * np.exp(1j * (self.A_live @ du))
            r[self.cplx_live] = E.real
            r[self.cplx_live + 1] = E.imag
        if self.shear_live.size:
            t = self.t0_live + self.S_live @ du
            if np.any(np.abs(t) > math.pi / 2):
                return None
            z = -np.tan(0.5 * t)
            r[self.shear_live] = z
            r[self.shear_live + 1] = np.sin(t)
            r[self.shear_live + 2] = z
        return r

    def program(self, reals: np.ndarray) -> Program:
        """The model's op program with the payload `reals`."""
        p = Program()
        p._ops, p._words, p._bytes = list(self.ops), list(self.words), list(self.bytes)
        p._reals = reals                      # (an ndarray: apply_piece slices it without a copy)
        p.n_prims, p.n_passes = self.n_prims, len(self.ops)
        p.pass_bytes = float(sum(self.bytes))
        p.final_pos, p.initial_pos = list(self.final_pos), list(self.initial_pos)
        return p

    def matches(self, prog: Program) -> bool:
        return list(prog._words) == self.words and list(prog._ops) == self.ops
