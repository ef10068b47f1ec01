"""Gate-fusion planner: primitives -> fused tile passes (the descriptors csrc/tile.cu executes).

The reference applies one gate per full sweep of the state (numpy_backend.py:545-570); its only fusion
precedent is the `1q_compaction` transpiler (backends/onequbitgate_transpiler.py:12-51).  Gate application is
HBM-bound, so this planner trades sweeps for on-chip work in three steps:

1. ``normalize``: chains of uncontrolled 1-qubit gates collapse into one 2x2 (as 1q_compaction does, but also
   across commuting controls); every diagonal gate becomes a *phase* ``amp *= w where (index & mask) == mask``
   (a diag(d0, d1) gate is d0 times a phase d1/d0; the scalars are collected into one global factor); phases on
   one or two qubits sharing a qubit are merged into *fans* ``amp *= [bit j] * w0 * prod_k (bit k ? w_k : 1)``
   -- the controlled-phase ladder of a QFT is one fan per target.
2. ``schedule``: greedy list scheduling into *passes*.  A pass owns a set T of `a` physical index bits (the
   lowest `c` bits always, for coalesced rows, plus chosen high bits); a CTA holds the 2^a amplitudes of one
   tile in shared memory.  Non-diagonal targets must lie in T; controls and phases may sit on any bit (bits
   outside T are constant per tile).  Inside a pass gates are grouped into *rounds*: in a round each thread
   keeps 2^r amplitudes in registers (the round's r register bits) and applies the round's gates there.
   Commutation rule: two gates may be reordered unless they share a qubit on which one is non-diagonal.
3. ``encode``: the descriptor words / payload reals documented in csrc/tile.cu.  Within a round the gates are
   regrouped (commuting moves only) into few, fat records: up to four rotations per LAYER record, every diagonal on
   the register bits in its table, diagonals that also involve up to three other bits as table variants (``_Group``).

Rounding differs from the gate-by-gate reference only by the re-association of products (~1e-16 per fused
gate), far inside the 1e-12 parity budget.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, FrozenSet, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from .lowering import Prim
from .program import Program

T_NONE, T_MATC, T_MATR, T_X = 0, 1, 2, 3          # target stage of a tile gate record
T_LAYER = 6                                         # ... up to one 3-shear gate per register bit (reg_mask says which)
D_NONE, D_FANT, D_FAN, D_PHASE = 0, 1, 2, 3         # diagonal stage (fan without / with register table, phase)
D_W, D_FIX = 4, 5                                   # ... fan without entries / with entries outside the tile only
D_TAB = 6                                           # ... one factor per register amplitude
MAX_GATES = 160          # per pass (record slots in csrc/tile.cu)
TAB_PITCH = 17            # complex entries between the variants of a record's table (csrc/tile.cu kTabPitch)
MAX_PAYLOAD = 2560       # float64 payload of one pass (shared-memory copy in csrc/tile.cu)
MAX_FAN_WORDS = 512      # fan directory + records of one pass (shared-memory copy in csrc/tile.cu)
MAX_FANS = 64            # per pass (shared-memory slots in csrc/tile.cu)


@dataclass
class PlanConfig:
    a: int = 12               # tile bits (2^a amplitudes of 16 B in shared memory)
    r: int = 4                # register bits per round (2^r amplitudes per thread)
    c: int = 4                # contiguous low bits always in the tile (2^c * 16 B rows)
    max_rounds: int = 6
    min_qubits: int = 9       # below this the state is tiny: one kernel per gate
    max_cost: float = 300.0   # DFMA-equivalents per amplitude per pass before a new pass is opened (a pass costs
                              # ~7 ms of data movement at 30 qubits whatever it computes: measured 676 -> 645 ms
                              # on the 30-qubit random-layer circuit going from 150 to 300)
    scan_limit: int = 4096    # how far ahead of the first unscheduled item a pass may look
    direct_io: bool = False   # flag the FIRST round as loadable straight from HBM (measured neutral at A = 12; the
                              # kernel accepts the flag and ignores it since its code was removed, csrc/tile.cu)
    direct_out: bool = False  # flag the LAST round as storable straight to HBM (round 1; the TMA pipeline of round 2
                              # stores every tile through the TMA unit and ignores the flag)
    kernel_tiles: Tuple[int, ...] = (9, 10, 11, 12)       # tile sizes csrc/tile.cu instantiates (r = 4)
    bank_pairs: bool = True   # keep both local bits of a shared-memory bank pair (0,3), (1,4), (2,5) out of one
                              # round's register set when the gates can wait for the next round (see _thread_order)
    exact_rotations: bool = False   # every rotation as three exact shears (no cosine factors left for the tables): then
                              # every table / fan / phase entry of a circuit of rotations and phase gates is a constant
                              # times exp(i a.theta), which is what param_replay.ParamModel relies on
    shears: bool = True       # 1-qubit gates as phase . rotation by three shears . phase (split_shears), and
                              # CX as H . CZ . H


DEFAULT = PlanConfig()


class Item:
    """One schedulable unit after normalisation."""
    __slots__ = ("kind", "t", "ctrls", "m", "bits", "w", "entries", "aux", "nd", "dg", "n_prims", "pos", "sh")

    def __init__(self, kind, t=-1, ctrls=(), m=None, bits=(), w=1.0 + 0j, entries=None, aux=-1, n_prims=1):
        self.kind = kind            # mat | x | phase | fan | swap
        self.t = t                  # mat/x: target; fan: target; swap: first qubit
        self.ctrls = tuple(ctrls)   # mat/x
        self.m = m                  # mat: 4 complex, row major
        self.bits = tuple(bits)     # phase: qubits that must all be 1
        self.w = complex(w)         # phase: factor; fan: w0
        self.entries = entries      # fan: {qubit: factor}
        self.aux = aux              # swap: second qubit
        self.n_prims = n_prims
        self.pos = -1
        self.sh = None              # mat: theta when m = rotation_matrix(theta) runs as in-place shears of a LAYER record
        self.refresh()

    def refresh(self):
        if self.kind in ("mat", "x"):
            self.nd = frozenset((self.t,))
            self.dg = frozenset(self.ctrls)
        elif self.kind in ("swap", "perm", "exchange"):
            self.nd = frozenset((self.t, self.aux))
            self.dg = frozenset()
        elif self.kind == "phase":
            self.nd = frozenset()
            self.dg = frozenset(self.bits)
        else:
            self.nd = frozenset()
            self.dg = frozenset((self.t,) + tuple(self.entries))

    def __repr__(self):
        if self.kind == "fan":
            return f"<fan t{self.t} {sorted(self.entries)}>"
        if self.kind == "phase":
            return f"<phase {list(self.bits)}>"
        if self.kind in ("swap", "perm", "exchange"):
            return f"<{self.kind} {self.t},{self.aux}>"
        return f"<{self.kind} t{self.t} c{list(self.ctrls)}>"


# ---------------------------------------------------------------------------------------------------------
# 1. normalisation
# ---------------------------------------------------------------------------------------------------------
_ID = np.array([1, 0, 0, 1], dtype=complex)
_XM = np.array([0, 1, 1, 0], dtype=complex)
_HM = np.array([1, 1, 1, -1], dtype=complex) / np.sqrt(2.0)


def _mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Row-major 2x2 product a @ b (in Python complex arithmetic: numpy scalars cost ten times as much per operation)."""
    a0, a1, a2, a3 = a.tolist()
    b0, b1, b2, b3 = b.tolist()
    return np.array([a0 * b0 + a1 * b2, a0 * b1 + a1 * b3, a2 * b0 + a3 * b2, a2 * b1 + a3 * b3], dtype=complex)


def _is_diag(m) -> bool:
    return m[1] == 0 and m[2] == 0


def _is_antidiag(m) -> bool:
    return m[0] == 0 and m[3] == 0 and m[1] != 0 and m[2] != 0


def _as_matrix(p: Prim) -> np.ndarray:
    if p.kind == "mat":
        return np.asarray(p.data, dtype=complex).reshape(4)
    if p.kind == "diag":
        return np.array([p.data[0], 0, 0, p.data[1]], dtype=complex)
    return _XM


def _unit(z: complex) -> complex:
    return z / abs(z)


def rotation_matrix(theta: float) -> np.ndarray:
    """R(theta) = [[cos, -sin], [sin, cos]], row major."""
    c, sn = math.cos(theta), math.sin(theta)
    return np.array([c, -sn, sn, c], dtype=complex)


_EXACT_SHEARS = False      # set by plan() from PlanConfig.exact_rotations for the duration of one call


def shear_params(theta: float, exact: bool = False) -> Tuple[float, float, float, float, float]:
    """(z, x, y, c_all, c_bit): R(theta), |theta| <= pi/2, as the in-place updates of a LAYER record (csrc/tile.cu)

        a0 += z a1;   a1 += x a0;   a0 += y a1   (skipped when y = 0)

    followed by a factor c_all on both amplitudes and c_bit on a1, which the record's diagonal stage applies.
    |theta| <= pi/4: two updates, z = -t, x = t / (1 + t^2) with t = tan(theta), leaving (b0 / c, b1 c) with
    c = cos(theta): c_all = c, c_bit = 1 + t^2, both within [1/2, 2].  Larger angles, or `exact`: the three shears of
    the lifting scheme, z = y = -tan(theta / 2), x = sin(theta), c_all = c_bit = 1 at the price of the third update
    (cheaper than two updates and a table when the record has no other use for a table)."""
    if not exact and not _EXACT_SHEARS and abs(theta) <= math.pi / 4 + 1e-15:
        t = math.tan(theta)
        return -t, t / (1.0 + t * t), 0.0, 1.0 / math.sqrt(1.0 + t * t), 1.0 + t * t
    z = -math.tan(theta / 2.0)
    return z, math.sin(theta), z, 1.0, 1.0


def shear_item(q: int, theta: float) -> "Item":
    """The rotation by theta, |theta| <= pi/2, on index bit q as the planner schedules it: a `mat` item flagged to run
    as the in-place updates of a LAYER record."""
    it = Item("mat", t=q, m=rotation_matrix(theta))
    it.sh = float(theta)
    return it


_ANTI = np.array([0, -1, 1, 0], dtype=complex)        # R(pi / 2)


def split_rotation(m, allow_anti: bool = True) -> Optional[Tuple[complex, complex, float, complex, bool]]:
    """A 2x2 unitary as

        m = g * diag(1, dl) . A^anti . R(theta) . diag(1, dr),   |g| = |dl| = |dr| = 1,

    with R(theta) = [[cos, -sin], [sin, cos]] and A = R(pi / 2) = [[0, -1], [1, 0]]; |theta| <= pi/4 when `allow_anti`,
    else |theta| <= pi/2 and anti = False.  R(theta) runs in place on the register tile (shear_params: two FMAs per
    component of an amplitude pair for |theta| <= pi/4, three beyond, against 8 multiply-adds and a copy for a complex
    2x2).  `dr` joins the phases in front of the gate, `g` the segment's global factor, and diag(1, dl) . A^anti stays
    on the wire as a pending diagonal or anti-diagonal factor: every control and phase commutes with it or is
    conjugated by it (emit_conjugated), and the next gate on the wire absorbs it.  A real rotation keeps dl = dr = 1; a
    real reflection (H) is a rotation times Z.
    Returns (g, dl, theta, dr, anti), or None when the pieces do not reproduce `m` to rounding (not unitary)."""
    a, b, c, d = np.asarray(m, dtype=complex).reshape(4).tolist()         # (plain Python arithmetic: this runs once
    isf = math.isfinite                                                  # per gate of a circuit)
    if not (isf(a.real) and isf(a.imag) and isf(b.real) and isf(b.imag) and isf(c.real) and isf(c.imag)
            and isf(d.real) and isf(d.imag)):
        return None
    if c == 0 or b == 0:
        return None                               # diagonal matrices never get here; anything else is not unitary
    if a.imag == 0 and b.imag == 0 and c.imag == 0 and d.imag == 0:
        det = a.real * d.real - b.real * c.real
        dl = 1.0 + 0j if det > 0 else -1.0 + 0j   # m = diag(1, dl) . R
        theta = math.atan2(c.real * dl.real, a.real)
        g, dr = 1.0 + 0j, 1.0 + 0j
        if abs(theta) > math.pi / 2:               # R(theta) = -R(theta -+ pi)
            theta -= math.copysign(math.pi, theta)
            g = -1.0 + 0j
    else:
        theta = math.atan2(abs(c), abs(a))         # in [0, pi/2]
        g = _unit(a) if a != 0 else 1.0 + 0j
        dr = -_unit(b) / g
        # the phases are read off the LARGE entries (those of small ones are only known to ~eps / |entry|)
        dl = _unit(d) / (g * dr) if abs(a) >= abs(c) else _unit(c) / g
    # (exact rotations need no |theta| <= pi/4: one branch of the plan less that moves with the angles)
    anti = allow_anti and not _EXACT_SHEARS and abs(theta) > math.pi / 4 + 1e-15
    if anti:                                       # R(theta) = +-A . R(theta -+ pi/2)
        if theta < 0:
            g = -g
        theta -= math.copysign(math.pi / 2, theta)
    r00 = r11 = math.cos(theta)                      # (rotation_matrix(theta), without the array)
    r10 = math.sin(theta)
    r01 = -r10
    if anti:
        r00, r01, r10, r11 = -r10, -r11, r00, r01
    err = max(abs(g * r00 - a), abs(g * r01 * dr - b), abs(g * dl * r10 - c), abs(g * dl * r11 * dr - d))
    if not err <= 3e-15:
        return None
    return g, dl, float(theta), dr, anti


def fold_conjugated_phases(prims: Sequence[Prim]) -> List[Prim]:
    """CX(c,t) . D(t) . CX(c,t) with D = diag(d0, d1) is diagonal: D(t) . diag_c(1, r) . CD_{c->t}(1, r^-2), r = d1/d0
    (the amplitude picks up d0 * r^(t xor c)).  This is how blueqat's time evolution emits exp(-i a Z.Z)
    (pauli.py:596-630 -> cx, rz, cx; also the rzz fallback, gate.py:1142-1147): a QAOA cost layer then needs no
    non-diagonal target at all and rides along with whatever pass comes next."""
    out: List[Prim] = []
    i, n = 0, len(prims)
    while i < n:
        p = prims[i]
        if (i + 2 < n and p.kind == "x" and len(p.controls) == 1):
            d, q = prims[i + 1], prims[i + 2]
            if (d.kind == "diag" and not d.controls and d.target == p.target and q.kind == "x"
                    and q.controls == p.controls and q.target == p.target and abs(complex(d.data[0])) > 0.5):
                d0, d1 = complex(d.data[0]), complex(d.data[1])
                r = d1 / d0
                c = p.controls[0]
                out.append(d)
                out.append(Prim("diag", c, (), np.array([1.0, r]), op_index=p.op_index))
                out.append(Prim("diag", p.target, (c,), np.array([1.0, 1.0 / (r * r)]), op_index=q.op_index))
                i += 3
                continue
        out.append(p)
        i += 1
    return out


def _prim_sets(p: Prim) -> Tuple[FrozenSet[int], FrozenSet[int]]:
    """(wires on which the primitive is non-diagonal, wires on which it is diagonal)."""
    if p.kind == "swap":
        return frozenset((p.target, p.aux)), frozenset()
    if p.kind == "x" or (p.kind == "mat" and not _is_diag(_as_matrix(p))):
        return frozenset((p.target,)), frozenset(p.controls)
    return frozenset(), frozenset((p.target,) + tuple(p.controls))


def shard_schedule(prims: Sequence[Prim], n: int, n_local: int, pos0: Sequence[int]) -> List[Prim]:
    """Reorder the primitives of a sharded program into *epochs* and insert the exchanges between them.

    Within an epoch the set of global wires is fixed; every primitive whose non-diagonal target is local and
    that commutes with all primitives deferred before it runs now (two primitives commute unless they share a
    wire on which one of them is non-diagonal), the others wait.  When nothing more can run, the global wires
    the waiting primitives need first come home, in exchange for the local wires whose next non-diagonal use is
    farthest away (Belady; wires on the low index bits stay local).  A brick-pattern circuit thus runs a whole
    light-cone triangle per epoch instead of paying one exchange per global wire per layer."""
    g = n - n_local
    if g <= 0:
        return list(prims)
    pos = list(pos0)
    glob = {w for w in range(n) if pos[w] >= n_local}
    sets = [_prim_sets(p) for p in prims]
    out: List[Prim] = []
    pending = list(range(len(prims)))
    while pending:
        skipped_nd: set = set()
        skipped_dg: set = set()
        deferred: List[int] = []
        frontier: List[int] = []                      # global wires that alone keep a primitive from running
        for i in pending:
            p = prims[i]
            nd, dg = sets[i]
            ok = not (any(q in skipped_nd or q in skipped_dg for q in nd) or any(q in skipped_nd for q in dg))
            if ok and p.kind != "swap" and nd & glob:
                ok = False
                frontier += [q for q in nd & glob if q not in frontier]
            if ok:
                out.append(p)
                if p.kind == "swap":                      # relabelling: the two names trade places
                    a, b = p.target, p.aux
                    pos[a], pos[b] = pos[b], pos[a]
                    if (a in glob) != (b in glob):
                        glob ^= {a, b}
            else:
                deferred.append(i)
                skipped_nd |= nd
                skipped_dg |= dg
        pending = deferred
        if not pending:
            break
        want: List[int] = frontier[:g]
        next_use = {}
        for rank_, i in enumerate(pending):
            for q in sets[i][0]:
                next_use.setdefault(q, rank_)
        if not want:
            raise RuntimeError("sharded planner is stuck without a global wire to bring home")
        far = len(pending) + 1
        cand = [w for w in range(n) if w not in glob]
        cand.sort(key=lambda w: (pos[w] < 4, -next_use.get(w, far), -pos[w]))
        for w_in, w_out in zip(want, cand):
            out.append(Prim("exchange", w_in, aux=w_out))
            pos[w_in], pos[w_out] = pos[w_out], pos[w_in]
            glob.discard(w_in)
            glob.add(w_out)
    return out


class Normalized:
    def __init__(self):
        self.items: List[Item] = []
        self.n_prims = 0
        self.n_relabelled = 0                 # swap gates absorbed into the wire -> index-bit map
        self.n_exchanges = 0                  # global <-> local index-bit exchanges (sharded states)
        self.final_pos: List[int] = []        # wire -> index bit at the end of the segment


def normalize(prims: Sequence[Prim], n: int, fuse_1q: bool = True, fans: bool = True,
              relabel_swaps: bool = True, n_local: Optional[int] = None,
              initial_pos: Optional[Sequence[int]] = None, shears: bool = False) -> Normalized:
    """`n_local` < n: the state is sharded on its top n - n_local index bits (one shard per GPU).  A gate that is
    non-diagonal on a wire living on such a *global* bit first brings the wire home: the global bit is exchanged
    with the top local bit (an `exchange` item: half of every shard crosses NVLink), after a local transposition
    (`perm`) has parked there the local wire whose next non-diagonal use is farthest away (Belady)."""
    out = Normalized()
    n_local = n if n_local is None else n_local
    items = out.items
    pending: List[Optional[np.ndarray]] = [None] * n
    pending_count = [0] * n
    scalar = [1.0 + 0j]
    # Items carry a sortable key (major, minor).  Gates arriving in program order are appended (next major);
    # a fused 1-qubit gate, emitted only when something forces it out, is *inserted* right after the last item
    # touching its qubit -- it commutes with everything in between, and placing it early keeps later phases
    # mergeable into the fans that precede it.
    start = (-1, 0)
    last_nd = [start] * n                # key of the last item that is non-diagonal on qubit q
    last_touch = [start] * n             # key of the last item involving qubit q at all
    open_fans: Dict[int, Item] = {}      # target qubit -> most recent fan with that target
    counter = [0, 0]

    def push(it: Item, after=None):
        if after is None:
            it.pos = (counter[0], 0)
            counter[0] += 1
        else:
            counter[1] += 1
            it.pos = (after[0], after[1] + counter[1])
        items.append(it)
        for q in it.nd:
            last_nd[q] = max(last_nd[q], it.pos)
        for q in it.nd | it.dg:
            last_touch[q] = max(last_touch[q], it.pos)

    def accepts(fan: Item, bits) -> bool:
        return all(last_nd[b] < fan.pos for b in bits)

    def touch(it: Item, bits):
        for q in bits:
            last_touch[q] = max(last_touch[q], it.pos)

    def emit_phase(bits: Tuple[int, ...], w: complex, count: int, early: bool = False):
        if w == 1.0:
            return
        if len(bits) == 0:
            scalar[0] *= w
            return
        if fans and len(bits) == 1:
            j = bits[0]
            f = open_fans.get(j)
            if f is not None and accepts(f, bits):
                f.w *= w
                f.n_prims += count
                return
            f = Item("fan", t=j, w=w, entries={}, n_prims=count)
            push(f, last_touch[j] if early else None)
            open_fans[j] = f
            return
        if fans and len(bits) == 2:
            k, j = bits
            for tgt, other in ((j, k), (k, j)):
                f = open_fans.get(tgt)
                if f is not None and accepts(f, bits):
                    f.entries[other] = f.entries.get(other, 1.0 + 0j) * w
                    f.n_prims += count
                    f.refresh()
                    touch(f, bits)
                    return
            f = Item("fan", t=j, w=1.0 + 0j, entries={k: complex(w)}, n_prims=count)
            push(f)
            open_fans[j] = f
            return
        push(Item("phase", bits=bits, w=w, n_prims=count))

    def emit_conjugated(bits: Tuple[int, ...], w: complex, count: int, done: Tuple[int, ...] = ()):
        """Phase on `bits` arriving after pending anti-diagonal gates M_k = [[0, a], [b, 0]] on some of them:
        phase(B + {k}, w) * M_k = M_k * phase(B, w) * phase(B + {k}, 1 / w)."""
        for k in bits:
            if k in done:
                continue
            m = pending[k]
            if m is not None and _is_antidiag(m):
                rest = tuple(b for b in bits if b != k)
                emit_conjugated(rest, w, count, done)
                emit_conjugated(bits, 1.0 / w, 0, done + (k,))
                return
        emit_phase(bits, w, count)

    def flush(q: int, keep_diag: bool = False):
        """Emit the pending 2x2 of index bit q.  With `shears` a non-diagonal one leaves a phase behind;
        `keep_diag` lets it stay pending (the caller only needs the wire diagonal, e.g. before using it as a
        control), otherwise it follows at once."""
        m, cnt = pending[q], pending_count[q]
        pending[q], pending_count[q] = None, 0
        if m is None:
            return
        if m[1] == 0 and m[2] == 0 and abs(m[0]) > 0.5:
            if m[0] != 1.0:
                scalar[0] *= m[0]
                emit_phase((q,), m[3] / m[0], cnt, early=True)
            else:
                emit_phase((q,), m[3], cnt, early=True)
            return
        if _is_antidiag(m):
            if shears:
                # [[0, a], [b, 0]] = diag(-a, b) . R(pi / 2): a rotation like any other (three in-place shears) instead
                # of a pair swap, which only the kernel built with the general target stages executes
                it = shear_item(q, math.pi / 2)
                it.n_prims = cnt
                push(it, last_touch[q])
                scalar[0] *= -m[1]
                emit_phase((q,), -m[2] / m[1], 0, early=True)
                return
            # [[0, a], [b, 0]] = diag(a, b) . X
            push(Item("x", t=q, n_prims=cnt), last_touch[q])
            if m[1] != 1.0:
                scalar[0] *= m[1]
            emit_phase((q,), m[2] / m[1], 0, early=True)
            return
        sp = split_rotation(m, allow_anti=keep_diag) if shears else None
        if sp is None:
            push(Item("mat", t=q, m=m, n_prims=cnt), last_touch[q])
            return
        g, dl, theta, dr, anti = sp
        scalar[0] *= g
        emit_phase((q,), dr, 0, early=True)
        if theta != 0.0:
            it = shear_item(q, theta)
            it.n_prims = cnt
            push(it, last_touch[q])
            cnt = 0
        if keep_diag:
            if anti:
                pending[q], pending_count[q] = np.array([0, -1, dl, 0], dtype=complex), cnt
            elif dl != 1.0:
                pending[q], pending_count[q] = np.array([1, 0, 0, dl], dtype=complex), cnt
        else:
            emit_phase((q,), dl, cnt, early=True)

    def settle(q: int):
        """Leave at most a DIAGONAL factor pending on index bit q (before the wire is used as the control of a
        non-diagonal gate, or leaves the shard)."""
        m = pending[q]
        if m is not None and not _is_diag(m):
            flush(q)

    prims = fold_conjugated_phases(prims)
    pos = list(initial_pos) if initial_pos is not None else list(range(n))   # wire -> index bit of its data
    if n_local < n:
        prims = shard_schedule(prims, n, n_local, pos)
    wire_at = [0] * n
    for w_, p_ in enumerate(pos):
        wire_at[p_] = w_
    uses: List[List[int]] = [[] for _ in range(n)]      # per wire: prims that are non-diagonal on it
    if n_local < n:
        for i, p in enumerate(prims):
            if p.kind in ("mat", "x") and not (p.kind == "mat" and _is_diag(_as_matrix(p))):
                uses[p.target].append(i)          # (bring_home below is only a safety net after shard_schedule)
    use_ptr = [0] * n

    def move(a: int, b: int):
        """Exchange the contents (wire, pending gate) of index bits a and b."""
        wa, wb = wire_at[a], wire_at[b]
        wire_at[a], wire_at[b] = wb, wa
        pos[wa], pos[wb] = b, a
        pending[a], pending[b] = pending[b], pending[a]
        pending_count[a], pending_count[b] = pending_count[b], pending_count[a]

    def bring_home(gpos: int, now: int) -> int:
        """Make the wire on global bit `gpos` local; returns its new index bit."""
        top = n_local - 1
        best, best_next = top, -1
        for q in range(n_local):
            w = wire_at[q]
            while use_ptr[w] < len(uses[w]) and uses[w][use_ptr[w]] <= now:
                use_ptr[w] += 1
            nxt = uses[w][use_ptr[w]] if use_ptr[w] < len(uses[w]) else len(prims) + 1
            if nxt > best_next or (nxt == best_next and q == top):
                best, best_next = q, nxt
        # a pending non-diagonal 2x2 must be applied while its wire is still local
        for q in (best, top):
            settle(q)
        if best != top:
            push(Item("perm", t=best, aux=top, n_prims=0))
            move(best, top)
        push(Item("exchange", t=top, aux=gpos, n_prims=0))
        move(top, gpos)
        out.n_exchanges += 1
        return top

    for i_prim, p in enumerate(prims):
        if p.kind == "exchange":                  # from shard_schedule: wire p.target comes home, p.aux leaves
            if i_prim > 0 and prims[i_prim - 1].kind == "exchange":
                continue                          # (handled with the first exchange of its run)
            run = []
            while i_prim + len(run) < len(prims) and prims[i_prim + len(run)].kind == "exchange":
                run.append(prims[i_prim + len(run)])
            # a non-diagonal 2x2 must be applied while its wire is local: settle EVERY leaving wire first, so that the
            # exchanges of the run stay adjacent and the sharded state executes them as one remap (exchange_many)
            for e in run:
                settle(pos[e.aux])
            first_ex = None
            for e in run:
                gpos, lbit = pos[e.target], pos[e.aux]
                it = Item("exchange", t=lbit, aux=gpos, n_prims=0)
                # (one major position for the whole run: an item placed "as early as its bits allow" lands behind the
                # run, not between two of its exchanges)
                push(it, None if first_ex is None else first_ex.pos)
                first_ex = first_ex or it
                move(lbit, gpos)
                out.n_exchanges += 1
            continue
        out.n_prims += 1
        if p.kind == "swap":
            a, b = p.target, p.aux
            if relabel_swaps:
                pos[a], pos[b] = pos[b], pos[a]
                wire_at[pos[a]], wire_at[pos[b]] = a, b
                out.n_relabelled += 1
                continue
            pending[a], pending[b] = pending[b], pending[a]
            pending_count[a], pending_count[b] = pending_count[b], pending_count[a]
            push(Item("swap", t=a, aux=b))
            continue
        tgt = pos[p.target]
        if tgt >= n_local and p.kind in ("mat", "x") and not (p.kind == "mat" and _is_diag(_as_matrix(p))):
            tgt = bring_home(tgt, i_prim)
        ctrls = tuple(pos[cq] for cq in p.controls)
        if not ctrls and fuse_1q:
            g = _as_matrix(p)
            pending[tgt] = g if pending[tgt] is None else _mul(g, pending[tgt])
            pending_count[tgt] += 1
            continue
        if not ctrls:
            pending[tgt] = _as_matrix(p)
            pending_count[tgt] = 1
            flush(tgt)
            continue
        # controlled gate: a diagonal pending factor commutes with a control, anything else is flushed first
        if p.kind == "diag":
            d0, d1 = complex(p.data[0]), complex(p.data[1])
            qubits = ctrls + (tgt,)
            if d0 != 1.0 and abs(d0) <= 0.5:            # not a phase pattern (non-unitary mat1-like payload)
                for q in qubits:
                    m = pending[q]
                    if m is not None and not _is_diag(m):
                        flush(q)
                flush(tgt)
                push(Item("mat", t=tgt, ctrls=ctrls, m=np.array([d0, 0, 0, d1], dtype=complex)))
                continue
            # a pending anti-diagonal 2x2 (X, Y, X*phase) need not be forced out either: P * M = M * P' with P'
            # the phase conjugated by M, which is again a product of phases (emit_conjugated)
            for q in qubits:
                m = pending[q]
                if m is not None and not _is_diag(m) and not _is_antidiag(m):
                    flush(q, keep_diag=True)
            if d0 == 1.0:
                emit_conjugated(qubits, d1, 1)
            else:
                emit_conjugated(ctrls, d0, 1)
                emit_conjugated(qubits, d1 / d0, 0)
            continue
        if shears and p.kind == "x" and len(ctrls) == 1:
            # CX = H_t . CZ . H_t: the Hadamards merge into the 1-qubit gates around them, and the CZ is a phase
            # (a table entry or a fan entry) instead of 2^R amplitude moves per thread
            pending[tgt] = _HM if pending[tgt] is None else _mul(_HM, pending[tgt])
            pending_count[tgt] += 1
            for q in (ctrls[0], tgt):
                m = pending[q]
                if m is not None and not _is_diag(m) and not _is_antidiag(m):
                    flush(q, keep_diag=True)
            emit_conjugated((ctrls[0], tgt), -1.0 + 0j, 0)
            pending[tgt] = _HM if pending[tgt] is None else _mul(_HM, pending[tgt])
            continue
        for cq in ctrls:
            settle(cq)
        flush(tgt)
        if p.kind == "x":
            push(Item("x", t=tgt, ctrls=ctrls))
        else:
            push(Item("mat", t=tgt, ctrls=ctrls, m=np.asarray(p.data, dtype=complex).reshape(4)))
    for q in range(n):
        flush(q)
    if scalar[0] != 1.0:
        g = Item("phase", bits=(), w=scalar[0], n_prims=0)
        g.pos = (-2, 0)                  # a global factor commutes with everything: ride along with the first pass
        items.append(g)
    items.sort(key=lambda it: it.pos)
    out.final_pos = pos
    return out


def restore_involutions(final_pos: Sequence[int]) -> List[List[Tuple[int, int]]]:
    """The data of wire w sits on index bit final_pos[w] and must end on bit w.  Returns at most two lists of
    disjoint transpositions to apply in order: a bit permutation M (bit p moves to M(p)) is tau2 * tau1 with,
    on every cycle c_0 -> c_1 -> ... -> c_(k-1) -> c_0, tau1: c_i <-> c_(-i) and tau2: c_i <-> c_(1-i)."""
    n = len(final_pos)
    move = [0] * n
    for w, p in enumerate(final_pos):
        move[p] = w
    seen = [False] * n
    t1: List[Tuple[int, int]] = []
    t2: List[Tuple[int, int]] = []
    for s0 in range(n):
        if seen[s0] or move[s0] == s0:
            seen[s0] = True
            continue
        cyc = []
        q = s0
        while not seen[q]:
            seen[q] = True
            cyc.append(q)
            q = move[q]
        k = len(cyc)
        for i in range(k):
            a, b = cyc[i], cyc[(-i) % k]
            if a < b:
                t1.append((a, b))
            a, b = cyc[i], cyc[(1 - i) % k]
            if a < b:
                t2.append((a, b))
    return [t for t in (t1, t2) if t]


# ---------------------------------------------------------------------------------------------------------
# 2. scheduling
# ---------------------------------------------------------------------------------------------------------
def _cost(it: Item) -> float:
    """Rough DFMA-equivalents per amplitude."""
    if it.kind == "mat":
        if it.sh is not None:
            return 2.0 if abs(it.sh) <= math.pi / 4 + 1e-15 else 3.0
        base = 4.0 if np.all(np.asarray(it.m).imag == 0) else 8.0
        return base / (1 << len(it.ctrls))
    if it.kind == "x":
        return 0.5
    if it.kind == "phase":
        return 4.0 / (1 << len(it.bits))
    if it.kind == "fan":
        return 5.0 if it.entries else 3.0
    return 0.0


def _payload(it: Item, a: int, r: int) -> int:
    """Upper bound of the float64 payload an item adds to its pass."""
    if it.kind == "mat":
        return 8 if it.sh is None else 48          # a layer record (16) and its table (32), if it opens a new one
    if it.kind == "phase":
        return 4
    if it.kind == "fan":
        if not it.entries:
            return 2
        return 2 * (2 + len(it.entries) + 32 + (1 << max(0, a - r - 5)) + 2 * (1 << r))
    return 0


@dataclass
class Round:
    reg: List[int]                       # physical qubits held in registers (r of them)
    items: List[Item] = field(default_factory=list)


@dataclass
class Pass:
    tile: List[int]                      # physical tile bits, ascending (first c are 0..c-1)
    rounds: List[Round] = field(default_factory=list)


def _conflict(it: Item, skipped_nd: set, skipped_dg: set) -> bool:
    for q in it.nd:
        if q in skipped_nd or q in skipped_dg:
            return True
    for q in it.dg:
        if q in skipped_nd:
            return True
    return False


class PassOverflow(Exception):
    """The encoded pass does not fit the kernel's per-pass shared-memory tables."""


def build_pass(remaining: List[Item], n: int, a: int, r: int, c: int, cfg: PlanConfig,
               pay_scale: float = 1.0) -> Tuple[Pass, List[Item]]:
    """Greedy: returns the pass and the items left over (program order preserved).  `pay_scale` < 1 trusts that
    the encoder will need less payload than the per-item upper bounds (most fans end up as table entries)."""
    tile = set(range(c))
    rounds: List[Round] = []
    done = [False] * len(remaining)
    cost = 0.0
    n_fans = 0
    n_gates = 0
    payload = 0
    fan_words = 0
    limit = min(len(remaining), cfg.scan_limit)
    # per item, once per pass: (schedulable, payload bound, fan with entries, fan words, cost, mat / x, plain fan)
    info = []
    for it in remaining[:limit]:
        kind = it.kind
        is_fan = kind == "fan" and bool(it.entries)
        info.append((kind not in ("swap", "perm", "exchange"), int(_payload(it, a, r) * pay_scale), is_fan,
                     3 + len(it.entries) if is_fan else (1 if kind in ("phase", "fan") else 0), _cost(it),
                     kind in ("mat", "x"), kind == "fan" and not it.entries))
    nd_left: Dict[int, int] = {}                   # wire -> non-diagonal gates on it still to be scheduled
    for idx in range(limit):
        if info[idx][5]:
            t = remaining[idx].t
            nd_left[t] = nd_left.get(t, 0) + 1
    alive = list(range(limit))                     # the items not yet placed, in program order
    for _ in range(cfg.max_rounds):
        reg: List[int] = []
        picked: List[int] = []
        skipped_nd: set = set()
        skipped_dg: set = set()
        for idx in alive:
            it = remaining[idx]
            sched, pay, is_fan, fw, it_cost, is_nd_kind, plain_fan = info[idx]
            nd, dg = it.nd, it.dg
            # (_conflict: a gate cannot pass a skipped one it does not commute with)
            ok = sched and nd.isdisjoint(skipped_nd) and nd.isdisjoint(skipped_dg) and dg.isdisjoint(skipped_nd)
            if ok and ((is_fan and n_fans >= MAX_FANS - 1) or n_gates + 2 >= MAX_GATES - 1
                       or payload + pay > MAX_PAYLOAD - 64 or fan_words + fw > MAX_FAN_WORDS - 8):
                ok = False
            # a 1-qubit phase in front of a gate on the same wire (the `dr` of split_shears) waits for the round in
            # which that wire is a register bit: there it is a table entry, anywhere else a record of its own
            wants_reg = bool(nd) or (plain_fan and nd_left.get(it.t, 0) > 0)
            if ok and wants_reg:
                t = it.t
                new_t = t not in tile
                new_r = t not in reg
                if (new_t and len(tile) >= a) or (new_r and len(reg) >= r) or cost + it_cost > cfg.max_cost:
                    ok = False
                elif new_r and cfg.bank_pairs and t < 6 and c >= 4 and (t + 3 if t < 3 else t - 3) in reg:
                    # (index bits 0..5 are local bits 0..5 whenever bits 4 and 5 end up in the tile: a guess that only
                    # costs bank conflicts when wrong)
                    ok = False
                else:
                    if new_t:
                        tile.add(t)
                    if new_r:
                        reg.append(t)
            if ok:
                picked.append(idx)
                if is_nd_kind:
                    nd_left[it.t] -= 1
                cost += it_cost
                n_fans += is_fan
                n_gates += 2 if it.kind == "fan" else 1
                payload += pay
                fan_words += fw
            else:
                skipped_nd |= nd
                skipped_dg |= dg
                if len(skipped_nd) >= n:
                    break
        if not picked:
            break
        if not reg and rounds:
            # only phases / fans became ready: they need no register bits of their own, fold into the last round
            rounds[-1].items += [remaining[i] for i in picked]
        else:
            rounds.append(Round(reg, [remaining[i] for i in picked]))
        for i in picked:
            done[i] = True
        alive = [i for i in alive if not done[i]]
    # complete the tile with the lowest unused bits, then every round's register set with unused tile bits
    q = c
    while len(tile) < a:
        while q in tile:
            q += 1
        tile.add(q)
    tl = sorted(tile)
    for rd in rounds:
        for q in reversed(tl):
            if len(rd.reg) >= r:
                break
            if q not in rd.reg:
                rd.reg.append(q)
    return Pass(tl, rounds), [it for i, it in enumerate(remaining) if not done[i]]


BANK_PAIRS = ((0, 3), (1, 4), (2, 5))


def _thread_order(tile: List[int], reg: List[int]) -> List[int]:
    """Local bit indices: register bits first, then thread bits.  The shared-memory slot of local index l is
    l ^ ((l >> 3) & 7) (the TMA unit's 128-byte swizzle for 16-byte elements, csrc/tile.cu), so local bits 0 / 3,
    1 / 4 and 2 / 5 each move an access by 1, 2 and 4 sixteen-byte bank groups and the others by none: the three
    lowest thread bits (lane bits 0..2, the eight threads of a quarter warp) are taken one from each pair when the
    register bits leave one, which makes the quarter warp's 16-byte accesses conflict free."""
    loc = {q: i for i, q in enumerate(tile)}
    regl = [loc[q] for q in reg]
    rest = [i for i in range(len(tile)) if i not in regl]
    first: List[int] = []
    for pair in BANK_PAIRS:
        for i in pair:
            if i in rest:
                first.append(i)
                break
    rest = first + [i for i in rest if i not in first]
    return regl + rest


# ---------------------------------------------------------------------------------------------------------
# 3. encoding
# ---------------------------------------------------------------------------------------------------------
def _pack8(vals: Sequence[int]) -> Tuple[int, int]:
    v = list(vals) + [0] * (16 - len(vals))
    lo = sum((x & 0xFF) << (8 * i) for i, x in enumerate(v[:8]))
    hi = sum((x & 0xFF) << (8 * i) for i, x in enumerate(v[8:16]))
    return lo, hi


class _Group:
    """One LAYER / TAB record being assembled by encode_pass: up to one shear gate per register bit, then one
    factor per register amplitude (any diagonal on the register bits), or instead of it one fan (`diag`);
    `after` holds the records of their own that must run behind it (general 2x2, pair swaps, fans with thread or
    fixed-bit entries).  encode_pass places every item of a round in the EARLIEST group the commutation rule
    allows (a shear on register bit k strictly after everything that touches k, a diagonal no earlier than the
    last non-diagonal gate on its bits), so gates of one circuit layer share a record even when the program
    lists them interleaved with the phases between them."""

    def __init__(self, r: int):
        self.r = r
        self.sh: Dict[int, float] = {}                           # register bit -> theta
        self.par: Dict[int, Tuple[float, float, float, float, float]] = {}   # ... -> shear_params, once settled
        self.tab = np.ones(1 << r, dtype=complex)
        self.has_tab = False
        self.diag: Optional[dict] = None                         # a fan fused as the record's diag stage
        self.after: List[dict] = []
        # a fan (or phase) whose qubits other than its register-bit target are NOT register bits depends, per thread
        # and per tile, on those bits only: for up to three such *selector* bits (index bits outside the tile,
        # thread bits of the tile as 48 + local position) the table is stored once per combination
        self.sel_bits: List[int] = []
        self.sel_fans: List[Tuple[Optional[int], Optional[int], complex, Dict[int, complex]]] = []

    def mul_phase(self, regs, w: complex) -> None:
        mask = sum(1 << k for k in regs)
        for rho in range(1 << self.r):
            if rho & mask == mask:
                self.tab[rho] *= w
        self.has_tab = True

    def take_sel(self, kt: Optional[int], tsel: Optional[int], w0: complex, entries: Dict[int, complex]) -> bool:
        """amp *= w0 * prod(entries whose selector bit is set) on the amplitudes with register bit kt set (kt None:
        all of them), if the selector bit `tsel` is set (None: always)."""
        want = ([tsel] if tsel is not None else []) + list(entries)
        bits = self.sel_bits + [q for q in want if q not in self.sel_bits]
        if len(bits) > 3 or self.diag is not None:
            return False
        self.sel_bits = bits
        self.sel_fans.append((kt, tsel, complex(w0), dict(entries)))
        self.has_tab = True
        return True

    def mul_fan(self, kt: int, w0: complex, entries: Dict[int, complex]) -> None:
        for rho in range(1 << self.r):
            if (rho >> kt) & 1:
                f = complex(w0)
                for k, w in entries.items():
                    if (rho >> k) & 1:
                        f *= w
                self.tab[rho] *= f
        self.has_tab = True

    def leftover(self) -> float:
        """What the record's diagonal stage cannot put right after the in-place rotations (shear_params): the fused
        fan only reaches the amplitudes with its bit set, so c_all is left for the pass to place."""
        if self.sh and self.diag is not None:
            (p,) = self.par.values()
            return p[3]
        return 1.0

    def settle_rotations(self) -> None:
        """Choose how the record's rotations run.  With a fused fan: two updates, the fan takes c_bit (fused_fan_scale)
        and the pass c_all.  With a table for other reasons: two updates where the angle allows, c_all and c_bit go
        into the table.  Otherwise three exact updates and no table (2k + 4 against 3k FP64 per amplitude, k <= 4)."""
        if not self.sh:
            return
        if self.diag is not None:
            return                                   # (par was set when the fan was fused)
        for k, theta in self.sh.items():
            p = shear_params(theta, exact=not self.has_tab)
            self.par[k] = p
            if p[3] != 1.0 or p[4] != 1.0:
                self.tab *= p[3]
                for rho in range(1 << self.r):
                    if (rho >> k) & 1:
                        self.tab[rho] *= p[4]

    def words_kw(self, prog: Program, r: int) -> Optional[dict]:
        kw: dict = {}
        p = self.sel_bits + [63] * (3 - len(self.sel_bits))          # 63: no such selector (reads as 0)
        sel = (p[0] << 4) | (p[1] << 10)
        if self.sh:
            pay: List[complex] = []
            for k in range(4):
                z, x, y = self.par.get(k, (0.0, 0.0, 0.0))[:3]
                pay += [complex(z, x), complex(y, 0.0)]
            kw = dict(tk=T_LAYER, rm=sum(1 << k for k in self.sh) | sel, roff_t=prog.add_reals(pay))
        if self.has_tab:
            assert self.diag is None
            m = len(self.sel_bits)
            # [variant][rho], variants TAB_PITCH entries apart (csrc/tile.cu: compile-time offsets, no bank conflicts)
            size = 1 << self.r
            base = self.tab.tolist()
            pad = [1.0 + 0j] * ((TAB_PITCH if m else size) - size)
            tabs: List[complex] = []
            for var in range(1 << m):
                on = {q for i, q in enumerate(self.sel_bits) if (var >> i) & 1}
                t = list(base)
                for kt, tsel, w0, entries in self.sel_fans:
                    if tsel is not None and tsel not in on:
                        continue
                    f = complex(w0)
                    for q, w in entries.items():
                        if q in on:
                            f *= w
                    if kt is None:
                        t = [x * f for x in t]
                    else:
                        for rho in range(size):
                            if (rho >> kt) & 1:
                                t[rho] *= f
                tabs += t
                tabs += pad
            kw.update(dk=D_TAB, roff_d=prog.add_reals(tabs), fid=p[2])
            kw.setdefault("rm", sel)
        elif self.diag is not None:
            kw.update(j=self.diag["j"], dk=self.diag["dk"], fid=self.diag["fid"], roff_d=self.diag["roff_d"])
        return kw or None


def encode_pass(prog: Program, ps: Pass, n: int, r: int, c: int, direct_io: bool = False,
                direct_out: bool = False, scale_in: complex = 1.0, final: bool = True) -> complex:
    """Appends the pass to `prog`.  `scale_in` is a factor on every amplitude handed down by earlier passes (the
    cos(theta) of rotations whose record had no table to take it, _Group.leftover); it goes into one of this pass's
    tables if there is one.  Returns what is still left (1 when `final`: the pass then spends a record on it)."""
    a = len(ps.tile)
    tile = ps.tile
    loc = {q: i for i, q in enumerate(tile)}
    nt = a - r
    fan_records: List[List[int]] = []          # per fan: [n_entries | roff << 32, masks...]
    round_words: List[int] = []
    n_prims = 0

    def split_mask(qubits, order) -> Tuple[int, int, int]:
        """-> (thread-part local mask, register-part rho mask, fixed physical mask)"""
        lm = rm = fm = 0
        for q in qubits:
            if q in loc:
                li = loc[q]
                k = order.index(li)
                if k < r:
                    rm |= 1 << k
                else:
                    lm |= 1 << li
            else:
                fm |= 1 << q
        return lm, rm, fm

    scalars: List[Tuple[int, complex]] = []     # (fixed-bit mask, factor): diagonal gates constant over a tile
    pay_lo = len(prog._reals)                   # everything this pass appends to `reals` is one contiguous range

    def gate_words(tk=T_NONE, j=None, dk=D_NONE, fid=0, lm=0, rm=0, fm=0, roff_t=0, roff_d=0) -> List[int]:
        j = r if j is None else j
        return [tk | (j << 8) | (dk << 16) | (fid << 24) | (lm << 32) | (rm << 48), fm, roff_t | (roff_d << 32)]

    def fan_fields(order, fid_entries, w0, lw, tgt) -> dict:
        """One fan: fixed (mask, w) entries, local per-bit factors `lw`, optional target qubit -> gate fields."""
        fid = len(fan_records)
        lm = fm = 0
        jreg = r                                   # register bit of the target (r = none)
        if tgt is not None:
            if tgt in loc:
                k = order.index(loc[tgt])
                if k < r:
                    jreg = k
                else:
                    lm = 1 << loc[tgt]
            else:
                fm = 1 << tgt
        if not fid_entries and not lw and tgt is not None:        # a 1-qubit phase: the factor is the payload
            return dict(j=jreg, dk=D_W, fid=0, lm=lm, fm=fm, roff_d=prog.add_reals([w0]))
        fr_off = prog.add_reals([w0] + [w for _, w in fid_entries])
        fan_records.append([len(fid_entries) | (fr_off << 32)] + [m for m, _ in fid_entries])
        if not lw:                                                # constant over a tile: no tables
            return dict(j=jreg, dk=D_FIX, fid=fid, lm=lm, fm=fm, roff_d=fr_off)
        tbits = order[r:]
        lane_t = np.ones(32, dtype=complex)
        for x in range(32):
            for i in range(min(5, nt)):
                if (x >> i) & 1 and tbits[i] in lw:
                    lane_t[x] *= lw[tbits[i]]
        warp_t = np.ones(1 << max(0, nt - 5), dtype=complex)
        for y in range(warp_t.size):
            for i in range(5, nt):
                if (y >> (i - 5)) & 1 and tbits[i] in lw:
                    warp_t[y] *= lw[tbits[i]]
        reg_t = np.ones(1 << r, dtype=complex)
        for rho in range(1 << r):
            for i in range(r):
                if (rho >> i) & 1 and order[i] in lw:
                    reg_t[rho] *= lw[order[i]]
        if np.any(reg_t != 1.0):
            return dict(j=jreg, dk=D_FAN, fid=fid, lm=lm, fm=fm,
                        roff_d=prog.add_reals(list(lane_t) + list(warp_t) + list(reg_t)))
        return dict(j=jreg, dk=D_FANT, fid=fid, lm=lm, fm=fm, roff_d=prog.add_reals(list(lane_t) + list(warp_t)))

    encoded_rounds: List[Tuple[List[int], List[int]]] = []
    round_slots: List[Tuple[List[int], List[_Group]]] = []
    for rd in ps.rounds:
        order = _thread_order(tile, rd.reg)
        slots: List[_Group] = [_Group(r)]
        last_any = [-1] * r          # per register bit: last slot holding anything that touches it
        last_nd = [-1] * r           # ... last slot holding (or followed by) a non-diagonal gate on it
        diag_min = [0] * r           # ... first slot whose table / `after` list may take a diagonal on it
        open_nd: Dict[int, dict] = {}     # register bit -> general 2x2 record still open for a fan on the same bit

        def reg_index(q) -> Optional[int]:
            li = loc.get(q)
            if li is None:
                return None
            k = order.index(li)
            return k if k < r else None

        def slot(i: int) -> _Group:
            while len(slots) <= i:
                slots.append(_Group(r))
            return slots[i]

        def touch(regs, i: int):
            for k in regs:
                last_any[k] = max(last_any[k], i)
                open_nd.pop(k, None)

        def table_slot(regs) -> _Group:
            """Earliest group whose table may take a diagonal on register bits `regs`."""
            i = max([diag_min[k] for k in regs] + [0])
            while slot(i).diag is not None:
                i += 1
            touch(regs, i)
            return slots[i]

        def attach(kw: dict, regs_dg, i: int):
            """A diagonal record of its own, behind group i."""
            slot(i).after.append(kw)
            touch(regs_dg, i)

        def emit_fan(t: int, w0: complex, entries: Dict[int, complex]):
            bits = (t,) + tuple(entries)
            if not any(q in loc for q in bits):                       # constant over a tile
                scalars.append((1 << t, w0))
                scalars.extend(((1 << t) | (1 << q), w) for q, w in entries.items())
                return
            regs = [reg_index(q) for q in bits]
            kt = regs[0]
            if all(k is not None for k in regs):
                # purely on register bits: one factor per register amplitude, shared by every thread
                table_slot(regs).mul_fan(kt, w0, {reg_index(q): w for q, w in entries.items()})
                return
            if kt is None and any(k is not None for k in regs[1:]):
                # target outside the registers, some entries inside: a 2-qubit phase is symmetric, so each such entry
                # becomes a fan on ITS register bit with the old target as its only entry
                rest = {}
                for q, w in entries.items():
                    if reg_index(q) is not None:
                        emit_fan(q, 1.0 + 0j, {t: w})
                    else:
                        rest[q] = w
                if rest or w0 != 1.0:
                    emit_fan(t, w0, rest)
                return
            touched = [k for k in regs if k is not None]
            i = max([diag_min[k] for k in touched] + [0])
            g = slot(i)
            nd_kw = open_nd.get(kt) if kt is not None else None
            if nd_kw is not None and all(last_nd[k] <= last_any[kt] for k in touched):
                # general 2x2 on a register bit followed by a fan on the same bit: one record, one dispatch
                fixed = [(1 << q, w) for q, w in entries.items() if q not in loc]
                lw = {loc[q]: w for q, w in entries.items() if q in loc}
                f = fan_fields(order, fixed, w0, lw, t)
                nd_kw.update(dk=f["dk"], fid=f["fid"], roff_d=f["roff_d"])
                touch(touched, last_any[kt])
                return
            has_reg_entries = any(k is not None for k in regs[1:])
            if (kt is not None and len(g.sh) == 1 and kt in g.sh and not g.has_tab and not g.sel_fans
                    and g.diag is None and not g.after and last_any[kt] == i):
                # one rotation, on bit kt, followed at once by a fan on kt (the H + controlled-phase ladder of a QFT):
                # the fan, with its register table if it has one, is the record's diag stage, and takes the 1 + t^2 the
                # in-place rotation leaves on the amplitudes it multiplies anyway
                fixed = [(1 << q, w) for q, w in entries.items() if q not in loc]
                lw = {loc[q]: w for q, w in entries.items() if q in loc}
                g.par[kt] = shear_params(g.sh[kt])
                g.diag = fan_fields(order, fixed, w0 * g.par[kt][4], lw, t)
                touch(touched, i)
                return
            if kt is not None and has_reg_entries:
                # the factors between register bits go into a table, the rest stays a fan
                reg_part = {reg_index(q): w for q, w in entries.items() if reg_index(q) is not None}
                table_slot([kt] + list(reg_part)).mul_fan(kt, 1.0 + 0j, reg_part)
                entries = {q: w for q, w in entries.items() if reg_index(q) is None}
                touched = [kt]
                if not entries and w0 == 1.0:
                    return
            def selector(q: int) -> int:
                if q in loc:
                    return 48 + loc[q]
                assert q < 48
                return q

            if all(reg_index(q) is None for q in entries):
                # every other qubit is a thread bit or outside the tile: table variants instead of a record of its own
                ents = {selector(q): w for q, w in entries.items()}
                if kt is not None:
                    slot(diag_min[kt])
                    for i in range(diag_min[kt], len(slots)):     # (a later existing group if this one is full)
                        if slots[i].take_sel(kt, None, w0, ents):
                            touch([kt], i)
                            return
                else:
                    for g2 in slots:
                        if (g2.has_tab or g2.sh) and g2.take_sel(None, selector(t), w0, ents):
                            return
            fixed = [(1 << q, w) for q, w in entries.items() if q not in loc]
            lw = {loc[q]: w for q, w in entries.items() if q in loc}
            attach(fan_fields(order, fixed, w0, lw, t), touched, max([diag_min[k] for k in touched] + [0]))

        for it in rd.items:
            n_prims += it.n_prims
            if it.kind == "mat" and it.sh is not None:
                k = reg_index(it.t)
                assert k is not None and not it.ctrls, "non-diagonal target outside the round's register bits"
                i = last_any[k] + 1
                while slot(i).diag is not None:      # (a record whose diag stage is a fused fan holds ONE rotation: the
                    i += 1                           # fan takes that rotation's scale factor, nobody would take ours)
                slot(i).sh[k] = it.sh
                touch([k], i)
                last_nd[k] = i
                diag_min[k] = i
            elif it.kind in ("mat", "x"):
                k = reg_index(it.t)
                assert k is not None, "non-diagonal target outside the round's register bits"
                lm, rm, fm = split_mask(it.ctrls, order)
                if it.kind == "mat":
                    real = bool(np.all(np.asarray(it.m).imag == 0))
                    kw = dict(tk=T_MATR if real else T_MATC, j=k, lm=lm, rm=rm, fm=fm, roff_t=prog.add_reals(it.m))
                else:
                    kw = dict(tk=T_X, j=k, lm=lm, rm=rm, fm=fm)
                cregs = [reg_index(q) for q in it.ctrls if reg_index(q) is not None]
                i = max([last_any[k]] + [last_nd[c] for c in cregs] + [0])
                slot(i).after.append(kw)
                touch([k] + cregs, i)
                last_nd[k] = i
                diag_min[k] = i + 1
                if it.kind == "mat" and lm == 0 and rm == 0 and fm == 0:
                    open_nd[k] = kw
            elif it.kind == "phase":
                bits = it.bits
                if not any(q in loc for q in bits):
                    scalars.append((sum(1 << q for q in bits), it.w))
                    continue
                regs = [reg_index(q) for q in bits]
                if all(k is not None for k in regs):
                    table_slot(regs).mul_phase(regs, it.w)
                    continue
                lm, rm, fm = split_mask(bits, order)
                touched = [k for k in regs if k is not None]
                attach(dict(dk=D_PHASE, lm=lm, rm=rm, fm=fm, roff_d=prog.add_reals([it.w])), touched,
                       max([diag_min[k] for k in touched] + [0]))
            else:
                emit_fan(it.t, it.w, dict(it.entries))
        round_slots.append((order, slots))
    # the rotations' scale factors: into the record's own table, else handed from record to record (and pass to pass)
    # until a table takes them
    left = complex(scale_in)
    for order, slots in round_slots:
        for g in slots:
            g.settle_rotations()
            left *= g.leftover()
    if left != 1.0:
        host = next((g for _, slots in round_slots for g in slots if g.has_tab), None)
        if host is not None:
            host.tab = host.tab * left
            left = 1.0 + 0j
        elif final or not 2.0 ** -40 < abs(left) < 2.0 ** 40:
            scalars.append((0, left))
            left = 1.0 + 0j
    for order, slots in round_slots:
        gates: List[int] = []
        for g in slots:
            kw = g.words_kw(prog, r)
            if kw:
                gates += gate_words(**kw)
            for kw in g.after:
                gates += gate_words(**kw)
        encoded_rounds.append((order, gates))
    if scalars:
        order, gates = encoded_rounds[-1]
        gates += gate_words(**fan_fields(order, scalars, 1.0 + 0j, {}, None))
    for order, gates in encoded_rounds:
        o_lo, o_hi = _pack8(order)
        round_words += [(len(gates) // 3) | ((3 + len(gates)) << 16), o_lo, o_hi] + gates
    t_lo, t_hi = _pack8(tile)
    n_fans = len(fan_records)
    directory: List[int] = []
    body: List[int] = []
    pay_len = len(prog._reals) - pay_lo
    if (pay_len > MAX_PAYLOAD or sum(len(g) // 3 for _, g in encoded_rounds) > MAX_GATES or n_fans > MAX_FANS
            or n_fans + sum(len(rec) for rec in fan_records) > MAX_FAN_WORDS):
        del prog._reals[pay_lo:]                  # nothing else of `prog` has been touched yet
        raise PassOverflow(f"payload {pay_len}, fans {n_fans}")
    base = 5 + n_fans
    for rec in fan_records:
        directory.append(base + len(body))
        body += rec
    rounds_off = base + len(body)
    # first / last round may bypass shared memory when none of the row bits (local 0..c-1) is a register bit:
    # the thread order then maps lanes 0..2^c-1 onto one contiguous row (see _thread_order)
    def _direct(order):
        return c >= 3 and all(b >= c for b in order[:r]) and order[r:r + c] == list(range(c))
    flags = 0
    if direct_io and _direct(encoded_rounds[0][0]):
        flags |= 1
    if (direct_io or direct_out) and _direct(encoded_rounds[-1][0]):
        flags |= 2
    words = [a | (c << 8) | (r << 16) | (len(ps.rounds) << 24) | (n_fans << 32) | (flags << 40), t_lo, t_hi,
             5 | (rounds_off << 32), pay_lo | (pay_len << 32)] + directory + body + round_words
    woff = prog.add_words(words)
    prog.add_op(_lib.OP_TILE, woff=woff, wlen=len(words), nbytes=32.0 * 2 ** n)
    prog.n_prims += n_prims
    return left


def _emit_single(prog: Program, it: Item, n: int) -> None:
    """One item as its own kernel launch (tiny states, swaps, and the fusion-off mode)."""
    dim = 2.0 ** n
    if it.kind == "mat":
        prog.add_op(_lib.OP_MAT, it.t, cmask=sum(1 << q for q in it.ctrls), roff=prog.add_reals(it.m),
                    nbytes=32.0 * dim / (1 << len(it.ctrls)))
    elif it.kind == "x":
        prog.add_op(_lib.OP_X, it.t, cmask=sum(1 << q for q in it.ctrls), nbytes=32.0 * dim / (1 << len(it.ctrls)))
    elif it.kind == "swap":
        prog.add_op(_lib.OP_SWAP, it.t, aux=it.aux, nbytes=16.0 * dim)
    elif it.kind == "phase":
        if not it.bits:
            prog.add_op(_lib.OP_SCALE, roff=prog.add_reals([it.w]), nbytes=32.0 * dim)
        else:
            prog.add_op(_lib.OP_DIAG, it.bits[-1], cmask=sum(1 << q for q in it.bits[:-1]),
                        roff=prog.add_reals([1.0, it.w]), nbytes=32.0 * dim / (1 << len(it.bits)))
    else:
        if it.w != 1.0:
            prog.add_op(_lib.OP_DIAG, it.t, roff=prog.add_reals([1.0, it.w]), nbytes=16.0 * dim)
        for q, w in it.entries.items():
            prog.add_op(_lib.OP_DIAG, it.t, cmask=1 << q, roff=prog.add_reals([1.0, w]), nbytes=8.0 * dim)
    prog.n_prims += it.n_prims


def cancelling_layout(prims: Sequence[Prim], n: int) -> List[int]:
    """The wire -> index-bit layout to START from so that the relabelling done by the program's swap gates ends on
    the identity (no restore pass).  Only meaningful when the initial state is |0...0>, which every layout
    represents equally: a QFT then runs on a bit-reversed register and its final swaps cost nothing at all."""
    perm = list(range(n))                       # perm[w]: the original position index now labelled w
    for p in prims:
        if p.kind == "swap":
            perm[p.target], perm[p.aux] = perm[p.aux], perm[p.target]
    pos0 = [0] * n
    for w, v in enumerate(perm):                # final[w] = pos0[perm[w]] must equal w
        pos0[v] = w
    return pos0


def plan(prims: Sequence[Prim], n_qubits: int, cfg: PlanConfig = DEFAULT, n_local: Optional[int] = None,
         initial_pos: Optional[Sequence[int]] = None, free_initial_layout: bool = False,
         on_pass: Optional[Callable[[Program], None]] = None) -> Program:
    """See _plan (this wrapper only scopes PlanConfig.exact_rotations)."""
    global _EXACT_SHEARS
    saved, _EXACT_SHEARS = _EXACT_SHEARS, bool(cfg.exact_rotations)
    try:
        return _plan(prims, n_qubits, cfg, n_local, initial_pos, free_initial_layout, on_pass)
    finally:
        _EXACT_SHEARS = saved


def _plan(prims: Sequence[Prim], n_qubits: int, cfg: PlanConfig = DEFAULT, n_local: Optional[int] = None,
          initial_pos: Optional[Sequence[int]] = None, free_initial_layout: bool = False,
          on_pass: Optional[Callable[[Program], None]] = None) -> Program:
    """Primitives of one unitary segment -> op program.

    `on_pass(prog)` is called whenever ops have been appended to `prog` (after every encoded pass): the backend uses
    it to launch a pass while the next one is still being planned (B200Backend, b200_apply_program_ex).

    Single GPU (n_local None): the program leaves every wire on its own index bit (relabelled swaps are paid back
    by at most two PERMUTE passes at the end).  Sharded (n_local < n_qubits): index bits >= n_local are the shard
    number; the program contains EXCHANGE ops for the host layer to run over NVLink, starts from `initial_pos`
    and ends in the layout `program.final_pos` (wire -> index bit), which is *not* restored."""
    n = n_qubits
    sharded = n_local is not None
    nl = n_local if sharded else n
    if free_initial_layout and initial_pos is None and prims:
        # the run starts from |0...0> (B200_OP_INIT): any layout is free, pick the one that cancels the swaps
        pos0 = cancelling_layout(prims, n)
        if pos0 != list(range(n)):
            initial_pos = pos0
    prog = Program()
    prog.initial_pos = list(initial_pos) if initial_pos is not None else list(range(n))
    prog.final_pos = list(prog.initial_pos)
    if not prims:
        return prog
    norm = normalize(prims, n, n_local=nl, initial_pos=initial_pos,
                     shears=cfg.shears and (sharded or n >= cfg.min_qubits))
    prog.final_pos = list(norm.final_pos)
    if not sharded and n < cfg.min_qubits:
        for it in norm.items:
            _emit_single(prog, it, n)
        _emit_restore(prog, norm, n)
        prog.final_pos = list(range(n))
        if on_pass is not None:
            on_pass(prog)
        return prog
    a = min(cfg.a, nl)
    if a not in cfg.kernel_tiles:
        fits = [t for t in cfg.kernel_tiles if t <= a]
        if not fits:
            raise ValueError(f"a sharded register needs at least {min(cfg.kernel_tiles)} local qubits per GPU for the fused "
                             f"tile pass, got {nl} (n_qubits = {n}): use fewer ranks or a single B200Backend")
        a = max(fits)
    r = min(cfg.r, a)
    c = min(cfg.c, a if a == nl else a - 1)       # a tile smaller than the register needs at least one free bit
    remaining = list(norm.items)
    carry = 1.0 + 0j                      # factor on every amplitude that no table has taken yet (encode_pass)
    while remaining:
        head = remaining[0]
        if head.kind == "swap":
            _emit_single(prog, remaining.pop(0), n)
            continue
        if head.kind == "perm":
            remaining.pop(0)
            woff = prog.add_words([min(head.t, head.aux) | (max(head.t, head.aux) << 8)])
            prog.add_op(_lib.OP_PERMUTE, woff=woff, wlen=1, nbytes=16.0 * 2 ** nl)
            continue
        if head.kind == "exchange":
            remaining.pop(0)
            prog.add_op(_lib.OP_EXCHANGE, head.t, aux=head.aux, nbytes=16.0 * 2 ** nl)
            prog.n_exchanges += 1
            continue
        for scale in (0.4, 0.7, 1.0):             # optimistic payload estimate first, the safe bound last
            ps, rest = build_pass(remaining, nl, a, r, c, cfg, pay_scale=scale)
            if not ps.rounds:                               # cannot happen: the head item always fits
                raise RuntimeError(f"planner made no progress at {remaining[0]}")
            last = all(it.kind in ("swap", "perm", "exchange") for it in rest)
            try:
                carry = encode_pass(prog, ps, nl, r, c, cfg.direct_io, cfg.direct_out, scale_in=carry, final=last)
                break
            except PassOverflow:
                if scale == 1.0:
                    raise
        remaining = rest
        if on_pass is not None:
            on_pass(prog)
    if sharded:
        prog.n_prims += norm.n_relabelled
    else:
        _emit_restore(prog, norm, n)
        prog.final_pos = list(range(n))
    if on_pass is not None:
        on_pass(prog)
    return prog


def _emit_restore(prog: Program, norm: Normalized, n: int) -> None:
    for pairs in restore_involutions(norm.final_pos):
        woff = prog.add_words([p | (q << 8) for p, q in pairs])
        prog.add_op(_lib.OP_PERMUTE, woff=woff, wlen=len(pairs), nbytes=32.0 * 2 ** n)
    prog.n_prims += norm.n_relabelled


def describe_tile_pass(words, woff: int, n: int, reals) -> dict:
    """What one encoded TILE pass asks of csrc/tile.cu: rounds, records by kind, and the FP64 instructions per amplitude
    the kernel executes for it (bench.py's FP64 roofline, tools/plan_stats.py)."""
    h0 = words[woff]
    a, r, n_rounds = h0 & 0xFF, (h0 >> 16) & 0xFF, (h0 >> 24) & 0xFF
    tile = [((words[woff + 1 + i // 8]) >> (8 * (i % 8))) & 0xFF for i in range(a)]
    runs, prev = 0, None
    for q in tile[3:]:
        runs += prev is None or q != prev + 1
        prev = q
    pos = woff + (words[woff + 3] >> 32)
    per = 1 << r
    out = dict(a=a, rounds=n_rounds, runs=runs, records=0, layers=0, rotations=0, rot3=0, tabs=0, fans=0, other=0, fp64=0.0,
               conflict_rounds=0, tile=tile)
    for _ in range(n_rounds):
        ng, nw = words[pos] & 0xFFFF, words[pos] >> 16
        order = [((words[pos + 1 + i // 8]) >> (8 * (i % 8))) & 0xFF for i in range(a)]
        lanes = order[r:r + 3]
        classes = {(b % 3) for b in lanes if b < 6}
        if len(classes) < 3:
            out["conflict_rounds"] += 1
        for g in range(ng):
            g0 = words[pos + 3 + 3 * g]
            tk, j, dk, rm = g0 & 0xFF, (g0 >> 8) & 0xFF, (g0 >> 16) & 0xFF, (g0 >> 48) & 0xFFFF
            out["records"] += 1
            f = 0.0
            if tk == T_LAYER:
                k = bin(rm & 15).count("1")
                out["layers"] += 1
                out["rotations"] += k
                roff = words[pos + 5 + 3 * g] & 0xFFFFFFFF
                for b in range(4):                # 4 or 6 DFMA per pair (third shear only when y != 0)
                    if (rm >> b) & 1:
                        three = reals[roff + 4 * b + 2] != 0.0
                        out["rot3"] += three
                        f += 3.0 if three else 2.0
            elif tk in (T_MATC, T_MATR):
                f += 8.0 if tk == T_MATC else 4.0
                out["other"] += 1
            elif tk == T_X:
                out["other"] += 1
            if dk == D_TAB:
                out["tabs"] += 1
                f += 4.0
            elif dk in (D_FAN, D_FANT, D_W, D_FIX):
                out["fans"] += 1
                frac = 0.5 if j < r else 1.0
                f += frac * (8.0 if dk == D_FAN else 4.0) + 8.0 / per
            elif dk == D_PHASE:
                f += 4.0 / (1 << bin(rm).count("1"))
            out["fp64"] += f
        pos += nw
    return out
