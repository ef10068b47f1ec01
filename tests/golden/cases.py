"""Golden-case definitions shared by make_golden.py (runs the real blueqat) and the tests.

A case is pure data so that it can be replayed on blueqat's ``Circuit`` and on
``qaqarot_b200.circuit.Circuit`` alike and committed as JSON:

    {"name": str, "n": int, "ops": [[gate, targets, params, options], ...],
     "initial_seed": int | None, "run": {"shots": int | None, "seed": int | None, "returns": str | None}}

``targets`` is JSON: an int, ``{"s": [start, stop, step]}`` for a slice, or a list of those for a tuple.
``params`` may contain ``{"mat": [[re, im] x 4]}`` for mat1.
"""
from __future__ import annotations

import math

import numpy as np

ONE_Q_PLAIN = ["x", "y", "z", "h", "s", "sdg", "t", "tdg", "sx", "sxdg", "i"]
ONE_Q_ANGLE = ["rx", "ry", "rz", "phase"]
TWO_Q_PLAIN = ["cx", "cz", "cy", "ch", "swap", "zz"]
TWO_Q_ANGLE = ["crx", "cry", "crz", "cphase", "rzz", "rxx", "ryy"]
THREE_Q = ["ccx", "ccz", "cswap"]


def enc_targets(t):
    if isinstance(t, slice):
        return {"s": [t.start, t.stop, t.step]}
    if isinstance(t, tuple):
        return [enc_targets(x) for x in t]
    return int(t)


def dec_targets(t):
    if isinstance(t, dict):
        return slice(*t["s"])
    if isinstance(t, list):
        return tuple(dec_targets(x) for x in t)
    return t


def enc_params(ps):
    out = []
    for p in ps:
        if isinstance(p, np.ndarray):
            out.append({"mat": [[float(z.real), float(z.imag)] for z in p.reshape(-1)]})
        else:
            out.append(float(p))
    return out


def dec_params(ps):
    out = []
    for p in ps:
        if isinstance(p, dict):
            out.append(np.array([complex(a, b) for a, b in p["mat"]]).reshape(2, 2))
        else:
            out.append(p)
    return out


def build(circuit_cls, case):
    """Replay a case on any Circuit class with blueqat's builder syntax."""
    c = circuit_cls(case["n"])
    for gate, targets, params, options in case["ops"]:
        w = getattr(c, gate)
        ps = dec_params(params)
        if ps or options:
            w = w(*ps, **(options or {}))
        w[dec_targets(targets)]
    return c


def initial_state(case):
    if case.get("initial_seed") is None:
        return None
    rng = np.random.default_rng(case["initial_seed"])
    v = rng.standard_normal(2 ** case["n"]) + 1j * rng.standard_normal(2 ** case["n"])
    return v / np.linalg.norm(v)


def random_unitary(rng):
    m = rng.standard_normal((2, 2)) + 1j * rng.standard_normal((2, 2))
    q, r = np.linalg.qr(m)
    return q * (np.diag(r) / abs(np.diag(r)))


def _op(gate, targets, params=(), options=None):
    return [gate, enc_targets(targets), enc_params(params), options]


def random_ops(rng, n, n_ops, allow_slices=True):
    ops = []
    for _ in range(n_ops):
        kind = rng.integers(0, 100)
        if kind < 30:
            g = ONE_Q_PLAIN[rng.integers(len(ONE_Q_PLAIN))]
            t = int(rng.integers(n))
            if allow_slices and rng.integers(8) == 0:
                t = slice(None) if rng.integers(2) else slice(int(rng.integers(n)), None, int(rng.integers(1, 3)))
            ops.append(_op(g, t))
        elif kind < 50:
            g = ONE_Q_ANGLE[rng.integers(len(ONE_Q_ANGLE))]
            ops.append(_op(g, int(rng.integers(n)), [rng.uniform(-2 * math.pi, 2 * math.pi)]))
        elif kind < 56:
            ps = list(rng.uniform(-math.pi, math.pi, 3 + int(rng.integers(2))))
            ops.append(_op("u", int(rng.integers(n)), ps))
        elif kind < 60:
            ops.append(_op("mat1", int(rng.integers(n)), [random_unitary(rng)]))
        elif n >= 2 and kind < 76:
            g = TWO_Q_PLAIN[rng.integers(len(TWO_Q_PLAIN))]
            a, b = (int(x) for x in rng.choice(n, 2, replace=False))
            ops.append(_op(g, (a, b)))
        elif n >= 2 and kind < 90:
            g = TWO_Q_ANGLE[rng.integers(len(TWO_Q_ANGLE))]
            a, b = (int(x) for x in rng.choice(n, 2, replace=False))
            ops.append(_op(g, (a, b), [rng.uniform(-2 * math.pi, 2 * math.pi)]))
        elif n >= 2 and kind < 94:
            a, b = (int(x) for x in rng.choice(n, 2, replace=False))
            ps = list(rng.uniform(-math.pi, math.pi, 3 + int(rng.integers(2))))
            ops.append(_op("cu", (a, b), ps))
        elif n >= 3:
            g = THREE_Q[rng.integers(len(THREE_Q))]
            a, b, c = (int(x) for x in rng.choice(n, 3, replace=False))
            ops.append(_op(g, (a, b, c)))
        else:
            ops.append(_op("h", int(rng.integers(n))))
    return ops


def random_layer_ops(n, depth, seed=1234):
    """SURVEY 8(d) C1/C4/C5 generator: H on all, RZ(theta ~ U[0, 2pi)) on all, brick CX."""
    rng = np.random.default_rng(seed)
    ops = []
    for d in range(depth):
        for q in range(n):
            ops.append(_op("h", q))
        for q in range(n):
            ops.append(_op("rz", q, [rng.uniform(0, 2 * math.pi)]))
        for q in range(d % 2, n - 1, 2):
            ops.append(_op("cx", (q, q + 1)))
    return ops


def qft_ops(n, x):
    """basis state x, then the textbook QFT with final bit-reversal swaps."""
    ops = [_op("x", q) for q in range(n) if (x >> q) & 1]
    for j in range(n - 1, -1, -1):
        ops.append(_op("h", j))
        for k in range(j - 1, -1, -1):
            ops.append(_op("cphase", (k, j), [math.pi / 2 ** (j - k)]))
    for q in range(n // 2):
        ops.append(_op("swap", (q, n - 1 - q)))
    return ops


def ring_edges(n):
    """Circulant C_n(1, n/2): (i, i+1 mod n) and (i, i+n/2) -- the MaxCut graph of SURVEY 8(d) C3."""
    e = [(i, (i + 1) % n) for i in range(n)]
    e += [(i, i + n // 2) for i in range(n // 2)]
    return e


def qaoa_ops(n, p=4, seed=0):
    """What vqe.QaoaAnsatz(sum Z_i Z_j, p).get_circuit(params) emits (vqe.py:130-144, pauli.py:596-630):
    H on all; per step, for every edge (i<j): cx[i,j] rz(-2*gamma)[j] cx[i,j]; then rx(beta) on all."""
    params = np.random.default_rng(seed).uniform(0, 1, 2 * p)
    betas, gammas = params[:p] * math.pi, params[p:] * 2 * math.pi
    edges = sorted(tuple(sorted(e)) for e in ring_edges(n))
    ops = [_op("h", slice(None))]
    for beta, gamma in zip(betas, gammas):
        for i, j in edges:
            ops += [_op("cx", (i, j)), _op("rz", j, [-2 * 1.0 * gamma]), _op("cx", (i, j))]
        ops.append(_op("rx", slice(None), [beta]))
    return ops, edges


def all_cases():
    cases = []

    def add(name, n, ops, initial_seed=None, shots=None, seed=None, returns=None, **extra):
        cases.append({"name": name, "n": n, "ops": ops, "initial_seed": initial_seed,
                      "run": {"shots": shots, "seed": seed, "returns": returns}, **extra})

    rng = np.random.default_rng(20261019)
    n = 5
    # every gate on every qubit position of a random 5-qubit state
    for g in ONE_Q_PLAIN:
        add(f"g1_{g}", n, [_op(g, t) for t in range(n)], initial_seed=11)
    for g in ONE_Q_ANGLE:
        add(f"g1_{g}", n, [_op(g, t, [rng.uniform(-6, 6)]) for t in range(n)], initial_seed=12)
    add("g1_u3", n, [_op("u", t, list(rng.uniform(-3, 3, 3))) for t in range(n)], initial_seed=13)
    add("g1_u4", n, [_op("u", t, list(rng.uniform(-3, 3, 4))) for t in range(n)], initial_seed=13)
    add("g1_mat1", n, [_op("mat1", t, [random_unitary(rng)]) for t in range(n)], initial_seed=14)
    pairs = [(a, b) for a in range(n) for b in range(n) if a != b]
    for g in TWO_Q_PLAIN:
        add(f"g2_{g}", n, [_op(g, p) for p in pairs], initial_seed=15)
    for g in TWO_Q_ANGLE:
        add(f"g2_{g}", n, [_op(g, p, [rng.uniform(-6, 6)]) for p in pairs], initial_seed=16)
    add("g2_cu", n, [_op("cu", p, list(rng.uniform(-3, 3, 4))) for p in pairs], initial_seed=17)
    triples = [(0, 1, 2), (2, 0, 4), (4, 3, 1), (1, 4, 0), (3, 2, 4), (4, 0, 3)]
    for g in THREE_Q:
        add(f"g3_{g}", n, [_op(g, t) for t in triples], initial_seed=18)
    # slices
    add("slice_all", 6, [_op("h", slice(None)), _op("rz", slice(1, None, 2), [0.7]),
                         _op("cx", (slice(0, 3), slice(3, 6))), _op("x", (0, -1)),
                         _op("cphase", (slice(None, None, 2), slice(1, None, 2)), [0.3]),
                         _op("ry", (slice(2), 5), [1.1]), _op("s", slice(None, None, -2))])
    # random circuits of growing size, from |0..0> and from random states
    for k, (nq, nops) in enumerate([(1, 20), (2, 40), (3, 60), (4, 80), (5, 100), (6, 120), (7, 120),
                                    (8, 150), (9, 150), (10, 200), (11, 200), (12, 250)]):
        add(f"rand_{nq}q", nq, random_ops(rng, nq, nops), initial_seed=None if k % 2 == 0 else 100 + k)
    # the three benchmark shapes at small size
    add("c1_layers_12q", 12, random_layer_ops(12, 20))
    add("c1_layers_14q", 14, random_layer_ops(14, 20))
    add("c2_qft_10q", 10, qft_ops(10, 123456789 % 1024))
    add("c2_qft_13q", 13, qft_ops(13, 123456789 % 8192))
    qops, edges = qaoa_ops(10)
    add("c3_qaoa_10q", 10, qops, zz_edges=edges)
    qops, edges = qaoa_ops(12)
    add("c3_qaoa_12q", 12, qops, zz_edges=edges)
    # measurement / shots: terminal, partial, scrambled order, mid-circuit, reset, keys
    m_all = _op("m", slice(None))
    add("shots_c1_10q", 10, random_layer_ops(10, 6) + [m_all], shots=300, seed=7)
    add("shots_scrambled", 8, random_ops(rng, 8, 60, False) + [_op("m", (5, 2, 7, 0, 3))], shots=200, seed=3)
    add("shots_partial", 6, random_ops(rng, 6, 40, False) + [_op("m", slice(1, None, 2))], shots=200, seed=4)
    add("shots_midcircuit", 5, random_ops(rng, 5, 25, False) + [_op("m", 2)] + random_ops(rng, 5, 25, False)
        + [_op("m", (0, 4))], shots=150, seed=5)
    add("shots_reset", 5, random_ops(rng, 5, 25, False) + [_op("reset", (1, 3))] + random_ops(rng, 5, 10, False)
        + [m_all], shots=150, seed=6)
    add("shots_nomeasure", 3, [_op("h", 0)], shots=10, seed=1)
    add("shots_twice", 4, [_op("h", slice(None)), m_all, _op("h", slice(None)), m_all], shots=100, seed=8)
    add("sv_and_shots", 4, random_ops(rng, 4, 30, False) + [_op("m", (1, 2))], shots=50, seed=9,
        returns="statevector_and_shots")
    add("sv_after_measure", 5, random_ops(rng, 5, 30, False) + [_op("m", (0, 3))], shots=None, seed=10)
    add("samples_keys", 4, [_op("h", slice(None)), _op("m", (0, 1), (), {"key": "a"}), _op("cx", (0, 2)),
                            _op("m", 2, (), {"key": "b"}), _op("m", 3, (), {"key": "a", "duplicated": "append"})],
        shots=40, seed=11, returns="samples")
    add("shots_initial", 5, random_ops(rng, 5, 20, False) + [m_all], initial_seed=77, shots=100, seed=12)
    return cases


def expectation_cases():
    """[(case name, [(coeff, [(letter, qubit), ...]), ...])] evaluated on that case's final state."""
    zz10 = [(1.0, [("Z", i), ("Z", j)]) for i, j in qaoa_ops(10)[1]]
    zz12 = [(1.0, [("Z", i), ("Z", j)]) for i, j in qaoa_ops(12)[1]]
    mixed = [(0.5, [("X", 0), ("Y", 3)]), (-1.25, [("Z", 1)]), (0.75, [("Y", 2), ("Y", 4), ("Z", 0)]),
             (2.0, []), (0.3, [("X", 5), ("X", 6), ("Z", 7)]), (-0.4, [("Y", 7)]), (1.5, [("X", 0), ("Y", 3)])]
    return [("c3_qaoa_10q", zz10), ("c3_qaoa_12q", zz12), ("rand_8q", mixed), ("rand_9q", mixed),
            ("c1_layers_12q", mixed)]
