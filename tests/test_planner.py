"""The fusion planner and the tile-pass encoding, executed by the numpy descriptor interpreter (no GPU)."""
import random
import warnings

import numpy as np
import pytest

import cases as C
from conftest import case_names
from helpers import TOL, maxabs, run_case
from oracle import sv_oracle as O
from oracle.program_interp import OracleState
from qaqarot_b200 import B200Backend, Circuit, planner as P, workloads as W
from qaqarot_b200 import _lib
from qaqarot_b200.lowering import lower

ALL = tuple(range(1, 20))
SMALL = [P.PlanConfig(a=5, r=2, c=2, min_qubits=5, kernel_tiles=ALL),
         P.PlanConfig(a=4, r=2, c=1, min_qubits=4, kernel_tiles=ALL, max_rounds=2),
         P.PlanConfig(a=6, r=3, c=0, min_qubits=6, kernel_tiles=ALL, max_cost=20.0),
         P.PlanConfig(a=7, r=4, c=3, min_qubits=7, kernel_tiles=ALL)]


def interp_backend(cfg):
    return B200Backend(state_factory=lambda n, d: OracleState(n, d), plan_config=cfg)


@pytest.mark.parametrize("cfg", SMALL, ids=lambda c: f"a{c.a}r{c.r}c{c.c}")
@pytest.mark.parametrize("name", case_names("statevector"))
def test_statevector_cases_through_tiles(name, cfg, gold):
    assert maxabs(run_case(gold.cases[name], interp_backend(cfg)), gold.sv(name)) <= TOL


@pytest.mark.parametrize("cfg", SMALL[:2], ids=lambda c: f"a{c.a}r{c.r}c{c.c}")
@pytest.mark.parametrize("name", case_names("shots"))
def test_shot_cases_through_tiles(name, cfg, gold):
    assert dict(run_case(gold.cases[name], interp_backend(cfg))) == gold.meta["cases"][name]["shots"]


@pytest.mark.parametrize("a", [9, 10, 11, 12, 13])
def test_kernel_tile_sizes(a):
    """The tile sizes the CUDA kernel instantiates, r = 4, on circuits just large enough to need them."""
    n = a + 1
    cfg = P.PlanConfig(a=a)
    rng = np.random.default_rng(a)
    ops = C.random_ops(rng, n, 120, True)
    circ = C.build(Circuit, {"n": n, "ops": ops})
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = O.OracleBackend().run(circ.ops, n)
    be = interp_backend(cfg)
    assert maxabs(circ.run(backend=be), ref) <= TOL
    kinds = [o[0] for o in be._compiled.prefix._ops]
    assert _lib.OP_TILE in kinds


def test_qft_becomes_fans():
    n = 16
    low = lower(W.qft(n).ops, n)
    norm = P.normalize(low.prims, n)
    kinds = [it.kind for it in norm.items]
    # one 2x2 per qubit (X merged into H), one fan per qubit that has lower neighbours; the swaps only relabel
    assert kinds.count("mat") == n and kinds.count("fan") == n - 1 and kinds.count("swap") == 0
    assert norm.n_relabelled == n // 2 and norm.final_pos == list(range(n))[::-1]
    assert sum(it.n_prims for it in norm.items) + norm.n_relabelled == len(low.prims)
    prog = P.plan(low.prims, n)
    kinds = [o[0] for o in prog._ops]
    assert kinds.count(_lib.OP_TILE) <= 3 and kinds.count(_lib.OP_PERMUTE) == 1 and len(kinds) <= 4
    assert prog.n_prims == len(low.prims)


def test_restore_involutions_compose_to_the_permutation():
    rng = np.random.default_rng(5)
    for n in (1, 2, 5, 9, 16):
        for _ in range(20):
            final_pos = [int(x) for x in rng.permutation(n)]
            bits = list(range(n))                      # bits[p] = wire whose data sits on index bit p
            for w, p_ in enumerate(final_pos):
                bits[p_] = w
            invs = P.restore_involutions(final_pos)
            assert len(invs) <= 2
            for pairs in invs:
                flat = [q for pq in pairs for q in pq]
                assert len(set(flat)) == len(flat)
                for a, b in pairs:
                    bits[a], bits[b] = bits[b], bits[a]
            assert bits == list(range(n))


def test_swaps_are_relabelled_not_executed():
    n = 10
    c = Circuit(n).h[:].swap[0, 7].cx[0, 3].rz(0.3)[7].swap[2, 7].h[2].cz[7, 1].swap[0, 1].swap[5, 6].swap[5, 6]
    ref = O.OracleBackend().run(c.ops, n)
    # from |0...0> the planner starts on the layout that cancels the swaps: nothing is ever moved
    be = interp_backend(P.PlanConfig(a=9))
    assert maxabs(c.run(backend=be), ref) <= TOL
    kinds = [o[0] for o in be._compiled.prefix._ops]
    assert _lib.OP_SWAP not in kinds and _lib.OP_PERMUTE not in kinds
    assert be._compiled.prefix.initial_pos != list(range(n)) and be._compiled.prefix.final_pos == list(range(n))
    # from a caller-supplied state the layout is given: the swaps are paid back by in-place involutions at the end
    rng = np.random.default_rng(3)
    init = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    init /= np.linalg.norm(init)
    be = interp_backend(P.PlanConfig(a=9))
    assert maxabs(c.run(backend=be, initial=init), O.OracleBackend().run(c.ops, n, initial=init)) <= TOL
    kinds = [o[0] for o in be._compiled.prefix._ops]
    assert _lib.OP_SWAP not in kinds and _lib.OP_PERMUTE in kinds


def test_cancelling_layout_undoes_the_swap_relabelling():
    rng = np.random.default_rng(9)
    for n in (2, 5, 12):
        c = Circuit(n)
        for _ in range(15):
            a, b = (int(x) for x in rng.choice(n, 2, replace=False))
            c.swap[a, b].h[a]
        low = lower(c.ops, n)
        pos0 = P.cancelling_layout(low.prims, n)
        norm = P.normalize(low.prims, n, initial_pos=pos0)
        assert norm.final_pos == list(range(n))


def test_one_qubit_chains_collapse():
    c = Circuit(10)
    for q in range(10):
        c.h[q].rz(0.1 * q)[q].rx(0.3)[q].t[q].x[q]
    low = lower(c.ops, 10)
    norm = P.normalize(low.prims, 10)
    assert len(norm.items) == 10 and all(it.kind == "mat" for it in norm.items)
    prog = P.plan(low.prims, 10)
    assert prog.n_passes == 1


def test_diagonal_pending_commutes_with_controls():
    """rz on a control may be delayed past the controlled gate; a non-diagonal gate may not."""
    c = Circuit(9).h[:].rz(0.4)[0].cx[0, 1].rz(0.3)[0].h[0].cx[1, 0]
    be = interp_backend(P.PlanConfig(a=9))
    ref = O.OracleBackend().run(c.ops, 9)
    assert maxabs(c.run(backend=be), ref) <= TOL


def test_non_unitary_mat1_is_not_turned_into_a_phase():
    proj = np.array([[0, 0], [0, 1]], dtype=complex)
    c = Circuit(9).h[:].mat1(proj)[3].cz[3, 4].mat1(proj)[3]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = O.OracleBackend().run(c.ops, 9)
    assert maxabs(c.run(backend=interp_backend(P.PlanConfig(a=9))), ref) <= TOL


def test_pass_bytes_accounting():
    n = 14
    low = lower(W.random_layers(n, 4).ops, n)
    prog = P.plan(low.prims, n)
    assert prog.pass_bytes == pytest.approx(prog.n_passes * 32.0 * 2 ** n)
    assert len(prog.op_bytes) == prog.n_passes


# ---- 1-qubit unitaries as phase . rotation-by-shears . phase, layer / table records ------------------------------
def _random_unitary(rng):
    z = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    q, _ = np.linalg.qr(z)
    return q.reshape(4)


def test_split_rotation_reproduces_unitaries_to_rounding():
    rng = np.random.default_rng(5)
    h = np.array([1, 1, 1, -1], dtype=complex) / np.sqrt(2)
    mats = [_random_unitary(rng) for _ in range(2000)] + [h, -h, 1j * h]
    for t in np.linspace(-7, 7, 58):                        # real rotations / reflections, near-diagonal and near-X
        c, s = np.cos(t), np.sin(t)
        mats += [np.array([c, -s, s, c], dtype=complex), np.array([c, s, s, -c], dtype=complex)]
    for t in 10.0 ** rng.uniform(-9, 0, 200):               # tiny off-diagonal entries with arbitrary phases
        a, b = rng.uniform(0, 6, 2)
        mats.append(np.array([np.cos(t), -np.exp(1j * a) * np.sin(t), np.exp(1j * b) * np.sin(t),
                              np.exp(1j * (a + b)) * np.cos(t)]) * np.exp(0.3j))
    for allow_anti in (True, False):
        n_anti = 0
        for m in mats:
            sp = P.split_rotation(m, allow_anti)
            assert sp is not None
            g, dl, theta, dr, anti = sp
            assert abs(abs(g) - 1) < 1e-15 and abs(abs(dl) - 1) < 1e-15 and abs(abs(dr) - 1) < 1e-15
            assert abs(theta) <= (np.pi / 4 if allow_anti else np.pi / 2) + 1e-12
            z, x, y, c_all, c_bit = P.shear_params(theta)
            assert max(abs(z), abs(x), abs(y)) <= 1.0 + 1e-12 and 0.5 - 1e-12 <= c_all <= 1 and 1 <= c_bit <= 2 + 1e-12
            assert (y == 0.0) == (abs(theta) <= np.pi / 4 + 1e-15)
            # the in-place updates followed by the two scale factors are the rotation
            a0, a1 = np.array([1.0, 0.0]), np.array([0.0, 1.0])                         # rows of the accumulated matrix
            a0 = a0 + z * a1
            a1 = a1 + x * a0
            if y:
                a0 = a0 + y * a1
            sm = np.array([a0[0] * c_all, a0[1] * c_all, a1[0] * c_all * c_bit, a1[1] * c_all * c_bit])
            assert np.max(np.abs(sm - P.rotation_matrix(theta))) <= 1e-15
            rm = P.rotation_matrix(theta)
            if anti:
                rm = P._mul(P._ANTI, rm)
                n_anti += 1
            rec = g * np.array([rm[0], rm[1] * dr, dl * rm[2], dl * rm[3] * dr])
            assert np.max(np.abs(rec - m)) <= 3e-15
            if np.all(m.imag == 0) and m[0] * m[3] - m[1] * m[2] > 0:
                assert dl == 1 and dr == 1                        # a real rotation needs no phases at all
        assert (0 < n_anti < len(mats)) if allow_anti else n_anti == 0
    assert P.split_rotation(h)[4] is False                    # H is exactly R(pi / 4) . Z: no anti-diagonal left over


def test_split_rotation_rejects_non_unitary_matrices():
    assert P.split_rotation(np.array([1, 1, 0, 1], dtype=complex)) is None
    assert P.split_rotation(np.array([1, 0.5, 0.5, 1], dtype=complex)) is None
    assert P.split_rotation(np.array([2, 1j, 1j, 2], dtype=complex) / 2) is None


def _records(prog):
    """(target kind, diag kind) of every gate slot of every tile pass."""
    out = []
    for op in prog._ops:
        if op[0] != _lib.OP_TILE:
            continue
        w = prog._words[op[5]:op[5] + op[3]]
        pos = w[3] >> 32
        for _ in range((w[0] >> 24) & 0xFF):
            for g in range(w[pos] & 0xFFFF):
                g0 = w[pos + 3 + 3 * g]
                out.append((g0 & 0xFF, (g0 >> 16) & 0xFF))
            pos += w[pos] >> 16
    return out


def test_layered_circuit_becomes_layer_and_table_records():
    """H / RZ / CX brick layers: every 1-qubit gate runs as shears inside a LAYER record, every CX as H . CZ . H
    with the CZ in a table or a fan -- no pair swap and no general 2x2 is left (the pass then runs on the kernel
    built without those stages, csrc/tile.cu k_tile<..., FULL = false>)."""
    n = 16
    low = lower(W.random_layers(n, 6).ops, n)
    prog = P.plan(low.prims, n)
    recs = _records(prog)
    kinds = {tk for tk, _ in recs}
    assert P.T_LAYER in kinds and not kinds & {P.T_MATC, P.T_MATR, P.T_X}
    assert any(dk == P.D_TAB for _, dk in recs)
    n_1q = sum(1 for p in low.prims if p.kind in ("mat", "diag") and not p.controls)
    assert sum(1 for tk, _ in recs if tk == P.T_LAYER) < n_1q // 2        # several gates per record
    base = P.plan(low.prims, n, P.PlanConfig(shears=False))
    assert {tk for tk, _ in _records(base)} & {P.T_MATC, P.T_X}            # the general path is still there
    ref = O.OracleBackend().run(W.random_layers(n, 6).ops, n)
    for cfg in (P.PlanConfig(), P.PlanConfig(shears=False)):
        assert maxabs(W.random_layers(n, 6).run(backend=interp_backend(cfg)), ref) <= TOL


def test_cx_chain_through_hadamard_rewrite():
    """GHZ ladder + rotations around the CX gates: the H . CZ . H rewrite keeps parity."""
    n = 10
    c = Circuit(n).h[0]
    for q in range(n - 1):
        c.ry(0.3 + q)[q].cx[q, q + 1].rx(0.2 * q)[q + 1]
    c.cx[n - 1, 0].cx[0, n - 1]
    ref = O.OracleBackend().run(c.ops, n)
    for cfg in SMALL + [P.PlanConfig(a=9)]:
        assert maxabs(c.run(backend=interp_backend(cfg)), ref) <= TOL


def test_pass_overflow_falls_back_to_the_safe_payload_bound(monkeypatch):
    """The optimistic payload estimate may overshoot the kernel's table: the pass is rebuilt, not truncated."""
    n = 14
    circ = W.random_layers(n, 10)
    low = lower(circ.ops, n)
    monkeypatch.setattr(P, "MAX_PAYLOAD", 400)
    prog = P.plan(low.prims, n, P.PlanConfig(a=10))
    for op in prog._ops:
        if op[0] == _lib.OP_TILE:
            assert (prog._words[op[5] + 4] >> 32) <= 400
    st = OracleState(n, 0)
    st.apply(prog)
    assert maxabs(st.download(), O.OracleBackend().run(circ.ops, n)) <= TOL


def test_table_variants_absorb_phases_on_thread_and_fixed_bits():
    """A CZ-like phase between a register bit and a thread bit / a bit outside the tile, and a 1-qubit phase on a
    thread bit, become variants of the layer record's table (selectors 48 + local bit / physical bit) instead of
    records of their own.  A hand-made round (register bits 4..7 of a 2^9 tile, 12 qubits) pins the encoding."""
    from qaqarot_b200.program import Program
    n, tile, reg = 12, list(range(9)), [4, 5, 6, 7]
    rng = np.random.default_rng(11)
    items, want_ops = [], []
    for q in reg:
        t = 1.5 - 0.5 * q                               # two- and three-update rotations
        items.append(P.shear_item(q, t))
        want_ops.append(("mat", q, P.rotation_matrix(t)))
    w = [np.exp(1j * a) for a in (0.4, 1.1, -0.8, 2.0)]
    items += [P.Item("fan", t=4, w=1.0, entries={0: w[0]}),        # register bit x thread bit
              P.Item("fan", t=11, w=1.0, entries={6: w[1]}),       # target outside the tile, entry on a register bit
              P.Item("fan", t=1, w=w[2], entries={}),              # 1-qubit phase on a thread bit
              P.Item("fan", t=5, w=w[3], entries={6: -1.0 + 0j})]  # register-local
    prog = Program()
    P.encode_pass(prog, P.Pass(tile, [P.Round(reg, items)]), n, 4, 4, False, True)
    recs = _records(prog)
    assert recs == [(P.T_LAYER, P.D_TAB)], recs
    g0 = prog._words[(prog._words[3] >> 32) + 3]
    sels = {((g0 >> 48) >> 4) & 63, ((g0 >> 48) >> 10) & 63, (g0 >> 24) & 0xFF}
    assert sels == {48 + 0, 11, 48 + 1}, sels
    psi = rng.standard_normal(2 ** n) + 1j * rng.standard_normal(2 ** n)
    st = OracleState(n, 0)
    st.upload(psi)
    st.apply(prog)
    idx = np.arange(2 ** n)
    ref = psi.copy()
    for _, q, m in want_ops:
        i0 = idx[(idx >> q) & 1 == 0]
        a0, a1 = ref[i0], ref[i0 | (1 << q)]
        ref[i0], ref[i0 | (1 << q)] = m[0] * a0 + m[1] * a1, m[2] * a0 + m[3] * a1
    bit = lambda q: (idx >> q) & 1 == 1
    ref[bit(4) & bit(0)] *= w[0]
    ref[bit(11) & bit(6)] *= w[1]
    ref[bit(1)] *= w[2]
    ref[bit(5)] *= w[3]
    ref[bit(5) & bit(6)] *= -1
    assert maxabs(st.download(), ref) <= 1e-13


def test_sharded_plan_with_too_few_local_qubits_fails_with_a_clear_message():
    """8 qubits on 4 ranks leave 6 local qubits: no tile size fits (the single-GPU path would fall back to one kernel per
    gate; a sharded plan cannot)."""
    from qaqarot_b200 import planner as P
    from qaqarot_b200.lowering import lower
    from qaqarot_b200 import Circuit
    c = Circuit(8).h[:].cx[0, 7]
    with pytest.raises(ValueError, match="local qubits"):
        P.plan(lower(list(c.ops), 8).prims, 8, P.DEFAULT, n_local=6)
