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
