"""Host-side logic on a CPU: lowering, program encoding, shot descent, cache handling, kwargs.

The device is replaced by oracle/program_interp.OracleState, an independent numpy executor of the encoded op
program, so these tests pin everything above the C ABI against the reference fixtures without a GPU.
"""
import random
import sys
import warnings
from collections import Counter

import numpy as np
import pytest

from conftest import case_names
from helpers import TOL, cpu_backend, maxabs, run_case
from qaqarot_b200 import Circuit, X, Y, Z
from qaqarot_b200.pauli import canonical_terms

MODES = [False, True]


@pytest.mark.parametrize("fusion", MODES)
@pytest.mark.parametrize("name", case_names("statevector"))
def test_statevector_cases(name, fusion, gold):
    out = run_case(gold.cases[name], cpu_backend(fusion))
    assert maxabs(out, gold.sv(name)) <= TOL


@pytest.mark.parametrize("fusion", MODES)
@pytest.mark.parametrize("name", case_names("shots"))
def test_shot_cases(name, fusion, gold):
    out = run_case(gold.cases[name], cpu_backend(fusion))
    assert dict(out) == gold.meta["cases"][name]["shots"]


@pytest.mark.parametrize("fusion", MODES)
def test_statevector_and_shots(fusion, gold):
    sv, cnt = run_case(gold.cases["sv_and_shots"], cpu_backend(fusion))
    assert dict(cnt) == gold.meta["cases"]["sv_and_shots"]["shots"]
    assert maxabs(sv, gold.sv("sv_and_shots")) <= TOL


def test_samples(gold):
    assert run_case(gold.cases["samples_keys"], cpu_backend()) == gold.meta["cases"]["samples_keys"]["samples"]


@pytest.mark.parametrize("fusion", MODES)
def test_hamiltonian_kwarg(fusion, gold):
    letter = {"X": X, "Y": Y, "Z": Z}
    for name, recs in gold.meta["expect"].items():
        for rec in recs:
            h = 0
            for coeff, letters in rec["terms"]:
                t = coeff
                for ch, q in letters:
                    t = t * letter[ch][q]
                h = h + t
            import cases as C
            circ = C.build(Circuit, gold.cases[name])
            init = C.initial_state(gold.cases[name])
            kw = {} if init is None else {"initial": init}
            val = circ.run(backend=cpu_backend(fusion), hamiltonian=h, **kw)
            assert isinstance(val, float)
            assert abs(val - rec["value"]) < 1e-10


def test_readme_known_answers():
    """README.md:83-89 and :72-76 of the reference."""
    assert Circuit(4).x[:].run(backend=cpu_backend(), hamiltonian=1 * Z[0] + 1 * Z[1]) == pytest.approx(-2.0)
    assert Circuit(20).x[:].m[:].run(backend=cpu_backend(), shots=100) == Counter({"1" * 20: 100})


def test_run_argument_handling():
    be = cpu_backend()
    c = Circuit().h[0].cx[0, 1]
    with pytest.raises(ValueError):
        c.run(backend=be, returns="nope")
    with pytest.warns(UserWarning):
        c.run(backend=be, foo=1)
    with pytest.warns(UserWarning):
        c.run(backend=be, shots=5, returns="statevector")
    with pytest.raises(ValueError):
        c.run(backend=be, initial=[1, 0, 0, 0])
    with pytest.raises(ValueError):
        c.run(backend=be, initial=np.array([1, 0]))
    assert c.run(backend=be, shots=7) == Counter({"00": 7})
    assert sum(Circuit().h[0].m[0].run(backend=be, returns="shots").values()) == 1024
    assert np.allclose(Circuit().run(backend=be), [1])
    assert Circuit().run(backend=be, shots=3) == Counter()
    v, s = Circuit().x[1].m[:].oneshot(backend=be)
    assert s == "01" and np.allclose(v, [0, 0, 1, 0])
    assert np.allclose(Circuit().x[0].statevector(backend=be), [0, 1])
    assert Circuit().x[0].m[0].shots(10, backend=be) == Counter({"1": 10})


def test_unknown_operation():
    class Weird:
        lowername = "weird"
        targets = 0
    with pytest.raises(ValueError):
        cpu_backend().run([Weird()], 1)


def test_cache_semantics():
    """save_cache / make_cache resume from the pre-measurement snapshot (numpy_backend.py:564-568, 610-626)."""
    be = cpu_backend()
    c = Circuit().h[0].cx[0, 1].m[:]
    c._backends["b200"] = be
    assert be.cache is None and be.cache_idx == -1
    c.make_cache(backend="b200")
    assert be.cache is not None and be.cache_idx == 1
    random.seed(3)
    got = c.run(backend="b200", shots=50)
    assert set(got) <= {"00", "11"} and sum(got.values()) == 50
    # appending after caching keeps working (tests/test_circuit.py: test_cache_then_append)
    c.x[1].m[1]
    got = c.run(backend="b200", shots=20)
    assert set(got) <= {"01", "10"} and be.cache is not None
    c2 = c.copy()
    assert c2._backends["b200"].cache is not None and c2._backends["b200"] is not be
    # a different qubit count invalidates the snapshot
    d = Circuit().x[0].m[0]
    d._backends["b200"] = be
    assert d.run(backend="b200", shots=3) == Counter({"1": 3})
    assert be.cache is None


def test_initial_disables_cache():
    be = cpu_backend()
    c = Circuit().h[0].m[0]
    init = np.array([0, 1], dtype=complex)
    with pytest.warns(UserWarning):
        c.run(backend=be, initial=init, save_cache=True, shots=2)
    assert be.cache is None


def test_ignore_global():
    be = cpu_backend()
    v = Circuit().x[0].rz(1.0)[0].run(backend=be, ignore_global=True)
    assert np.allclose(v, [0, 1])


def test_canonical_terms_accepts_mirror_algebra():
    h = 2 * X[0] * Y[1] - Z[2] + 0.5 * X[0] * X[0]
    terms = {tuple(l): c for c, l in canonical_terms(h)}
    assert terms[(("X", 0), ("Y", 1))] == 2
    assert terms[(("Z", 2),)] == -1
    assert terms[()] == 0.5


@pytest.mark.reference
def test_blueqat_objects_are_accepted():
    """The same backend instance consumes genuine blueqat operations and pauli expressions."""
    import sys
    sys.path.insert(0, "/root/reference")
    warnings.simplefilter("ignore", SyntaxWarning)
    import blueqat
    from blueqat import pauli
    import qaqarot_b200
    qaqarot_b200.register()
    assert "b200" in blueqat.backends.BACKENDS
    c = blueqat.Circuit().h[0].cx[0, 1].rz(0.3)[1].ccx[0, 1, 2].swap[0, 2].rxx(0.2)[0, 1].m[:]
    be = cpu_backend()
    random.seed(11)
    ref = c.run(backend="numpy", shots=40)
    random.seed(11)
    assert c.run(backend=be, shots=40) == ref
    c2 = blueqat.Circuit().h[:3].cx[0, 1].ry(0.4)[2].cz[1, 2]
    assert maxabs(c2.run(backend=be), c2.run(backend="numpy")) <= TOL
    h = 1.5 * pauli.Z[0] * pauli.Z[1] + pauli.X[2] - 0.3 * pauli.Y[0] * pauli.X[1]
    from blueqat import vqe
    ref_e = vqe.sparse_expectation(h.to_expr().simplify().to_matrix(3, sparse="csc"), c2.run(backend="numpy"))
    assert abs(c2.run(backend=be, hamiltonian=h) - ref_e) < 1e-12


def test_plan_cache_is_keyed_on_gate_values():
    """A circuit rebuilt with the same angles (vqe.py:67-83 builds one per objective call) reuses the plan; a changed
    angle, target or measurement key plans again and gives the new circuit's result."""
    from qaqarot_b200 import workloads as W
    be = cpu_backend()
    a, _, _ = W.qaoa_maxcut(8, seed=3)
    b, _, _ = W.qaoa_maxcut(8, seed=3)
    c, _, _ = W.qaoa_maxcut(8, seed=4)
    assert all(x is not y for x, y in zip(a.ops, b.ops))
    va = a.run(backend=be)
    plan_a = be._compiled
    vb = b.run(backend=be)
    assert be._compiled is plan_a and np.array_equal(va, vb)
    vc = c.run(backend=be)
    assert be._compiled is not plan_a
    from oracle import sv_oracle as O
    assert np.abs(vc - O.OracleBackend().run(c.ops, 8)).max() <= 1e-12 and np.abs(vc - va).max() > 1e-3
    d = Circuit(3).h[0].rz(0.25)[1].m[0]
    e = Circuit(3).h[0].rz(0.25)[2].m[0]
    d.run(backend=be, shots=2)
    plan_d = be._compiled
    e.run(backend=be, shots=2)
    assert be._compiled is not plan_d


def test_newly_planned_circuits_are_launched_pass_by_pass():
    """a circuit that has to be planned goes to the device in pieces while the planner works
    (planner.plan's on_pass -> apply_piece); a piece may only read the pool tails appended for it (the interpreter poisons
    the rest).  Same result as planning first and launching afterwards; the cached plan of a second call goes in one go."""
    from qaqarot_b200 import workloads as W
    from helpers import cpu_backend as interp_backend
    n = 13
    calls = []
    streamed, plain = interp_backend(True), interp_backend(True)
    plain.stream_plans = False
    streamed.param_replay = plain.param_replay = False      # (the speculative launch has its own test below)
    st = streamed._state(n)
    orig = st.apply_piece

    def spy(*a, **kw):
        calls.append(a[1:3])
        return orig(*a, **kw)
    st.apply_piece = spy
    for seed in range(3):
        circ, ham, _ = W.qaoa_maxcut(n, p=3, seed=seed)
        a = circ.run(backend=streamed)
        b = circ.run(backend=plain)
        assert np.abs(a - b).max() <= 1e-15
        n_pieces = len(calls)
        assert n_pieces >= 2 and calls[0][0] == 0 and all(x[1] == y[0] for x, y in zip(calls, calls[1:]))
        assert abs(circ.run(backend=streamed, hamiltonian=ham) - circ.run(backend=plain, hamiltonian=ham)) <= 1e-12
        assert len(calls) == n_pieces            # same gates by value: the cached plan, one apply()
        del calls[:]


def test_objective_loop_launches_from_the_planner_model_and_checks_it():
    """SURVEY 8(f) rank 1, param_replay: the third call in a row with the same gate structure and other angles builds a model of
    the planner's payload; from then on the circuit is launched with the predicted payload before it is planned, and the
    real plan only confirms it.  Energies against the oracle; a model that predicts wrongly costs a second run, never a
    wrong answer."""
    from qaqarot_b200 import workloads as W
    from qaqarot_b200 import param_replay as R
    from helpers import cpu_backend
    from oracle import sv_oracle as O
    n = 12
    be = cpu_backend(True)
    for seed in range(5):
        circ, ham, edges = W.qaoa_maxcut(n, p=3, seed=seed)
        psi = O.OracleBackend().run(circ.ops, n)
        want = O.expectation([(1.0, [("Z", i), ("Z", j)]) for i, j in edges], psi)
        assert abs(circ.run(backend=be, hamiltonian=ham) - want) < 1e-10
        assert maxabs(circ.run(backend=be), psi) <= 1e-12               # (same angles: the cached plan)
    assert be.replay_stats == {"built": 1, "hits": 3, "misses": 0, "unmodelable": 0}
    # a model that is wrong: the real plan disagrees, the circuit runs again, the model is dropped
    (model,) = [m for m in be._models.values() if m is not None]
    model.E0_live = model.E0_live * np.exp(0.3j)
    circ, ham, edges = W.qaoa_maxcut(n, p=3, seed=77)
    psi = O.OracleBackend().run(circ.ops, n)
    assert maxabs(circ.run(backend=be), psi) <= 1e-12
    assert be.replay_stats["misses"] == 1 and all(m is None for m in be._models.values())
    # a structure whose payload is not of the model's form (a general 2x2 that moves with the angle) is left alone
    from qaqarot_b200 import Circuit
    be2 = cpu_backend(True)
    for th in (0.3, 0.9, 1.4, 2.0):
        c = Circuit(n).h[:].u(th, 0.2, 0.5)[3].cx[3, 4].rx(th)[:]
        assert maxabs(c.run(backend=be2), O.OracleBackend().run(c.ops, n)) <= 1e-12
    assert be2.replay_stats["hits"] == 0


def test_objective_loop_that_walks_out_of_the_models_range_gets_a_new_model():
    """A ParamModel holds within a quarter turn of the rotations it was built at.  An optimiser that moves an rx angle past
    pi leaves that range: the model declines (the backend plans first, as always), and after three such calls in a row a
    new model is built around the angles the loop has reached.  Every statevector against the oracle."""
    from qaqarot_b200 import Circuit, param_replay as R, workloads as W
    from helpers import cpu_backend
    from oracle import sv_oracle as O
    n = 12
    edges = W.ring_edges(n)

    def ansatz(beta, gamma):
        c = Circuit(n).h[:]
        for i, j in edges:
            c.cx[i, j].rz(-2.0 * gamma)[j].cx[i, j]
        return c.rx(beta)[:]
    be = cpu_backend(True)
    stats = []
    for k, (beta, gamma) in enumerate([(0.5, 0.3), (0.6, 0.35), (0.7, 0.4), (0.8, 0.45),         # built on the third, hit
                                       (3.5, 0.5), (3.6, 0.55), (3.7, 0.6),                     # out of range: declined
                                       (3.8, 0.65), (3.9, 0.7)]):                               # re-centred, hit
        c = ansatz(beta, gamma)
        assert maxabs(c.run(backend=be), O.OracleBackend().run(c.ops, n)) <= 1e-12, k
        stats.append((be.replay_stats["built"], be.replay_stats["hits"]))
    assert be.replay_stats["misses"] == 0 and be.replay_stats["unmodelable"] == 0
    assert stats == [(0, 0), (0, 0), (1, 1), (1, 2), (1, 2), (1, 2), (1, 2), (2, 3), (2, 4)], stats
    (model,) = [m for m in be._models.values() if m is not None]
    assert model.generation == 1 and model.declined == 0


@pytest.mark.reference
def test_objective_loop_with_blueqats_own_qaoa_ansatz():
    """The reference's own call path (vqe.py:130-144 QaoaAnsatz.get_circuit, rebuilt for every angle vector): genuine
    blueqat gate objects go through structure_key / with_angles (Operation.create), the model is built on the third call
    and confirmed on every later one; statevectors against blueqat's numpy backend."""
    import warnings
    from conftest import REFERENCE
    sys.path.insert(0, REFERENCE)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            from blueqat import vqe
            from blueqat.pauli import Z as BZ
    finally:
        sys.path.remove(REFERENCE)
    from helpers import cpu_backend
    n = 12
    ham = sum((BZ[i] * BZ[(i + 1) % n] for i in range(1, n)), BZ[0] * BZ[1])
    ansatz = vqe.QaoaAnsatz(ham, 2)
    be = cpu_backend(True)
    rng = np.random.default_rng(3)
    for _ in range(5):
        circ = ansatz.get_circuit(rng.uniform(0, 1, 4))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = circ.run(backend="numpy")
        assert maxabs(circ.run(backend=be), want) <= 1e-12
    assert be.replay_stats["built"] == 1 and be.replay_stats["hits"] == 3 and be.replay_stats["misses"] == 0


def test_param_model_declines_what_it_cannot_describe():
    """ParamModel.evaluate returns None -- the backend then plans first, as always -- when gates that shared an angle no
    longer do, or when a rotation leaves the range in which its slot is affine in the gate angle (rx(beta) beyond pi: the
    planner's angle turns around there); structure_key ignores angle values and nothing else."""
    import dataclasses
    from qaqarot_b200 import param_replay as R, planner as P, workloads as W
    from qaqarot_b200.lowering import lower
    n = 12
    cfg = dataclasses.replace(P.DEFAULT, exact_rotations=True)

    def plan_prefix(gates):
        low = lower(tuple(gates), n, 0)
        return P.plan(low.prims, n, cfg, free_initial_layout=True)
    c1, c2 = W.qaoa_maxcut(n, p=2, seed=1)[0], W.qaoa_maxcut(n, p=2, seed=2)[0]
    assert R.structure_key(c1.ops, n) == R.structure_key(c2.ops, n)
    assert R.structure_key(c1.ops, n) != R.structure_key(W.qaoa_maxcut(n, p=3, seed=1)[0].ops, n)
    model = R.ParamModel(list(c2.ops), R.angle_values(c1.ops), plan_prefix)
    assert len(model.members) == 4                                       # 2 betas, 2 gammas
    v = R.angle_values(c2.ops)
    assert np.abs(model.evaluate(v) - np.asarray(plan_prefix(c2.ops)._reals)).max() <= 1e-13
    w = v.copy()
    w[model.members[0][0]] += 0.1                                        # one gate of a group goes its own way
    assert model.evaluate(w) is None
    for g, ix in enumerate(model.members):                               # an rx angle beyond pi
        w = v.copy()
        w[ix] = 3.5
        got = model.evaluate(w)
        if got is None:
            break
    else:
        raise AssertionError("no group is a rotation angle")
    # the rebuilt gates are new objects with the new angles, the others are shared
    gates = R.with_angles(c2.ops, v + 0.25)
    assert np.allclose(R.angle_values(gates), v + 0.25) and gates[0] is c2.ops[0]
