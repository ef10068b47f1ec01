"""Pin the CPU oracle (oracle/sv_oracle.py) against the real reference.

(a) committed golden fixtures generated from blueqat's NumPyBackend (tests/golden/make_golden.py);
(b) the reference's own golden vector and analytic cases (tests/test_circuit.py), restated;
(c) the live reference, when /root/reference exists (build container).
"""
import math
import random
import warnings
from collections import Counter

import numpy as np
import pytest

import cases as C
from conftest import case_names, golden, REFERENCE
from oracle import sv_oracle as O
from qaqarot_b200.circuit import Circuit


def run_oracle(case, **over):
    circ = C.build(Circuit, case)
    run = dict(case["run"])
    run.update(over)
    kw = {}
    init = C.initial_state(case)
    if init is not None:
        kw["initial"] = init
    if run["shots"] is not None:
        kw["shots"] = run["shots"]
    if run["returns"] is not None:
        kw["returns"] = run["returns"]
    if run["seed"] is not None:
        random.seed(run["seed"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return O.OracleBackend().run(circ.ops, circ.n_qubits, **kw)


@pytest.mark.parametrize("name", case_names("statevector"))
def test_statevector_matches_reference_fixture(name, gold):
    out = run_oracle(gold.cases[name])
    # same arithmetic in the same order: agreement is at the rounding level, far below 1e-12
    assert np.abs(out - gold.sv(name)).max() <= 1e-15


@pytest.mark.parametrize("name", case_names("shots"))
def test_shots_match_reference_fixture(name, gold):
    out = run_oracle(gold.cases[name])
    assert dict(out) == gold.meta["cases"][name]["shots"]


def test_statevector_and_shots(gold):
    sv, cnt = run_oracle(gold.cases["sv_and_shots"])
    assert dict(cnt) == gold.meta["cases"]["sv_and_shots"]["shots"]
    assert np.abs(sv - gold.sv("sv_and_shots")).max() <= 1e-15


def test_samples(gold):
    assert run_oracle(gold.cases["samples_keys"]) == gold.meta["cases"]["samples_keys"]["samples"]


def test_expectation_matrix_route(gold):
    for name, recs in gold.meta["expect"].items():
        psi = gold.sv(name)
        for rec in recs:
            terms = [(c, [(l, q) for l, q in ls]) for c, ls in rec["terms"]]
            assert abs(O.expectation(terms, psi) - rec["value"]) < 1e-12


def test_reference_golden_vector():
    """tests/test_circuit.py:697-747 (test_complicated_circuit), 16 amplitudes to 9 digits."""
    hp = 1.5707963267948966
    a, b, c_, d = 0.095491506289, 1.1856905316303521e-08, 1.2142490216037756e-08, math.pi / 2
    c = Circuit()
    c.x[0].h[0].rx(-hp)[2].cx[0, 2].rz(a)[2]
    c.cx[0, 2].h[0].rx(hp)[2].h[0].ry(-hp)[2]
    c.cx[0, 2].cx[2, 3].rz(a)[3].cx[2, 3].cx[0, 2].h[0]
    c.rx(hp)[2].h[0].ry(-hp)[2].cx[0, 1]
    c.cx[1, 2].rz(a)[2].cx[1, 2].cx[0, 1].h[0].u(d, 1.58, -0.62)[2]
    c.h[0].rx(-hp)[2].cx[0, 1].cx[1, 2].cx[2, 3]
    c.rz(a)[3].cx[2, 3].cx[1, 2].cx[0, 1].h[0]
    c.rx(hp)[2].u(0.42, -hp, 1.64)[2].h[2]
    c.cx[0, 2].rz(-a)[2].cx[0, 2].rx(hp)[0].h[2]
    c.rx(-hp)[0].h[2].cx[0, 2].cx[2, 3].rz(-a)[3]
    c.cx[2, 3].cx[0, 2].rx(hp)[0].h[2]
    c.rx(-hp)[0].h[2].cx[0, 1].cx[1, 2].rz(-a)[2]
    c.cx[1, 2].cx[0, 1].rx(hp)[0].h[2]
    c.rx(-hp)[0].t[2].s[2].cx[0, 1].cx[1, 2].sdg[1].cx[2, 3]
    c.rz(-a)[3].cx[2, 3].cx[1, 2].cx[0, 1]
    c.rx(hp)[0].h[2].h[0].rx(-hp)[1].h[2]
    c.cx[0, 1].cx[1, 2].rz(b)[2].cx[1, 2].cx[0, 1].h[0]
    c.rx(hp)[1].h[2].rx(-hp)[0]
    c.rx(-hp)[1].rx(-hp)[2].cx[0, 1].cx[1, 2]
    c.rz(b)[2].cx[1, 2].cx[0, 1].rx(hp)[0]
    c.rx(hp)[1].rx(hp)[2]
    c.rx(-hp)[1].cx[1, 3].rz(c_)[3]
    c.cx[1, 3].rx(hp)[1].rx(-hp)[1].cx[0, 1]
    c.cx[1, 2].rz(-c_)[2].cx[1, 2].cx[0, 1]
    c.rx(hp)[1]
    c.cx[0, 3].u(0.23, 1.24, -0.65)[3].cx[3, 1].cx[3, 0]
    vec = O.OracleBackend().run(c.ops, c.n_qubits, ignore_global=True)
    expect = np.array([
        5.88423813e-01 + 0.00000000e+00j, -3.82057626e-02 - 5.70122617e-02j, -2.52821022e-17 - 5.09095967e-17j,
        -1.21188626e-11 + 5.63063568e-10j, -2.19604047e-01 - 2.85449458e-01j, -2.59211189e-03 + 4.58219688e-02j,
        3.08617333e-09 - 3.56619861e-09j, 4.48946755e-18 - 3.62425819e-19j, 4.64439684e-09 - 1.48402425e-09j,
        4.61321871e-18 - 4.67197922e-18j, -3.59382904e-01 + 4.73135946e-01j, 2.20759589e-02 + 6.42836440e-02j,
        -1.55912415e-17 - 3.57403200e-17j, 5.05381446e-10 + 2.03362289e-10j, 3.82475330e-01 - 1.07620677e-01j,
        2.29456407e-02 - 3.47003613e-02j])
    assert np.allclose(vec, expect)


def test_reference_analytic_cases():
    """tests/test_circuit.py:35-63 (H), :102-173 (CX ordering, issue #76), :367-399 (global phases)."""
    run = lambda c: O.OracleBackend().run(c.ops, c.n_qubits)
    r = 1 / math.sqrt(2)
    assert np.allclose(run(Circuit().h[0]), [r, r])
    assert np.allclose(run(Circuit().x[0].h[0]), [r, -r])
    assert np.allclose(run(Circuit().h[0].h[1]), [0.5] * 4)
    assert np.allclose(run(Circuit().x[1].cx[0, 1]), [0, 0, 1, 0])
    assert np.allclose(run(Circuit().x[0].cx[0, 1]), [0, 0, 0, 1])
    assert np.allclose(run(Circuit().h[0].cx[0, 1]), [r, 0, 0, r])
    assert np.allclose(run(Circuit().x[2].cx[:4:2, 1:4:2]), [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0])
    th = 1.23
    assert np.allclose(run(Circuit().rz(th)[0]), [np.exp(-0.5j * th), 0])
    assert np.allclose(run(Circuit().x[0].phase(th)[0]), [0, np.exp(1j * th)])
    assert np.allclose(run(Circuit(1)), [1, 0])
    assert np.allclose(run(Circuit()), [1])


@pytest.mark.reference
def test_live_reference_random_circuits():
    import sys
    sys.path.insert(0, REFERENCE)
    warnings.simplefilter("ignore", SyntaxWarning)
    from blueqat import Circuit as RefCircuit
    rng = np.random.default_rng(5)
    for n in (2, 4, 7, 9):
        for rep in range(3):
            case = {"name": "live", "n": n, "ops": C.random_ops(rng, n, 80) + [C._op("m", slice(None))],
                    "initial_seed": None, "run": {"shots": 30, "seed": 40 + rep, "returns": "statevector_and_shots"}}
            random.seed(case["run"]["seed"])
            rsv, rcnt = C.build(RefCircuit, case).run(backend="numpy", shots=30, returns="statevector_and_shots")
            osv, ocnt = run_oracle(case)
            assert rcnt == ocnt
            assert np.abs(rsv - osv).max() <= 1e-15


def test_mirror_matches_blueqat_builder_semantics():
    """Target expansion / fallback shapes of the host mirror (gate.py:1283-1332, fallbacks 3.5)."""
    c = Circuit().h[:3].cx[0:2, 1:3].m[-1]
    assert c.n_qubits == 3
    assert [o.lowername for o in c.ops] == ["h", "cx", "measure"]
    assert list(c.ops[0].target_iter(3)) == [0, 1, 2]
    assert list(c.ops[1].control_target_iter(3)) == [(0, 1), (1, 2)]
    assert list(c.ops[2].target_iter(3)) == [2]
    with pytest.raises(ValueError):
        list(Circuit().cx[0, 0].ops[0].control_target_iter(2))
    with pytest.raises(ValueError):
        list(Circuit().cx[0:2, 2].ops[0].control_target_iter(3))
    sw = Circuit().swap[0, 1].ops[0].fallback(2)
    assert [(g.lowername, g.targets) for g in sw] == [("cx", (0, 1)), ("cx", (1, 0)), ("cx", (0, 1))]
    assert Circuit().u(1, 2, 3)[0].ops[0].gamma == 0.0
    d = Circuit().rx(0.3)[0].s[1].dagger()
    assert [(g.lowername, g.params) for g in d.ops] == [("sdg", ()), ("rx", (-0.3,))]
