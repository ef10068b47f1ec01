"""Device marginal distribution / non-sampling sampler: qaqarot_b200.samplers against the oracle's
restatement of blueqat's ``vqe.expect`` and, in the build container, against blueqat's own function."""
import os
import sys

import numpy as np
import pytest

from conftest import REFERENCE
from helpers import maxabs
from oracle import sv_oracle as O
from oracle.program_interp import OracleState
from qaqarot_b200 import B200Backend, Circuit, samplers, workloads as W
from qaqarot_b200.planner import PlanConfig


def _interp_backend():
    return B200Backend(state_factory=lambda n, d: OracleState(n, d),
                       plan_config=PlanConfig(a=5, r=2, c=2, min_qubits=5, kernel_tiles=tuple(range(1, 20))))


CASES = [(W.random_layers(9, 4), (0, 3, 8)), (W.qft(8), (7, 1)), (W.qaoa_maxcut(8)[0], (2, 2, 5)),
         (Circuit(6).h[0].cx[0, 1].cx[1, 2].x[4], (0, 1, 2, 4, 5)), (Circuit(4).h[:], ())]


@pytest.mark.parametrize("k", range(len(CASES)))
def test_marginal_matches_the_oracle_restatement(k):
    circ, meas = CASES[k]
    psi = O.OracleBackend().run(circ.ops, circ.n_qubits)
    want = O.expect_marginal(psi, meas)
    got = samplers.non_sampling_sampler(circ, meas, backend=_interp_backend())
    assert set(got) == set(want) or all(abs(want.get(key, 0.0) - got.get(key, 0.0)) < 1e-13 for key in set(got) | set(want))
    for key in want:
        assert abs(got.get(key, 0.0) - want[key]) < 1e-13
    assert abs(sum(got.values()) - 1.0) < 1e-12


def test_ghz_marginal_has_exactly_two_outcomes():
    circ = Circuit(7).h[0]
    for q in range(6):
        circ.cx[q, q + 1]
    got = samplers.non_sampling_sampler(circ, (0, 3, 6), backend=_interp_backend())
    assert set(got) == {(0, 0, 0), (1, 1, 1)} and abs(got[(0, 0, 0)] - 0.5) < 1e-14


def test_state_vector_sampler_draws_from_the_marginal():
    circ = Circuit(5).h[0].cx[0, 3]
    np.random.seed(3)
    got = samplers.get_state_vector_sampler(1000, backend=_interp_backend())(circ, (0, 3))
    assert set(got) <= {(0, 0), (1, 1)} and abs(sum(got.values()) - 1.0) < 1e-12 and abs(got[(0, 0)] - 0.5) < 0.1


def test_marginal_rejects_qubits_outside_the_register():
    with pytest.raises(ValueError):
        _interp_backend().marginal(Circuit(3).h[0].ops, 3, (0, 3))


@pytest.mark.reference
def test_marginal_matches_blueqat_expect():
    """blueqat's own vqe.expect on the oracle's statevector (the reference is importable in the build container)."""
    sys.path.insert(0, REFERENCE)
    try:
        from blueqat.vqe import expect
    finally:
        sys.path.remove(REFERENCE)
    for circ, meas in CASES:
        psi = O.OracleBackend().run(circ.ops, circ.n_qubits)
        want = expect(psi, meas)
        got = samplers.non_sampling_sampler(circ, meas, backend=_interp_backend())
        assert set(k for k, v in want.items() if v > 1e-30) <= set(got)
        for key, v in want.items():
            assert abs(got.get(key, 0.0) - v) < 1e-13


@pytest.mark.gpu
def test_marginal_on_the_device():
    for n, meas in ((14, (0, 5, 13, 9)), (18, tuple(range(0, 18, 2))), (12, (3,))):
        circ = W.random_layers(n, 5)
        psi = O.OracleBackend().run(circ.ops, n)
        want = O.expect_marginal(psi, meas)
        be = B200Backend()
        try:
            got = samplers.non_sampling_sampler(circ, meas, backend=be)
        finally:
            be.close()
        for key in set(want) | set(got):
            assert abs(got.get(key, 0.0) - want.get(key, 0.0)) < 1e-13
