"""``run(backend="b200_fuse")``: the planner's normal form returned as a circuit.  The fused
circuit must be the same unitary as the original, global phase included, on ANY backend -- checked with the oracle."""
import random
import sys
import warnings

import numpy as np
import pytest

import cases as C
from conftest import REFERENCE, case_names
from helpers import TOL, maxabs
from oracle import sv_oracle as O
from qaqarot_b200 import Circuit, workloads as W


def _oracle(circ, **kw):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return O.OracleBackend().run(circ.ops, circ.n_qubits, **kw)


@pytest.mark.parametrize("name", case_names("statevector"))
def test_fused_circuit_has_the_same_statevector(name, gold):
    case = gold.cases[name]
    circ = C.build(Circuit, case)
    if case.get("run", {}).get("initial") is not None:
        pytest.skip("initial= is a run option, not part of the circuit")
    fused = circ.run(backend="b200_fuse")
    assert isinstance(fused, Circuit) and fused.n_qubits == circ.n_qubits
    random.seed(11)                      # (cases with a measurement in the middle collapse by the seeded draws)
    want = _oracle(circ)
    random.seed(11)
    assert maxabs(_oracle(fused), want) <= TOL


def test_measurements_and_resets_keep_their_place():
    rng = np.random.default_rng(12)
    for _ in range(10):
        n = int(rng.integers(2, 7))
        ops = C.random_ops(rng, n, 40, True)
        for _ in range(2):
            ops.insert(int(rng.integers(0, len(ops) + 1)), C._op("reset" if rng.integers(2) else "m", int(rng.integers(n))))
        ops.append(C._op("m", tuple(range(n)) if n > 1 else 0))
        circ = C.build(Circuit, {"n": n, "ops": ops})
        fused = circ.run(backend="b200_fuse")
        random.seed(5)
        want = _oracle(circ, shots=40)
        random.seed(5)
        assert _oracle(fused, shots=40) == want


def test_fusion_shrinks_layered_circuits_and_keeps_ladders():
    lay = W.random_layers(10, 6)
    fused = lay.run(backend="b200_fuse")
    assert len(fused.ops) < 0.7 * len(lay.ops)
    names = {op.lowername for op in fused.ops}
    assert names <= {"mat1", "phase", "cphase", "cx", "x", "swap"}
    qft = W.qft(9).run(backend="b200_fuse")
    assert sum(op.lowername == "mat1" for op in qft.ops) == 9          # one per Hadamard, nothing else to merge
    assert maxabs(_oracle(qft), _oracle(W.qft(9))) <= TOL


@pytest.mark.reference
def test_blueqat_circuits_come_back_as_blueqat_circuits():
    sys.path.insert(0, REFERENCE)
    try:
        import blueqat
        import qaqarot_b200
        qaqarot_b200.register()
        c = blueqat.Circuit(4).h[:].rz(0.3)[1].cx[0, 1].t[1].cx[0, 1].ry(0.2)[3].cz[2, 3].swap[0, 3]
        fused = c.run(backend="b200_fuse")
        assert isinstance(fused, blueqat.Circuit)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert maxabs(fused.run(backend="numpy"), c.run(backend="numpy")) <= TOL
    finally:
        sys.path.remove(REFERENCE)
