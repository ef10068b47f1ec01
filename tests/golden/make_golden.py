"""Generate the committed golden fixtures by running the REAL reference (blueqat from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Writes next to this file:
    cases.json    -- the case definitions (pure data, see cases.py)
    golden.npz    -- for each case: the reference's statevector (``sv/<name>``)
    golden.json   -- shot Counters / samples / expectation values / reference test vectors

Every value is produced by ``blueqat.Circuit(...).run(backend="numpy", ...)`` (numpy_backend.py:572-630)
with Python's global ``random`` seeded as the case says; expectation values by the in-tree matrix route
``vqe.sparse_expectation(expr.to_matrix(n, sparse="csc"), psi)`` (vqe.py:355-366, pauli.py:893-935).
"""
import json
import os
import random
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, "/root/reference")

warnings.simplefilter("ignore", SyntaxWarning)
import blueqat                                            # noqa: E402
from blueqat import Circuit, pauli, vqe                   # noqa: E402
import cases as C                                         # noqa: E402


def main():
    all_cases = C.all_cases()
    arrays, meta = {}, {"blueqat_version": getattr(blueqat, "__version__", "?"), "numpy": np.__version__,
                        "cases": {}, "expect": {}, "reference_vectors": {}}
    finals = {}
    for case in all_cases:
        name, run = case["name"], case["run"]
        circ = C.build(Circuit, case)
        kwargs = {}
        init = C.initial_state(case)
        if init is not None:
            kwargs["initial"] = init
        if run["shots"] is not None:
            kwargs["shots"] = run["shots"]
        if run["returns"] is not None:
            kwargs["returns"] = run["returns"]
        if run["seed"] is not None:
            random.seed(run["seed"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            out = circ.run(backend="numpy", **kwargs)
        rec = {}
        returns = run["returns"] or ("statevector" if run["shots"] is None else "shots")
        if returns == "statevector":
            arrays["sv/" + name] = np.array(out)
            finals[name] = np.array(out)
        elif returns == "shots":
            rec["shots"] = dict(out)
        elif returns == "statevector_and_shots":
            arrays["sv/" + name] = np.array(out[0])
            rec["shots"] = dict(out[1])
        elif returns == "samples":
            rec["samples"] = out
        meta["cases"][name] = rec

    # expectation values through the reference's matrix route
    letter = {"X": pauli.X, "Y": pauli.Y, "Z": pauli.Z}
    for name, terms in C.expectation_cases():
        n = next(c["n"] for c in all_cases if c["name"] == name)
        expr = 0
        for coeff, letters in terms:
            t = coeff * pauli.I
            for ch, q in letters:
                t = t * letter[ch][q]
            expr = expr + t
        mat = expr.to_expr().simplify().to_matrix(n, sparse="csc")
        meta["expect"].setdefault(name, []).append(
            {"terms": terms, "value": float(vqe.sparse_expectation(mat, finals[name]))})

    # sanity: our QAOA generator is the circuit vqe.QaoaAnsatz emits (same state up to rounding)
    for nq in (10, 12):
        ops, edges = C.qaoa_ops(nq)
        h = sum((pauli.Z[i] * pauli.Z[j] for i, j in edges[1:]), pauli.Z[edges[0][0]] * pauli.Z[edges[0][1]])
        params = np.random.default_rng(0).uniform(0, 1, 8)
        ref = vqe.QaoaAnsatz(h, 4).get_circuit(params).run(backend="numpy")
        assert np.abs(ref - finals[f"c3_qaoa_{nq}q"]).max() < 1e-12, "qaoa generator drifted from QaoaAnsatz"

    # the reference's own analytic / golden expectations that tests restate are kept in the tests
    # themselves (they are literals in /root/reference/tests/test_circuit.py); README known answers:
    meta["reference_vectors"]["readme_hamiltonian"] = {"circuit": "Circuit(4).x[:]", "hamiltonian": "Z0+Z1",
                                                       "value": -2.0}

    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(all_cases, f)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **arrays)
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f)
    tot = sum(a.nbytes for a in arrays.values())
    print(f"{len(all_cases)} cases, {len(arrays)} vectors, {tot/1e6:.2f} MB raw")


if __name__ == "__main__":
    main()
