"""Generate tests/golden/json_circuits.json with the REAL reference (build container only):

    python tests/golden/make_json_fixtures.py

For each case: the document blueqat's own ``serialize`` writes (circuit_funcs/json_serializer.py:57-91) and the
statevector ``deserialize(document).run(backend="numpy")`` returns, plus the hand-written version "1" documents of
the reference's tests (tests/test_json.py:85-118) with their statevectors.
"""
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
warnings.simplefilter("ignore")
from blueqat import Circuit                                                        # noqa: E402
from blueqat.circuit_funcs.json_serializer import deserialize, serialize          # noqa: E402


def circuits():
    yield "slices_u", Circuit().h[:3].x[4].u(1.2, 3.4, 2.3, 1.0)[0].h[:].u(1.2, 3.4, 2.3)[1]
    c = Circuit().h[:3].x[4].cx[1, 3].u(1.2, 3.4, 2.3, 1.0)[0]
    c.h[:].u(1.2, 3.4, 2.3)[1:].rzz(0.3)[0, 2].cphase(0.7)[0:2, 2:4].crx(0.4)[4, 0]
    yield "two_qubit_mix", c
    yield "phase_alias", Circuit().r(0.2)[0, 2].h[:].rx(0.5)[1].ry(-0.25)[2].rz(1.5)[0].sdg[1].tdg[2].sx[0]
    rng = np.random.default_rng(5)
    c = Circuit(9)
    for _ in range(6):
        c.h[:]
        for q in range(9):
            c.rz(float(rng.uniform(0, 6)))[q]
        c.cx[0:8:2, 1:9:2].cz[1:8:2, 2:9:2]
    yield "layers9", c


def main():
    out = {"documents": {}, "v1_documents": {}}
    for name, c in circuits():
        d = serialize(c)
        sv = deserialize(json.loads(json.dumps(d))).run(backend="numpy")
        out["documents"][name] = {"doc": d, "sv_re": sv.real.tolist(), "sv_im": sv.imag.tolist()}
    v1 = {
        "flat": {"schema": {"name": "blueqat-circuit", "version": "1"}, "n_qubits": 3, "ops": [
            {"name": "phase", "targets": [0], "params": [0.2]}, {"name": "phase", "targets": [2], "params": [0.2]},
            {"name": "h", "targets": [0], "params": []}, {"name": "h", "targets": [1], "params": []},
            {"name": "h", "targets": [2], "params": []}, {"name": "cx", "targets": [0, 1], "params": []}]},
        "unflat": {"schema": {"name": "blueqat-circuit", "version": "1"}, "n_qubits": 3, "ops": [
            {"name": "phase", "targets": [0, 2], "params": [0.2]}, {"name": "h", "targets": [0, 1, 2], "params": []},
            {"name": "rx", "targets": [1], "params": [0.3]}]},
    }
    for name, d in v1.items():
        c = deserialize(d)
        sv = c.run(backend="numpy")
        out["v1_documents"][name] = {"doc": d, "repr": repr(c), "sv_re": sv.real.tolist(), "sv_im": sv.imag.tolist()}
    # keyed measurements: the document only (serializeV2, tests/test_json.py:121-160)
    c = Circuit().m(key="foo")[0].h[:].m(key="foo", duplicated="replace")[1]
    out["keyed_measurements"] = serialize(c)
    with open(os.path.join(HERE, "json_circuits.json"), "w") as f:
        json.dump(out, f)
    print("wrote", len(out["documents"]), "+", len(out["v1_documents"]), "documents")


if __name__ == "__main__":
    main()
