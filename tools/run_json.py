#!/usr/bin/env python
"""Run a circuit stored in blueqat's JSON schema (circuit_funcs/json_serializer.py) on the B200 backend.

    python tools/run_json.py circuit.json [--shots N] [--seed S] [--out state.npy]

Without --shots the statevector is written to --out (or its norm and first amplitudes printed); with --shots the
Counter is printed as JSON.  The file is what ``json.dump(serialize(circuit), f)`` writes, schema version 1 or 2.
"""
import argparse
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np                                                # noqa: E402

from qaqarot_b200 import B200Backend                              # noqa: E402
from qaqarot_b200.json_io import deserialize                      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("path")
    ap.add_argument("--shots", type=int, default=0)
    ap.add_argument("--seed", type=int, default=None)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    with open(args.path) as f:
        circ = deserialize(json.load(f))
    be = B200Backend()
    if args.shots:
        if args.seed is not None:
            random.seed(args.seed)
        print(json.dumps(dict(circ.run(backend=be, shots=args.shots))))
    else:
        sv = circ.run(backend=be)
        if args.out:
            np.save(args.out, sv)
        print(json.dumps({"n_qubits": circ.n_qubits, "ops": len(circ.ops), "norm2": float(np.vdot(sv, sv).real),
                          "first_amplitudes": [[float(a.real), float(a.imag)] for a in sv[:4]]}))
    be.close()


if __name__ == "__main__":
    main()
