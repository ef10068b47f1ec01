#!/usr/bin/env python
"""Per-record cost of the fused tile kernel, by record kind (needs a B200).

Builds one tile pass by hand -- a single round holding K copies of the same item -- and times it for K = 1, 5, 9, 17
on a 2^n-amplitude state; the slope is what one record of that kind adds to a pass.  This separates the fixed cost of a
record (dispatch, payload loads, the warp's dependent chain) from its FP64 work.

    python tools/record_cost.py [--qubits 28] > gpurun_out/record_cost.json
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from qaqarot_b200 import planner as P                     # noqa: E402
from qaqarot_b200.device import DeviceState               # noqa: E402
from qaqarot_b200.program import Program                  # noqa: E402


def ry(t):
    return np.array([np.cos(t), -np.sin(t), np.sin(t), np.cos(t)], dtype=complex)


def shear_item(q):
    return P.shear_item(q, 0.29)


def items_of(kind: str, k: int, n: int):
    """K items on register bits 4..7 (the round's register bits), cycling through them."""
    out = []
    for i in range(k):
        q = 4 + (i % 4)
        if kind == "x":
            out.append(P.Item("x", t=q))
        elif kind == "layer1":                  # one 3-shear gate per record (a CZ on thread bits keeps them apart)
            out.append(shear_item(q))
            out.append(P.Item("fan", t=4, w=np.exp(0.1j), entries={5: -1.0 + 0j, 6: 1j, 7: -1j}))
        elif kind == "layer4":                  # four gates (one per register bit) + a table per record
            for qq in (4, 5, 6, 7):
                out.append(shear_item(qq))
            out.append(P.Item("fan", t=4, w=np.exp(0.1j), entries={5: -1.0 + 0j}))
        elif kind == "matr":
            out.append(P.Item("mat", t=q, m=ry(0.3)))
        elif kind == "matc":
            out.append(P.Item("mat", t=q, m=ry(0.3) * np.exp(0.2j) * np.array([1, 1j, 1j, 1])))
        elif kind == "w":                       # 1-qubit phase on a register bit (D_W)
            out.append(P.Item("fan", t=q, w=np.exp(0.1j), entries={}))
        elif kind == "layer1+fan":              # fused record
            out.append(shear_item(q))
            out.append(P.Item("fan", t=q, w=np.exp(0.1j), entries={4 + ((i + 1) % 4): np.exp(0.01j), 9: np.exp(0.02j)}))
        elif kind == "fant":                    # fan with thread-bit entries only
            out.append(P.Item("fan", t=q, w=np.exp(0.1j), entries={0: np.exp(0.01j), 9: np.exp(0.02j)}))
        elif kind == "fan":                     # fan with a register-bit entry
            out.append(P.Item("fan", t=q, w=np.exp(0.1j), entries={4 + ((i + 1) % 4): np.exp(0.01j), 9: np.exp(0.02j)}))
        elif kind == "matr+fan":
            out.append(P.Item("mat", t=q, m=ry(0.3)))
            out.append(P.Item("fan", t=q, w=np.exp(0.1j), entries={4 + ((i + 1) % 4): np.exp(0.01j), 9: np.exp(0.02j)}))
        else:
            raise ValueError(kind)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=28)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    n = args.qubits
    st = DeviceState(n, 0)
    tile = list(range(12))
    res = {"n_qubits": n, "tiles": 1 << (n - 12), "kinds": {}}
    for kind in ("x", "layer1", "layer4", "layer1+fan", "matr", "matc", "fant", "fan", "matr+fan"):
        row = {}
        ks = (1, 5, 9, 13) if kind in ("matr+fan", "layer1+fan", "layer4") else (1, 5, 9, 17)   # payload bound
        for k in ks:
            prog = Program()
            ps = P.Pass(tile, [P.Round([4, 5, 6, 7], items_of(kind, k, n))])
            P.encode_pass(prog, ps, n, 4, 4, False, True)
            st.set_basis(0)
            for _ in range(2):
                st.apply(prog)
            st.sync()
            ts = []
            for _ in range(args.reps):
                st.apply(prog)
                st.sync()
                ts.append(st.last_program_ms())
            row[k] = float(np.median(ts))
        slope = (row[ks[-1]] - row[1]) / (ks[-1] - 1.0)
        res["kinds"][kind] = {"ms_by_records": row, "ms_per_record": slope,
                              "sm_cycles_per_record_per_tile_per_cta": slope * 1e-3 * 1.965e9 / (res["tiles"] / 296.0)}
        print(kind, row, f"slope {slope:.4f} ms/record", file=sys.stderr)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
