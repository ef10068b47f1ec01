"""Shared helpers for the parity tests."""
import random
import warnings

import numpy as np

import cases as C
from qaqarot_b200.backend import B200Backend
from qaqarot_b200.circuit import Circuit

TOL = 1e-12      # north_star: max-abs amplitude error vs NumPyBackend in complex128


def run_case(case, backend, **over):
    """Replay a golden case on `backend` (a B200Backend instance)."""
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
        return circ.run(backend=backend, **kw)


def cpu_backend(fusion=True, **kw):
    """B200Backend whose device is replaced by the numpy program interpreter (host-logic tests only)."""
    from oracle.program_interp import OracleState
    return B200Backend(fusion=fusion, state_factory=lambda n, dev: OracleState(n, dev), **kw)


def maxabs(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max())
