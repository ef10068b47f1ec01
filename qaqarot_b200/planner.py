"""Fusion planner (placeholder: one pass per primitive until the tile pass lands)."""
from .program import unfused_program


def plan(prims, n_qubits):
    return unfused_program(prims)
