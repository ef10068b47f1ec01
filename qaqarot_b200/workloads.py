"""Circuit generators for the BASELINE.json configurations, on the host-mirror ``Circuit``.

Used by bench.py and the tests; every generator is deterministic (fixed seeds) so the CPU reference and the
device run the same gate list.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import numpy as np

from .circuit import Circuit


def random_layers(n: int, depth: int, seed: int = 1234) -> Circuit:
    """C1/C4/C5: per layer H on all qubits, RZ(theta ~ U[0, 2pi)) on all qubits, brick-pattern CX."""
    rng = np.random.default_rng(seed)
    c = Circuit(n)
    for d in range(depth):
        for q in range(n):
            c.h[q]
        for q in range(n):
            c.rz(float(rng.uniform(0, 2 * math.pi)))[q]
        for q in range(d % 2, n - 1, 2):
            c.cx[q, q + 1]
    return c


def qft(n: int, x: int = 123456789) -> Circuit:
    """C2: basis state x (mod 2^n) prepared with X gates, textbook QFT, final bit-reversal swaps."""
    x %= 1 << n
    c = Circuit(n)
    for q in range(n):
        if (x >> q) & 1:
            c.x[q]
    for j in range(n - 1, -1, -1):
        c.h[j]
        for k in range(j - 1, -1, -1):
            c.cphase(math.pi / 2 ** (j - k))[k, j]
    for q in range(n // 2):
        c.swap[q, n - 1 - q]
    return c


def qft_analytic_amplitude(n: int, x: int, y: np.ndarray) -> np.ndarray:
    """Amplitudes <y| QFT |x> = exp(2 pi i (x*y mod N) / N) / sqrt(N), for integer index arrays `y`."""
    x %= 1 << n
    mod = (x * y.astype(object)) % (1 << n) if n > 31 else (x * y.astype(np.int64)) % (1 << n)
    return np.exp(2j * np.pi * np.asarray(mod, dtype=np.float64) / float(1 << n)) / math.sqrt(float(1 << n))


def ring_edges(n: int) -> List[Tuple[int, int]]:
    """Circulant C_n(1, n/2): the MaxCut graph of C3."""
    e = [(i, (i + 1) % n) for i in range(n)]
    e += [(i, i + n // 2) for i in range(n // 2)]
    return sorted(tuple(sorted(p)) for p in e)


def qaoa_maxcut(n: int, p: int = 4, seed: int = 0):
    """C3: the circuit vqe.QaoaAnsatz(sum Z_i Z_j, p).get_circuit(params) emits (vqe.py:130-144,
    pauli.py:596-630) and the Hamiltonian.  Returns (circuit, hamiltonian, edges)."""
    from .pauli import Z
    params = np.random.default_rng(seed).uniform(0, 1, 2 * p)
    betas, gammas = params[:p] * math.pi, params[p:] * 2 * math.pi
    edges = ring_edges(n)
    c = Circuit(n)
    c.h[:]
    for beta, gamma in zip(betas, gammas):
        for i, j in edges:
            c.cx[i, j].rz(-2.0 * float(gamma))[j].cx[i, j]
        c.rx(float(beta))[:]
    h = Z[edges[0][0]] * Z[edges[0][1]]
    for i, j in edges[1:]:
        h = h + Z[i] * Z[j]
    return c, h, edges
