"""Samplers feeding blueqat's ``get_energy``: the marginal distribution of the measured qubits
straight from the device state.

The reference's ``vqe.expect`` (blueqat/vqe.py:231-252) walks all 2^n amplitudes in a Python loop after
``circuit.run(returns="statevector")`` has copied them to the host -- unusable beyond ~22 qubits.  Here the state
stays in HBM and one ``b200_cond_probs`` reduction (include/b200sv.h; the reduction the batched sampler already uses)
returns the 2^k probabilities of the k measured qubits.  ``non_sampling_sampler`` / ``get_state_vector_sampler`` have
the reference's signatures (vqe.py:255-292): ``sampler(circuit, meas) -> {bits tuple: probability}``.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np


def marginal_table(st, meas: Sequence[int]) -> Tuple[Tuple[int, ...], np.ndarray]:
    """(distinct measured qubits, probabilities[2^k]) with bit j of the table index = qubit distinct[j]."""
    distinct = tuple(dict.fromkeys(int(q) for q in meas))
    if not distinct:
        return distinct, np.array([st.norm2()])
    return distinct, np.asarray(st.marginal(distinct), dtype=np.float64)


NOISE_FLOOR = 1e-28      # probabilities below (1e-14)^2 are rounding residue of the fused gate arithmetic


def expect_on_state(st, meas: Sequence[int]) -> Dict[Tuple[int, ...], float]:
    """``vqe.expect(statevector, meas)`` (vqe.py:231-252) evaluated on the device: keys are the outcomes of the
    qubits in `meas` (in that order, repeats allowed).  The reference leaves out outcomes of probability exactly 0;
    here amplitudes that are exactly 0 on the gate-by-gate numpy path may carry ~1e-16 of rounding residue, so
    outcomes below NOISE_FLOOR are left out instead (a GHZ state then has its two keys, as on the reference)."""
    meas = tuple(int(q) for q in meas)
    distinct, table = marginal_table(st, meas)
    where = {q: j for j, q in enumerate(distinct)}
    out: Dict[Tuple[int, ...], float] = {}
    for idx in np.nonzero(table > NOISE_FLOOR)[0]:
        out[tuple((int(idx) >> where[q]) & 1 for q in meas)] = float(table[idx])
    return out


def non_sampling_sampler(circuit, meas, backend=None) -> Dict[Tuple[int, ...], float]:
    """Drop-in for ``blueqat.vqe.non_sampling_sampler`` (vqe.py:255-259) on a B200 backend instance."""
    be = _backend(backend)
    return be.marginal(circuit.ops, circuit.n_qubits, meas)


def get_state_vector_sampler(n_sample: int, backend=None):
    """Drop-in for ``blueqat.vqe.get_state_vector_sampler`` (vqe.py:282-292): multinomial draws (numpy's global
    generator, as in the reference) from the exact marginal distribution."""
    def sampling_by_measurement(circuit, meas):
        e = non_sampling_sampler(circuit, meas, backend)
        bits, probs = zip(*e.items())
        dists = np.random.multinomial(n_sample, probs) / n_sample
        return dict(zip(tuple(bits), dists))

    return sampling_by_measurement


_default = None


def _backend(backend):
    global _default
    if backend is not None:
        return backend
    if _default is None:
        from .backend import B200Backend
        _default = B200Backend()
    return _default
