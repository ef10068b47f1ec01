"""numpy interpreter of the device op program.  TEST INFRASTRUCTURE -- never imported by the product.

``OracleState`` has the method surface of ``qaqarot_b200.device.DeviceState`` and executes the *encoded*
program arrays (struct b200_op[], words, reals -- exactly what crosses the C ABI) with numpy, so that the
host-side logic (lowering, fusion planning, descriptor encoding, shot descent, cache handling) can be checked
on a CPU against the reference fixtures, and so that GPU tests can compare the CUDA executor with an
independent executor of the same program.  Tests inject it with ``B200Backend(state_factory=...)``.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence

import numpy as np

from oracle.sv_oracle import sub

OP_MAT, OP_DIAG, OP_X, OP_SWAP, OP_PERMUTE, OP_TILE, OP_SCALE = 1, 2, 3, 4, 5, 16, 17


def _bits(mask: int):
    return [b for b in range(mask.bit_length()) if mask >> b & 1]


class OracleState:
    def __init__(self, n_qubits: int, device: int = 0, index_offset: int = 0):
        self.n_qubits = n_qubits
        self.device = device
        self.index_offset = index_offset       # shard number << n_qubits (qaqarot_b200/sharded.py)
        self.psi = np.zeros(1 << n_qubits, dtype=complex)
        self.psi[0] = 1.0
        self.launches = 0

    # -- lifetime / contents --------------------------------------------------------------------
    def close(self):
        pass

    @property
    def dim(self):
        return 1 << self.n_qubits

    def set_basis(self, index=0):
        self.psi[:] = 0.0
        self.psi[index] = 1.0

    def upload(self, vec, offset=0):
        self.psi[offset:offset + vec.size] = vec

    def download(self, offset=0, count=None, out=None):
        count = self.dim - offset if count is None else count
        res = self.psi[offset:offset + count].copy()
        if out is not None:
            out[:] = res
            return out
        return res

    MARGINAL_LOW_BITS = (8, 13)

    def marginal_low(self, k):
        """Counterpart of b200_marginal_low: table[v] = sum |amp[i]|^2 over i mod 2^k = v."""
        p = self.psi.real ** 2 + self.psi.imag ** 2
        return p.reshape(-1, 1 << k).sum(axis=0)

    def copy_from(self, other):
        self.psi[:] = other.psi

    def clone(self):
        c = OracleState(self.n_qubits)
        c.psi[:] = self.psi
        return c

    # -- program execution ----------------------------------------------------------------------
    def apply_piece(self, program, op_from, op_to, words_from, reals_from, init_basis=None, first=True):
        """Counterpart of DeviceState.apply_piece (b200_apply_program_ex): ops [op_from, op_to) of a program that is still
        being planned.  Like the library, it may only look at the pool tails the piece refers to: everything below
        words_from / reals_from is hidden (poisoned) while the piece runs."""
        from qaqarot_b200.program import Program
        piece = Program()
        piece._ops = list(program._ops[op_from:op_to])
        piece._bytes = list(program._bytes[op_from:op_to])
        piece._words = [0xDEADBEEFDEADBEEF] * words_from + list(program._words[words_from:])
        piece._reals = [float("nan")] * reals_from + list(program._reals[reals_from:])
        self.apply(piece, init_basis=init_basis)

    def apply(self, program, init_basis=None):
        if init_basis is not None:
            self.psi[:] = 0.0
            if (init_basis & ~(self.dim - 1)) == self.index_offset:
                self.psi[init_basis & (self.dim - 1)] = 1.0
        if len(program) == 0:
            return
        ops, n_ops, words, reals = program.frozen()
        n = self.n_qubits
        for i in range(n_ops):
            o = ops[i]
            ctrl = {c: 1 for c in _bits(o.cmask)}
            if o.kind == OP_MAT:
                m = reals[o.roff:o.roff + 8].view(complex)
                q0 = sub(self.psi, n, {**ctrl, o.target: 0})
                q1 = sub(self.psi, n, {**ctrl, o.target: 1})
                n0 = m[0] * q0 + m[1] * q1
                n1 = m[2] * q0 + m[3] * q1
                q0[...] = n0
                q1[...] = n1
            elif o.kind == OP_X:
                q0 = sub(self.psi, n, {**ctrl, o.target: 0})
                q1 = sub(self.psi, n, {**ctrl, o.target: 1})
                t = q0.copy()
                q0[...] = q1
                q1[...] = t
            elif o.kind == OP_DIAG:
                d = reals[o.roff:o.roff + 4].view(complex)
                sub(self.psi, n, {**ctrl, o.target: 0})[...] *= d[0]
                sub(self.psi, n, {**ctrl, o.target: 1})[...] *= d[1]
            elif o.kind == OP_SWAP:
                a = sub(self.psi, n, {o.target: 1, o.aux: 0})
                b = sub(self.psi, n, {o.target: 0, o.aux: 1})
                t = a.copy()
                a[...] = b
                b[...] = t
            elif o.kind == OP_PERMUTE:
                pairs = [(int(w) & 0xFF, (int(w) >> 8) & 0xFF) for w in words[o.woff:o.woff + o.wlen]]
                seen = [q for pq in pairs for q in pq]
                assert len(set(seen)) == len(seen) and all(q < n for q in seen), "transpositions must be disjoint"
                idx = np.arange(self.psi.size, dtype=np.int64)
                src = idx.copy()
                for p_, q_ in pairs:
                    diff = ((src >> p_) ^ (src >> q_)) & 1
                    src ^= (diff << p_) | (diff << q_)
                self.psi[:] = self.psi[src]
            elif o.kind == OP_SCALE:
                self.psi *= complex(reals[o.roff], reals[o.roff + 1])
            elif o.kind == OP_TILE:
                from oracle.tile_interp import run_tile_pass
                run_tile_pass(self.psi, n, words[o.woff:o.woff + o.wlen], reals, self.index_offset)
            else:
                raise ValueError(f"unknown op kind {o.kind}")
            self.launches += 1

    def last_program_ms(self):
        return 0.0

    def launch_count(self):
        return self.launches

    # -- measurement ----------------------------------------------------------------------------
    def prob_zero(self, target):
        v = sub(self.psi, self.n_qubits, {target: 0})
        return float(np.sum(v.real ** 2 + v.imag ** 2))

    def collapse(self, target, outcome, p, reset=False):
        s = 1.0 / math.sqrt(p)
        q0 = sub(self.psi, self.n_qubits, {target: 0})
        q1 = sub(self.psi, self.n_qubits, {target: 1})
        if outcome == 0:
            q0[...] *= s
            q1[...] = 0.0
        elif reset:
            q0[...] = q1 * s
            q1[...] = 0.0
        else:
            q1[...] *= s
            q0[...] = 0.0

    def norm2(self):
        return float(np.vdot(self.psi, self.psi).real)

    def scale(self, z):
        self.psi *= complex(z)

    def first_above(self, eps):
        hit = np.nonzero(np.abs(self.psi) > eps)[0]
        if hit.size == 0:
            return None
        return int(hit[0]), complex(self.psi[hit[0]])

    def first_above_wire(self, eps, wire_of_bit):
        hit = np.nonzero(np.abs(self.psi) > eps)[0].astype(np.int64) | np.int64(self.index_offset)
        if hit.size == 0:
            return None
        key = np.zeros_like(hit)
        for p, w in enumerate(wire_of_bit):
            key |= ((hit >> p) & 1) << int(w)
        return int(key.min())

    def sample(self, order: Sequence[int], uniforms: np.ndarray) -> np.ndarray:
        """Same descent as csrc/sample.cu: conditional probabilities on the uncollapsed state."""
        order = list(order)
        uniforms = np.asarray(uniforms, dtype=float).reshape(-1, max(1, len(order)))
        prob = self.psi.real ** 2 + self.psi.imag ** 2
        out = np.zeros(uniforms.shape[0], dtype=np.uint64)
        memo = {}
        for s in range(uniforms.shape[0]):
            fixed = {}
            bits = 0
            for j, q in enumerate(order):
                if q in fixed:
                    p0 = 0.0 if fixed[q] else 1.0
                else:
                    key = (tuple(sorted(fixed.items())), q)
                    if key not in memo:
                        s0 = float(np.sum(sub(prob, self.n_qubits, {**fixed, q: 0})))
                        s1 = float(np.sum(sub(prob, self.n_qubits, {**fixed, q: 1})))
                        memo[key] = s0 / (s0 + s1)
                    p0 = memo[key]
                bit = 0 if uniforms[s, j] < p0 else 1
                fixed[q] = bit
                bits |= bit << j
            out[s] = bits
        return out

    def set_index_offset(self, offset):
        self.index_offset = offset

    def cond_probs(self, fixed_mask, qubit, group_bits):
        prob = self.psi.real ** 2 + self.psi.imag ** 2
        idx = np.arange(self.dim, dtype=np.int64)
        out = np.zeros((len(group_bits), 2))
        for g, gb in enumerate(group_bits):
            sel = (idx & int(fixed_mask)) == int(gb)
            if qubit < 0:
                out[g, 0] = prob[sel].sum()
            else:
                bit = (idx >> qubit) & 1
                out[g, 0] = prob[sel & (bit == 0)].sum()
                out[g, 1] = prob[sel & (bit == 1)].sum()
        return out

    def marginal(self, qubits):
        qubits = [int(q) for q in qubits]
        idx = np.arange(1 << len(qubits), dtype=np.uint64)
        gb = np.zeros_like(idx)
        for j, q in enumerate(qubits):
            gb |= ((idx >> np.uint64(j)) & np.uint64(1)) << np.uint64(q)
        return self.cond_probs(sum(1 << q for q in qubits), -1, gb)[:, 0].copy()

    def expect_group(self, xmask, zymasks, coeffs):
        idx = np.arange(self.dim, dtype=np.int64)
        w = np.zeros(self.dim, dtype=complex)
        for zy, c in zip(zymasks, coeffs):
            par = np.zeros(self.dim, dtype=np.int64)
            for b in _bits(int(zy)):
                par ^= ((idx | self.index_offset) >> b) & 1
            w += complex(c) * (1 - 2 * par)
        return complex(np.sum(np.conj(self.psi[idx ^ int(xmask)]) * self.psi * w))

    def sync(self):
        pass
