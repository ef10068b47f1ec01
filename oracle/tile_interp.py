"""numpy interpreter of one fused tile pass descriptor.  TEST INFRASTRUCTURE -- never imported by the product.

Executes exactly the words / reals that cross the C ABI for a ``B200_OP_TILE`` op (format: csrc/tile.cu and
qaqarot_b200/planner.py::encode_pass), gate by gate on the whole state, evaluating masks, fan tables and
fixed-bit products the way the CUDA kernel does per thread.  It also checks the structural promises the
kernel relies on (tile bits ascending with the low `c` bits contiguous, orders being permutations, payload
offsets in range).
"""
from __future__ import annotations

import numpy as np

T_NONE, T_MATC, T_MATR, T_X = 0, 1, 2, 3
T_LAYER = 6
D_NONE, D_FANT, D_FAN, D_PHASE, D_W, D_FIX, D_TAB = 0, 1, 2, 3, 4, 5, 6


def _unpack8(lo: int, hi: int, count: int):
    v = [(lo >> (8 * i)) & 0xFF for i in range(8)] + [(hi >> (8 * i)) & 0xFF for i in range(8)]
    return v[:count]


def _pext(idx: np.ndarray, positions) -> np.ndarray:
    out = np.zeros_like(idx)
    for i, p in enumerate(positions):
        out |= ((idx >> np.uint64(p)) & np.uint64(1)) << np.uint64(i)
    return out


def run_tile_pass(psi: np.ndarray, n: int, words: np.ndarray, reals: np.ndarray, index_offset: int = 0) -> None:
    w = [int(x) for x in words]
    a, c, r = w[0] & 0xFF, (w[0] >> 8) & 0xFF, (w[0] >> 16) & 0xFF
    n_rounds, n_fans = (w[0] >> 24) & 0xFF, (w[0] >> 32) & 0xFF
    tile = _unpack8(w[1], w[2], a)
    assert a <= n and tile == sorted(set(tile)) and tile[:c] == list(range(c)) and all(q < n for q in tile)
    dir_off, rounds_off = w[3] & 0xFFFFFFFF, w[3] >> 32
    pay_lo, pay_len = w[4] & 0xFFFFFFFF, w[4] >> 32
    assert dir_off == 5 and pay_lo + pay_len <= reals.size
    cre = reals.view(np.complex128) if reals.size % 2 == 0 else reals[:-1].view(np.complex128)

    def cplx(roff, count=1):
        assert roff % 2 == 0 and pay_lo <= roff and roff + 2 * count <= pay_lo + pay_len, "payload offset out of range"
        return cre[roff // 2: roff // 2 + count]

    idx = np.arange(psi.size, dtype=np.uint64)
    tile_mask = sum(1 << q for q in tile)
    fixed = (idx | np.uint64(index_offset)) & ~np.uint64(tile_mask)
    local = _pext(idx, tile)

    # per-tile fixed products of every fan (the kernel computes these once per tile in its setup phase)
    fan_fixed = []
    for f in range(n_fans):
        off = w[dir_off + f]
        n_e, roff = w[off] & 0xFFFFFFFF, w[off] >> 32
        ws = cplx(roff, 1 + n_e)
        val = np.full(psi.size, ws[0], dtype=complex)
        for e in range(n_e):
            m = np.uint64(w[off + 1 + e])
            val = np.where((fixed & m) == m, val * ws[1 + e], val)
        fan_fixed.append(val)

    pos = rounds_off
    for _ in range(n_rounds):
        n_gates, n_words = w[pos] & 0xFFFF, w[pos] >> 16
        order = _unpack8(w[pos + 1], w[pos + 2], a)
        assert sorted(order) == list(range(a)), "round order is not a permutation of the local bits"
        rho = _pext(local, order[:r])
        tid = _pext(local, order[r:])
        lane, warp = tid & np.uint64(31), tid >> np.uint64(5)
        g = pos + 3
        for _g in range(n_gates):
            g0, fm, roffs = w[g], np.uint64(w[g + 1]), w[g + 2]
            g += 3
            tk, j, dk, fid = g0 & 0xFF, (g0 >> 8) & 0xFF, (g0 >> 16) & 0xFF, (g0 >> 24) & 0xFF
            lm, rm = np.uint64((g0 >> 32) & 0xFFFF), np.uint64((g0 >> 48) & 0xFFFF)
            roff_t, roff_d = roffs & 0xFFFFFFFF, roffs >> 32
            cond = ((fixed & fm) == fm) & ((local & lm) == lm)
            if tk == T_LAYER:
                # up to one rotation per register bit (reg mask says which), each as in-place shears
                # a0 += z a1; a1 += x a0; a0 += y a1 (y = 0: two updates, the record's diagonal stage rescales);
                # payload 4 x (z, x, y, 0)
                assert int(fm) == 0 and int(lm) == 0 and 0 < (int(rm) & 15) < (1 << r)
                assert dk in (D_NONE, D_FANT, D_FAN, D_W, D_FIX, D_TAB)
                pay = cplx(roff_t, 8)
                for k in range(r):
                    if not (int(rm) >> k) & 1:
                        continue
                    z, x, y = pay[2 * k].real, pay[2 * k].imag, pay[2 * k + 1].real
                    tbit = np.uint64(1 << tile[order[k]])
                    i0 = idx[(idx & tbit) == 0]
                    i1 = i0 | tbit
                    a0, a1 = psi[i0], psi[i1]
                    a0 = a0 + z * a1
                    a1 = a1 + x * a0
                    if y != 0.0:
                        a0 = a0 + y * a1
                    psi[i0], psi[i1] = a0, a1
            elif tk != T_NONE:
                assert j < r and dk in (D_NONE, D_FANT, D_FAN, D_W, D_FIX)
                assert dk == D_NONE or (int(fm) == 0 and int(lm) == 0), "a fused fan must be unconditional"
                assert dk == D_NONE or (tk != T_X and int(rm) == 0), "only an uncontrolled 2x2 is fused with a fan"
                c2 = cond & ((rho & rm) == rm)
                tbit = np.uint64(1 << tile[order[j]])
                i0 = idx[c2 & ((idx & tbit) == 0)]
                i1 = i0 | tbit
                a0, a1 = psi[i0], psi[i1]
                if tk == T_X:
                    psi[i0], psi[i1] = a1, a0
                else:
                    assert tk in (T_MATC, T_MATR)
                    m = cplx(roff_t, 4)
                    if tk == T_MATR:
                        assert np.all(m.imag == 0)
                    psi[i0] = m[0] * a0 + m[1] * a1
                    psi[i1] = m[2] * a0 + m[3] * a1
            if dk in (D_FANT, D_FAN):
                nt = a - r
                n_warp = 1 << max(0, nt - 5)
                tab = cplx(roff_d, 32 + n_warp + (1 << r if dk == D_FAN else 0))
                lane_t, warp_t, reg_t = tab[:32], tab[32:32 + n_warp], tab[32 + n_warp:]
                keep = cond if j >= r else cond & (((rho >> np.uint64(j)) & np.uint64(1)) == 1)
                factor = fan_fixed[fid] * lane_t[lane.astype(np.int64)] * warp_t[warp.astype(np.int64)]
                if dk == D_FAN:
                    factor = factor * reg_t[rho.astype(np.int64)]
                psi[keep] *= factor[keep]
            elif dk in (D_W, D_FIX):
                keep = cond if j >= r else cond & (((rho >> np.uint64(j)) & np.uint64(1)) == 1)
                if dk == D_W:
                    psi[keep] *= cplx(roff_d)[0]
                else:
                    psi[keep] *= fan_fixed[fid][keep]
            elif dk == D_TAB:
                assert int(fm) == 0 and int(lm) == 0, "a table applies to every amplitude"
                # table[variant][rho], variants 17 entries apart; variant = up to three selector bits: reg mask bits 4..9 and 10..15, fan byte.
                # A selector < 48 is an index bit outside the tile, 48 + l the tile-local thread bit l, 63 none.
                ps = [(int(rm) >> 4) & 63, (int(rm) >> 10) & 63, fid]
                assert ps == sorted(ps, key=lambda p: p == 63), "used selector bits come first"
                m = sum(1 for p in ps if p != 63)
                var = np.zeros_like(fixed)
                for i, p in enumerate(ps[:m]):
                    if p < 48:
                        assert p not in tile and p < n + 32
                        bit = (fixed >> np.uint64(p)) & np.uint64(1)
                    else:
                        assert p - 48 in order[r:], "a selector inside the tile must be a thread bit of the round"
                        bit = (local >> np.uint64(p - 48)) & np.uint64(1)
                    var |= bit << np.uint64(i)
                tab = cplx(roff_d, (17 << m) if m else (1 << r))
                psi *= tab[var.astype(np.int64) * 17 + rho.astype(np.int64)]
            elif dk == D_PHASE:
                assert tk == T_NONE
                keep = cond & ((rho & rm) == rm)
                psi[keep] *= cplx(roff_d)[0]
            else:
                assert dk == D_NONE and tk != T_NONE, "empty gate record"
        assert g == pos + n_words, "round length mismatch"
        pos += n_words
    assert pos == len(w), "descriptor length mismatch"
