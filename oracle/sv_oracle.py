"""CPU oracle for the statevector hot path.  TEST INFRASTRUCTURE -- never imported by the product.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this module; ``qaqarot_b200`` must not (and fails loudly without its CUDA library instead).

What it is: a numpy restatement of blueqat's ``NumPyBackend`` (reference:
blueqat/backends/numpy_backend.py, cited per function below).  It performs, gate by gate, the same
complex128 arithmetic in the same order as the reference -- the same products, the same sums, ``H`` scaled
after the add by ``1/sqrt(2)``, ``S`` as an exact ``1j``, ``CCX`` as ``H * CCZ * H``, measurement as
``norm(...)**2`` against one ``random.random()`` draw per measured qubit followed by a division of the
whole vector by ``sqrt(p)`` -- but addresses amplitudes through reshaped strided *views* of one array
instead of the reference's boolean masks over an index array and double buffering.  Element-wise results
are therefore the reference's to the last bit while running 5-30x faster, which keeps the CPU test suite
short and makes the reported CPU baseline conservative.

Pinning (parity is NOT unpinned): tests/test_oracle_golden.py checks this module against
(a) tests/golden/*.npz, produced by tests/golden/make_golden.py which imports the real blueqat from
/root/reference and runs ``Circuit.run(backend="numpy")`` (statevectors, seeded shot Counters, samples),
(b) the reference's own golden vector (tests/test_circuit.py:697-747) and analytic cases, restated in the
tests, and (c) -- when /root/reference is present, i.e. in the build container only -- the live reference
on randomised circuits.

Operations are duck typed (``lowername``, ``targets``, ``target_iter``, ``control_target_iter``,
``fallback``, ``theta/phi/lam/gamma/mat/key/duplicated``) so both blueqat ``Operation`` objects and
``qaqarot_b200.circuit.Op`` records are accepted.
"""
from __future__ import annotations

import cmath
import math
import random
import warnings
from collections import Counter
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

DEFAULT_SHOTS = 1024          # numpy_backend.py:33


def bit_view(psi: np.ndarray, n: int, bits: Tuple[int, ...]) -> Tuple[np.ndarray, Tuple[int, ...]]:
    """Reshape `psi` (2^n,) so that each qubit in `bits` gets its own axis of length 2.

    Returns (view, axes) with axes[i] = axis of bits[i].  Little endian: qubit k is bit k of the index
    (numpy_backend.py:82-83), so higher qubits are earlier axes.
    """
    order = sorted(bits, reverse=True)
    shape, prev = [], n
    for b in order:
        shape += [1 << (prev - 1 - b), 2]
        prev = b
    shape.append(1 << prev)
    v = psi.reshape(shape)
    return v, tuple(2 * order.index(b) + 1 for b in bits)


def sub(psi: np.ndarray, n: int, fixed: Dict[int, int]) -> np.ndarray:
    """Strided view of the amplitudes whose qubit `k` equals `fixed[k]`, in index order."""
    bits = tuple(fixed)
    v, axes = bit_view(psi, n, bits)
    idx: List[Any] = [slice(None)] * v.ndim
    for b, ax in zip(bits, axes):
        idx[ax] = fixed[b]
    return v[tuple(idx)]


class Ctx:
    """Per-run state (reference: _NumPyBackendContext, numpy_backend.py:36-70)."""

    def __init__(self, n_qubits: int, cache: Optional[np.ndarray] = None, cache_idx: int = -1):
        self.n_qubits = n_qubits
        self.qubits = np.zeros(2 ** n_qubits, dtype=complex)
        self.save_ctx_cache = True
        self.cache = cache
        self.cache_idx = cache_idx
        self.shots_result: Counter = Counter()
        self.cregs = [0] * n_qubits
        self.sample: Dict[str, List[int]] = {}

    def prepare(self, initial: Optional[np.ndarray]) -> None:           # numpy_backend.py:52-62
        if self.cache is not None:
            np.copyto(self.qubits, self.cache)
        elif initial is not None:
            np.copyto(self.qubits, initial)
        else:
            self.qubits.fill(0.0)
            self.qubits[0] = 1.0
        self.cregs = [0] * self.n_qubits
        self.sample = {}

    def store_shot(self) -> None:                                       # numpy_backend.py:64-70
        key = "".join(str(b) for b in self.cregs)
        self.shots_result[key] = self.shots_result.get(key, 0) + 1


# ---------------------------------------------------------------------------------------------------
# gate actions.  Each takes (op, ctx) and mutates ctx.qubits in place.
# ---------------------------------------------------------------------------------------------------

def _pairs(ctx: Ctx, t: int, ctrl: Tuple[int, ...] = ()):
    """Views (q0, q1) of the target-bit-0 / target-bit-1 halves, restricted to all controls = 1."""
    fixed0 = {c: 1 for c in ctrl}
    fixed1 = dict(fixed0)
    fixed0[t], fixed1[t] = 0, 1
    return sub(ctx.qubits, ctx.n_qubits, fixed0), sub(ctx.qubits, ctx.n_qubits, fixed1)


def _apply_2x2(ctx: Ctx, t: int, a00, a01, a10, a11, ctrl=()):
    """new0 = a00*q0 + a01*q1 ; new1 = a10*q0 + a11*q1   (numpy_backend.py:150-151, 286-287)."""
    q0, q1 = _pairs(ctx, t, ctrl)
    n0 = a00 * q0 + a01 * q1
    n1 = a10 * q0 + a11 * q1
    q0[...] = n0
    q1[...] = n1


def _apply_2x2_post(ctx: Ctx, t: int, a00, a01, a10, a11):
    """Operand order of gate_u / gate_mat1: q*a accumulated with += (numpy_backend.py:415-418, 437-440)."""
    q0, q1 = _pairs(ctx, t)
    n0 = q0 * a00
    n0 += q1 * a01
    n1 = q0 * a10
    n1 += q1 * a11
    q0[...] = n0
    q1[...] = n1


def _u_entries(theta, phi, lam, gamma):
    """Matrix entries exactly as gate_u / gate_cu build them (numpy_backend.py:275-280, 405-410)."""
    gp = cmath.exp(1j * gamma)
    a00 = math.cos(theta * 0.5) * gp
    a11 = a00 * cmath.exp(1j * (phi + lam))
    a01 = a10 = math.sin(theta * 0.5) * gp
    a01 *= -cmath.exp(1j * lam)
    a10 *= cmath.exp(1j * phi)
    return a00, a01, a10, a11


def g_x(op, ctx):                                                       # :75-89
    for t in op.target_iter(ctx.n_qubits):
        q0, q1 = _pairs(ctx, t)
        tmp = q0.copy()
        q0[...] = q1
        q1[...] = tmp


def g_y(op, ctx):                                                       # :92-106
    for t in op.target_iter(ctx.n_qubits):
        q0, q1 = _pairs(ctx, t)
        n0 = -1.0j * q1
        n1 = 1.0j * q0
        q0[...] = n0
        q1[...] = n1


def g_z(op, ctx):                                                       # :109-116
    for t in op.target_iter(ctx.n_qubits):
        sub(ctx.qubits, ctx.n_qubits, {t: 1})[...] *= -1


def g_h(op, ctx):                                                       # :119-134
    for t in op.target_iter(ctx.n_qubits):
        q0, q1 = _pairs(ctx, t)
        n0 = q0 + q1
        n1 = q0 - q1
        q0[...] = n0
        q1[...] = n1
        ctx.qubits *= 1.0 / math.sqrt(2)


def g_rx(op, ctx):                                                      # :137-155
    h = op.theta * 0.5
    a00 = a11 = math.cos(h)
    a01 = a10 = -1j * math.sin(h)
    for t in op.target_iter(ctx.n_qubits):
        _apply_2x2(ctx, t, a00, a01, a10, a11)


def g_ry(op, ctx):                                                      # :158-177
    h = op.theta * 0.5
    a00 = a11 = math.cos(h)
    a10 = math.sin(h)
    a01 = -a10
    for t in op.target_iter(ctx.n_qubits):
        _apply_2x2(ctx, t, a00, a01, a10, a11)


def g_rz(op, ctx):                                                      # :180-192
    h = op.theta * 0.5
    a0 = complex(math.cos(h), -math.sin(h))
    a1 = complex(math.cos(h), math.sin(h))
    for t in op.target_iter(ctx.n_qubits):
        q0, q1 = _pairs(ctx, t)
        q0[...] *= a0
        q1[...] *= a1


def _phase_on(ctx, bits, factor):
    sub(ctx.qubits, ctx.n_qubits, {b: 1 for b in bits})[...] *= factor


def g_phase(op, ctx):                                                   # :195-205
    a = complex(math.cos(op.theta), math.sin(op.theta))
    for t in op.target_iter(ctx.n_qubits):
        _phase_on(ctx, (t,), a)


def g_t(op, ctx):                                                       # :208-219
    r = 1 / math.sqrt(2)
    for t in op.target_iter(ctx.n_qubits):
        _phase_on(ctx, (t,), complex(r, r))


def g_s(op, ctx):                                                       # :222-229
    for t in op.target_iter(ctx.n_qubits):
        _phase_on(ctx, (t,), 1.j)


def g_cz(op, ctx):                                                      # :232-241
    for c, t in op.control_target_iter(ctx.n_qubits):
        _phase_on(ctx, (c, t), -1)


def g_cx(op, ctx):                                                      # :244-261
    for c, t in op.control_target_iter(ctx.n_qubits):
        q0, q1 = _pairs(ctx, t, (c,))
        tmp = q0.copy()
        q0[...] = q1
        q1[...] = tmp


def g_cu(op, ctx):                                                      # :264-291
    a = _u_entries(op.theta, op.phi, op.lam, op.gamma)
    for c, t in op.control_target_iter(ctx.n_qubits):
        _apply_2x2(ctx, t, *a, ctrl=(c,))


def g_crx(op, ctx):                                                     # :294-314
    h = op.theta * 0.5
    a00 = a11 = math.cos(h)
    a01 = a10 = -1j * math.sin(h)
    for c, t in op.control_target_iter(ctx.n_qubits):
        _apply_2x2(ctx, t, a00, a01, a10, a11, ctrl=(c,))


def g_cry(op, ctx):                                                     # :317-338
    h = op.theta * 0.5
    a00 = a11 = math.cos(h)
    a10 = math.sin(h)
    a01 = -a10
    for c, t in op.control_target_iter(ctx.n_qubits):
        _apply_2x2(ctx, t, a00, a01, a10, a11, ctrl=(c,))


def g_crz(op, ctx):                                                     # :341-355
    h = op.theta * 0.5
    a0 = complex(math.cos(h), -math.sin(h))
    a1 = complex(math.cos(h), math.sin(h))
    for c, t in op.control_target_iter(ctx.n_qubits):
        q0, q1 = _pairs(ctx, t, (c,))
        q0[...] *= a0
        q1[...] *= a1


def g_cphase(op, ctx):                                                  # :358-369
    a = complex(math.cos(op.theta), math.sin(op.theta))
    for c, t in op.control_target_iter(ctx.n_qubits):
        _phase_on(ctx, (c, t), a)


def g_ccz(op, ctx):                                                     # :372-382
    _phase_on(ctx, tuple(op.targets), -1)


class _One:
    """Single-target stand-in used by g_ccx."""
    def __init__(self, t): self.t = t
    def target_iter(self, n): return iter((self.t,))


def g_ccx(op, ctx):                                                     # :385-392  (H, CCZ, H)
    t = op.targets[2]
    g_h(_One(t), ctx)
    g_ccz(op, ctx)
    g_h(_One(t), ctx)


def g_u(op, ctx):                                                       # :395-422
    a = _u_entries(op.theta, op.phi, op.lam, op.gamma)
    for t in op.target_iter(ctx.n_qubits):
        _apply_2x2_post(ctx, t, *a)


def g_mat1(op, ctx):                                                    # :425-444
    m = op.mat
    for t in op.target_iter(ctx.n_qubits):
        _apply_2x2_post(ctx, t, m[0, 0], m[0, 1], m[1, 0], m[1, 1])


def _p_zero(ctx, t) -> float:
    """np.linalg.norm(q[bit t == 0])**2 over the same element order as the reference (:455, :487)."""
    return np.linalg.norm(sub(ctx.qubits, ctx.n_qubits, {t: 0}).ravel()) ** 2


def g_measure(op, ctx):                                                 # :447-477
    measured = []
    for t in op.target_iter(ctx.n_qubits):
        p0 = _p_zero(ctx, t)
        if random.random() < p0:
            sub(ctx.qubits, ctx.n_qubits, {t: 1})[...] = 0.0
            ctx.qubits /= math.sqrt(p0)
            bit = 0
        else:
            sub(ctx.qubits, ctx.n_qubits, {t: 0})[...] = 0.0
            ctx.qubits /= math.sqrt(1.0 - p0)
            bit = 1
        ctx.cregs[t] = bit
        measured.append(bit)
    key = getattr(op, "key", None)
    if key is not None:
        if key in ctx.sample:
            if op.duplicated == "replace":
                ctx.sample[key] = measured
            elif op.duplicated == "append":
                ctx.sample[key] += measured
            else:
                raise ValueError("Measurement key is duplicated.")
        else:
            ctx.sample[key] = measured


def g_reset(op, ctx):                                                   # :480-497
    for t in op.target_iter(ctx.n_qubits):
        p0 = _p_zero(ctx, t)
        q0, q1 = _pairs(ctx, t)
        if random.random() < p0:
            q1[...] = 0.0
            ctx.qubits /= math.sqrt(p0)
        else:
            q0[...] = q1
            q1[...] = 0.0
            ctx.qubits /= math.sqrt(1.0 - p0)


ACTIONS = {name[2:]: fn for name, fn in list(globals().items()) if name.startswith("g_")}
NON_CACHEABLE = ("measure", "reset")                                    # OPS_DISABLE_CTX_CACHE, :31


def check_initial(initial, n_qubits: int) -> np.ndarray:                # :500-512
    if not isinstance(initial, np.ndarray):
        raise ValueError(f"`initial` must be a np.ndarray, but {type(initial)}")
    if initial.shape != (2 ** n_qubits,):
        raise ValueError(f"`initial.shape` is not matched. Expected: {(2**n_qubits,)}, Actual: {initial.shape}")
    return initial if initial.dtype == complex else initial.astype(complex)


def ignore_global_phase(vec: np.ndarray) -> np.ndarray:                 # utils.py:130-144
    for q in vec:
        if abs(q) > 0.0000001:
            vec *= abs(q) / q
            break
    return vec


def run_gate(ctx: Ctx, op) -> None:                                     # run_single_gate, :548-560
    fn = ACTIONS.get(op.lowername)
    if fn is not None:
        fn(op, ctx)
    elif hasattr(op, "fallback"):
        for g in op.fallback(ctx.n_qubits):
            run_gate(ctx, g)
    else:
        raise ValueError(f"Unknown operation {op.lowername}.")


def run_inner(ctx: Ctx, ops, initial) -> Ctx:                           # _run_inner, :545-570
    ctx.prepare(initial)
    for i, op in enumerate(ops[ctx.cache_idx + 1:]):
        if op.lowername in NON_CACHEABLE and ctx.save_ctx_cache:
            ctx.cache = ctx.qubits.copy()
            ctx.cache_idx += i
            ctx.save_ctx_cache = False
        run_gate(ctx, op)
    return ctx


class OracleBackend:
    """Same call surface as NumPyBackend (numpy_backend.py:515-685)."""

    RETURNS = ("statevector", "shots", "statevector_and_shots", "_inner_ctx", "samples")

    def __init__(self):
        self.cache: Optional[np.ndarray] = None
        self.cache_idx = -1

    def copy(self):
        import copy
        return copy.deepcopy(self)

    def run(self, gates, n_qubits, shots=None, initial=None, returns=None, ignore_global=False,
            save_cache=False, **kwargs):
        if returns is None:
            returns = "statevector" if shots is None else "shots"
        if returns not in self.RETURNS:
            raise ValueError(f"Unknown returns type '{returns}'")
        if shots is None:
            shots = 1 if returns in ("statevector", "_inner_ctx") else DEFAULT_SHOTS
        if returns == "statevector" and shots > 1:
            shots = 1
            warnings.warn("When `returns` = 'statevector', `shots` is ignored.")
        if kwargs:
            warnings.warn(f"Unknown run arguments: {kwargs}")
        if returns == "samples":
            return self.samples(gates, n_qubits, shots, initial)
        if initial is not None:
            initial = check_initial(initial, n_qubits)
            if save_cache:
                warnings.warn("When initial is not None, saving cache is disabled.")
                save_cache = False
            self.cache, self.cache_idx = None, -1
        elif self.cache is None or len(self.cache) != 2 ** n_qubits:
            self.cache, self.cache_idx = None, -1
        ctx = Ctx(n_qubits, self.cache, self.cache_idx)
        for _ in range(shots):
            run_inner(ctx, gates, initial)
            if ctx.cregs:
                ctx.store_shot()
            if save_cache:
                self.cache, self.cache_idx = ctx.cache, ctx.cache_idx
        if ignore_global:
            ignore_global_phase(ctx.qubits)
        return {"statevector": ctx.qubits, "shots": ctx.shots_result,
                "statevector_and_shots": (ctx.qubits, ctx.shots_result), "_inner_ctx": ctx}[returns]

    @staticmethod
    def statevector(gates, n_qubits, initial=None):                    # :632-640
        if initial is not None:
            initial = check_initial(initial, n_qubits)
        return run_inner(Ctx(n_qubits), gates, initial).qubits

    @staticmethod
    def shots(gates, n_qubits, shots, initial=None):                   # :642-654
        if initial is not None:
            initial = check_initial(initial, n_qubits)
        ctx = Ctx(n_qubits)
        for _ in range(shots):
            run_inner(ctx, gates, initial)
            if ctx.cregs:
                ctx.store_shot()
        return ctx.shots_result

    @staticmethod
    def oneshot(gates, n_qubits, initial=None):                        # :656-667
        if initial is not None:
            initial = check_initial(initial, n_qubits)
        ctx = run_inner(Ctx(n_qubits), gates, initial)
        if ctx.cregs:
            ctx.store_shot()
        return ctx.qubits, ctx.shots_result.most_common()[0][0]

    @staticmethod
    def samples(gates, n_qubits, shots, initial=None):                 # :669-682
        if initial is not None:
            initial = check_initial(initial, n_qubits)
        ctx = Ctx(n_qubits)
        out = []
        for _ in range(shots):
            run_inner(ctx, gates, initial)
            out.append(ctx.sample)
        return out

    def make_cache(self, gates, n_qubits):                             # :684-685
        self.run(gates, n_qubits, save_cache=True)


# ---------------------------------------------------------------------------------------------------
# Pauli expectation oracle: the in-tree matrix route  Re <psi| sum_k c_k P_k |psi>
# (reference: pauli.py:893-935 `Expr.to_matrix` + vqe.py:355-366 `sparse_expectation`), restated
# matrix-free: P|psi> is a bit-flip gather times a sign/phase vector.
# ---------------------------------------------------------------------------------------------------

def pauli_apply(letters, psi: np.ndarray) -> np.ndarray:
    """(product of Pauli letters [(letter, qubit), ...]) applied to psi."""
    n = psi.size.bit_length() - 1
    idx = np.arange(psi.size, dtype=np.int64)
    out = psi
    flip = 0
    phase = np.ones(psi.size, dtype=complex)
    for letter, q in letters:
        bit = (idx >> q) & 1
        if letter == "X":
            flip |= 1 << q
        elif letter == "Y":                      # Y|0> = i|1>, Y|1> = -i|0>: factor by the *source* bit
            flip |= 1 << q
            phase *= np.where(bit == 0, 1j, -1j)
        elif letter == "Z":
            phase *= np.where(bit == 0, 1.0, -1.0)
        elif letter != "I":
            raise ValueError(letter)
    assert n >= 0
    res = np.empty_like(psi)
    res[idx ^ flip] = phase * out
    return res


def expectation(terms, psi: np.ndarray) -> float:
    """terms = [(coeff, [(letter, qubit), ...]), ...] -> Re <psi|H|psi>."""
    acc = 0.0 + 0.0j
    for coeff, letters in terms:
        acc += coeff * np.vdot(psi, pauli_apply(letters, psi))
    return float(acc.real)



def expect_marginal(psi: np.ndarray, meas) -> dict:                     # vqe.py:231-252 (`expect`)
    """{outcome tuple over `meas`: probability}, as blueqat's non-sampling VQE sampler computes it from a
    statevector: amplitudes of probability exactly 0 are skipped, so outcomes that never occur have no key."""
    meas = tuple(int(m) for m in meas)
    mask = 0
    for m in meas:
        mask |= 1 << m
    prob = psi.real ** 2 + psi.imag ** 2
    out: dict = {}
    for i in np.nonzero(prob)[0]:
        k = int(i) & mask
        key = tuple(1 if k & (1 << m) else 0 for m in meas)
        out[key] = out.get(key, 0.0) + float(prob[i])
    return out
