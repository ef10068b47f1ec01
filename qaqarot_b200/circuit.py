"""Host-side mirror of the slice of blueqat's circuit DSL that feeds the statevector hot path.

blueqat itself is not installed on the GPU box, so the package carries its own small, table-driven
stand-in for the objects that cross the drop-in boundary:

* ``Circuit``   -- method-chain builder, ``run / statevector / shots / oneshot / make_cache / copy / dagger``
                   (reference: blueqat/circuit.py:36-237)
* ``Op``        -- one record type for every operation, exposing the attributes a backend reads from a
                   blueqat ``Operation``: ``lowername, targets, params, target_iter, control_target_iter,
                   fallback, matrix, dagger`` plus ``theta/phi/lam/gamma/mat/key/duplicated``
                   (reference: blueqat/gate.py:17-155, 1220-1268)
* ``BackendRegistry`` -- ``register_backend / unregister_backend / set_default_backend``
                   (reference: blueqat/circuit.py:371-439)

The reference has one class per gate (38 classes); here a gate is a row of ``GATES``.  ``B200Backend``
consumes either these ``Op`` records or genuine blueqat ``Operation`` objects (duck typed), so when blueqat
is importable the backend is registered there too (see ``qaqarot_b200.register``) and this module is only
used to build circuits where blueqat is absent.
"""
from __future__ import annotations

import cmath
import math
import warnings
from typing import Any, Callable, Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np

PI = math.pi

# ----------------------------------------------------------------------------------------------------
# target expansion (reference semantics: gate.py:1283-1332)
# ----------------------------------------------------------------------------------------------------


def _expand_one(arg, length: int) -> Iterator[int]:
    if isinstance(arg, slice):
        yield from range(*arg.indices(length))
        return
    try:
        i = arg.__index__()
    except AttributeError:
        raise TypeError("indices must be integers or slices, not " + type(arg).__name__) from None
    yield i + length if i < 0 else i


def expand_targets(targets, length: int) -> Iterator[int]:
    """int | slice | tuple of those -> qubit indices (negative ints wrap, slices clip to `length`)."""
    if isinstance(targets, tuple):
        for a in targets:
            yield from _expand_one(a, length)
    else:
        yield from _expand_one(targets, length)


def expand_pairs(targets, length: int) -> List[Tuple[int, int]]:
    """(controls, targets) -> zipped (control, target) pairs with the reference's validation."""
    if not isinstance(targets, tuple) or len(targets) != 2:
        raise ValueError("Control and target qubits pair(s) are required.")
    cs = list(expand_targets(targets[0], length))
    ts = list(expand_targets(targets[1], length))
    if len(cs) != len(ts):
        raise ValueError("The number of control qubits and target qubits are must be same.")
    if any(c == t for c, t in zip(cs, ts)):
        raise ValueError("Control qubit and target qubit are must be different.")
    return list(zip(cs, ts))


def highest_index(targets) -> int:
    """Largest qubit index named by `targets` (used to grow Circuit.n_qubits; gate.py:1335-1351)."""
    def one(t):
        if isinstance(t, slice):
            lo = -1 if t.start is None else t.start.__index__()
            hi = 0 if t.stop is None else t.stop.__index__()
            return max(lo, hi - 1)
        return t.__index__()
    if isinstance(targets, tuple):
        return max((one(t) for t in targets), default=-1)
    return one(targets)


# ----------------------------------------------------------------------------------------------------
# gate table
# ----------------------------------------------------------------------------------------------------


def u_matrix(theta: float, phi: float, lam: float, gamma: float = 0.0) -> np.ndarray:
    """U(theta, phi, lam, gamma) as blueqat defines it (gate.py:554-600)."""
    c, s = math.cos(0.5 * theta), math.sin(0.5 * theta)
    return np.array([[c, -cmath.exp(1j * lam) * s],
                     [cmath.exp(1j * phi) * s, cmath.exp(1j * (phi + lam)) * c]],
                    dtype=complex) * cmath.exp(1j * gamma)


def _controlled(m: np.ndarray) -> np.ndarray:
    """4x4 matrix in blueqat's |t c> ordering (control = low bit): rows/cols 1 and 3 carry the 2x2."""
    out = np.eye(4, dtype=complex)
    out[1, 1], out[1, 3], out[3, 1], out[3, 3] = m[0, 0], m[0, 1], m[1, 0], m[1, 1]
    return out


class GateSpec:
    """One row of the gate table."""

    def __init__(self, name: str, arity: int, pnames: Sequence[str] = (), pdefaults: Sequence[float] = (),
                 matrix: Optional[Callable[..., np.ndarray]] = None,
                 decompose: Optional[Callable[..., List["Op"]]] = None,
                 inverse: Optional[Callable[["Op"], "Op"]] = None,
                 kind: str = "gate", accepts_options: bool = False):
        self.name = name
        self.arity = arity                    # 1: target list, 2: (controls, targets) pairs, 3: fixed triple
        self.pnames = tuple(pnames)
        self.pdefaults = tuple(pdefaults)
        self.matrix = matrix
        self.decompose = decompose            # per single target / pair / triple -> list of Ops
        self.inverse = inverse
        self.kind = kind                      # "gate" | "measure" | "reset"
        self.accepts_options = accepts_options


class Op:
    """A circuit operation.  Quacks like blueqat.gate.Operation for everything a backend touches."""

    __slots__ = ("spec", "lowername", "targets", "params", "options", "key", "duplicated")

    def __init__(self, spec: GateSpec, targets, params: tuple = (), options: Optional[dict] = None):
        self.spec = spec
        self.lowername = spec.name
        self.targets = targets
        self.params = tuple(params)
        self.options = options
        self.key = None
        self.duplicated = None
        if spec.kind == "measure":
            opts = options or {}
            self.key = None if opts.get("key") is None else str(opts["key"])
            self.duplicated = None if opts.get("duplicated") is None else str(opts["duplicated"])

    # -- attribute view of params (theta / phi / lam / gamma / mat) --
    def __getattr__(self, item):
        spec = object.__getattribute__(self, "spec")
        if item in spec.pnames:
            return self.params[spec.pnames.index(item)]
        raise AttributeError(item)

    @property
    def uppername(self) -> str:
        return self.lowername.upper()

    @property
    def n_qargs(self) -> int:
        return self.spec.arity

    def target_iter(self, n_qubits: int) -> Iterator[int]:
        return expand_targets(self.targets, n_qubits)

    def control_target_iter(self, n_qubits: int):
        return iter(expand_pairs(self.targets, n_qubits))

    def matrix(self) -> np.ndarray:
        if self.spec.matrix is None:
            raise NotImplementedError(f"{self.lowername} has no matrix")
        return self.spec.matrix(*self.params)

    def dagger(self) -> "Op":
        if self.spec.kind != "gate":
            raise ValueError(f"{self.lowername} has no Hermitian conjugate")
        if self.spec.inverse is None:
            return self
        return self.spec.inverse(self)

    def fallback(self, n_qubits: int) -> List["Op"]:
        """Decomposition into simpler operations (reference: IFallbackOperation, gate.py:85-91)."""
        dec = self.spec.decompose
        if dec is None:
            raise NotImplementedError(f"fallback of {self.lowername} is not defined.")
        out: List[Op] = []
        if self.spec.arity == 1:
            for t in self.target_iter(n_qubits):
                out += dec(self, t)
        elif self.spec.arity == 2:
            for c, t in self.control_target_iter(n_qubits):
                out += dec(self, c, t)
        else:
            out += dec(self, *self.targets)
        return out

    def __repr__(self) -> str:
        def show(t):
            if isinstance(t, slice):
                s = "" if t.start is None else str(t.start)
                e = "" if t.stop is None else str(t.stop)
                return f"{s}:{e}" if t.step is None else f"{s}:{e}:{t.step}"
            return str(t)
        tg = ", ".join(show(t) for t in self.targets) if isinstance(self.targets, tuple) else show(self.targets)
        ps = "(" + ", ".join(str(p) for p in self.params) + ")" if self.params else ""
        return f"{self.lowername}{ps}[{tg}]"


GATES: Dict[str, GateSpec] = {}
ALIASES: Dict[str, str] = {"r": "phase", "p": "phase", "cr": "cphase", "cp": "cphase", "cnot": "cx",
                           "toffoli": "ccx", "m": "measure"}


def make(name: str, targets, *params, **options) -> Op:
    """Create an operation by (possibly aliased) name."""
    spec = GATES.get(ALIASES.get(name, name))
    if spec is None:
        raise ValueError(f"Unknown operation `{name}'.")
    if options and not spec.accepts_options:
        raise ValueError(f"{name} doesn't take options")
    n_req = len(spec.pnames) - len(spec.pdefaults)
    if not n_req <= len(params) <= len(spec.pnames):
        if not spec.pnames:
            raise ValueError(f"{name} doesn't take parameters")
        raise TypeError(f"{name} takes {n_req}..{len(spec.pnames)} parameters, got {len(params)}")
    full = tuple(params) + spec.pdefaults[len(params) - n_req:]
    return Op(spec, targets, full, options or None)


def _row(name, arity, **kw):
    GATES[name] = GateSpec(name, arity, **kw)


def _neg(name):
    return lambda op: make(name, op.targets, *[-p for p in op.params])


def _to(name):
    return lambda op: make(name, op.targets, *op.params)


def _via_u(*uparams):
    """Decomposition through U with fixed parameters."""
    return lambda op, t: [make("u", t, *uparams)]


def _via_cu(*cuparams):
    return lambda op, c, t: [make("cu", (c, t), *cuparams)]


_S2 = 1 / math.sqrt(2)
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
_Z = np.array([[1, 0], [0, -1]], dtype=complex)
_H = np.array([[1, 1], [1, -1]], dtype=complex) * _S2


def _rx(th):
    c, s = math.cos(th / 2), math.sin(th / 2)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=complex)


def _ry(th):
    c, s = math.cos(th / 2), math.sin(th / 2)
    return np.array([[c, -s], [s, c]], dtype=complex)


def _rz(th):
    return np.array([[cmath.exp(-0.5j * th), 0], [0, cmath.exp(0.5j * th)]], dtype=complex)


def _ph(th):
    return np.array([[1, 0], [0, cmath.exp(1j * th)]], dtype=complex)


def _perm(n, mapping):
    m = np.zeros((n, n), dtype=complex)
    for src in range(n):
        m[mapping.get(src, src), src] = 1
    return m


# one-qubit gates ------------------------------------------------------------------------------------
_row("i", 1, matrix=lambda: np.eye(2, dtype=complex), decompose=lambda op, t: [])
_row("x", 1, matrix=lambda: _X.copy())
_row("y", 1, matrix=lambda: _Y.copy())
_row("z", 1, matrix=lambda: _Z.copy())
_row("h", 1, matrix=lambda: _H.copy())
_row("s", 1, matrix=lambda: np.diag([1, 1j]), inverse=_to("sdg"),
     decompose=lambda op, t: [make("phase", t, PI / 2)])
_row("sdg", 1, matrix=lambda: np.diag([1, -1j]), inverse=_to("s"),
     decompose=lambda op, t: [make("phase", t, -PI / 2)])
_row("t", 1, matrix=lambda: _ph(PI / 4), inverse=_to("tdg"),
     decompose=lambda op, t: [make("phase", t, PI / 4)])
_row("tdg", 1, matrix=lambda: _ph(-PI / 4), inverse=_to("t"),
     decompose=lambda op, t: [make("phase", t, -PI / 4)])
_row("sx", 1, matrix=lambda: u_matrix(PI / 2, -PI / 2, PI / 2, PI / 4), inverse=_to("sxdg"),
     decompose=_via_u(PI / 2, -PI / 2, PI / 2, PI / 4))
_row("sxdg", 1, matrix=lambda: u_matrix(-PI / 2, -PI / 2, PI / 2, -PI / 4), inverse=_to("sx"),
     decompose=_via_u(-PI / 2, -PI / 2, PI / 2, -PI / 4))
_row("rx", 1, pnames=("theta",), matrix=_rx, inverse=_neg("rx"),
     decompose=lambda op, t: [make("u", t, op.theta, -PI / 2, PI / 2, 0.0)])
_row("ry", 1, pnames=("theta",), matrix=_ry, inverse=_neg("ry"),
     decompose=lambda op, t: [make("u", t, op.theta, 0.0, 0.0, 0.0)])
_row("rz", 1, pnames=("theta",), matrix=_rz, inverse=_neg("rz"),
     decompose=lambda op, t: [make("u", t, 0.0, 0.0, op.theta, -op.theta / 2)])
_row("phase", 1, pnames=("theta",), matrix=_ph, inverse=_neg("phase"))
_row("u", 1, pnames=("theta", "phi", "lam", "gamma"), pdefaults=(0.0,), matrix=u_matrix,
     inverse=lambda op: make("u", op.targets, -op.theta, -op.lam, -op.phi, -op.gamma))
_row("mat1", 1, pnames=("mat",), matrix=lambda m: m,
     inverse=lambda op: make("mat1", op.targets, np.asarray(op.mat).T.conjugate()))

# two-qubit gates (targets = (controls, targets)) ------------------------------------------------------
_row("cx", 2, matrix=lambda: _controlled(_X))
_row("cz", 2, matrix=lambda: _controlled(_Z))
_row("cy", 2, matrix=lambda: _controlled(_Y), decompose=_via_cu(PI, PI / 2, PI / 2, 0.0))
_row("ch", 2, matrix=lambda: _controlled(_H), decompose=_via_cu(-PI / 2, PI, 0.0, 0.0))
_row("crx", 2, pnames=("theta",), matrix=lambda th: _controlled(_rx(th)), inverse=_neg("crx"),
     decompose=lambda op, c, t: [make("cu", (c, t), op.theta, -PI / 2, PI / 2, 0.0)])
_row("cry", 2, pnames=("theta",), matrix=lambda th: _controlled(_ry(th)), inverse=_neg("cry"),
     decompose=lambda op, c, t: [make("cu", (c, t), op.theta, 0.0, 0.0, 0.0)])
_row("crz", 2, pnames=("theta",), matrix=lambda th: _controlled(_rz(th)), inverse=_neg("crz"),
     decompose=lambda op, c, t: [make("cu", (c, t), 0.0, 0.0, op.theta, -op.theta / 2)])
_row("cphase", 2, pnames=("theta",), matrix=lambda th: _controlled(_ph(th)), inverse=_neg("cphase"),
     decompose=lambda op, c, t: [make("cu", (c, t), 0.0, op.theta, 0.0, 0.0)])


def _cu_network(op, c, t):
    th, ph, la, ga = op.params
    return [make("phase", c, ga + 0.5 * (la + ph)), make("phase", t, 0.5 * (la - ph)), make("cx", (c, t)),
            make("u", t, -0.5 * th, 0.0, -0.5 * (ph + la)), make("cx", (c, t)), make("u", t, 0.5 * th, ph, 0.0)]


_row("cu", 2, pnames=("theta", "phi", "lam", "gamma"), pdefaults=(0.0,),
     matrix=lambda th, ph, la, ga=0.0: _controlled(u_matrix(th, ph, la, ga)),
     inverse=lambda op: make("cu", op.targets, -op.theta, -op.lam, -op.phi, -op.gamma),
     decompose=_cu_network)
_row("swap", 2, matrix=lambda: _perm(4, {1: 2, 2: 1}),
     decompose=lambda op, a, b: [make("cx", (a, b)), make("cx", (b, a)), make("cx", (a, b))])
_row("rzz", 2, pnames=("theta",), inverse=_neg("rzz"),
     matrix=lambda th: np.diag([cmath.exp(-0.5j * th), cmath.exp(0.5j * th),
                                cmath.exp(0.5j * th), cmath.exp(-0.5j * th)]),
     decompose=lambda op, a, b: [make("cx", (a, b)), make("rz", b, op.theta), make("cx", (a, b))])
_row("rxx", 2, pnames=("theta",), inverse=_neg("rxx"),
     decompose=lambda op, a, b: [make("h", a), make("h", b), make("rzz", (a, b), op.theta),
                                 make("h", a), make("h", b)])
_row("ryy", 2, pnames=("theta",), inverse=_neg("ryy"),
     decompose=lambda op, a, b: [make("rx", a, -PI / 2), make("rx", b, -PI / 2),
                                 make("rzz", (a, b), op.theta),
                                 make("rx", a, PI / 2), make("rx", b, PI / 2)])
_row("zz", 2, matrix=lambda: np.diag([1, 1j, 1j, 1]).astype(complex),
     decompose=lambda op, a, b: [make("ry", b, PI / 2), make("cx", (a, b)), make("rz", a, PI / 2),
                                 make("u", b, -PI / 2, PI / 2, 0.0, PI / 4)])

# three-qubit gates (targets = fixed triple, no slicing) -----------------------------------------------
_row("ccz", 3, matrix=lambda: np.diag([1, 1, 1, 1, 1, 1, 1, -1]).astype(complex),
     decompose=lambda op, c1, c2, t: [
         make("cx", (c2, t)), make("tdg", t), make("cx", (c1, t)), make("t", t), make("cx", (c2, t)),
         make("tdg", t), make("cx", (c1, t)), make("t", c2), make("t", t), make("cx", (c1, c2)),
         make("t", c1), make("tdg", c2), make("cx", (c1, c2))])
_row("ccx", 3, matrix=lambda: _perm(8, {3: 7, 7: 3}),
     decompose=lambda op, c1, c2, t: [make("h", t), make("ccz", (c1, c2, t)), make("h", t)])
_row("cswap", 3, matrix=lambda: _perm(8, {3: 5, 5: 3}),
     decompose=lambda op, c, a, b: [make("cx", (b, a)), make("ccx", (c, a, b)), make("cx", (b, a))])

# non-unitary ------------------------------------------------------------------------------------------
_row("measure", 1, kind="measure", accepts_options=True)
_row("reset", 1, kind="reset")


# ----------------------------------------------------------------------------------------------------
# backend registry and Circuit
# ----------------------------------------------------------------------------------------------------

BACKENDS: Dict[str, Callable[[], Any]] = {}
_default_backend = "b200"


class BackendRegistry:
    """Counterpart of BlueqatGlobalSetting's backend functions (circuit.py:371-439)."""

    @staticmethod
    def register_backend(name: str, backend: Callable[[], Any], allow_overwrite: bool = False) -> None:
        if not allow_overwrite and name in BACKENDS:
            raise ValueError(f"Backend '{name}' is already registered as backend.")
        BACKENDS[name] = backend

    @staticmethod
    def unregister_backend(name: str) -> None:
        if name not in BACKENDS:
            raise ValueError(f"Backend '{name}' is not registered.")
        del BACKENDS[name]

    @staticmethod
    def set_default_backend(name: str) -> None:
        if name not in BACKENDS:
            raise ValueError(f"Backend '{name}' is not registered.")
        global _default_backend
        _default_backend = name

    @staticmethod
    def get_default_backend_name() -> str:
        return _default_backend


class _Pending:
    """`circuit.rz(0.3)` -- waits for `[targets]`."""

    def __init__(self, circuit: "Circuit", name: str):
        self.circuit, self.name, self.params, self.options = circuit, name, (), {}

    def __call__(self, *params, **options):
        self.params, self.options = params, options
        return self

    def __getitem__(self, targets) -> "Circuit":
        c = self.circuit
        c.ops.append(make(self.name, targets, *self.params, **self.options))
        c.n_qubits = max(highest_index(targets) + 1, c.n_qubits)
        return c


class Circuit:
    """Store operations and hand them to a backend (reference: blueqat/circuit.py:36-237)."""

    def __init__(self, n_qubits: int = 0, ops: Optional[List[Op]] = None):
        self.ops: List[Op] = ops or []
        self.n_qubits = n_qubits
        self._backends: Dict[str, Any] = {}

    def __getattr__(self, name: str):
        if name.startswith("_"):
            raise AttributeError(name)
        if ALIASES.get(name, name) in GATES:
            return _Pending(self, name)
        if name.startswith("run_with_") and name[9:] in BACKENDS:
            be = self._backend(name[9:])
            return lambda *a, **k: be.run(self.ops, self.n_qubits, *a, **k)
        raise AttributeError(f"'circuit' object has no attribute or gate '{name}'")

    def __repr__(self):
        return f"Circuit({self.n_qubits})." + ".".join(repr(o) for o in self.ops)

    def _backend(self, which):
        if which is None:
            which = _default_backend
        if not isinstance(which, str):
            return which
        if which not in self._backends:
            factory = BACKENDS.get(which)
            if factory is None:
                raise ValueError(f"Backend {which} doesn't exist.")
            self._backends[which] = factory()
        return self._backends[which]

    def __iadd__(self, other):
        if not isinstance(other, Circuit):
            return NotImplemented
        self.ops += other.ops
        self.n_qubits = max(self.n_qubits, other.n_qubits)
        return self

    def __add__(self, other):
        if not isinstance(other, Circuit):
            return NotImplemented
        c = self.copy()
        c += other
        return c

    def copy(self, copy_backends: bool = True) -> "Circuit":
        c = Circuit(self.n_qubits, list(self.ops))
        if copy_backends:
            c._backends = {k: v.copy() for k, v in self._backends.items()}
        return c

    def dagger(self, ignore_measurement: bool = False) -> "Circuit":
        ops = []
        for g in reversed(self.ops):
            try:
                ops.append(g.dagger())
            except ValueError:
                if not ignore_measurement:
                    raise ValueError("Cannot make the Hermitian conjugate of this circuit because "
                                     "the circuit contains measurement.") from None
        return Circuit(self.n_qubits, ops)

    def run(self, *args, backend=None, **kwargs):
        return self._backend(backend).run(self.ops, self.n_qubits, *args, **kwargs)

    def make_cache(self, backend=None) -> None:
        return self._backend(backend).make_cache(self.ops, self.n_qubits)

    def statevector(self, backend=None, **kwargs) -> np.ndarray:
        if kwargs.get("returns"):
            raise ValueError("Circuit.statevector has no argument `returns`.")
        be = self._backend(backend)
        if hasattr(be, "statevector"):
            return be.statevector(self.ops, self.n_qubits, **kwargs)
        return be.run(self.ops, self.n_qubits, returns="statevector", **kwargs)

    def shots(self, shots: int, backend=None, **kwargs):
        if kwargs.get("returns"):
            raise ValueError("Circuit.shots has no argument `returns`.")
        be = self._backend(backend)
        if hasattr(be, "shots"):
            return be.shots(self.ops, self.n_qubits, shots=shots, **kwargs)
        return be.run(self.ops, self.n_qubits, shots=shots, returns="shots", **kwargs)

    def oneshot(self, backend=None, **kwargs):
        if kwargs.get("returns"):
            raise ValueError("Circuit.oneshot has no argument `returns`.")
        be = self._backend(backend)
        if hasattr(be, "oneshot"):
            return be.oneshot(self.ops, self.n_qubits, **kwargs)
        v, cnt = be.run(self.ops, self.n_qubits, shots=1, returns="statevector_and_shots", **kwargs)
        return v, cnt.most_common()[0][0]
