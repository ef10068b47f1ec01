"""Minimal Pauli-operator algebra: the input type of ``run(hamiltonian=...)``.

Mirrors the part of blueqat.pauli a backend reads (reference: blueqat/pauli.py:326-371 for X/Y/Z/I,
:374-582 Term, :667-891 Expr): ``X[i] Y[i] Z[i] I``, products and sums with numbers, ``Term.ops`` (each
with ``.op`` in "XYZ" and ``.n``), ``Term.coeff``, ``Expr.terms``, ``to_expr()``, ``simplify()``,
``max_n()``.  ``B200Backend`` accepts these objects and blueqat's own interchangeably (duck typed through
``canonical_terms``).  The algebra is kept deliberately small: a Term is a coefficient times a
{qubit: letter} dictionary, multiplication is a dictionary merge with the single-qubit product table.
"""
from __future__ import annotations

from numbers import Number
from typing import Dict, Iterable, List, Tuple

# single-qubit products: (left, right) -> (phase, letter or "")
_PROD = {}
for _a, _b, _c in (("X", "Y", "Z"), ("Y", "Z", "X"), ("Z", "X", "Y")):
    _PROD[_a, _b] = (1j, _c)
    _PROD[_b, _a] = (-1j, _c)
for _a in "XYZ":
    _PROD[_a, _a] = (1.0, "")


class Pauli:
    """A single Pauli letter on one qubit."""
    __slots__ = ("op", "n")

    def __init__(self, op: str, n: int):
        self.op, self.n = op, int(n)

    def to_term(self) -> "Term":
        return Term({self.n: self.op} if self.op != "I" else {}, 1.0)

    def to_expr(self) -> "Expr":
        return self.to_term().to_expr()

    def __repr__(self):
        return "I" if self.op == "I" else f"{self.op}[{self.n}]"

    def __eq__(self, other):
        return isinstance(other, Pauli) and (self.op, self.n) == (other.op, other.n)

    def __hash__(self):
        return hash((self.op, self.n))

    # arithmetic is delegated to Term
    def __mul__(self, o): return self.to_term() * o
    def __rmul__(self, o): return _as_term(o) * self.to_term()
    def __add__(self, o): return self.to_expr() + o
    def __radd__(self, o): return _as_expr(o) + self.to_expr()
    def __sub__(self, o): return self.to_expr() - o
    def __rsub__(self, o): return _as_expr(o) - self.to_expr()
    def __neg__(self): return -self.to_term()
    def __truediv__(self, o): return self.to_term() / o


class _Letter:
    """`X[3]` / `X(3)` constructor."""

    def __init__(self, op: str):
        self._op = op

    def __getitem__(self, n: int) -> Pauli:
        return Pauli(self._op, n)

    __call__ = __getitem__


X, Y, Z = _Letter("X"), _Letter("Y"), _Letter("Z")
I = Pauli("I", 0)


class Term:
    """coeff * product of Pauli letters on distinct qubits (always stored simplified)."""
    __slots__ = ("_letters", "coeff")

    def __init__(self, letters: Dict[int, str], coeff=1.0):
        self._letters = dict(sorted(letters.items()))
        self.coeff = coeff

    @property
    def ops(self) -> Tuple[Pauli, ...]:
        return tuple(Pauli(c, n) for n, c in self._letters.items())

    def key(self) -> Tuple[Tuple[int, str], ...]:
        return tuple(self._letters.items())

    def to_term(self): return self
    def to_expr(self): return Expr([self])
    def simplify(self): return self

    def max_n(self) -> int:
        return max(self._letters, default=-1)

    @property
    def n_qubits(self) -> int:
        return self.max_n() + 1

    def __mul__(self, other):
        if isinstance(other, Expr):
            return self.to_expr() * other
        other = _as_term(other)
        letters, coeff = dict(self._letters), self.coeff * other.coeff
        for n, c in other._letters.items():
            if n in letters:
                ph, res = _PROD[letters[n], c]
                coeff = coeff * ph
                if res:
                    letters[n] = res
                else:
                    del letters[n]
            else:
                letters[n] = c
        if isinstance(coeff, complex) and coeff.imag == 0:
            coeff = coeff.real
        return Term(letters, coeff)

    def __rmul__(self, other): return _as_term(other) * self
    def __truediv__(self, other): return Term(self._letters, self.coeff / other)
    def __neg__(self): return Term(self._letters, -self.coeff)
    def __add__(self, o): return self.to_expr() + o
    def __radd__(self, o): return _as_expr(o) + self.to_expr()
    def __sub__(self, o): return self.to_expr() - o
    def __rsub__(self, o): return _as_expr(o) - self.to_expr()

    def __repr__(self):
        body = "*".join(repr(p) for p in self.ops) or "I"
        return body if self.coeff == 1 else f"{self.coeff}*{body}"


class Expr:
    """Sum of Terms."""
    __slots__ = ("terms",)

    def __init__(self, terms: Iterable[Term]):
        self.terms = tuple(terms)

    def to_expr(self): return self
    def __iter__(self): return iter(self.terms)

    def simplify(self) -> "Expr":
        acc: Dict[tuple, complex] = {}
        for t in self.terms:
            acc[t.key()] = acc.get(t.key(), 0.0) + t.coeff
        return Expr(Term(dict(k), v) for k, v in sorted(acc.items(), key=lambda kv: repr(kv[0])) if v)

    def max_n(self) -> int:
        return max((t.max_n() for t in self.terms), default=-1)

    @property
    def n_qubits(self) -> int:
        return self.max_n() + 1

    def __add__(self, o): return Expr(self.terms + _as_expr(o).terms)
    def __radd__(self, o): return Expr(_as_expr(o).terms + self.terms)
    def __neg__(self): return Expr(-t for t in self.terms)
    def __sub__(self, o): return self + (-_as_expr(o))
    def __rsub__(self, o): return _as_expr(o) + (-self)

    def __mul__(self, o):
        if isinstance(o, Number):
            return Expr(t * o for t in self.terms)
        return Expr(a * b for a in self.terms for b in _as_expr(o).terms)

    def __rmul__(self, o):
        if isinstance(o, Number):
            return Expr(o * t for t in self.terms)
        return _as_expr(o) * self

    def __truediv__(self, o): return Expr(t / o for t in self.terms)

    def __repr__(self):
        return " + ".join(repr(t) for t in self.terms) or "0"


def _as_term(x) -> Term:
    if isinstance(x, Term):
        return x
    if isinstance(x, Pauli):
        return x.to_term()
    if isinstance(x, Number):
        return Term({}, x)
    raise TypeError(f"cannot use {type(x).__name__} as a Pauli term")


def _as_expr(x) -> Expr:
    if isinstance(x, Expr):
        return x
    return _as_term(x).to_expr()


def canonical_terms(hamiltonian) -> List[Tuple[complex, List[Tuple[str, int]]]]:
    """Any Pauli / Term / Expr (this module's or blueqat's) -> [(coeff, [(letter, qubit), ...]), ...].

    Follows the convention of the reference's only consumer of ``hamiltonian=`` (quimb.py:36-72):
    the expression is simplified first, identity factors are dropped.
    """
    if isinstance(hamiltonian, Number):
        return [(hamiltonian, [])]
    if not hasattr(hamiltonian, "to_expr"):
        raise ValueError(f"`hamiltonian` must be a Pauli expression, but {type(hamiltonian)}")
    expr = hamiltonian.to_expr().simplify()
    out = []
    for term in expr.terms:
        letters = [(p.op, int(p.n)) for p in term.ops if p.op != "I"]
        out.append((term.coeff, letters))
    return out
