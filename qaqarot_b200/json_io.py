"""blueqat's JSON circuit schema (wire formats) read and written without blueqat.

Mirrors ``blueqat/circuit_funcs/json_serializer.py:57-111`` (schema ``blueqat-circuit``, versions "1" and "2") and the
flattening it applies first (``circuit_funcs/flatten.py:5-32``): the documents this module writes are the ones the
reference's ``serialize`` writes for the same circuit, and every document the reference writes loads here into a
circuit with the same operations.  ``tools/run_json.py`` is the host tool around it: a circuit file in, statevector /
shots / energy from the B200 backend out.
"""
from __future__ import annotations

from typing import Any, Dict, List

from .circuit import Circuit, Op, make

SCHEMA_NAME = "blueqat-circuit"
AVAILABLE_SCHEMA_VERSIONS = ["1", "2"]
LATEST_SCHEMA_VERSION = "2"


def flatten(c: Circuit) -> Circuit:
    """Slices and multiple targets expanded into one operation per target (flatten.py:5-32); a keyed measurement stays
    one operation over the tuple of its targets."""
    n = c.n_qubits
    ops: List[Op] = []
    for op in c.ops:
        kind, arity = op.spec.kind, op.spec.arity
        if kind == "measure":
            if op.key is None:
                ops += [make(op.lowername, t, *op.params) for t in op.target_iter(n)]
            else:
                options = {"key": op.key}
                if op.duplicated is not None:
                    options["duplicated"] = op.duplicated
                ops.append(make(op.lowername, tuple(op.target_iter(n)), *op.params, **options))
        elif kind == "reset" or (kind == "gate" and arity == 1):
            ops += [make(op.lowername, t, *op.params) for t in op.target_iter(n)]
        elif kind == "gate" and arity == 2:
            ops += [make(op.lowername, t, *op.params) for t in op.control_target_iter(n)]
        else:
            raise ValueError(f"Cannot process operation {op.lowername}.")
    return Circuit(n, ops)


def serialize(c: Circuit) -> Dict[str, Any]:
    """json_serializer.py:57-91."""
    def one(op: Op) -> Dict[str, Any]:
        targets = op.targets
        if isinstance(targets, slice):
            raise TypeError("Not flatten circuit.")
        if isinstance(targets, int):
            targets = [targets]
        if isinstance(targets, tuple):
            targets = list(targets)
        options: Dict[str, Any] = {}
        if op.spec.kind == "measure":
            if op.key is not None:
                options["key"] = op.key
            if op.duplicated is not None:
                options["duplicated"] = op.duplicated
        return {"name": str(op.lowername), "params": [float(p) for p in op.params], "options": options,
                "targets": targets}

    c = flatten(c)
    return {"schema": {"name": SCHEMA_NAME, "version": LATEST_SCHEMA_VERSION}, "n_qubits": c.n_qubits,
            "ops": [one(op) for op in c.ops]}


def deserialize(data: Dict[str, Any]) -> Circuit:
    """json_serializer.py:94-111.  Version "1" documents have no `options`."""
    schema = data.get("schema", {})
    if schema.get("name", "") != SCHEMA_NAME:
        raise ValueError("Invalid schema")
    if schema.get("version", "") not in AVAILABLE_SCHEMA_VERSIONS:
        raise ValueError("Unknown schema version")
    ops = [make(d["name"], tuple(d["targets"]), *(float(p) for p in d["params"]), **(d.get("options") or {}))
           for d in data["ops"]]
    return Circuit(data["n_qubits"], ops)
