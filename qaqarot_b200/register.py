"""Register backend "b200" with the host mirror and, when it is importable, with blueqat itself.

blueqat side: ``BlueqatGlobalSetting.register_backend(name, factory)`` (blueqat/circuit.py:371-394) stores a
zero-argument callable returning a ``Backend``; afterwards ``Circuit(...).run(backend="b200")`` and
``circuit.run_with_b200(...)`` (circuit.py:73-77) reach ``B200Backend.run``.
"""
from __future__ import annotations

from .backend import B200Backend
from .circuit import BACKENDS, BackendRegistry
from .transpile import FuseTranspiler

NAME = "b200"
FUSE_NAME = "b200_fuse"          # transpiler backend: run() returns the planner's normal form as a circuit


def register(name: str = NAME, allow_overwrite: bool = True) -> bool:
    """Returns True when blueqat was found and the backend registered there as well."""
    BackendRegistry.register_backend(name, B200Backend, allow_overwrite=True)
    BackendRegistry.register_backend(FUSE_NAME, FuseTranspiler, allow_overwrite=True)
    try:
        from blueqat import BlueqatGlobalSetting
        from blueqat.backends import BACKENDS as BQ
    except Exception:
        return False
    if name in BQ and not allow_overwrite:
        return True
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        BlueqatGlobalSetting.register_backend(name, B200Backend, allow_overwrite=True)
        BlueqatGlobalSetting.register_backend(FUSE_NAME, FuseTranspiler, allow_overwrite=True)
    return True
