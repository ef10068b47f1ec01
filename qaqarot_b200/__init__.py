"""qaqarot_b200 -- B200-native complex128 statevector backend for blueqat (Qaqarot/qaqarot).

``import qaqarot_b200`` registers backend ``"b200"`` (with blueqat when it is importable, and always with the
bundled host mirror), so ``Circuit(...).run(backend="b200")`` is a drop-in for ``run(backend="numpy")``.
The CUDA library is loaded lazily, on the first run; there is no CPU fallback.
"""
from .backend import B200Backend, ShardedB200Backend
from .circuit import BackendRegistry, Circuit
from .pauli import I, X, Y, Z
from .register import register
from . import samplers

__all__ = ["B200Backend", "ShardedB200Backend", "BackendRegistry", "Circuit", "X", "Y", "Z", "I", "register", "samplers"]

register()
