"""pytest plugin used by tests/test_reference_suite.py: makes "b200" blueqat's default backend before the
reference's own tests/conftest.py reads it.  The device is the numpy interpreter of the op program
(no GPU in the build container); with B200_REF_SUITE_DEVICE=cuda it is the CUDA library (GPU box, blueqat from
baseline/_ref through B200_REF_ROOT)."""
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.environ.get("B200_REF_ROOT", "/root/reference"))
warnings.simplefilter("ignore", SyntaxWarning)

from blueqat import BlueqatGlobalSetting  # noqa: E402
from oracle.program_interp import OracleState  # noqa: E402
from qaqarot_b200 import B200Backend  # noqa: E402


def _factory():
    if os.environ.get("B200_REF_SUITE_DEVICE") == "cuda":
        return B200Backend()
    return B200Backend(state_factory=lambda n, d: OracleState(n, d))


BlueqatGlobalSetting.register_backend("b200", _factory, allow_overwrite=True)
BlueqatGlobalSetting.set_default_backend("b200")
