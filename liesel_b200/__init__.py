"""
liesel_b200 — B200-native log-posterior / gradient / IWLS engine behind Liesel's model interface.

Importing the package does not touch the GPU; ``liesel_b200.plan`` / ``liesel_b200.interface``
load ``csrc/libliesel_b200.so`` and raise if it has not been built (there is no fallback).
"""

from . import spec, splines, synth  # noqa: F401  (NumPy-only modules)
from .spec import ModelSpec  # noqa: F401

__version__ = "0.1.0"


def __getattr__(name):
    if name in ("Plan", "chol", "iwls_propose", "iwls_propose_diag", "mh_step"):
        from . import plan as _plan

        return getattr(_plan, name)
    if name == "B200Interface":
        from .interface import B200Interface

        return B200Interface
    raise AttributeError(name)
