"""`do_mpc` look-alike exposing the controller path of do_mpc_b200 under the reference's module names
(reference do_mpc/__init__.py:103-128).  Subsystems outside the hot path (simulator, estimators, graphics, ...)
are not provided; accessing them raises with a clear message."""
import do_mpc_b200 as _b
from do_mpc_b200 import model, controller, data, simulator, estimator, differentiator, approximateMPC, tools  # noqa: F401

__version__ = _b.__version__


class _Missing:
    def __init__(self, name): self._n = name
    def __getattr__(self, item):
        raise NotImplementedError(f"do_mpc.{self._n}.{item} is outside the scope of do_mpc_b200 (controller hot path only)")


graphics = _Missing("graphics")
from do_mpc_b200 import sampling  # noqa: F401,E402
