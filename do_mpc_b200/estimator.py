"""State-feedback estimator (reference do_mpc/estimator/_base.py:28-71): make_step(y) returns y — needed only so that the
closed loop of the reference's examples/tests (MPC -> Simulator -> StateFeedback) can be written unchanged."""
import numpy as np

from .data import Data


class Estimator:
    def __init__(self, model):
        self.model = model
        self.data = Data(model)
        self._x0 = model._x(0.0)

    @property
    def x0(self): return self._x0
    @x0.setter
    def x0(self, v): self._x0 = self.model._x(np.asarray(v.full() if hasattr(v, "full") else v, dtype=float).ravel())
    def reset_history(self): self.data.init_storage()


from .mhe import MHE, MHESettings  # noqa: E402,F401  (moving horizon estimator: do_mpc_b200/mhe.py)


class StateFeedback(Estimator):
    def make_step(self, y0):
        return y0
