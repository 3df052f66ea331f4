"""do_mpc_b200 — B200-native batched nonlinear MPC behind the do-mpc Python API.

Public surface mirrors `do_mpc` for the controller path (reference do_mpc/__init__.py:103-128):
`do_mpc_b200.model.Model`, `do_mpc_b200.controller.MPC`, `do_mpc_b200.data`.  The NLP solve
(`nlpsol('ipopt')` in the reference) runs as hand-written CUDA for sm_100a (do_mpc_b200/csrc) behind
a ctypes C-ABI (include/dompc_b200.h); there is no CPU fallback.
"""
from . import sym
from . import model as _model_mod
from . import data
from . import controller as _controller_mod
from . import casadi_api
from . import sampling
from . import approximateMPC
from . import rollout
from . import tools
from . import differentiator as _diff_mod
from . import simulator as _sim_mod
from . import estimator as _est_mod

__version__ = "0.1.0"


class _NS:
    def __init__(self, **kw):
        self.__dict__.update(kw)


model = _NS(Model=_model_mod.Model)
controller = _NS(MPC=_controller_mod.MPC, MPCSettings=_controller_mod.MPCSettings)
simulator = _NS(Simulator=_sim_mod.Simulator)
estimator = _NS(StateFeedback=_est_mod.StateFeedback, Estimator=_est_mod.Estimator, MHE=_est_mod.MHE, MHESettings=_est_mod.MHESettings)
differentiator = _NS(DoMPCDifferentiator=_diff_mod.DoMPCDifferentiator, NLPDifferentiatorSettings=_diff_mod.NLPDifferentiatorSettings,
                     NLPDifferentiatorStatus=_diff_mod.NLPDifferentiatorStatus, indexf=_diff_mod.indexf)
