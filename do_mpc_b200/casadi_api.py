"""The slice of the `casadi` namespace that do-mpc model / controller templates use.

`from do_mpc_b200.casadi_api import *` plays the role of `from casadi import *` in the reference's
example templates (e.g. examples/CSTR/template_model.py:24-25); `compat/casadi` re-exports it
under the original module name so those templates run unchanged.
"""
from .sym.sx import (SX, DM, Function, vertcat, horzcat, veccat, vec, reshape, sum1, sum2, sumsqr, norm_2, mtimes, dot,
                     transpose, diag, jacobian, gradient, hessian, tangent, substitute, depends_on, symvar,
                     collocation_points, if_else, exp, log, sqrt, sin, cos, tan, tanh, fabs, atan, asin, acos, sinh,
                     cosh, sign, floor, ceil, arctan, arcsin, arccos, fmin, fmax, atan2, arctan2, power)
from .sym.struct import entry, struct_symSX, struct_SX, struct_symMX, struct_MX, SymStruct, NumStruct

def vertsplit(x, incr=1):
    """casadi.vertsplit: the rows of x in blocks of `incr` (list of DM / SX column blocks)."""
    import numpy as _np
    if isinstance(x, SX):
        return [x[i:i + incr, :] for i in range(0, x.shape[0], incr)]
    a = _np.asarray(x.full() if hasattr(x, "full") else x, dtype=float)
    a = a.reshape(a.shape[0], -1)
    return [DM(a[i:i + incr]) for i in range(0, a.shape[0], incr)]


def horzsplit(x, incr=1):
    import numpy as _np
    if isinstance(x, SX):
        return [x[:, i:i + incr] for i in range(0, x.shape[1], incr)]
    a = _np.asarray(x.full() if hasattr(x, "full") else x, dtype=float)
    a = a.reshape(a.shape[0], -1)
    return [DM(a[:, i:i + incr]) for i in range(0, a.shape[1], incr)]


MX = SX
inf = float("inf")
pi = 3.141592653589793

__all__ = [n for n in dir() if not n.startswith("_")]
