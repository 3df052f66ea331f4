"""Parametric NLP sensitivities of the controller — `do_mpc.differentiator.DoMPCDifferentiator` on the GPU.

Reference: do_mpc/differentiator/_nlpdifferentiator.py (`DoMPCDifferentiator.differentiate()` :838-856 after
`mpc.make_step()`, `sens_num['dxdp', indexf[...], indexf[...]]` :773-790) and helper.py (settings / status).
The linear algebra runs in the solver's CUDA library (csrc/ipm_sens.cuh through `dmb_sens_host`, include/dompc_b200.h):
same active-set reduction and the same linear system as the reference, solved with the structured factorisation of the
interior-point method instead of a sparse LU, for a whole batch of instances at once.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class NLPDifferentiatorSettings:
    """Subset of the reference's settings that has a meaning here (helper.py:30-80)."""
    active_set_tol: float = 1e-6
    check_SC: bool = True
    track_residuals: bool = False       # residuals are reduced on the device (iterative refinement, residual_ratio_max)
    check_LICQ: bool = False            # not evaluated (needs a rank computation of the active constraint Jacobian)
    check_rank: bool = False
    lstsq_fallback: bool = False
    lin_solver: str = "structured"
    set_lam_zero: bool = False
    # extension: measure activity against the bounds relaxed by bound_relax_factor (what the IPM converged to)
    relaxed_bounds_rule: bool = False
    active_weight: float = 1e10


@dataclass
class NLPDifferentiatorStatus:
    LICQ: Optional[bool] = None
    SC: Optional[bool] = None
    residuals: Optional[float] = None
    lse_solved: bool = False
    full_rank: Optional[bool] = None
    sym_KKT: bool = True
    reduced_nlp: bool = True
    indefinite: Optional[bool] = None   # extension: reduced Hessian on the active set is not positive definite


class _IndexF:
    """`casadi.tools.indexf[...]`: a deferred power index."""
    class Key:
        def __init__(self, key): self.key = key

    def __getitem__(self, key):
        return _IndexF.Key(key if isinstance(key, tuple) else (key,))


indexf = _IndexF()


class _SensNum:
    """`nlp_diff.sens_num['dxdp', indexf['_u', 0, 0], indexf['_x0']]` (reference :773-790)."""

    def __init__(self, opt_x, opt_p):
        self._x, self._p = opt_x, opt_p
        self._dxdp = np.zeros((opt_x.size, opt_p.size))

    def __setitem__(self, key, val):
        assert key == "dxdp"
        self._dxdp = np.asarray(val, dtype=float).reshape(self._x.size, self._p.size)

    def __getitem__(self, key):
        if key == "dxdp":
            return self._dxdp
        name, kx, kp = key
        assert name == "dxdp"
        ix = self._x.f_index(kx.key if isinstance(kx, _IndexF.Key) else kx)
        ip = self._p.f_index(kp.key if isinstance(kp, _IndexF.Key) else kp)
        return self._dxdp[np.ix_(np.asarray(ix, dtype=int), np.asarray(ip, dtype=int))]


class DoMPCDifferentiator:
    def __init__(self, optimizer, **kwargs):
        assert optimizer.flags["setup"], "call setup() on the optimizer first"
        if getattr(optimizer.recipe, "n_free0", 0):
            raise NotImplementedError("sensitivities for problems with a free root state (MHE, nl_cons_single_slack)")
        self.optimizer = optimizer
        self.settings = NLPDifferentiatorSettings(**kwargs)
        self.status = NLPDifferentiatorStatus()
        self.x_scaling_factors = optimizer.opt_x_scaling.full().ravel()
        self._sens_num = _SensNum(optimizer._opt_x, optimizer._opt_p)
        self.n_x, self.n_p, self.n_g = optimizer.n_opt_x, optimizer.n_opt_p, optimizer.recipe.n_g

    @property
    def sens_num(self):
        return self._sens_num

    def _call(self, x, lam_g, P, status=None, cols=None, u0_only=False, want_dlam=True):
        st = self.settings
        return self.optimizer.S.sensitivities(x, lam_g, P, status=status, cols=cols, u0_only=u0_only, want_dlam=want_dlam,
                                              active_set_tol=st.active_set_tol, active_weight=st.active_weight,
                                              relaxed_bounds_rule=st.relaxed_bounds_rule)

    def differentiate(self):
        """Sensitivities at the solution of the last make_step(): returns (dx_dp_num [n_opt_x, n_opt_p] of the UNSCALED
        decision variables, dlam_dp_num [n_g + n_opt_x, n_opt_p]) like the reference (:838-856)."""
        o = self.optimizer
        x = o.opt_x_num.full().ravel(); lam_g = np.asarray(o.lam_g_num.full()).ravel(); p = o.opt_p_num.full().ravel()
        r = self._call(x[None], lam_g[None], p[None])
        sst = int(r["sens_status"][0])
        self.status.lse_solved = sst in (0, 1)
        self.status.indefinite = sst == 1
        self.status.full_rank = sst != 2
        dx = r["dxdp"][0] * self.x_scaling_factors[:, None]
        dlam = r["dlamdp"][0]
        if self.settings.check_SC:
            lam = np.concatenate([lam_g, np.asarray(o.lam_x_num.full()).ravel()])
            act = np.abs(dlam).max(axis=1) > 0
            self.status.SC = bool(np.all(np.abs(lam[act]) >= self.settings.active_set_tol)) if act.any() else True
        self._sens_num["dxdp"] = dx
        return dx, dlam

    def differentiate_batch(self, out, cols=None, u0_only=True):
        """Batch counterpart for the result of `mpc.make_step_batch(..., return_full=True)`:
        returns dict(du0dp or dxdp [B, rows, ncols] of the unscaled variables, sens_status [B])."""
        r = self._call(out["x"], out["lam_g"], out["p"], status=out["status"], cols=cols, u0_only=u0_only, want_dlam=False)
        rec = self.optimizer.recipe
        scale = rec.u_scaling if u0_only else self.x_scaling_factors
        return {("du0dp" if u0_only else "dxdp"): r["dxdp"] * scale[None, :, None], "sens_status": r["sens_status"]}
