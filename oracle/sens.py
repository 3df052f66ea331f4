"""CPU restatement of the reference's NLP differentiator — TEST INFRASTRUCTURE ONLY (see oracle/README.md).

Follows do_mpc/differentiator/_nlpdifferentiator.py:
  :178-229  unused symbolic variables are removed (here: FlatNLP.used_mask),
  :289-304  A = d(grad_z L)/dz with z = (x, lam_g, lam_x), L = f + lam_g'g + lam_x'x ;  B = d(grad_z L)/dp,
  :345-375  active set from the primal solution, |g - bound| <= active_set_tol / |x - bound| <= active_set_tol,
  :445-465  rows/columns of inactive constraints deleted,
  :469-505  A_red S = -B_red (sparse LU, scipy, like the reference's default lin_solver),
  :555-590  S mapped back to dx/dp (n_w x n_p) and dlam/dp ((n_g + n_w) x n_p).
The reference evaluates A and B from CasADi symbolics; this restatement uses FlatNLP's exact Hessian / Jacobian for A and
central differences in p for B (the parameter derivatives are not part of FlatNLP).  Parity unpinned: the reference
ships no golden sensitivities and CasADi is not available here; the independent check is a finite difference of the
*solution map* itself (tests/test_sensitivities_hostsim.py)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spl


def _B_matrix(nlp, x, lam_g, p, cols, h_rel=1e-6):
    n_w, n_g = x.size, lam_g.size
    B = np.zeros((n_w + n_g, len(cols)))

    def F(pp):
        return np.concatenate([nlp.grad_f(x, pp) + sp.csr_matrix(nlp.jac_g(x, pp)).T @ lam_g, nlp.g(x, pp)])
    for k, j in enumerate(cols):
        h = h_rel * max(1.0, abs(p[j]))
        pp = p.copy(); pp[j] += h; Fp = F(pp)
        pp[j] -= 2 * h; Fm = F(pp)
        B[:, k] = (Fp - Fm) / (2 * h)
    return B


def nlp_sensitivity(nlp, x, lam_g, p, lbx, ubx, lbg, ubg, active_set_tol=1e-6, cols=None):
    """Returns dxdp (n_w, n_cols) and dlamdp (n_g + n_w, n_cols) in the variables of the (scaled) NLP; cols = opt_p indices
    (default all)."""
    x, lam_g, p = (np.asarray(v, dtype=float).ravel() for v in (x, lam_g, p))
    used = nlp.used_mask()
    iu = np.where(used)[0]
    n_w, n_g = x.size, lam_g.size
    H = sp.csr_matrix(nlp.hess_l(x, p, lam_g))
    J = sp.csr_matrix(nlp.jac_g(x, p))
    g = nlp.g(x, p)
    cols = list(range(p.size)) if cols is None else list(cols)
    B = _B_matrix(nlp, x, lam_g, p, cols)
    act_g = np.where((np.abs(g - lbg) <= active_set_tol) | (np.abs(g - ubg) <= active_set_tol))[0]
    xa = used & ((np.abs(x - lbx) <= active_set_tol) | (np.abs(x - ubx) <= active_set_tol))
    act_x = np.where(xa)[0]
    nu, na, nb = len(iu), len(act_g), len(act_x)
    pos = -np.ones(n_w, dtype=int); pos[iu] = np.arange(nu)
    Hr = H[iu][:, iu]
    Jr = J[act_g][:, iu]
    E = sp.csr_matrix((np.ones(nb), (np.arange(nb), pos[act_x])), shape=(nb, nu))       # bound rows: lam_x' x
    A = sp.bmat([[Hr, Jr.T, E.T], [Jr, None, None], [E, None, None]], format="csc")
    rhs = np.zeros((nu + na + nb, len(cols)))
    rhs[:nu] = -B[iu]
    rhs[nu:nu + na] = -B[n_w + act_g]
    try:
        S = spl.splu(A).solve(rhs)                     # the reference: scipy spsolve on the same reduced matrix (:486-488)
        if not np.all(np.isfinite(S)):
            raise RuntimeError("singular")
    except RuntimeError:
        # the reference returns NaN here unless settings.lstsq_fallback (:506-511); dense least squares only for small systems
        if A.shape[0] > 4000:
            raise np.linalg.LinAlgError("reduced KKT matrix of the reference algorithm is singular (LICQ violated)")
        S = np.linalg.lstsq(A.toarray(), rhs, rcond=None)[0]
    dxdp = np.zeros((n_w, len(cols))); dxdp[iu] = S[:nu]
    dlam = np.zeros((n_g + n_w, len(cols)))
    dlam[act_g] = S[nu:nu + na]
    dlam[n_g + act_x] = S[nu + na:]
    return dxdp, dlam
