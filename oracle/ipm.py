"""CPU restatement of the solver behind nlpsol('ipopt') (TEST INFRASTRUCTURE, see oracle/README.md).

The reference calls CasADi's nlpsol with plugin 'ipopt' (do_mpc/controller/_mpc.py:1328) and runs it in
do_mpc/optimizer.py:770.  IPOPT itself is a third-party dependency that is NOT in /root/reference
(requirements.txt:1 `casadi>=3.6.0`, bundles IPOPT 3.14.x + MUMPS); its source cannot be consulted here.
This file restates the published algorithm — A. Wächter, L. T. Biegler, "On the implementation of an
interior-point filter line-search algorithm for large-scale nonlinear programming", Math. Program. 106
(2006) — with IPOPT's documented default option values, full-space and sparse, on the flat NLP of
oracle/nlp.py.  It is a *different linear-algebra path* from the product (which eliminates stage by
stage on the GPU): here the whole augmented KKT matrix is factorised with LAPACK dsytrf, whose
Bunch-Kaufman pivots give the exact inertia.

Parity is pinned against the reference's own outputs:
  * goldens testing/results/*.pkl: restarting from the golden point converges in <= 3 iterations to it;
  * shipped 1000-sample nominal-CSTR dataset: cold starts reproduce u0 / success (tests/test_oracle_golden.py).

Deliberate deviations from IPOPT (documented in DESIGN.md §Parity):
  * variables that appear in no constraint and no objective term (the reference's dummy collocation
    slots) are removed before the solve and keep their initial-guess value;
    IPOPT instead keeps them and regularises every iteration with delta_w = 1e-4/3^k — that
    schedule is reproduced here (`mimic_dummy_reg`) so iterates stay close;
  * the restoration phase is replaced by a simpler feasibility restoration (`_restoration`: primal barrier Newton steps
    on the proximal feasibility problem; the CUDA kernel implements the same substitute);
  * no watchdog, no "tiny step" heuristics.
"""
import time
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp
from scipy.linalg import lapack

from .blockldl import BlockLDL, BlockPartition


@dataclass
class IpmOptions:
    tol: float = 1e-8
    max_iter: int = 3000
    dual_inf_tol: float = 1.0
    constr_viol_tol: float = 1e-4
    compl_inf_tol: float = 1e-4
    acceptable_tol: float = 1e-6
    acceptable_iter: int = 15
    acceptable_dual_inf_tol: float = 1e10
    acceptable_constr_viol_tol: float = 1e-2
    acceptable_compl_inf_tol: float = 1e-2
    s_max: float = 100.0
    mu_init: float = 0.1
    mu_min: float = 1e-11
    kappa_mu: float = 0.2
    theta_mu: float = 1.5
    kappa_eps: float = 10.0
    tau_min: float = 0.99
    bound_push: float = 1e-2
    bound_frac: float = 1e-2
    bound_relax_factor: float = 1e-8
    constr_mult_init_max: float = 1e3
    bound_mult_init_val: float = 1.0
    kappa_sigma: float = 1e10
    kappa_d: float = 1e-5
    # filter line search
    alpha_red_factor: float = 0.5
    eta_phi: float = 1e-8
    theta_max_fact: float = 1e4
    theta_min_fact: float = 1e-4
    s_phi: float = 2.3
    s_theta: float = 1.1
    gamma_phi: float = 1e-8
    gamma_theta: float = 1e-5
    alpha_min_frac: float = 0.05
    delta: float = 1.0
    max_soc: int = 4
    kappa_soc: float = 0.99
    # inertia correction
    delta_w_first: float = 1e-4
    delta_w_min: float = 1e-20
    delta_w_max: float = 1e20
    kappa_w_plus_first: float = 100.0
    kappa_w_plus: float = 8.0
    kappa_w_minus: float = 1.0 / 3.0
    delta_c_bar: float = 1e-8
    kappa_c: float = 0.25
    # nlp scaling
    nlp_scaling_max_gradient: float = 100.0
    nlp_scaling_min_value: float = 1e-8
    # restoration substitute
    resto_max_iter: int = 60
    kappa_resto: float = 0.9
    mimic_dummy_reg: bool = True
    verbose: bool = False
    # KKT factorisation: "block" = block LDL^T on the scenario-tree partition (oracle/blockldl.py, exact inertia, sparse),
    # "dense" = one LAPACK dsytrf of the whole matrix (the round-1 path; kept as the cross-check of the block path)
    linear_solver: str = "block"
    # warm start (IPOPT warm_start_init_point = yes; defaults of the warm_start_* options)
    warm_start_init_point: bool = False
    warm_start_bound_push: float = 1e-3
    warm_start_bound_frac: float = 1e-3
    warm_start_mult_bound_push: float = 1e-3
    warm_start_mult_init_max: float = 1e6


def _inertia_from_sytrf(ldu, ipiv):
    n = ldu.shape[0]
    npos = nneg = nzero = 0
    k = 0
    while k < n:
        if ipiv[k] > 0:
            d = ldu[k, k]
            if d > 0: npos += 1
            elif d < 0: nneg += 1
            else: nzero += 1
            k += 1
        else:
            a, b, c = ldu[k, k], ldu[k + 1, k], ldu[k + 1, k + 1]
            det = a * c - b * b
            tr = a + c
            if det < 0: npos += 1; nneg += 1
            elif det > 0:
                if tr > 0: npos += 2
                else: nneg += 2
            else:
                nzero += 1
                if tr > 0: npos += 1
                elif tr < 0: nneg += 1
                else: nzero += 1
            k += 2
    return npos, nneg, nzero


class _KKT:
    """Symmetric indefinite solve with exact inertia (LAPACK dsytrf/dsytrs on the dense lower triangle)."""

    def __init__(self, K):
        A = K.toarray() if sp.issparse(K) else np.array(K)
        self.n = A.shape[0]
        ldu, ipiv, info = lapack.dsytrf(A, lower=1)
        self.ok = (info == 0)
        self.ldu, self.ipiv = ldu, ipiv
        self.inertia = _inertia_from_sytrf(ldu, ipiv) if info >= 0 else (0, 0, self.n)
        if info > 0:
            self.inertia = (self.inertia[0], self.inertia[1], max(self.inertia[2], 1))

    def solve(self, rhs):
        x, info = lapack.dsytrs(self.ldu, self.ipiv, rhs, lower=1)
        return x


class IPMSolver:
    """Solve  min f(x) s.t. lbg<=g(x)<=ubg, lbx<=x<=ubx  for a FlatNLP-like object.

    `nlp` must provide n_w, n_g, lbx, ubx, lbg, ubg and f/g/grad_f/jac_g/hess_l(x, p[, lam, sigma]),
    plus used_mask().
    """

    def __init__(self, nlp, options=None):
        self.nlp = nlp
        self.o = options or IpmOptions()
        o = self.o
        used = nlp.used_mask()
        self.iv = np.where(used)[0]                 # variables that take part in the solve
        n = self.n = self.iv.size
        lbx = nlp.lbx[self.iv].copy(); ubx = nlp.ubx[self.iv].copy()
        # bound relaxation (bound_relax_factor)
        rl = np.minimum(o.bound_relax_factor * np.maximum(1.0, np.abs(lbx)), o.constr_viol_tol)
        ru = np.minimum(o.bound_relax_factor * np.maximum(1.0, np.abs(ubx)), o.constr_viol_tol)
        self.xL = np.where(np.isfinite(lbx), lbx - rl, -np.inf)
        self.xU = np.where(np.isfinite(ubx), ubx + ru, np.inf)
        self.iL = np.where(np.isfinite(self.xL))[0]
        self.iU = np.where(np.isfinite(self.xU))[0]
        eq = (nlp.lbg == nlp.ubg)
        self.ic = np.where(eq)[0]                   # equality rows  c(x) = g - lbg
        self.id = np.where(~eq)[0]                  # inequality rows d_L <= d(x) <= d_U, slack s
        self.m_c, self.m_d = self.ic.size, self.id.size
        self.c_rhs = nlp.lbg[self.ic]
        self.bp = self._partition() if o.linear_solver == "block" else None

    def _partition(self):
        """Block partition of the KKT unknowns [x | s | y_c | y_d] for oracle/blockldl.py: one block per scenario-tree
        node (k, s) holding the node's own variables (_x[k,s,:], _u[k,s], _eps[k,s]) together with the collocation unknowns,
        algebraic states, constraint rows and inequality slacks of the edge that ENDS in that node; leaves are eliminated
        first, the root last.  With that grouping every diagonal block is a small KKT matrix with a square nonsingular
        constraint Jacobian, so the block factorisation never meets a structurally singular pivot block.  Returns None for
        NLP objects without the stage layout (then the dense factorisation is used)."""
        nlp = self.nlp
        if not all(hasattr(nlp, a) for a in ("ix_x", "ix_u", "ix_z", "edges", "n_scenarios")):
            return None
        N = nlp.N
        off = np.concatenate([[0], np.cumsum(nlp.n_scenarios)]).astype(int)
        n_nodes = int(off[-1])
        nid = lambda k, s_: int(off[k] + min(s_, nlp.n_scenarios[k] - 1))
        vb = np.zeros(nlp.n_w, dtype=np.int64)                     # node of every variable (dummies: irrelevant)
        S = nlp.ix_x.shape[1]
        for k in range(N + 1):
            for s_ in range(S):
                vb[nlp.ix_x[k, s_].ravel()] = nid(k, s_)
        for k in range(N):
            for s_ in range(S):
                if nlp.ix_z.size: vb[nlp.ix_z[k, s_].ravel()] = nid(k + 1, s_)
            for s_ in range(nlp.ix_u.shape[1]):
                vb[nlp.ix_u[k, s_]] = nid(k, s_)
        if nlp.ix_eps.size:
            for k in range(nlp.ix_eps.shape[0]):
                for s_ in range(S):
                    vb[nlp.ix_eps[k, s_]] = nid(k, s_)
        if hasattr(nlp, "partition_extra"):
            for idx, (k, s_) in nlp.partition_extra():
                vb[np.asarray(idx)] = nid(k, s_)
        gb = np.zeros(nlp.n_g, dtype=np.int64)                     # node of every constraint row (init rows: root)
        starts = [e[4] for e in nlp.edges] + [nlp.n_g]
        for (k, s_, b, ch, row), nxt in zip(nlp.edges, starts[1:]):
            gb[row:nxt] = nid(k + 1, ch)
        node_of = np.concatenate([vb[self.iv], gb[self.id], gb[self.ic], gb[self.id]])
        return BlockPartition(n_nodes - 1 - node_of)               # deepest nodes first, root last

    def _factor(self, K):
        """KKT factorisation with exact inertia: block LDL^T on the tree partition, dense dsytrf as fallback / cross-check."""
        if self.bp is not None and self.bp.part.size == K.shape[0]:
            kk = BlockLDL(K, self.bp)
            if kk.ok or K.shape[0] > 3000:
                return kk
        return _KKT(K)

    # ---- problem functions on the reduced variable set, scaled ---------------------------------
    def _full(self, x):
        xf = self.x_full.copy(); xf[self.iv] = x
        return xf

    def _eval_f(self, x): return self.df * self.nlp.f(self._full(x), self.p)
    def _eval_grad(self, x): return self.df * self.nlp.grad_f(self._full(x), self.p)[self.iv]
    def _eval_g(self, x):
        g = self.nlp.g(self._full(x), self.p)
        return self.dc * (g[self.ic] - self.c_rhs), self.dd * g[self.id]
    def _eval_jac(self, x):
        J = self.nlp.jac_g(self._full(x), self.p)[:, self.iv]
        return sp.diags(self.dc) @ J[self.ic], sp.diags(self.dd) @ J[self.id]
    def _eval_hess(self, x, yc, yd):
        lam = np.zeros(self.nlp.n_g)
        lam[self.ic] = self.dc * yc; lam[self.id] = self.dd * yd
        H = self.nlp.hess_l(self._full(x), self.p, lam, self.df)
        return H[self.iv][:, self.iv]

    # ---- main entry ---------------------------------------------------------------------------
    def solve(self, x0_full, p, lam_g0=None, lam_x0=None):
        o = self.o
        t_start = time.perf_counter()
        self.p = np.asarray(p, dtype=float)
        self.x_full = np.array(x0_full, dtype=float)
        n, m_c, m_d = self.n, self.m_c, self.m_d
        iL, iU, xL, xU = self.iL, self.iU, self.xL, self.xU
        x = self.x_full[self.iv].copy()

        # -- gradient-based scaling at the user's starting point -----------------------------------
        self.df, self.dc, self.dd = 1.0, np.ones(m_c), np.ones(m_d)
        g0 = np.abs(self._eval_grad(x)).max() if n else 0.0
        Jc0, Jd0 = self._eval_jac(x)
        def rowscale(J, m):
            if m == 0: return np.ones(0)
            mx = np.asarray(abs(J).max(axis=1).todense()).ravel()
            sc = np.where(mx > o.nlp_scaling_max_gradient, o.nlp_scaling_max_gradient / np.maximum(mx, 1e-300), 1.0)
            return np.maximum(sc, o.nlp_scaling_min_value)
        dc, dd = rowscale(Jc0, m_c), rowscale(Jd0, m_d)
        if g0 > o.nlp_scaling_max_gradient:
            self.df = max(o.nlp_scaling_max_gradient / g0, o.nlp_scaling_min_value)
        self.dc, self.dd = dc, dd
        dL = self.dd * self.nlp.lbg[self.id]; dU = self.dd * self.nlp.ubg[self.id]
        rl = np.minimum(o.bound_relax_factor * np.maximum(1.0, np.abs(dL)), o.constr_viol_tol)
        ru = np.minimum(o.bound_relax_factor * np.maximum(1.0, np.abs(dU)), o.constr_viol_tol)
        dL = np.where(np.isfinite(dL), dL - rl, -np.inf); dU = np.where(np.isfinite(dU), dU + ru, np.inf)
        jL = np.where(np.isfinite(dL))[0]; jU = np.where(np.isfinite(dU))[0]
        self.dL, self.dU, self.jL, self.jU = dL, dU, jL, jU

        # -- starting point ---------------------------------------------------------------------
        ws = bool(o.warm_start_init_point) and lam_g0 is not None and lam_x0 is not None
        b_push, b_frac = (o.warm_start_bound_push, o.warm_start_bound_frac) if ws else (o.bound_push, o.bound_frac)

        def push(v, lo, hi):
            v = v.copy()
            both = np.isfinite(lo) & np.isfinite(hi)
            onlyl = np.isfinite(lo) & ~np.isfinite(hi)
            onlyu = ~np.isfinite(lo) & np.isfinite(hi)
            pl = np.minimum(b_push * np.maximum(1.0, np.abs(lo[both])), b_frac * (hi[both] - lo[both]))
            pu = np.minimum(b_push * np.maximum(1.0, np.abs(hi[both])), b_frac * (hi[both] - lo[both]))
            v[both] = np.minimum(np.maximum(v[both], lo[both] + pl), hi[both] - pu)
            v[onlyl] = np.maximum(v[onlyl], lo[onlyl] + b_push * np.maximum(1.0, np.abs(lo[onlyl])))
            v[onlyu] = np.minimum(v[onlyu], hi[onlyu] - b_push * np.maximum(1.0, np.abs(hi[onlyu])))
            return v
        x = push(x, xL, xU)
        c, d = self._eval_g(x)
        s = push(d, dL, dU)
        zL = np.full(iL.size, o.bound_mult_init_val); zU = np.full(iU.size, o.bound_mult_init_val)
        vL = np.full(jL.size, o.bound_mult_init_val); vU = np.full(jU.size, o.bound_mult_init_val)
        self.onlyL = np.isfinite(xL) & ~np.isfinite(xU); self.onlyU = ~np.isfinite(xL) & np.isfinite(xU)
        self.s_onlyL = np.isfinite(dL) & ~np.isfinite(dU); self.s_onlyU = ~np.isfinite(dL) & np.isfinite(dU)

        grad = self._eval_grad(x)
        Jc, Jd = self._eval_jac(x)
        # least-squares multiplier estimate (cold start) / user multipliers (warm start)
        yc, yd = np.zeros(m_c), np.zeros(m_d)
        if ws:
            lam_g0 = np.asarray(lam_g0, dtype=float).ravel(); lam_x0 = np.asarray(lam_x0, dtype=float).ravel()
            clip = lambda v: np.minimum(np.maximum(v, o.warm_start_mult_bound_push), o.warm_start_mult_init_max)
            lx = self.df * lam_x0[self.iv]                       # lam_x = zU - zL of the unscaled problem
            zL, zU = clip(-lx[iL]), clip(lx[iU])
            yc = np.clip(self.df * lam_g0[self.ic] / self.dc, -o.warm_start_mult_init_max, o.warm_start_mult_init_max)
            yd = np.clip(self.df * lam_g0[self.id] / self.dd, -o.warm_start_mult_init_max, o.warm_start_mult_init_max)
            vL, vU = clip(-yd[jL]), clip(yd[jU])
        elif m_c + m_d > 0:
            rx = grad.copy(); rx[iL] -= zL; rx[iU] += zU
            rs = np.zeros(m_d); rs[jL] -= vL; rs[jU] += vU
            K = self._assemble(sp.csr_matrix((n, n)), np.ones(n), np.ones(m_d), Jc, Jd, 0.0)
            kk = self._factor(K)
            if kk.ok and kk.inertia[2] == 0:
                sol = kk.solve(np.concatenate([rx, rs, np.zeros(m_c + m_d)]))
                yc, yd = -sol[n + m_d:n + m_d + m_c], -sol[n + m_d + m_c:]
                if max(np.abs(yc).max(initial=0.0), np.abs(yd).max(initial=0.0)) > o.constr_mult_init_max:
                    yc[:] = 0.0; yd[:] = 0.0
        mu = o.mu_init
        tau = max(o.tau_min, 1.0 - mu)
        filt = []                                        # list of (theta, phi) corners
        theta0 = np.abs(c).sum() + np.abs(d - s).sum()
        theta_max = o.theta_max_fact * max(1.0, theta0)
        theta_min = o.theta_min_fact * max(1.0, theta0)
        delta_w_last = 0.0
        it = 0
        n_accept = 0
        status = "Maximum_Iterations_Exceeded"
        hist = []
        dummies_present = self.iv.size < self.nlp.n_w
        degenerate_reg = o.mimic_dummy_reg and dummies_present
        resto_count = 0

        def slacks(x, s):
            return x[iL] - xL[iL], xU[iU] - x[iU], s[jL] - dL[jL], dU[jU] - s[jU]

        def barrier_phi(x, s, mu, f):
            sl = slacks(x, s)
            if any((v <= 0).any() for v in sl):
                return np.inf
            val = f - mu * sum(np.log(v).sum() for v in sl)
            if o.kappa_d > 0:
                val += o.kappa_d * mu * ((x - xL)[self.onlyL].sum() + (xU - x)[self.onlyU].sum()
                                         + (s - dL)[self.s_onlyL].sum() + (dU - s)[self.s_onlyU].sum())
            return val

        def errors(x, s, yc, yd, zL, zU, vL, vU, grad, Jc, Jd, c, d, mu):
            rx = grad + Jc.T @ yc + Jd.T @ yd
            rx[iL] -= zL; rx[iU] += zU
            rs = -yd.copy(); rs[jL] -= vL; rs[jU] += vU
            sxl, sxu, ssl, ssu = slacks(x, s)
            comp = [sxl * zL - mu, sxu * zU - mu, ssl * vL - mu, ssu * vU - mu]
            n_z = zL.size + zU.size + vL.size + vU.size
            sum_z = np.abs(zL).sum() + np.abs(zU).sum() + np.abs(vL).sum() + np.abs(vU).sum()
            sum_y = np.abs(yc).sum() + np.abs(yd).sum()
            s_d = max(o.s_max, (sum_y + sum_z) / max(m_c + m_d + n_z, 1)) / o.s_max
            s_c = max(o.s_max, sum_z / max(n_z, 1)) / o.s_max
            dual = max(np.abs(rx).max(initial=0.0), np.abs(rs).max(initial=0.0))
            prim = max(np.abs(c).max(initial=0.0), np.abs(d - s).max(initial=0.0))
            cmp_ = max([np.abs(v).max(initial=0.0) for v in comp])
            return dual, prim, cmp_, s_d, s_c

        f = self._eval_f(x)
        acceptable_cnt = 0
        while True:
            dual, prim, cmp0, s_d, s_c = errors(x, s, yc, yd, zL, zU, vL, vU, grad, Jc, Jd, c, d, 0.0)
            E0 = max(dual / s_d, prim, cmp0 / s_c)
            # unscaled quantities for the secondary tolerances
            dual_u = dual / self.df
            prim_u = max((np.abs(c) / self.dc).max(initial=0.0), (np.abs(d - s) / self.dd).max(initial=0.0))
            cmp_u = cmp0 / self.df
            hist.append((it, f / self.df, prim, dual, mu))
            if o.verbose:
                print(f"{it:4d} f={f/self.df: .8e} inf_pr={prim:.2e} inf_du={dual:.2e} lg(mu)={np.log10(mu):5.1f} E0={E0:.2e}")
            if E0 <= o.tol and dual_u <= o.dual_inf_tol and prim_u <= o.constr_viol_tol and cmp_u <= o.compl_inf_tol:
                status = "Solve_Succeeded"; break
            acceptable_now = (E0 <= o.acceptable_tol and dual_u <= o.acceptable_dual_inf_tol
                              and prim_u <= o.acceptable_constr_viol_tol and cmp_u <= o.acceptable_compl_inf_tol)
            if acceptable_now:
                acceptable_cnt += 1
                if acceptable_cnt >= o.acceptable_iter:
                    status = "Solved_To_Acceptable_Level"; break
            else:
                acceptable_cnt = 0
            if it >= o.max_iter:
                status = "Maximum_Iterations_Exceeded"; break
            # -- barrier parameter update (monotone, fast decrease allowed) ----------------------------
            while True:
                dual_m, prim_m, cmp_m, s_d, s_c = errors(x, s, yc, yd, zL, zU, vL, vU, grad, Jc, Jd, c, d, mu)
                Emu = max(dual_m / s_d, prim_m, cmp_m / s_c)
                if Emu <= o.kappa_eps * mu and mu > max(o.mu_min, o.tol / (o.kappa_eps + 1) if False else o.mu_min):
                    new_mu = max(o.tol / 10.0, min(o.kappa_mu * mu, mu ** o.theta_mu))
                    new_mu = max(new_mu, o.mu_min)
                    if new_mu >= mu:
                        break
                    mu = new_mu
                    tau = max(o.tau_min, 1.0 - mu)
                    filt = []
                else:
                    break
            # -- search direction ---------------------------------------------------------------------
            sxl, sxu, ssl, ssu = slacks(x, s)
            Sx = np.zeros(n); Sx[iL] += zL / sxl; Sx[iU] += zU / sxu
            Ss = np.zeros(m_d); Ss[jL] += vL / ssl; Ss[jU] += vU / ssu
            gphi_x = grad.copy(); gphi_x[iL] -= mu / sxl; gphi_x[iU] += mu / sxu
            gphi_s = np.zeros(m_d); gphi_s[jL] -= mu / ssl; gphi_s[jU] += mu / ssu
            if o.kappa_d > 0:
                gphi_x[self.onlyL] += o.kappa_d * mu; gphi_x[self.onlyU] -= o.kappa_d * mu
                gphi_s[self.s_onlyL] += o.kappa_d * mu; gphi_s[self.s_onlyU] -= o.kappa_d * mu
            W = self._eval_hess(x, yc, yd)
            rhs = -np.concatenate([gphi_x + Jc.T @ yc + Jd.T @ yd, gphi_s - yd, c, d - s])
            # inertia correction (Alg. IC)
            delta_w, delta_c = 0.0, 0.0
            if degenerate_reg:
                delta_w = o.delta_w_first if delta_w_last == 0.0 else max(o.delta_w_min, o.kappa_w_minus * delta_w_last)
            tried_zero = degenerate_reg
            kk = None
            fail = False
            while True:
                K = self._assemble(W, Sx + delta_w, Ss + delta_w, Jc, Jd, delta_c)
                kk = self._factor(K)
                npos, nneg, nzero = kk.inertia
                if kk.ok and nzero == 0 and nneg == m_c + m_d:
                    break
                if nzero > 0 and delta_c == 0.0:
                    delta_c = o.delta_c_bar * mu ** o.kappa_c
                if delta_w == 0.0:
                    delta_w = o.delta_w_first if delta_w_last == 0.0 else max(o.delta_w_min, o.kappa_w_minus * delta_w_last)
                else:
                    delta_w = (o.kappa_w_plus_first if delta_w_last == 0.0 else o.kappa_w_plus) * delta_w
                if delta_w > o.delta_w_max:
                    fail = True; break
            if fail:
                status = "Solved_To_Acceptable_Level" if acceptable_now else "Error_In_Step_Computation"; break
            if delta_w > 0:
                delta_w_last = delta_w
            sol = kk.solve(rhs)
            # one step of iterative refinement keeps the residual at round-off level
            res = rhs - K @ sol
            sol = sol + kk.solve(res)
            dx, ds = sol[:n], sol[n:n + m_d]
            dyc, dyd = sol[n + m_d:n + m_d + m_c], sol[n + m_d + m_c:]
            dzL = mu / sxl - zL - zL / sxl * dx[iL]
            dzU = mu / sxu - zU + zU / sxu * dx[iU]
            dvL = mu / ssl - vL - vL / ssl * ds[jL]
            dvU = mu / ssu - vU + vU / ssu * ds[jU]

            def ftb(v, dv, tau):                      # fraction to the boundary for v>0
                neg = dv < 0
                if not neg.any(): return 1.0
                return min(1.0, float((-tau * v[neg] / dv[neg]).min()))
            a_max = min(ftb(sxl, dx[iL], tau), ftb(sxu, -dx[iU], tau), ftb(ssl, ds[jL], tau), ftb(ssu, -ds[jU], tau))
            a_z = min(ftb(zL, dzL, tau), ftb(zU, dzU, tau), ftb(vL, dvL, tau), ftb(vU, dvU, tau))

            theta = np.abs(c).sum() + np.abs(d - s).sum()
            phi = barrier_phi(x, s, mu, f)
            dphi = gphi_x @ dx + gphi_s @ ds

            def filter_ok(th, ph):
                return all((th < ft or ph < fp) for ft, fp in filt)

            def acceptable(th_t, ph_t, alpha):
                """returns (ok, armijo_type)"""
                if not np.isfinite(ph_t) or th_t > theta_max:
                    return False, False
                switching = (dphi < 0) and (alpha * (-dphi) ** o.s_phi > o.delta * theta ** o.s_theta)
                eps_r = 10 * np.finfo(float).eps
                if theta <= theta_min and switching:
                    ok = ph_t - phi - eps_r * abs(phi) <= o.eta_phi * alpha * dphi
                    arm = True
                else:
                    ok = (th_t - (1 - o.gamma_theta) * theta <= eps_r * abs(theta)) or \
                         (ph_t - (phi - o.gamma_phi * theta) <= eps_r * abs(phi))
                    arm = False
                if ok and not filter_ok(th_t, ph_t):
                    ok = False
                return ok, arm

            if dphi < 0 and theta <= theta_min:
                a_min = min(o.gamma_theta, o.gamma_phi * theta / (-dphi), o.delta * theta ** o.s_theta / (-dphi) ** o.s_phi)
            elif dphi < 0:
                a_min = min(o.gamma_theta, o.gamma_phi * theta / (-dphi))
            else:
                a_min = o.gamma_theta
            a_min *= o.alpha_min_frac

            alpha = a_max
            accepted = False
            armijo = False
            ls = 0
            soc_used = False
            while alpha >= a_min or ls == 0:
                ls += 1
                xt, st = x + alpha * dx, s + alpha * ds
                ct, dt = self._eval_g(xt)
                ft = self._eval_f(xt)
                th_t = np.abs(ct).sum() + np.abs(dt - st).sum()
                ph_t = barrier_phi(xt, st, mu, ft)
                ok, arm = acceptable(th_t, ph_t, alpha)
                if ok:
                    accepted, armijo = True, arm
                    break
                # second-order correction on the first trial only
                if ls == 1 and th_t >= theta and o.max_soc > 0:
                    c_soc, dms_soc = alpha * c + ct, alpha * (d - s) + (dt - st)
                    th_old = th_t
                    for _ in range(o.max_soc):
                        rhs_soc = rhs.copy()
                        rhs_soc[n + m_d:n + m_d + m_c] = -c_soc
                        rhs_soc[n + m_d + m_c:] = -dms_soc
                        sol2 = kk.solve(rhs_soc)
                        sol2 = sol2 + kk.solve(rhs_soc - K @ sol2)
                        dx2, ds2 = sol2[:n], sol2[n:n + m_d]
                        a_soc = min(ftb(sxl, dx2[iL], tau), ftb(sxu, -dx2[iU], tau), ftb(ssl, ds2[jL], tau), ftb(ssu, -ds2[jU], tau))
                        xt2, st2 = x + a_soc * dx2, s + a_soc * ds2
                        ct2, dt2 = self._eval_g(xt2)
                        ft2 = self._eval_f(xt2)
                        th2 = np.abs(ct2).sum() + np.abs(dt2 - st2).sum()
                        ph2 = barrier_phi(xt2, st2, mu, ft2)
                        ok2, arm2 = acceptable(th2, ph2, alpha)
                        if ok2:
                            accepted, armijo, soc_used = True, arm2, True
                            xt, st, ct, dt, ft = xt2, st2, ct2, dt2, ft2
                            sol = sol2
                            dyc, dyd = sol2[n + m_d:n + m_d + m_c], sol2[n + m_d + m_c:]
                            # IPOPT replaces the whole search direction by the corrected one: the bound-multiplier
                            # steps and their fraction-to-the-boundary step size follow the corrected primal step
                            dzL = mu / sxl - zL - zL / sxl * dx2[iL]
                            dzU = mu / sxu - zU + zU / sxu * dx2[iU]
                            dvL = mu / ssl - vL - vL / ssl * ds2[jL]
                            dvU = mu / ssu - vU + vU / ssu * ds2[jU]
                            a_z = min(ftb(zL, dzL, tau), ftb(zU, dzU, tau), ftb(vL, dvL, tau), ftb(vU, dvU, tau))
                            break
                        if th2 > o.kappa_soc * th_old:
                            break
                        th_old = th2
                        c_soc, dms_soc = a_soc * c_soc + ct2, a_soc * dms_soc + (dt2 - st2)
                    if accepted:
                        break
                alpha *= o.alpha_red_factor
            if o.verbose:
                print(f"     alpha={alpha:.3e} a_max={a_max:.3e} a_z={a_z:.3e} ls={ls} dphi={dphi:.3e} phi={phi:.12e} accepted={int(accepted)} "
                      f"armijo={int(armijo)} soc={int(soc_used)} delta_w={delta_w:.1e}")
            if not accepted:
                # ---- restoration substitute ----
                resto_count += 1
                if acceptable_now:               # IPOPT: "restoration phase called at acceptable point" -> acceptable termination
                    status = "Solved_To_Acceptable_Level"; break
                filt.append(((1 - o.gamma_theta) * theta, phi - o.gamma_phi * theta))    # IPOPT augments the filter first
                r = self._restoration(x, s, theta, filt, theta_max, mu, barrier_phi)
                if r is None:
                    status = "Infeasible_Problem_Detected" if theta > 1e-6 else "Restoration_Failed"
                    break
                x, s = r
                c, d = self._eval_g(x); f = self._eval_f(x); grad = self._eval_grad(x); Jc, Jd = self._eval_jac(x)
                # multipliers: bound multipliers capped, least-squares y (as at the starting point)
                zL = np.minimum(zL, 1e3); zU = np.minimum(zU, 1e3); vL = np.minimum(vL, 1e3); vU = np.minimum(vU, 1e3)
                rx = grad.copy(); rx[iL] -= zL; rx[iU] += zU
                rs = np.zeros(m_d); rs[jL] -= vL; rs[jU] += vU
                Kls = self._assemble(sp.csr_matrix((n, n)), np.ones(n), np.ones(m_d), Jc, Jd, 0.0)
                kls = self._factor(Kls)
                yc = np.zeros(m_c); yd = np.zeros(m_d)
                if kls.ok and kls.inertia[2] == 0:
                    sol = kls.solve(np.concatenate([rx, rs, np.zeros(m_c + m_d)]))
                    yc_, yd_ = -sol[n + m_d:n + m_d + m_c], -sol[n + m_d + m_c:]
                    if max(np.abs(yc_).max(initial=0.0), np.abs(yd_).max(initial=0.0)) <= o.constr_mult_init_max:
                        yc, yd = yc_, yd_
                it += 1
                continue
            # -- accept the step -------------------------------------------------------------------------
            if not armijo:
                filt.append(((1 - o.gamma_theta) * theta, phi - o.gamma_phi * theta))
            a_used = a_soc if soc_used else alpha
            x, s = xt, st
            yc = yc + a_used * dyc; yd = yd + a_used * dyd
            zL = zL + a_z * dzL; zU = zU + a_z * dzU; vL = vL + a_z * dvL; vU = vU + a_z * dvU
            sxl, sxu, ssl, ssu = slacks(x, s)
            ks = o.kappa_sigma
            zL = np.maximum(np.minimum(zL, ks * mu / sxl), mu / (ks * sxl))
            zU = np.maximum(np.minimum(zU, ks * mu / sxu), mu / (ks * sxu))
            vL = np.maximum(np.minimum(vL, ks * mu / ssl), mu / (ks * ssl))
            vU = np.maximum(np.minimum(vU, ks * mu / ssu), mu / (ks * ssu))
            c, d, f = ct, dt, ft
            grad = self._eval_grad(x)
            Jc, Jd = self._eval_jac(x)
            it += 1

        # ---- un-scale and scatter back ---------------------------------------------------------------------
        x_out = self._full(x)
        lam_g = np.zeros(self.nlp.n_g)
        lam_g[self.ic] = self.dc * yc / self.df
        lam_g[self.id] = self.dd * yd / self.df
        lam_x = np.zeros(self.nlp.n_w)
        lx = np.zeros(n); lx[iL] -= zL; lx[iU] += zU
        lam_x[self.iv] = lx / self.df
        g_out = self.nlp.g(x_out, self.p)
        return {
            "x": x_out, "f": f / self.df, "g": g_out, "lam_g": lam_g, "lam_x": lam_x,
            "success": status in ("Solve_Succeeded", "Solved_To_Acceptable_Level"),
            "return_status": status, "iter_count": it, "mu": mu, "restorations": resto_count,
            "t_wall_total": time.perf_counter() - t_start, "history": hist,
        }

    # ---- helpers ----------------------------------------------------------------------------------------
    def _assemble(self, W, Dx, Ds, Jc, Jd, delta_c):
        n, m_c, m_d = self.n, self.m_c, self.m_d
        blocks = [[W + sp.diags(Dx), None, Jc.T, Jd.T],
                  [None, sp.diags(Ds) if m_d else None, None, -sp.identity(m_d) if m_d else None],
                  [Jc, None, -delta_c * sp.identity(m_c), None],
                  [Jd, -sp.identity(m_d) if m_d else None, None, -delta_c * sp.identity(m_d) if m_d else None]]
        if m_d == 0:
            blocks = [[blocks[0][0], blocks[0][2]], [blocks[2][0], blocks[2][2]]]
        if m_c == 0 and m_d == 0:
            return (W + sp.diags(Dx)).tocsc()
        return sp.bmat(blocks, format="csc")

    def _restoration(self, x, s, theta_ref, filt, theta_max, mu, barrier_phi):
        """Feasibility restoration substitute (IPOPT's own restoration phase — a nested interior-point solve of
        min rho*||c||_1 + zeta/2*||D_R (x - x_R)||^2 — is third-party code that is not restated).  Primal barrier Newton
        iteration on
            min  zeta/2 ||D_R (x - x_R)||^2 + zeta/2 ||D_Rs (s - s_R)||^2 - mu sum log(bound slacks)   s.t.  c(x) = 0, d(x) - s = 0
        with IPOPT's zeta = sqrt(mu), D_R = diag(1 / max(1, |x_R|)): every step satisfies the linearised constraints, so it is
        a descent direction for theta = ||(c, d - s)||_1; fraction-to-the-boundary (outer tau) and Armijo backtracking on
        theta.  Returns when theta <= kappa_resto * theta_ref and the point is acceptable to the (already augmented) filter
        of the outer problem; None if theta cannot be reduced (local infeasibility) or after resto_max_iter steps.
        The Hessian of this problem is a positive diagonal, so the KKT matrix always has the correct inertia."""
        o = self.o
        n, m_c, m_d = self.n, self.m_c, self.m_d
        iL, iU, xL, xU, jL, jU, dL, dU = self.iL, self.iU, self.xL, self.xU, self.jL, self.jU, self.dL, self.dU
        xr, sr = x.copy(), s.copy()
        zeta = np.sqrt(mu)
        DRx = zeta / np.maximum(1.0, np.abs(x)) ** 2
        DRs = zeta / np.maximum(1.0, np.abs(s)) ** 2
        tau = max(o.tau_min, 1.0 - mu)
        c, d = self._eval_g(xr)
        th = np.abs(c).sum() + np.abs(d - sr).sum()
        W0 = sp.csr_matrix((n, n))

        def ftb(v, dv):
            neg = dv < 0
            if not neg.any(): return 1.0
            return min(1.0, float((-tau * v[neg] / dv[neg]).min()))
        for _ in range(o.resto_max_iter):
            Jc, Jd = self._eval_jac(xr)
            sxl, sxu, ssl, ssu = xr[iL] - xL[iL], xU[iU] - xr[iU], sr[jL] - dL[jL], dU[jU] - sr[jU]
            Sx = DRx.copy(); Sx[iL] += mu / sxl ** 2; Sx[iU] += mu / sxu ** 2
            Ss = DRs.copy(); Ss[jL] += mu / ssl ** 2; Ss[jU] += mu / ssu ** 2
            gx = DRx * (xr - x); gx[iL] -= mu / sxl; gx[iU] += mu / sxu
            gs = DRs * (sr - s); gs[jL] -= mu / ssl; gs[jU] += mu / ssu
            K = self._assemble(W0, Sx, Ss, Jc, Jd, 0.0)
            kk = self._factor(K)
            if not (kk.ok and kk.inertia[2] == 0):
                return None
            sol = kk.solve(-np.concatenate([gx, gs, c, d - sr]))
            dx, ds = sol[:n], sol[n:n + m_d]
            if not np.all(np.isfinite(sol)):
                return None
            a = min(ftb(sxl, dx[iL]), ftb(sxu, -dx[iU]), ftb(ssl, ds[jL]), ftb(ssu, -ds[jU]))
            improved = False
            while a > 1e-10:
                xt, st = xr + a * dx, sr + a * ds
                ct, dt = self._eval_g(xt)
                tht = np.abs(ct).sum() + np.abs(dt - st).sum()
                if tht < (1 - 1e-4 * a) * th:
                    improved = True
                    break
                a *= 0.5
            if not improved:
                return None
            xr, sr, c, d, th = xt, st, ct, dt, tht
            if th <= o.kappa_resto * theta_ref and th <= theta_max:
                ph = barrier_phi(xr, sr, mu, self._eval_f(xr))
                if np.isfinite(ph) and all((th < a_ or ph < b_) for a_, b_ in filt):
                    return xr, sr
        return None


def solve_mpc(nlp, x0, u_prev, z0=None, tvp=None, p=None, options=None, solver=None):
    """Cold-start solve exactly as the reference's sampler drives it (set_initial_guess + make_step)."""
    solver = solver or IPMSolver(nlp, options)
    w0 = nlp.initial_guess(x0, u_prev, z0)
    pp = nlp.pack_p(x0, u_prev, tvp=tvp, p=p)
    res = solver.solve(w0, pp)
    res["u0"] = res["x"][nlp.ix_u[0, 0]] * nlp.d.u_scaling
    return res
