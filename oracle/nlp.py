"""CPU restatement of the NLP that do-mpc's MPC controller hands to nlpsol.

TEST INFRASTRUCTURE ONLY (see oracle/README.md): nothing under do_mpc_b200/ imports this.

What is restated (reference file:line):
  * variable / parameter / constraint layout .......... do_mpc/controller/_mpc.py:1126-1165, 1189-1245
  * orthogonal collocation block gk ..................... do_mpc/optimizer.py:891-993
  * discrete-time block ................................. do_mpc/optimizer.py:820-824
  * scenario tree tables ................................ do_mpc/optimizer.py:1011-1039
  * objective (lterm, mterm, eps penalty, rterm) ........ do_mpc/controller/_mpc.py:1196-1275
  * soft / hard nonlinear constraints ................... do_mpc/optimizer.py:553-585, _mpc.py:1228-1245
  * bounds on opt_x ..................................... do_mpc/controller/_mpc.py:1061-1095
  * initial guess ....................................... do_mpc/controller/_mpc.py:969-971

The model is given as sympy expressions (oracle/models.py); derivatives come from sympy
(independent of the product's own symbolic engine) and are evaluated with numpy.  The flat
vectors use exactly the reference ordering so that golden `_opt_x_num`,
`_lam_g_num` and `opt_p_num` rows can be plugged in directly.

Lagrangian convention (CasADi / IPOPT, also do_mpc/differentiator/_nlpdifferentiator.py:287):
    L = f + lam_g' g + lam_x' x.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import scipy.sparse as sp
import sympy as sy

from .collocation import collocation_matrices


# --------------------------------------------------------------------------------------
# problem description
# --------------------------------------------------------------------------------------
@dataclass
class ModelDef:
    model_type: str                      # 'continuous' | 'discrete'
    x: List[sy.Symbol]
    u: List[sy.Symbol]
    z: List[sy.Symbol] = field(default_factory=list)
    tvp: List[sy.Symbol] = field(default_factory=list)
    p: List[sy.Symbol] = field(default_factory=list)
    rhs: List[sy.Expr] = field(default_factory=list)
    alg: List[sy.Expr] = field(default_factory=list)
    aux: Dict[str, sy.Expr] = field(default_factory=dict)

    @property
    def n_x(self): return len(self.x)
    @property
    def n_u(self): return len(self.u)
    @property
    def n_z(self): return len(self.z)
    @property
    def n_tvp(self): return len(self.tvp)
    @property
    def n_p(self): return len(self.p)


@dataclass
class NlCons:
    name: str
    expr: sy.Expr
    ub: float = np.inf
    soft: bool = False
    penalty: float = 1.0
    max_violation: float = np.inf


@dataclass
class MPCDef:
    model: ModelDef
    n_horizon: int
    t_step: float
    n_robust: int = 0
    open_loop: bool = False
    state_discretization: str = "collocation"
    collocation_type: str = "radau"
    collocation_deg: int = 2
    collocation_ni: int = 1
    nl_cons_check_colloc_points: bool = False
    nl_cons_single_slack: bool = False
    cons_check_colloc_points: bool = True
    use_terminal_bounds: bool = False
    lterm: sy.Expr = sy.Integer(0)
    mterm: sy.Expr = sy.Integer(0)
    rterm_factor: Optional[np.ndarray] = None          # default quadratic penalty on scaled du
    nl_cons: List[NlCons] = field(default_factory=list)
    x_scaling: Optional[np.ndarray] = None
    u_scaling: Optional[np.ndarray] = None
    z_scaling: Optional[np.ndarray] = None
    x_lb: Optional[np.ndarray] = None
    x_ub: Optional[np.ndarray] = None
    u_lb: Optional[np.ndarray] = None
    u_ub: Optional[np.ndarray] = None
    z_lb: Optional[np.ndarray] = None
    z_ub: Optional[np.ndarray] = None
    x_terminal_lb: Optional[np.ndarray] = None
    x_terminal_ub: Optional[np.ndarray] = None
    p_scenarios: Optional[np.ndarray] = None           # (n_comb, n_p); row 0 = nominal

    def __post_init__(self):
        m = self.model
        def dflt(v, n, val):
            return np.full(n, val, dtype=float) if v is None else np.asarray(v, dtype=float).reshape(n)
        self.x_scaling = dflt(self.x_scaling, m.n_x, 1.0)
        self.u_scaling = dflt(self.u_scaling, m.n_u, 1.0)
        self.z_scaling = dflt(self.z_scaling, m.n_z, 1.0)
        self.x_lb = dflt(self.x_lb, m.n_x, -np.inf)
        self.x_ub = dflt(self.x_ub, m.n_x, np.inf)
        self.u_lb = dflt(self.u_lb, m.n_u, -np.inf)
        self.u_ub = dflt(self.u_ub, m.n_u, np.inf)
        self.z_lb = dflt(self.z_lb, m.n_z, -np.inf)
        self.z_ub = dflt(self.z_ub, m.n_z, np.inf)
        self.x_terminal_lb = dflt(self.x_terminal_lb, m.n_x, -np.inf)
        self.x_terminal_ub = dflt(self.x_terminal_ub, m.n_x, np.inf)
        # _mpc.py:915-918 — default terminal bounds are the state bounds
        if self.use_terminal_bounds:
            if np.all(self.x_terminal_ub == np.inf):
                self.x_terminal_ub = self.x_ub.copy()
            if np.all(self.x_terminal_lb == -np.inf):
                self.x_terminal_lb = self.x_lb.copy()
        self.rterm_factor = dflt(self.rterm_factor, m.n_u, 0.0)
        if self.p_scenarios is None:
            self.p_scenarios = np.zeros((1, m.n_p))
        self.p_scenarios = np.asarray(self.p_scenarios, dtype=float).reshape(-1, m.n_p) if m.n_p else np.zeros((1, 0))


# --------------------------------------------------------------------------------------
# atoms: small vector functions with sympy derivatives, evaluated vectorised over instances
# --------------------------------------------------------------------------------------
class Atom:
    """out(v): R^n_in -> R^n_out.  val, jac and hess(lam) are numpy-vectorised over instances."""

    def __init__(self, name, inputs, outputs):
        self.name = name
        self.n_in = len(inputs)
        self.n_out = len(outputs)
        outputs = [sy.sympify(o) for o in outputs]
        lam = sy.symbols(f"lam_{name}_0:{max(self.n_out, 1)}")
        jac = [[sy.diff(o, v) for v in inputs] for o in outputs]
        lag = sum(l * o for l, o in zip(lam, outputs)) if outputs else sy.Integer(0)
        glag = [sy.diff(lag, v) for v in inputs]
        hes = [[sy.diff(g, v) for v in inputs] for g in glag]
        self.jac_nz = [(i, j) for i in range(self.n_out) for j in range(self.n_in) if jac[i][j] != 0]
        self.hes_nz = [(i, j) for i in range(self.n_in) for j in range(self.n_in) if hes[i][j] != 0]
        mods = ["numpy"]
        self._val = sy.lambdify([inputs], outputs, modules=mods, cse=True)
        self._jac = sy.lambdify([inputs], [jac[i][j] for i, j in self.jac_nz], modules=mods, cse=True)
        self._hes = sy.lambdify([inputs, lam], [hes[i][j] for i, j in self.hes_nz], modules=mods, cse=True)

    @staticmethod
    def _stack(vals, n):
        out = np.empty((len(vals), n))
        for i, v in enumerate(vals):
            out[i, :] = v
        return out

    def val(self, V):                     # V: (n_in, n_inst) -> (n_out, n_inst)
        return self._stack(self._val(V), V.shape[1])

    def jac(self, V):                     # -> (len(jac_nz), n_inst)
        return self._stack(self._jac(V), V.shape[1])

    def hess(self, V, LAM):               # -> (len(hes_nz), n_inst)
        return self._stack(self._hes(V, LAM), V.shape[1])


@dataclass
class _Inst:
    """All instances of one atom: gather indices into w=[opt_x;opt_p], rows in g (or -1 for objective)."""
    atom: Atom
    in_idx: np.ndarray        # (n_in, n_inst) indices into w
    out_row: np.ndarray       # (n_out, n_inst) rows of g; -1 -> objective with weight
    weight: np.ndarray        # (n_inst,) objective weight (only when out_row == -1)


# --------------------------------------------------------------------------------------
class FlatNLP:
    """min f(w_x, p) s.t. lbg <= g(w_x, p) <= ubg, lbx <= w_x <= ubx in the reference's flat layout."""

    def __init__(self, d: MPCDef):
        self.d = d
        m = d.model
        self.n_x, self.n_u, self.n_z, self.n_tvp, self.n_pm = m.n_x, m.n_u, m.n_z, m.n_tvp, m.n_p
        N = self.N = d.n_horizon
        self.n_comb = d.p_scenarios.shape[0]
        self.S = self.n_comb ** d.n_robust
        self.S_u = 1 if d.open_loop else self.S
        self.discrete = (m.model_type == "discrete")
        if self.discrete:
            self.P = 0
        else:
            assert d.state_discretization == "collocation"
            self.P = (d.collocation_deg + 1) * d.collocation_ni
        self.Pz = max(self.P, 1)
        # slack variables (optimizer.py:553-571)
        self.soft = [c for c in d.nl_cons if c.soft]
        self.n_eps = len(self.soft)
        self.n_nl = len(d.nl_cons)
        self.N_eps = 1 if d.nl_cons_single_slack else N
        self._tree()
        self._layout()
        self._build_terms()
        self._bounds()

    # -- scenario tree (optimizer.py:1011-1039) ---------------------------------------
    def _tree(self):
        N, nr, nc = self.N, self.d.n_robust, self.n_comb
        self.n_branches = [nc if k < nr else 1 for k in range(N)]
        self.n_scenarios = [nc ** min(k, nr) for k in range(N + 1)]
        S = self.n_scenarios[-1]
        self.child = -np.ones((N, S, self.n_branches[0]), dtype=int)
        self.parent = -np.ones((N + 1, S), dtype=int)
        self.branch_offset = -np.ones((N, S), dtype=int)
        for k in range(N):
            cnt = 0
            for s in range(self.n_scenarios[k]):
                for b in range(self.n_branches[k]):
                    self.child[k, s, b] = cnt
                    self.parent[k + 1, cnt] = s
                    cnt += 1
                self.branch_offset[k, s] = 0 if (nr == 0 or k < nr) else s % self.n_branches[0]

    # -- flat layouts (_mpc.py:1126-1134, 1160-1165) ------------------------------------
    def _layout(self):
        N, S, P, Pz = self.N, self.S, self.P, self.Pz
        o = 0
        self.ix_x = o + np.arange((N + 1) * S * (P + 1) * self.n_x).reshape(N + 1, S, P + 1, self.n_x); o = self.ix_x.size
        self.ix_z = o + np.arange(N * S * Pz * self.n_z).reshape(N, S, Pz, self.n_z); o += self.ix_z.size
        self.ix_u = o + np.arange(N * self.S_u * self.n_u).reshape(N, self.S_u, self.n_u); o += self.ix_u.size
        self.ix_eps = o + np.arange(self.N_eps * S * self.n_eps).reshape(self.N_eps, S, self.n_eps); o += self.ix_eps.size
        self.n_w = o
        # parameters, appended behind opt_x in the stacked vector w
        q = 0
        self.ip_x0 = q + np.arange(self.n_x); q += self.n_x
        self.ip_tvp = q + np.arange((N + 1) * self.n_tvp).reshape(N + 1, self.n_tvp); q += self.ip_tvp.size
        self.ip_p = q + np.arange(self.n_comb * self.n_pm).reshape(self.n_comb, self.n_pm); q += self.ip_p.size
        self.ip_uprev = q + np.arange(self.n_u); q += self.n_u
        self.n_p = q
        sc = np.ones(self.n_w)
        sc[self.ix_x] = self.d.x_scaling
        if self.n_z:
            sc[self.ix_z] = self.d.z_scaling
        sc[self.ix_u] = self.d.u_scaling
        self.scaling = sc

    # -- terms -------------------------------------------------------------------------
    def _build_terms(self):
        d, m = self.d, self.d.model
        nx, nu, nz, ntv, npm = self.n_x, self.n_u, self.n_z, self.n_tvp, self.n_pm
        sx, su, sz = d.x_scaling, d.u_scaling, d.z_scaling
        # scaled symbols
        X = sy.symbols(f"X0:{nx}"); U = sy.symbols(f"U0:{nu}"); Z = sy.symbols(f"Z0:{nz}") if nz else ()
        TV = sy.symbols(f"TV0:{ntv}") if ntv else (); PP = sy.symbols(f"PP0:{npm}") if npm else ()
        EPS = sy.symbols(f"EPS0:{self.n_eps}") if self.n_eps else ()
        UP = sy.symbols(f"UPREV0:{nu}")
        sub = {}
        for s_, S_, c_ in ((m.x, X, sx), (m.u, U, su), (m.z, Z, sz)):
            for a, b, c in zip(s_, S_, c_):
                sub[a] = b * c if c != 1.0 else b
        for a, b in zip(m.tvp, TV): sub[a] = b
        for a, b in zip(m.p, PP): sub[a] = b
        S = lambda e: sy.sympify(e).subs(sub, simultaneous=True)
        pt_in = list(X) + list(U) + list(Z) + list(TV) + list(PP)
        # optimizer.py:814-818: rhs on unscaled values, divided by x_scaling; alg unscaled
        rhs_s = [S(e) / c if c != 1.0 else S(e) for e, c in zip(m.rhs, sx)]
        alg_s = [S(e) for e in m.alg]
        self.a_rhs = Atom("rhs", pt_in, rhs_s)
        self.a_alg = Atom("alg", pt_in, alg_s) if nz else None
        self.a_lterm = Atom("lterm", pt_in, [S(d.lterm)])
        self.a_mterm = Atom("mterm", list(X) + list(TV) + list(PP), [S(d.mterm)])
        if self.n_nl:
            outs = []
            for c in d.nl_cons:
                e = S(c.expr)
                if c.soft:
                    e = e - EPS[self.soft.index(c)]
                outs.append(e)
            self.a_nl = Atom("nlcons", pt_in + list(EPS), outs)
        if self.n_eps:
            self.a_eps = Atom("epsterm", list(EPS), [sum(c.penalty * e for c, e in zip(self.soft, EPS))])
        self.a_rterm = Atom("rterm", list(U) + list(UP),
                            [sum(float(r) * (u - up) ** 2 for r, u, up in zip(d.rterm_factor, U, UP))])
        self.a_aux = Atom("aux", pt_in, [S(e) for e in m.aux.values()]) if m.aux else None

        N, P, deg, ni = self.N, self.P, d.collocation_deg, d.collocation_ni
        W0 = self.n_w                     # parameter offset inside w
        if not self.discrete:
            _, C, D = collocation_matrices(deg, d.collocation_type)
            h = d.t_step / ni
        lin_r, lin_c, lin_v = [], [], []   # linear part of g:  g += A_lin @ w
        def lin(rows, cols, coef):
            rows = np.asarray(rows).ravel(); cols = np.asarray(cols).ravel()
            lin_r.extend(rows); lin_c.extend(cols); lin_v.extend(np.broadcast_to(coef, rows.shape))
        inst = {k: ([], [], []) for k in ("rhs", "alg", "lterm", "mterm", "nl", "eps", "rterm")}
        self._rhs_coef = []                 # multiplier h for collocation rows, 1 for discrete
        def add(kind, in_idx, out_row, weight=1.0):
            inst[kind][0].append(np.asarray(in_idx)); inst[kind][1].append(np.asarray(out_row)); inst[kind][2].append(weight)

        row = 0
        # initial condition (_mpc.py:1189): _x[0,0,-1] - x0/x_scaling = 0.  The division by the
        # (constant) scaling is folded into the linear coefficient on the parameter.
        r0 = row + np.arange(nx)
        lin(r0, self.ix_x[0, 0, -1], 1.0)
        lin(r0, W0 + self.ip_x0, -1.0 / sx)
        row += nx
        self.edges = []
        omega = [1.0 / self.n_scenarios[k + 1] for k in range(N)]
        for k in range(N):
            for s in range(self.n_scenarios[k]):
                s_u = 0 if d.open_loop else s
                for b in range(self.n_branches[k]):
                    cs = b + self.branch_offset[k, s]          # parameter realisation of this branch
                    ch = self.child[k, s, b]
                    xk = self.ix_x[k, s, -1]
                    uk = self.ix_u[k, s_u]
                    tvk = W0 + self.ip_tvp[k]
                    pk = W0 + self.ip_p[cs]
                    col_x = self.ix_x[k + 1, ch, :-1] if P else None   # (P, nx)
                    col_z = self.ix_z[k, ch]                            # (Pz, nz)
                    self.edges.append((k, s, b, ch, row))
                    if self.discrete:
                        # optimizer.py:823: ifcn -> [alg, rhs/x_scaling]; cons: g=alg, then xf - x_next
                        pin = np.concatenate([xk, uk, col_z[0], tvk, pk])
                        if nz:
                            add("alg", pin, row + np.arange(nz)); row += nz
                        rr = row + np.arange(nx)
                        add("rhs", pin, rr); self._rhs_coef.append(1.0)
                        lin(rr, self.ix_x[k + 1, ch, -1], -1.0)
                        row += nx
                    else:
                        # ik_split / zk_split (optimizer.py:913-933)
                        xs = [[None] * (deg + 1) for _ in range(ni)]
                        zs = [[None] * (deg + 1) for _ in range(ni)]
                        xs[0][0] = xk
                        off, offz = 0, 0
                        zs[0][0] = col_z[offz]; offz += 1
                        first_j = 1
                        for i in range(ni):
                            for j in range(first_j, deg + 1):
                                xs[i][j] = col_x[off]; off += 1
                                zs[i][j] = col_z[offz]; offz += 1
                            first_j = 0
                        xkf = col_x[off]; off += 1
                        assert off == P and offz == P
                        for i in range(ni):
                            pin0 = np.concatenate([xs[i][0], uk, zs[i][0], tvk, pk])
                            if nz:
                                add("alg", pin0, row + np.arange(nz)); row += nz
                            for j in range(1, deg + 1):
                                rr = row + np.arange(nx)
                                pin = np.concatenate([xs[i][j], uk, zs[i][j], tvk, pk])
                                add("rhs", pin, rr); self._rhs_coef.append(h)
                                for r in range(deg + 1):
                                    lin(rr, xs[i][r], -C[r, j])
                                row += nx
                                if nz:
                                    add("alg", pin, row + np.arange(nz)); row += nz
                            rr = row + np.arange(nx)
                            x_next = xs[i + 1][0] if i + 1 < ni else xkf
                            lin(rr, x_next, 1.0)
                            for r in range(deg + 1):
                                lin(rr, xs[i][r], -D[r])
                            row += nx
                        # node continuity (_mpc.py:1224)
                        rr = row + np.arange(nx)
                        lin(rr, xkf, 1.0); lin(rr, self.ix_x[k + 1, ch, -1], -1.0)
                        row += nx
                    # nonlinear constraints (_mpc.py:1228-1245) — indices literally as in the reference
                    k_eps = min(k, self.N_eps - 1)
                    if self.n_nl:
                        epsk = self.ix_eps[k_eps, s]
                        if d.nl_cons_check_colloc_points:
                            for i in range(P):
                                pin = np.concatenate([self.ix_x[k + 1, s, i], uk, self.ix_z[k, s, i], tvk, pk, epsk])
                                add("nl", pin, row + np.arange(self.n_nl)); row += self.n_nl
                        else:
                            pin = np.concatenate([xk, uk, self.ix_z[k, s, 0], tvk, pk, epsk])
                            add("nl", pin, row + np.arange(self.n_nl)); row += self.n_nl
                    # objective (_mpc.py:1251-1275)
                    zl = self.ix_z[k, s, -1]
                    add("lterm", np.concatenate([xk, uk, zl, tvk, pk]), [-1], omega[k])
                    if self.n_eps:
                        add("eps", self.ix_eps[k_eps, s], [-1], 1.0)
                    if k == N - 1:
                        add("mterm", np.concatenate([self.ix_x[k + 1, s, -1], W0 + self.ip_tvp[k + 1], pk]), [-1], omega[k])
                    if k == 0:
                        # u_prev/u_scaling: handled by passing the already-divided value (see _w)
                        add("rterm", np.concatenate([uk, W0 + self.n_p + np.arange(nu)]), [-1], omega[k])
                    else:
                        add("rterm", np.concatenate([uk, self.ix_u[k - 1, self.parent[k, s_u]]]), [-1], omega[k])
        self.n_g = row
        self.A_lin = sp.csr_matrix((np.array(lin_v, dtype=float), (np.array(lin_r), np.array(lin_c))),
                                   shape=(self.n_g, self.n_w + self.n_p))
        atoms = {"rhs": self.a_rhs, "alg": self.a_alg, "lterm": self.a_lterm, "mterm": self.a_mterm,
                 "nl": getattr(self, "a_nl", None), "eps": getattr(self, "a_eps", None), "rterm": self.a_rterm}
        self.terms = []
        for kind, (ii, rr, ww) in inst.items():
            if not ii:
                continue
            a = atoms[kind]
            in_idx = np.stack(ii, axis=1)
            out_row = np.stack([np.asarray(r).reshape(-1) for r in rr], axis=1)
            t = _Inst(a, in_idx, out_row, np.asarray(ww, dtype=float))
            t.kind = kind
            t.coef = np.asarray(self._rhs_coef) if kind == "rhs" else np.ones(in_idx.shape[1])
            self.terms.append(t)
        # constraint bounds
        self.lbg = np.zeros(self.n_g); self.ubg = np.zeros(self.n_g)
        for t in self.terms:
            if t.kind == "nl":
                ub = np.array([c.ub for c in self.d.nl_cons], dtype=float)
                self.lbg[t.out_row] = -np.inf
                self.ubg[t.out_row] = ub[:, None]

    # -- bounds (_mpc.py:1061-1095 through the scaling-dividing setters optimizer.py:233-265) --
    def _bounds(self):
        d = self.d
        lb = np.full(self.n_w, -np.inf); ub = np.full(self.n_w, np.inf)
        N = self.N
        if d.cons_check_colloc_points:
            lb[self.ix_x[1:N]] = d.x_lb; ub[self.ix_x[1:N]] = d.x_ub
            if self.n_z:
                lb[self.ix_z] = d.z_lb; ub[self.ix_z] = d.z_ub
        else:
            lb[self.ix_x[1:N, :, -1]] = d.x_lb; ub[self.ix_x[1:N, :, -1]] = d.x_ub
            if self.n_z:
                lb[self.ix_z[:, :, 0]] = d.z_lb; ub[self.ix_z[:, :, 0]] = d.z_ub
        lb[self.ix_x[N, :, -1]] = d.x_terminal_lb; ub[self.ix_x[N, :, -1]] = d.x_terminal_ub
        lb[self.ix_u] = d.u_lb; ub[self.ix_u] = d.u_ub
        if self.n_eps:
            lb[self.ix_eps] = 0.0
            ub[self.ix_eps] = np.array([c.max_violation for c in self.soft], dtype=float)
        self.lbx = lb / self.scaling
        self.ubx = ub / self.scaling

    # -- parameters / initial guess -----------------------------------------------------
    def pack_p(self, x0, u_prev, tvp=None, p=None):
        """opt_p in reference order [_x0 | _tvp(N+1) | _p(n_comb) | _u_prev] (_mpc.py:1160-1165)."""
        out = np.zeros(self.n_p)
        out[self.ip_x0] = np.asarray(x0, dtype=float).ravel()
        if self.n_tvp:
            out[self.ip_tvp] = np.asarray(tvp, dtype=float).reshape(self.N + 1, self.n_tvp)
        if self.n_pm:
            out[self.ip_p] = self.d.p_scenarios if p is None else np.asarray(p, dtype=float).reshape(self.n_comb, self.n_pm)
        out[self.ip_uprev] = np.asarray(u_prev, dtype=float).ravel()
        return out

    def initial_guess(self, x0, u0, z0=None):
        """set_initial_guess (_mpc.py:969-971): broadcast x0,u0,z0 (scaled); eps stays 0."""
        w = np.zeros(self.n_w)
        w[self.ix_x] = np.asarray(x0, dtype=float).ravel() / self.d.x_scaling
        w[self.ix_u] = np.asarray(u0, dtype=float).ravel() / self.d.u_scaling
        if self.n_z:
            w[self.ix_z] = np.asarray(z0, dtype=float).ravel() / self.d.z_scaling
        return w

    # -- evaluation -----------------------------------------------------------------------
    def _w(self, x, p):
        # [opt_x | opt_p | u_prev/u_scaling] — the last block feeds the k=0 rterm (_mpc.py:1273)
        return np.concatenate([x, p, p[self.ip_uprev] / self.d.u_scaling])

    def f(self, x, p):
        w = self._w(x, p)
        val = 0.0
        for t in self.terms:
            if t.out_row[0, 0] == -1:
                val += float(np.dot(t.weight, t.atom.val(w[t.in_idx])[0]))
        return val

    def g(self, x, p):
        w = self._w(x, p)
        g = self.A_lin @ w[: self.n_w + self.n_p]
        for t in self.terms:
            if t.out_row[0, 0] != -1:
                g[t.out_row] += t.atom.val(w[t.in_idx]) * t.coef
        return g

    def grad_f(self, x, p):
        w = self._w(x, p)
        gr = np.zeros(w.size)
        for t in self.terms:
            if t.out_row[0, 0] == -1:
                J = t.atom.jac(w[t.in_idx])
                for n, (_, j) in enumerate(t.atom.jac_nz):
                    np.add.at(gr, t.in_idx[j], J[n] * t.weight)
        return gr[: self.n_w]

    def jac_g(self, x, p):
        w = self._w(x, p)
        rows, cols, vals = [], [], []
        for t in self.terms:
            if t.out_row[0, 0] != -1:
                J = t.atom.jac(w[t.in_idx])
                for n, (i, j) in enumerate(t.atom.jac_nz):
                    rows.append(t.out_row[i]); cols.append(t.in_idx[j]); vals.append(J[n] * t.coef)
        if rows:
            rows = np.concatenate(rows); cols = np.concatenate(cols); vals = np.concatenate(vals)
            keep = cols < self.n_w
            Jn = sp.csr_matrix((vals[keep], (rows[keep], cols[keep])), shape=(self.n_g, self.n_w))
        else:
            Jn = sp.csr_matrix((self.n_g, self.n_w))
        return (self.A_lin[:, : self.n_w] + Jn).tocsr()

    def hess_l(self, x, p, lam_g, sigma=1.0):
        """Hessian of sigma*f + lam_g' g w.r.t. opt_x (full symmetric, sparse)."""
        w = self._w(x, p)
        rows, cols, vals = [], [], []
        for t in self.terms:
            ninst = t.in_idx.shape[1]
            if t.out_row[0, 0] == -1:
                LAM = (sigma * t.weight)[None, :]
            else:
                LAM = lam_g[t.out_row] * t.coef
            H = t.atom.hess(w[t.in_idx], LAM)
            for n, (i, j) in enumerate(t.atom.hes_nz):
                rows.append(t.in_idx[i]); cols.append(t.in_idx[j]); vals.append(H[n])
        if not rows:
            return sp.csr_matrix((self.n_w, self.n_w))
        rows = np.concatenate(rows); cols = np.concatenate(cols); vals = np.concatenate(vals)
        keep = (rows < self.n_w) & (cols < self.n_w)
        return sp.csr_matrix((vals[keep], (rows[keep], cols[keep])), shape=(self.n_w, self.n_w))

    # -- which variables actually appear (App. B 'dummy' variables) ----------------------------
    def used_mask(self):
        used = np.zeros(self.n_w, dtype=bool)
        A = self.A_lin[:, : self.n_w].tocoo()
        used[A.col] = True
        for t in self.terms:
            idx = t.in_idx[t.in_idx < self.n_w]
            used[idx] = True
        return used

    def aux(self, x, p):
        """opt_aux: aux expressions at (k, s) (_mpc.py:1278-1284), shape (N, S, n_aux)."""
        if self.a_aux is None:
            return np.zeros((self.N, self.S, 0))
        w = self._w(x, p)
        out = np.zeros((self.N, self.S, self.a_aux.n_out))
        for k in range(self.N):
            last = None
            for s in range(self.n_scenarios[k]):
                s_u = 0 if self.d.open_loop else s
                cs = (self.n_branches[k] - 1) + self.branch_offset[k, s]
                pin = np.concatenate([self.ix_x[k, s, -1], self.ix_u[k, s_u], self.ix_z[k, s, -1],
                                      self.n_w + self.ip_tvp[k], self.n_w + self.ip_p[cs]])
                last = self.a_aux.val(w[pin][:, None])[:, 0]
                out[k, s] = last
            out[k, self.n_scenarios[k]:] = last
        return out
