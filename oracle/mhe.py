"""CPU restatement of the NLP that do-mpc's moving horizon estimator hands to nlpsol (TEST INFRASTRUCTURE ONLY,
see oracle/README.md; nothing under do_mpc_b200/ imports this).

Restated (reference file:line):
  * variable / parameter / constraint layout ........ do_mpc/estimator/_mhe.py:1051-1059 (opt_x), :1082-1088 (opt_p)
  * arrival cost, stage cost, slack penalty ........... _mhe.py:1114-1123, :1184-1189
  * collocation / continuity / measurement rows ....... _mhe.py:1128-1161 (ifcn from do_mpc/optimizer.py:891-993)
  * nonlinear constraints, appended TWICE at the end .. _mhe.py:1163-1182 (`cons.append(nl_cons_k)` after the if/else as well)
  * bounds: states at ALL k (unlike the controller) ... _mhe.py:995-1028
  * make_step parameter fill .......................... _mhe.py:941-950

The class re-uses the evaluation machinery of oracle/nlp.py (atoms with sympy derivatives, stacked index vectors); only the
layout and the list of terms are MHE-specific.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import scipy.sparse as sp
import sympy as sy

from .collocation import collocation_matrices
from .nlp import Atom, FlatNLP, ModelDef, NlCons, _Inst


@dataclass
class MHEDef:
    model: ModelDef
    n_horizon: int
    t_step: float
    p_est: List[sy.Symbol]                    # estimated parameters (subset of model.p, model order)
    meas: List[sy.Expr]                       # measurement functions h(x, u, z, tvp, p); y = h + v (noise on every measurement)
    stage_cost: sy.Expr                       # in (w..., v..., tvp, p)
    arrival_cost: sy.Expr                     # in (x, x_prev, p_est, p_est_prev, p_set)
    v: List[sy.Symbol]
    x_prev: List[sy.Symbol]
    p_est_prev: List[sy.Symbol]
    collocation_type: str = "radau"
    collocation_deg: int = 2
    collocation_ni: int = 1
    nl_cons_check_colloc_points: bool = False
    cons_check_colloc_points: bool = True
    nl_cons: List[NlCons] = field(default_factory=list)
    x_scaling: Optional[np.ndarray] = None
    u_scaling: Optional[np.ndarray] = None
    x_lb: Optional[np.ndarray] = None
    x_ub: Optional[np.ndarray] = None
    u_lb: Optional[np.ndarray] = None
    u_ub: Optional[np.ndarray] = None
    p_est_lb: Optional[np.ndarray] = None
    p_est_ub: Optional[np.ndarray] = None


class FlatMHE(FlatNLP):
    """min f(opt_x, opt_p) s.t. lbg <= g <= ubg, lbx <= opt_x <= ubx in the reference's flat MHE layout (no algebraic states,
    no process noise, no soft constraints: what the reference's regression example uses)."""

    def __init__(self, d: MHEDef):
        self.d = d
        m = d.model
        assert m.n_z == 0 and m.model_type == "continuous"
        nx, nu, ntv = m.n_x, m.n_u, m.n_tvp
        N = self.N = d.n_horizon
        P = self.P = (d.collocation_deg + 1) * d.collocation_ni
        ny = len(d.meas); nv = len(d.v)
        assert ny == nv
        pe = list(d.p_est); ps = [p for p in m.p if p not in pe]
        npe, nps = len(pe), len(ps)
        dflt = lambda v, n, val: np.full(n, val, dtype=float) if v is None else np.asarray(v, dtype=float).reshape(n)
        sx, su = dflt(d.x_scaling, nx, 1.0), dflt(d.u_scaling, nu, 1.0)
        # ---- layouts --------------------------------------------------------------------------------
        o = 0
        self.ix_x = o + np.arange((N + 1) * (P + 1) * nx).reshape(N + 1, P + 1, nx); o += self.ix_x.size
        self.ix_u = o + np.arange(N * nu).reshape(N, nu); o += self.ix_u.size
        self.ix_v = o + np.arange(N * nv).reshape(N, nv); o += self.ix_v.size
        self.ix_pe = o + np.arange(npe); o += npe
        self.n_w = o
        q = 0
        self.ip_xprev = q + np.arange(nx); q += nx
        self.ip_peprev = q + np.arange(npe); q += npe
        self.ip_pset = q + np.arange(nps); q += nps
        self.ip_tvp = q + np.arange(N * ntv).reshape(N, ntv); q += N * ntv
        self.ip_ymeas = q + np.arange(N * ny).reshape(N, ny); q += N * ny
        self.n_p = q
        sc = np.ones(self.n_w); sc[self.ix_x] = sx; sc[self.ix_u] = su
        self.scaling = sc
        W0 = self.n_w
        # ---- atoms (scaled symbols) -----------------------------------------------------------------------
        X = sy.symbols(f"X0:{nx}"); U = sy.symbols(f"U0:{nu}"); TV = sy.symbols(f"TV0:{ntv}") if ntv else ()
        PE = sy.symbols(f"PE0:{npe}") if npe else (); PS = sy.symbols(f"PS0:{nps}") if nps else ()
        V = sy.symbols(f"V0:{nv}")
        XP = sy.symbols(f"XP0:{nx}"); PEP = sy.symbols(f"PEP0:{npe}") if npe else ()
        sub = {}
        for a, b, c in zip(m.x, X, sx): sub[a] = b * c if c != 1.0 else b
        for a, b, c in zip(m.u, U, su): sub[a] = b * c if c != 1.0 else b
        for a, b in zip(m.tvp, TV): sub[a] = b
        for a, b in zip(pe, PE): sub[a] = b
        for a, b in zip(ps, PS): sub[a] = b
        for a, b in zip(d.v, V): sub[a] = b
        for a, b in zip(d.x_prev, XP): sub[a] = b
        for a, b in zip(d.p_est_prev, PEP): sub[a] = b
        S = lambda e: sy.sympify(e).subs(sub, simultaneous=True)
        pt = list(X) + list(U) + list(TV) + list(PE) + list(PS)
        a_rhs = Atom("mhe_rhs", pt, [S(e) / c if c != 1.0 else S(e) for e, c in zip(m.rhs, sx)])
        a_meas = Atom("mhe_meas", pt + list(V), [S(h) + v for h, v in zip(d.meas, V)])
        a_stage = Atom("mhe_stage", list(V) + list(TV) + list(PE) + list(PS), [S(d.stage_cost)])
        a_arr = Atom("mhe_arr", list(X) + list(XP) + list(PE) + list(PEP) + list(PS), [S(d.arrival_cost)])
        a_nl = Atom("mhe_nl", pt, [S(c.expr) for c in d.nl_cons]) if d.nl_cons else None
        n_nl = len(d.nl_cons)
        _, C, D = collocation_matrices(d.collocation_deg, d.collocation_type)
        deg, ni = d.collocation_deg, d.collocation_ni
        h = d.t_step / ni
        lin_r, lin_c, lin_v = [], [], []

        def lin(rows, cols, coef):
            rows = np.asarray(rows).ravel(); cols = np.asarray(cols).ravel()
            lin_r.extend(rows); lin_c.extend(cols); lin_v.extend(np.broadcast_to(coef, rows.shape))
        inst = {k: ([], [], []) for k in ("rhs", "meas", "stage", "arr", "nl")}
        self._rhs_coef = []

        def add(kind, in_idx, out_row, weight=1.0):
            inst[kind][0].append(np.asarray(in_idx)); inst[kind][1].append(np.asarray(out_row)); inst[kind][2].append(weight)
        pe_i, ps_i = self.ix_pe, W0 + self.ip_pset
        row = 0
        self.edges = []
        add("arr", np.concatenate([self.ix_x[0, -1], W0 + self.ip_xprev, pe_i, W0 + self.ip_peprev, ps_i]), [-1], 1.0)
        for k in range(N):
            self.edges.append((k, 0, 0, 0, row))
            xk, uk, tvk = self.ix_x[k, -1], self.ix_u[k], W0 + self.ip_tvp[k]
            col_x = self.ix_x[k + 1, :-1]
            xs = [[None] * (deg + 1) for _ in range(ni)]
            xs[0][0] = xk
            off = 0; first_j = 1
            for i in range(ni):
                for j in range(first_j, deg + 1):
                    xs[i][j] = col_x[off]; off += 1
                first_j = 0
            xkf = col_x[off]; off += 1
            for i in range(ni):
                for j in range(1, deg + 1):
                    rr = row + np.arange(nx)
                    add("rhs", np.concatenate([xs[i][j], uk, tvk, pe_i, ps_i]), rr); self._rhs_coef.append(h)
                    for r in range(deg + 1):
                        lin(rr, xs[i][r], -C[r, j])
                    row += nx
                rr = row + np.arange(nx)
                x_next = xs[i + 1][0] if i + 1 < ni else xkf
                lin(rr, x_next, 1.0)
                for r in range(deg + 1):
                    lin(rr, xs[i][r], -D[r])
                row += nx
            rr = row + np.arange(nx)                                                  # continuity (_mhe.py:1153)
            lin(rr, xkf, 1.0); lin(rr, self.ix_x[k + 1, -1], -1.0)
            row += nx
            rr = row + np.arange(ny)                                                  # measurement (_mhe.py:1158): h(x_{k+1}, u_k, ...) + v - y_meas
            add("meas", np.concatenate([self.ix_x[k + 1, -1], uk, tvk, pe_i, ps_i, self.ix_v[k]]), rr)
            lin(rr, W0 + self.ip_ymeas[k], -1.0)
            row += ny
            if n_nl:
                pts = [self.ix_x[k + 1, i] for i in range(P)] if d.nl_cons_check_colloc_points else [xk]
                for xp in pts + [pts[-1]]:                                            # the last block is appended twice (_mhe.py:1180-1182)
                    add("nl", np.concatenate([xp, uk, tvk, pe_i, ps_i]), row + np.arange(n_nl)); row += n_nl
            add("stage", np.concatenate([self.ix_v[k], tvk, pe_i, ps_i]), [-1], 1.0)
        self.n_g = row
        self.A_lin = sp.csr_matrix((np.array(lin_v, dtype=float), (np.array(lin_r), np.array(lin_c))), shape=(self.n_g, self.n_w + self.n_p))
        atoms = {"rhs": a_rhs, "meas": a_meas, "stage": a_stage, "arr": a_arr, "nl": a_nl}
        self.terms = []
        for kind, (ii, rr, ww) in inst.items():
            if not ii:
                continue
            in_idx = np.stack(ii, axis=1)
            out_row = np.stack([np.asarray(r).reshape(-1) for r in rr], axis=1)
            t = _Inst(atoms[kind], in_idx, out_row, np.asarray(ww, dtype=float))
            t.kind = kind
            t.coef = np.asarray(self._rhs_coef) if kind == "rhs" else np.ones(in_idx.shape[1])
            self.terms.append(t)
        self.lbg = np.zeros(self.n_g); self.ubg = np.zeros(self.n_g)
        for t in self.terms:
            if t.kind == "nl":
                ub = np.array([c.ub for c in d.nl_cons], dtype=float)
                self.lbg[t.out_row] = -np.inf
                self.ubg[t.out_row] = ub[:, None]
        # ---- bounds (_mhe.py:995-1028) ----------------------------------------------------------------------------
        lb = np.full(self.n_w, -np.inf); ub = np.full(self.n_w, np.inf)
        x_lb, x_ub = dflt(d.x_lb, nx, -np.inf), dflt(d.x_ub, nx, np.inf)
        if d.cons_check_colloc_points:
            lb[self.ix_x] = x_lb; ub[self.ix_x] = x_ub
        else:
            lb[self.ix_x[1:N, -1]] = x_lb; ub[self.ix_x[1:N, -1]] = x_ub
        lb[self.ix_u] = dflt(d.u_lb, nu, -np.inf); ub[self.ix_u] = dflt(d.u_ub, nu, np.inf)
        lb[self.ix_pe] = dflt(d.p_est_lb, npe, -np.inf); ub[self.ix_pe] = dflt(d.p_est_ub, npe, np.inf)
        self.lbx, self.ubx = lb / sc, ub / sc
        # attributes the block-LDL partition of oracle/ipm.py looks at (one "scenario"; no z / eps)
        self.n_scenarios = [1] * (N + 1)
        self.ix_z = np.zeros((N, 1, 1, 0), dtype=int); self.ix_eps = np.zeros((0, 1, 0), dtype=int)
        self.ix_x = self.ix_x[:, None, :, :]; self.ix_u = self.ix_u[:, None, :]

    def partition_extra(self):
        """variables the generic partition does not know: the measurement noise of stage k belongs to the block of node k + 1
        (with its measurement rows); the estimated parameters stay in the root block."""
        return [(self.ix_v[k], (k + 1, 0)) for k in range(self.N)]

    def _w(self, x, p):
        return np.concatenate([x, p])

    def initial_guess(self, x0, u0, p_est0):
        """set_initial_guess (_mhe.py:864-883): x0, u0, p_est0 (scaled) broadcast; v stays 0."""
        w = np.zeros(self.n_w)
        w[self.ix_x] = np.asarray(x0, dtype=float).ravel() / self.scaling[self.ix_x[0, 0, 0]]
        w[self.ix_u] = np.asarray(u0, dtype=float).ravel() / self.scaling[self.ix_u[0, 0]]
        w[self.ix_pe] = np.asarray(p_est0, dtype=float).ravel()
        return w
