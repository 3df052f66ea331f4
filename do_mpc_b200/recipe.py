"""Stage-structured NLP recipe: what MPC.setup() hands to the CUDA code generator.

The reference builds one flat symbolic NLP (do_mpc/controller/_mpc.py:1097-1301) and lets CasADi
differentiate all n_w variables.  Because that NLP is stage-repetitive, this package instead keeps
  * the per-stage atoms as small expression graphs in *scaled* stage symbols, and
  * integer tables describing the scenario tree (do_mpc/optimizer.py:998-1048) and the flat
    opt_x / opt_p / g layouts,
and generates kernels that evaluate every (instance, edge) with the same code.

Vocabulary: a *node* is (k, s) — the state `_x[k, s, -1]` with its input `_u[k, s]` and slack `_eps[k, s]`;
an *edge* is (k, s, b) — one discretised interval from node (k, s) to child node (k+1, child[k][s][b]),
owning the collocation unknowns y = [_x[k+1, ch, 0..P-1] | _z[k, ch, 0..Pz-1]].
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np

from .sym import expr as E
from .sym.sx import SX, Function, collocation_points, vertcat


def collocation_constants(deg: int, scheme: str):
    """tau_root, C, D exactly as do_mpc/optimizer.py:843-888 (Lagrange basis through [0]+points)."""
    from numpy.polynomial import polynomial as nppoly
    tau = np.array([0.0] + collocation_points(deg, scheme))
    C = np.zeros((deg + 1, deg + 1)); D = np.zeros(deg + 1)
    for j in range(deg + 1):
        p = np.array([1.0])
        for r in range(deg + 1):
            if r != j:
                p = nppoly.polymul(p, np.array([-tau[r], 1.0]) / (tau[j] - tau[r]))
        D[j] = nppoly.polyval(1.0, p)
        dp = nppoly.polyder(p)
        for r in range(deg + 1):
            C[j, r] = nppoly.polyval(tau[r], dp)
    return tau, C, D


@dataclass
class Recipe:
    # sizes
    n_x: int; n_u: int; n_z: int; n_tvp: int; n_pm: int; n_comb: int
    n_eps: int; n_nl: int; N: int; n_robust: int; P: int; Pz: int; discrete: bool
    n_y: int = 0
    nl_rows_per_edge: int = 0
    # tree
    n_branches: List[int] = field(default_factory=list)
    n_scenarios: List[int] = field(default_factory=list)
    node_k: np.ndarray = None; node_s: np.ndarray = None
    node_first_edge: np.ndarray = None; node_n_edges: np.ndarray = None; node_parent_edge: np.ndarray = None
    edge_k: np.ndarray = None; edge_parent: np.ndarray = None; edge_child: np.ndarray = None
    edge_pidx: np.ndarray = None; edge_omega: np.ndarray = None
    # flat layout (reference order)
    n_w: int = 0; n_g: int = 0; n_p: int = 0
    node_x_flat: np.ndarray = None; node_u_flat: np.ndarray = None; node_eps_flat: np.ndarray = None
    edge_yx_flat: np.ndarray = None; edge_yz_flat: np.ndarray = None; edge_g_flat: np.ndarray = None
    p_x0: int = 0; p_tvp: int = 0; p_p: int = 0; p_uprev: int = 0
    # symbolic atoms in scaled stage symbols
    q: SX = None; tvp: SX = None; pp: SX = None; uprev: SX = None; xN: SX = None
    G: SX = None            # (n_y) edge equality residuals gk (same row order as the reference's ifcn)
    lterm: SX = None; epscost: SX = None; nl: SX = None; nl_ub: np.ndarray = None
    rterm: SX = None; mterm: SX = None; aux: SX = None
    xf_sel: np.ndarray = None   # indices inside y of the end-of-interval state (E y = x_child)
    x_scaling: np.ndarray = None; u_scaling: np.ndarray = None; z_scaling: np.ndarray = None
    opt_x_scaling: np.ndarray = None
    # ---- generalisations used by the moving-horizon estimator / lifted global variables (do_mpc_b200/mhe.py) ----------------
    n_free0: int = 0             # the LAST n_free0 components of the root state are free (no initial-condition row)
    arrival: SX = None           # root cost term in the root-state symbols `x0sym` (MHE arrival cost), with tvp / pp
    x0sym: SX = None
    w2flat: np.ndarray = None    # explicit internal -> flat opt_x map (NWI entries, -1 = no flat counterpart); None: MPC layout
    lam2flat: np.ndarray = None  # explicit internal multiplier row -> flat g row map (NLAM entries)
    w_init_from: np.ndarray = None   # initial value of an entry without flat counterpart: copy of another internal entry (-1: none)
    has_dummies: bool = None     # opt_x holds variables that appear nowhere (IPOPT's delta_w schedule is mimicked); None: MPC rule
    n_lift: int = 0              # the last n_lift state components are a lifted global variable (nl_cons_single_slack), not model states


def build_recipe(mpc) -> Recipe:
    """Translate a configured MPC object into the stage recipe (reference _prepare_nlp, restated stage-wise)."""
    m, st = mpc.model, mpc.settings
    n_x, n_u, n_z, n_tvp, n_pm = m.n_x, m.n_u, m.n_z, m.n_tvp, m.n_p
    N, n_robust, n_comb = st.n_horizon, st.n_robust, mpc.n_combinations
    if st.open_loop and n_comb ** n_robust > 1:
        # with a single scenario the flag has no effect (one input sequence either way, reference _mpc.py:1113-1117)
        raise NotImplementedError("open_loop=True couples the inputs of all scenarios; not supported by the batched solver")
    lift = bool(st.nl_cons_single_slack and mpc.n_eps > 0)
    if lift and n_comb ** n_robust > 1:
        # `_eps[0, s]` is shared by the nodes (k, s) of all stages (reference _mpc.py:1228-1254), which are not on one branch of the tree
        raise NotImplementedError("nl_cons_single_slack together with a scenario tree (the shared slack couples different branches)")
    if m.n_w > 0:
        pass  # process noise symbols are fixed to zero in the controller (reference _mpc.py:1166)
    discrete = (m.model_type == "discrete")
    deg, ni = st.collocation_deg, st.collocation_ni
    P = 0 if discrete else (deg + 1) * ni
    Pz = max(P, 1)
    n_eps, n_nl = mpc.n_eps, mpc.n_nl
    # nl_cons_single_slack: ONE slack vector for the whole horizon (reference _mpc.py:1120-1123).  It becomes n_l extra state
    # components with the trivial dynamics eps+ = eps and a free root (the mechanism of the MHE's estimated parameters): every
    # stage sees its own copy, the linear consistency rows keep the copies equal, only the root copy is bounded / mapped to opt_x.
    n_l = n_eps if lift else 0
    n_eps_flat = n_eps
    if lift: n_eps = 0
    sx, su, sz = (np.asarray(v.full()).ravel() for v in (mpc._x_scaling.cat, mpc._u_scaling.cat, mpc._z_scaling.cat))
    if n_robust > 0 and st.nl_cons_check_colloc_points and mpc.n_nl > 0:
        # same quirk for the collocation states: the reference checks `_x[k+1, s, i]` of the stage-k scenario s (_mpc.py:1233)
        raise NotImplementedError("nl_cons_check_colloc_points together with n_robust>0 (reference indexes the collocation states of the "
                                  "stage-k scenario, SURVEY App. F.5)")

    # ---- scaled stage symbols ----------------------------------------------------------------------------
    Xa = SX.sym("Xn", n_x + n_l); Xn = Xa[0:n_x] if n_l else Xa
    Un = SX.sym("Un", n_u); EPS = SX.sym("EPS", n_eps) if n_eps else SX.zeros(0, 1)
    EPSv = Xa[n_x:n_x + n_l] if n_l else EPS          # the slack the stage functions see
    TV = SX.sym("TV", n_tvp) if n_tvp else SX.zeros(0, 1)
    PP = SX.sym("PP", n_pm) if n_pm else SX.zeros(0, 1)
    UPREV = SX.sym("UPREV", n_u)
    W0 = SX.zeros(m.n_w, 1)
    sxS, suS, szS = SX(sx.reshape(-1, 1)), SX(su.reshape(-1, 1)), SX(sz.reshape(-1, 1)) if n_z else SX.zeros(0, 1)

    # scaling['_p', ...] multiplies the parameters seen by the model equations only (optimizer.py:811-815); cost and constraints get opt_p['_p']
    sp = np.asarray(mpc._p_scaling.cat.full()).ravel() if n_pm else np.zeros(0)
    PPm = PP * SX(sp.reshape(-1, 1)) if n_pm and np.any(sp != 1.0) else PP

    def rhs_scaled(x, u, z):          # optimizer.py:807-818
        return m._rhs_fun(x * sxS, u * suS, z * szS if n_z else z, TV, PPm, W0) / sxS

    def alg(x, u, z):
        return m._alg_fun(x * sxS, u * suS, z * szS if n_z else z, TV, PPm, W0)

    # ---- edge unknowns and residuals gk (optimizer.py:891-993 / :820-824) ---------------------------------
    if discrete:
        n_y = n_z + n_x
        Ya = SX.sym("Y", n_y + n_l); Y = Ya
        zk = Y[0:n_z] if n_z else SX.zeros(0, 1)
        xf = Y[n_z:n_y]
        rows = []
        if n_z:
            rows.append(alg(Xn, Un, zk))
        rows.append(rhs_scaled(Xn, Un, zk) - xf)        # aux variable xf (not part of the reference opt_x)
        G = vertcat(*rows)
        xf_sel = np.arange(n_z, n_y)
        z_first = zk; z_last = zk
        yx_count, yz_count = 0, n_z
    else:
        n_y = P * (n_x + n_z)
        Ya = SX.sym("Y", n_y + n_l); Y = Ya
        _, C, D = collocation_constants(deg, st.collocation_type)
        C[np.abs(C) < 1e-14] = 0.0; D[np.abs(D) < 1e-14] = 0.0
        h = st.t_step / ni
        yx = [Y[i * n_x:(i + 1) * n_x] for i in range(P)]
        yz = [Y[P * n_x + i * n_z:P * n_x + (i + 1) * n_z] if n_z else SX.zeros(0, 1) for i in range(P)]
        xs = [[None] * (deg + 1) for _ in range(ni)]; zs = [[None] * (deg + 1) for _ in range(ni)]
        xs[0][0] = Xn
        off, offz = 0, 0
        zs[0][0] = yz[offz]; offz += 1
        first_j = 1
        for i in range(ni):
            for j in range(first_j, deg + 1):
                xs[i][j] = yx[off]; off += 1
                zs[i][j] = yz[offz]; offz += 1
            first_j = 0
        xkf = yx[off]; off += 1
        assert off == P and offz == P
        rows = []
        for i in range(ni):
            if n_z:
                rows.append(alg(xs[i][0], Un, zs[i][0]))
            for j in range(1, deg + 1):
                xp = SX.zeros(n_x, 1)
                for r in range(deg + 1):
                    if C[r, j] != 0.0:
                        xp = xp + float(C[r, j]) * xs[i][r]
                rows.append(h * rhs_scaled(xs[i][j], Un, zs[i][j]) - xp)
                if n_z:
                    rows.append(alg(xs[i][j], Un, zs[i][j]))
            xf_i = SX.zeros(n_x, 1)
            for r in range(deg + 1):
                if D[r] != 0.0:
                    xf_i = xf_i + float(D[r]) * xs[i][r]
            x_next = xs[i + 1][0] if i + 1 < ni else xkf
            rows.append(x_next - xf_i)
        G = vertcat(*rows)
        xf_sel = np.arange((P - 1) * n_x, P * n_x)
        z_first = yz[0]; z_last = yz[P - 1]
    assert G.shape[0] == n_y
    if n_l:
        G = vertcat(G, EPSv - Ya[n_y:n_y + n_l])        # lifted slack: child copy = parent copy
        xf_sel = np.concatenate([xf_sel, n_y + np.arange(n_l)])
    n_ym = n_y; n_y = n_y + n_l

    q = vertcat(Xa, Un, EPS, Ya)
    # ---- cost / constraints per edge (_mpc.py:1228-1275), evaluated on unscaled values ----------------------
    xu, uu = Xn * sxS, Un * suS
    lterm = mpc.lterm_fun(xu, uu, z_last * szS if n_z else z_last, TV, PP)
    epscost = mpc.epsterm_fun(EPS) if n_eps else SX(0.0)
    if n_l:
        lterm = lterm + mpc.epsterm_fun(EPSv)          # added once per edge like the reference does (_mpc.py:1254); one scenario: omega = 1
    nl_rows = []
    if n_nl:
        if st.nl_cons_check_colloc_points:
            for i in range(P):
                nl_rows.append(mpc._nl_cons_fun(yx[i] * sxS, uu, yz[i] * szS if n_z else yz[i], TV, PP, EPSv))
        else:
            nl_rows.append(mpc._nl_cons_fun(xu, uu, z_first * szS if n_z else z_first, TV, PP, EPSv))
    nl = vertcat(*nl_rows) if nl_rows else SX.zeros(0, 1)
    nl_ub = np.tile(np.asarray(mpc._nl_cons_ub.cat.full()).ravel(), len(nl_rows)) if nl_rows else np.zeros(0)
    XN = SX.sym("XN", n_x + n_l)
    mterm = mpc.mterm_fun((XN[0:n_x] if n_l else XN) * sxS, TV, PP)
    if mpc.flags["rterm_fun"]:
        rterm = mpc.rterm_fun(xu, uu, UPREV, SX.zeros(n_z, 1), TV, PP)   # u_prev enters scaled (reference :1265-1269)
        if E.depends_on(rterm.ids()[0], Xn.ids()):
            raise NotImplementedError("user rterm depending on the states")
        _zs = SX.sym("_rterm_z_probe", n_z) if n_z else None
        if n_z and E.depends_on(mpc.rterm_fun(xu, uu, UPREV, _zs, TV, PP).ids()[0], _zs.ids()):
            raise NotImplementedError("user rterm depending on the algebraic states")
    else:
        rf = np.asarray(mpc.rterm_factor.cat.full()).ravel()
        rterm = SX(0.0)
        for i in range(n_u):
            if rf[i] != 0.0:
                rterm = rterm + float(rf[i]) * (Un[i] - UPREV[i]) ** 2
    aux = m._aux_expression_fun(xu, uu, z_last * szS if n_z else z_last, TV, PP)
    if n_z and n_comb ** n_robust > 1:
        # the reference evaluates lterm / nl_cons / aux of edge (k, s, b) with `_z[k, s, .]` of the stage-k scenario s although
        # `_z[k, .]` is stored per CHILD scenario (_mpc.py:1213 vs :1233,1241,1252): for k < n_robust the edges of one node would read
        # the algebraic states of another edge.  Where none of the three depends on z the quirk has no effect and the tree is edge-local.
        zids = list(zk.ids()) if discrete else [i_ for z_ in yz for i_ in z_.ids()]
        for nm_, ex_ in (("lterm", lterm), ("nl_cons", nl), ("aux expressions", aux)):
            if ex_.shape[0] and any(E.depends_on(i_, zids) for i_ in ex_.ids()):
                raise NotImplementedError(f"{nm_} depending on algebraic states together with n_robust>0 "
                                          "(reference indexing quirk, SURVEY App. F.5)")

    # ---- tree tables (optimizer.py:1011-1039) ----------------------------------------------------------------
    n_branches = [n_comb if k < n_robust else 1 for k in range(N)]
    n_scen = [n_comb ** min(k, n_robust) for k in range(N + 1)]
    S = n_comb ** n_robust
    node_off = np.concatenate([[0], np.cumsum(n_scen)]).astype(int)
    n_node = int(node_off[-1])
    node_k = np.zeros(n_node, int); node_s = np.zeros(n_node, int)
    node_first_edge = -np.ones(n_node, int); node_n_edges = np.zeros(n_node, int); node_parent_edge = -np.ones(n_node, int)
    ek, ep, ec, epi, eom, egf = [], [], [], [], [], []
    for k in range(N + 1):
        for s in range(n_scen[k]):
            node_k[node_off[k] + s] = k; node_s[node_off[k] + s] = s
    nl_rows_per_edge = nl.shape[0]
    g_row = n_x
    for k in range(N):
        cnt = 0
        for s in range(n_scen[k]):
            n = node_off[k] + s
            node_first_edge[n] = len(ek); node_n_edges[n] = n_branches[k]
            for b in range(n_branches[k]):
                boff = 0 if (n_robust == 0 or k < n_robust) else s % n_branches[0]
                e = len(ek)
                ek.append(k); ep.append(n); ec.append(node_off[k + 1] + cnt); epi.append(b + boff)
                eom.append(1.0 / n_scen[k + 1]); egf.append(g_row)
                node_parent_edge[node_off[k + 1] + cnt] = e
                g_row += (P * (n_x + n_z) if not discrete else n_z) + n_x + nl_rows_per_edge
                cnt += 1
    n_edge = len(ek)

    # ---- flat layout -------------------------------------------------------------------------------
    o_x = 0
    sz_x = (N + 1) * S * (P + 1) * n_x
    o_z = o_x + sz_x; sz_zz = N * S * Pz * n_z
    o_u = o_z + sz_zz; sz_u = N * S * n_u
    o_e = o_u + sz_u; sz_e = (1 if lift else N) * S * n_eps_flat
    n_w = o_e + sz_e
    node_x_flat = np.array([o_x + ((node_k[n] * S + node_s[n]) * (P + 1) + P) * n_x for n in range(n_node)])
    node_u_flat = np.array([o_u + (node_k[n] * S + node_s[n]) * n_u if node_k[n] < N else -1 for n in range(n_node)])
    node_eps_flat = np.array([o_e + (node_k[n] * S + node_s[n]) * n_eps if node_k[n] < N else -1 for n in range(n_node)])
    edge_yx_flat = np.array([o_x + (((ek[e] + 1) * S + node_s[ec[e]]) * (P + 1)) * n_x for e in range(n_edge)])
    edge_yz_flat = np.array([o_z + ((ek[e] * S + node_s[ec[e]]) * Pz) * n_z for e in range(n_edge)])
    n_p = n_x + (N + 1) * n_tvp + n_comb * n_pm + n_u
    osc = np.ones(n_w)
    osc[o_x:o_x + sz_x] = np.tile(sx, (N + 1) * S * (P + 1))
    if n_z: osc[o_z:o_z + sz_zz] = np.tile(sz, N * S * Pz)
    osc[o_u:o_u + sz_u] = np.tile(su, N * S)

    extra = {}
    if n_l:
        # explicit internal <-> flat maps: node blocks (x, eps_copy, u), edge blocks (y of the model, eps_child)
        w2f, init_from, nvp = [], [], n_x + n_l + n_u
        for n in range(n_node):
            w2f += [int(node_x_flat[n]) + i for i in range(n_x)] + ([o_e + i for i in range(n_l)] if n == 0 else [-1] * n_l)
            w2f += [int(node_u_flat[n]) + i if node_u_flat[n] >= 0 else -1 for i in range(n_u)]
            init_from += [-1] * n_x + ([-1] * n_l if n == 0 else [n_x + i for i in range(n_l)]) + [-1] * n_u
        for e in range(n_edge):
            if discrete:
                w2f += [int(edge_yz_flat[e]) + i for i in range(n_z)] + [-1] * n_x
            else:
                w2f += [int(edge_yx_flat[e]) + i for i in range(P * n_x)] + [int(edge_yz_flat[e]) + i for i in range(P * n_z)]
            w2f += [-1] * n_l
            init_from += [-1] * n_ym + [n_x + i for i in range(n_l)]
        l2f = list(range(n_x)) + [-1] * n_l
        for e in range(n_edge):
            g0 = int(egf[e])
            if discrete:
                l2f += [g0 + i for i in range(n_z)] + [-1] * n_x + [-1] * n_l
                l2f += [g0 + n_z + i for i in range(n_x)] + [-1] * n_l
                l2f += [g0 + n_z + n_x + i for i in range(nl_rows_per_edge)]
            else:
                l2f += [g0 + i for i in range(n_ym)] + [-1] * n_l
                l2f += [g0 + n_ym + i for i in range(n_x)] + [-1] * n_l
                l2f += [g0 + n_ym + n_x + i for i in range(nl_rows_per_edge)]
        extra = dict(n_free0=n_l, w2flat=np.array(w2f), lam2flat=np.array(l2f), w_init_from=np.array(init_from), n_lift=n_l,
                     has_dummies=bool(P > 0))
        sx = np.concatenate([sx, np.ones(n_l)])
    n_xm = n_x; n_x = n_x + n_l

    return Recipe(
        **extra,
        n_x=n_x, n_u=n_u, n_z=n_z, n_tvp=n_tvp, n_pm=n_pm, n_comb=n_comb, n_eps=n_eps, n_nl=n_nl, N=N,
        n_robust=n_robust, P=P, Pz=Pz, discrete=discrete, n_y=n_y, nl_rows_per_edge=nl_rows_per_edge,
        n_branches=n_branches, n_scenarios=n_scen,
        node_k=node_k, node_s=node_s, node_first_edge=node_first_edge, node_n_edges=node_n_edges,
        node_parent_edge=node_parent_edge,
        edge_k=np.array(ek), edge_parent=np.array(ep), edge_child=np.array(ec), edge_pidx=np.array(epi),
        edge_omega=np.array(eom), n_w=n_w, n_g=g_row, n_p=n_p,
        node_x_flat=node_x_flat, node_u_flat=node_u_flat, node_eps_flat=node_eps_flat,
        edge_yx_flat=edge_yx_flat, edge_yz_flat=edge_yz_flat, edge_g_flat=np.array(egf),
        p_x0=0, p_tvp=n_xm, p_p=n_xm + (N + 1) * n_tvp, p_uprev=n_xm + (N + 1) * n_tvp + n_comb * n_pm,
        q=q, tvp=TV, pp=PP, uprev=UPREV, xN=XN, G=G, lterm=lterm, epscost=epscost, nl=nl, nl_ub=nl_ub,
        rterm=rterm, mterm=mterm, aux=aux, xf_sel=xf_sel,
        x_scaling=sx, u_scaling=su, z_scaling=sz, opt_x_scaling=osc,
    )
