"""Code generation: stage recipe -> `problem_gen.h` (dimensions, index tables, straight-line model functions).

This is the B200 counterpart of what CasADi does inside nlpsol for the reference
(do_mpc/controller/_mpc.py:1326-1328: AD of the flat NLP + an SX virtual machine): here the edge
Lagrangian  L_e(q) = sig_l*lterm + sig_e*epscost + lam'G + lam_nl'nl  is differentiated symbolically
once per problem (NQ variables, not n_w), and residuals / Jacobian nonzeros / Hessian nonzeros are
emitted as CUDA device functions that the generic interior-point kernel (csrc/ipm_core.cuh) calls for
every (instance, edge).
"""
from __future__ import annotations

import hashlib
from typing import List

import numpy as np

from ..recipe import Recipe
from ..sym import expr as E
from ..sym.sx import SX

_QUAL = "DMB_DEV"


def _ctab(ctype, name, vals, fmt="{}"):
    vals = list(vals)
    body = ", ".join(fmt.format(v) for v in vals) if vals else ("0" if ctype == "int" else "0.0")
    return f"DMB_TABLE const {ctype} {name}[{max(len(vals), 1)}] = {{{body}}};"


def _dbl(v):
    return E._fmt_const(float(v))


def _fn(name, args, body, ret="void"):
    return f"{_QUAL} {ret} {name}({args}) {{\n{body}\n}}\n"


def generate(r: Recipe) -> str:
    q_ids = r.q.ids()
    tv_ids, pp_ids, up_ids, xn_ids = r.tvp.ids(), r.pp.ids(), r.uprev.ids(), r.xN.ids()
    n_q = len(q_ids)
    n_vp = r.n_x + r.n_u + r.n_eps
    u_ids = q_ids[r.n_x:r.n_x + r.n_u]
    inp = {}
    for i, s in enumerate(q_ids): inp[s] = f"q[{i}]"
    for i, s in enumerate(tv_ids): inp[s] = f"tvp[{i}]"
    for i, s in enumerate(pp_ids): inp[s] = f"pp[{i}]"
    sig_args = "const double* __restrict__ q, const double* __restrict__ tvp, const double* __restrict__ pp"

    out: List[str] = []
    G_ids = r.G.ids()
    nl_ids = r.nl.ids() if r.nl.numel() else []
    n_nl = len(nl_ids)
    lt, ec = r.lterm.ids()[0], r.epscost.ids()[0]

    # ---- residuals, cost ------------------------------------------------------------------------------------
    out.append(_fn("gen_G", sig_args + ", double* __restrict__ g",
                   E.emit_c([(f"g[{i}]", e) for i, e in enumerate(G_ids)], inp)))
    out.append(_fn("gen_cost", sig_args + ", double* lt, double* ec",
                   E.emit_c([("*lt", lt), ("*ec", ec)], inp)))
    out.append(_fn("gen_nl", sig_args + ", double* __restrict__ d",
                   E.emit_c([(f"d[{i}]", e) for i, e in enumerate(nl_ids)], inp) if n_nl else "  (void)d;"))

    # ---- Jacobians (structural nonzeros, row-major order) ---------------------------------------------------------
    theta_ids = list(tv_ids) + list(pp_ids)        # parameters the stage functions depend on (sensitivity right-hand sides)
    n_th = len(theta_ids)

    def jac_nz(rows):
        nz, r_, c_, dth = [], [], [], []
        for i, ex in enumerate(rows):
            g = E.gradient(ex, list(q_ids) + theta_ids)
            for j, d in enumerate(g[:n_q]):
                if d != E.ZERO:
                    nz.append(d); r_.append(i); c_.append(j)
            dth.append(g[n_q:])
        return nz, r_, c_, dth
    jn, jr, jc, dG_dth = jac_nz(G_ids)
    out.append(_fn("gen_J", sig_args + ", double* __restrict__ jnz",
                   E.emit_c([(f"jnz[{k}]", e) for k, e in enumerate(jn)], inp)))
    dn, dr, dcol, dD_dth = jac_nz(nl_ids)
    out.append(_fn("gen_nl_jac", sig_args + ", double* __restrict__ dnz",
                   E.emit_c([(f"dnz[{k}]", e) for k, e in enumerate(dn)], inp) if dn else "  (void)dnz;"))

    # ---- edge Lagrangian gradient / Hessian ----------------------------------------------------------------------------
    lam = [E.mk_sym(f"lam{i}") for i in range(len(G_ids))]
    lnl = [E.mk_sym(f"lnl{i}") for i in range(n_nl)]
    sig_l, sig_e = E.mk_sym("sig_l"), E.mk_sym("sig_e")
    L = E.mk(E.ADD, E.mk(E.MUL, sig_l, lt), E.mk(E.MUL, sig_e, ec))
    for l_, g_ in zip(lam, G_ids): L = E.mk(E.ADD, L, E.mk(E.MUL, l_, g_))
    for l_, g_ in zip(lnl, nl_ids): L = E.mk(E.ADD, L, E.mk(E.MUL, l_, g_))
    inpL = dict(inp)
    for i, s in enumerate(lam): inpL[s] = f"lam[{i}]"
    for i, s in enumerate(lnl): inpL[s] = f"lnl[{i}]"
    inpL[sig_l] = "sig_l"; inpL[sig_e] = "sig_e"
    gL = E.gradient(L, q_ids)
    lag_args = sig_args + ", const double* __restrict__ lam, const double sig_l, const double sig_e, const double* __restrict__ lnl"
    out.append(_fn("gen_lag_grad", lag_args + ", double* __restrict__ gq",
                   "  (void)lam; (void)lnl;\n" + E.emit_c([(f"gq[{i}]", e) for i, e in enumerate(gL)], inpL)))
    hn, hr, hc, dgL_dth = [], [], [], []
    for i in range(n_q):
        row = E.gradient(gL[i], list(q_ids[:i + 1]) + theta_ids) if gL[i] != E.ZERO else [E.ZERO] * (i + 1 + n_th)
        for j, d in enumerate(row[:i + 1]):
            if d != E.ZERO:
                hn.append(d); hr.append(i); hc.append(j)
        dgL_dth.append(row[i + 1:])
    out.append(_fn("gen_lag_hess", lag_args + ", double* __restrict__ hnz",
                   "  (void)lam; (void)lnl; (void)sig_l; (void)sig_e;\n" +
                   (E.emit_c([(f"hnz[{k}]", e) for k, e in enumerate(hn)], inpL) if hn else "  (void)hnz;")))

    # ---- fused derivative evaluation at one point (shares all common sub-expressions) ----------------------------------
    # objective-only gradient: L with lam = lnl = 0
    Lobj = E.mk(E.ADD, E.mk(E.MUL, sig_l, lt), E.mk(E.MUL, sig_e, ec))
    gO = E.gradient(Lobj, q_ids)
    gC = [E.mk(E.SUB, a, b) for a, b in zip(gL, gO)]
    fused = ([(f"gobj[{i}]", e) for i, e in enumerate(gO)] + [(f"gcon[{i}]", e) for i, e in enumerate(gC)] +
             [(f"jnz[{k}]", e) for k, e in enumerate(jn)] + [(f"hnz[{k}]", e) for k, e in enumerate(hn)] +
             [(f"dnz[{k}]", e) for k, e in enumerate(dn)])
    d_args = lag_args + ", double* __restrict__ gobj, double* __restrict__ gcon, double* __restrict__ jnz, double* __restrict__ hnz, double* __restrict__ dnz"
    d_void = "  (void)q; (void)tvp; (void)pp; (void)lam; (void)lnl; (void)sig_l; (void)sig_e; (void)gobj; (void)gcon; (void)jnz; (void)hnz; (void)dnz;\n"
    out.append(_fn("gen_derivs", d_args, d_void + E.emit_c(fused, inpL)))
    # row-wise (CSR) views: J rows (nonzeros are emitted row-major), full symmetric H rows
    jr_start = [0] * (len(G_ids) + 1)
    for r_ in jr: jr_start[r_ + 1] += 1
    for i in range(len(G_ids)): jr_start[i + 1] += jr_start[i]
    hrows = [[] for _ in range(n_q)]
    for k, (r_, c_) in enumerate(zip(hr, hc)):
        hrows[r_].append((c_, k))
        if r_ != c_: hrows[c_].append((r_, k))
    hr_start, hr_col, hr_idx = [0], [], []
    for row in hrows:
        for c_, k in sorted(row):
            hr_col.append(c_); hr_idx.append(k)
        hr_start.append(len(hr_col))
    # straight-line products of one column of Yext with the sparse Hessian rows (factor phase, n_y <= 16): mc = (H_yy + Sigma_y) yc,
    # nc = H_vy yc; same summation order as the CSR loops they replace (row entries by ascending column, Sigma term last)
    body = []
    for rq in range(n_q):
        terms = [f"hnz[{hr_idx[t]}] * yc[{hr_col[t] - n_vp}]" for t in range(hr_start[rq], hr_start[rq + 1]) if hr_col[t] >= n_vp]
        if rq >= n_vp:
            j = rq - n_vp
            body.append(f"  {{ double s = 0.0; " + " ".join(f"s += {t_};" for t_ in terms) + f" s += sigy[{j}] * yc[{j}]; mc[{j}] = s; }}")
        else:
            body.append(f"  {{ double s = 0.0; " + " ".join(f"s += {t_};" for t_ in terms) + f" nc[{rq}] = s; }}")
    out.append(_fn("gen_hy_products", "const double* __restrict__ hnz, const double* __restrict__ sigy, const double* __restrict__ yc, "
                   "double* __restrict__ mc, double* __restrict__ nc", "  (void)hnz; (void)sigy; (void)yc; (void)mc; (void)nc;\n" + "\n".join(body)))
    # column-wise (CSC) view of the Jacobian nonzeros: J' lam products with one lane per column
    jcols = [[] for _ in range(n_q)]
    for k, c_ in enumerate(jc):
        jcols[c_].append(k)
    jc_start, jc_idx = [0], []
    for col in jcols:
        jc_idx += col
        jc_start.append(len(jc_idx))
    dr_start = [0] * (n_nl + 1)
    for r_ in dr: dr_start[r_ + 1] += 1
    for i in range(n_nl): dr_start[i + 1] += dr_start[i]

    # ---- H_e * v with the static sparsity of the edge Hessian (straight-line; hnz = lower-triangle nonzeros) ----------
    lines = []
    for i in range(n_q):
        terms = [f"hnz[{k}] * v[{c_}]" for (c_, k) in sorted(hrows[i])]
        lines.append(f"  out[{i}] = " + (" + ".join(terms) if terms else "0.0") + ";")
    out.append(_fn("gen_hmul", "const double* __restrict__ hnz, const double* __restrict__ v, double* __restrict__ out",
                   "  (void)hnz; (void)v;\n" + "\n".join(lines)))

    # ---- rterm (u, u_prev, tvp) ------------------------------------------------------------------------------------------
    inpR = {}
    for i, s in enumerate(u_ids): inpR[s] = f"u[{i}]"
    for i, s in enumerate(up_ids): inpR[s] = f"up[{i}]"
    for i, s in enumerate(tv_ids): inpR[s] = f"tvp[{i}]"
    rt = r.rterm.ids()[0]
    if E.depends_on(rt, pp_ids) or E.depends_on(rt, [s for s in q_ids if s not in u_ids]):
        raise NotImplementedError("rterm may only depend on (u, u_prev, tvp)")
    ruu = list(u_ids) + list(up_ids)
    gR = E.gradient(rt, ruu)
    HR = [E.gradient(g, ruu) for g in gR]
    r_args = "const double* __restrict__ u, const double* __restrict__ up, const double* __restrict__ tvp"
    out.append(_fn("gen_rterm", r_args, "  (void)u; (void)up; (void)tvp;\n  double r;\n" + E.emit_c([("r", rt)], inpR) + "\n  return r;", "double"))
    out.append(_fn("gen_rterm_grad", r_args + ", double* __restrict__ gr",
                   "  (void)u; (void)up; (void)tvp;\n" + E.emit_c([(f"gr[{i}]", e) for i, e in enumerate(gR)], inpR)))
    out.append(_fn("gen_rterm_hess", r_args + ", double* __restrict__ hr",
                   "  (void)u; (void)up; (void)tvp;\n" +
                   E.emit_c([(f"hr[{i * len(ruu) + j}]", HR[i][j]) for i in range(len(ruu)) for j in range(len(ruu))], inpR)))

    # ---- mterm (x_N, tvp_N, p) ------------------------------------------------------------------------------------------------
    inpM = {}
    for i, s in enumerate(xn_ids): inpM[s] = f"x[{i}]"
    for i, s in enumerate(tv_ids): inpM[s] = f"tvp[{i}]"
    for i, s in enumerate(pp_ids): inpM[s] = f"pp[{i}]"
    mt = r.mterm.ids()[0]
    gM = E.gradient(mt, xn_ids)
    HM = [E.gradient(g, xn_ids) for g in gM]
    m_args = "const double* __restrict__ x, const double* __restrict__ tvp, const double* __restrict__ pp"
    out.append(_fn("gen_mterm", m_args, "  (void)x; (void)tvp; (void)pp;\n  double r;\n" + E.emit_c([("r", mt)], inpM) + "\n  return r;", "double"))
    out.append(_fn("gen_mterm_grad", m_args + ", double* __restrict__ gm",
                   "  (void)x; (void)tvp; (void)pp;\n" + E.emit_c([(f"gm[{i}]", e) for i, e in enumerate(gM)], inpM)))
    out.append(_fn("gen_mterm_hess", m_args + ", double* __restrict__ hm",
                   "  (void)x; (void)tvp; (void)pp;\n" +
                   E.emit_c([(f"hm[{i * r.n_x + j}]", HM[i][j]) for i in range(r.n_x) for j in range(r.n_x)], inpM)))
    # ---- root cost (MHE arrival cost) in the root-state symbols ------------------------------------------------------------------
    if r.arrival is not None:
        x0_ids = r.x0sym.ids()
        inpA = {}
        for i, s_ in enumerate(x0_ids): inpA[s_] = f"x[{i}]"
        for i, s_ in enumerate(tv_ids): inpA[s_] = f"tvp[{i}]"
        for i, s_ in enumerate(pp_ids): inpA[s_] = f"pp[{i}]"
        ar = r.arrival.ids()[0]
        gA = E.gradient(ar, x0_ids)
        HA = [E.gradient(g_, x0_ids) for g_ in gA]
        out.append(_fn("gen_arrival", m_args, "  (void)x; (void)tvp; (void)pp;\n  double r;\n" + E.emit_c([("r", ar)], inpA) + "\n  return r;", "double"))
        out.append(_fn("gen_arrival_grad", m_args + ", double* __restrict__ ga",
                       "  (void)x; (void)tvp; (void)pp;\n" + E.emit_c([(f"ga[{i}]", e) for i, e in enumerate(gA)], inpA)))
        out.append(_fn("gen_arrival_hess", m_args + ", double* __restrict__ ha",
                       "  (void)x; (void)tvp; (void)pp;\n" +
                       E.emit_c([(f"ha[{i * r.n_x + j}]", HA[i][j]) for i in range(r.n_x) for j in range(r.n_x)], inpA)))
    # ---- parameter derivatives for the sensitivity right-hand sides (differentiator, _nlpdifferentiator.py:289-304) -------------
    def switch(cases):
        if not cases:
            return "  (void)j;"
        return "  switch (j) {\n" + "\n".join(f"  case {j}: {{\n{body}\n  }} break;" for j, body in enumerate(cases)) + "\n  default: break;\n  }"
    cases = []
    for j in range(n_th):
        outs = ([(f"dq[{i}]", dgL_dth[i][j]) for i in range(n_q)] + [(f"dG[{i}]", dG_dth[i][j]) for i in range(len(G_ids))] +
                [(f"dD[{i}]", dD_dth[i][j]) for i in range(n_nl)])
        cases.append(E.emit_c(outs, inpL))
    out.append(_fn("gen_dtheta", "const int j, " + lag_args + ", double* __restrict__ dq, double* __restrict__ dG, double* __restrict__ dD",
                   "  (void)q; (void)tvp; (void)pp; (void)lam; (void)lnl; (void)sig_l; (void)sig_e; (void)dq; (void)dG; (void)dD;\n" + switch(cases)))
    cases = []
    for j in range(len(tv_ids)):
        cases.append(E.emit_c([(f"o[{i}]", E.gradient(g, [tv_ids[j]])[0]) for i, g in enumerate(gR)], inpR))
    out.append(_fn("gen_rterm_dtvp", "const int j, " + r_args + ", double* __restrict__ o", "  (void)u; (void)up; (void)tvp; (void)o;\n" + switch(cases)))
    cases = []
    for j in range(n_th):
        cases.append(E.emit_c([(f"o[{i}]", E.gradient(g, [theta_ids[j]])[0]) for i, g in enumerate(gM)], inpM))
    out.append(_fn("gen_mterm_dtheta", "const int j, " + m_args + ", double* __restrict__ o", "  (void)x; (void)tvp; (void)pp; (void)o;\n" + switch(cases)))

    # ---- aux -------------------------------------------------------------------------------------------------------------------------
    aux_ids = r.aux.ids()
    out.append(_fn("gen_aux", sig_args + ", double* __restrict__ a",
                   E.emit_c([(f"a[{i}]", e) for i, e in enumerate(aux_ids)], inp)))

    # ---- header with dimensions and tables ------------------------------------------------------------------------------------------------
    n_node = len(r.node_k)
    node_rw = [1.0 / r.n_scenarios[k] if k < r.N else 0.0 for k in r.node_k]
    has_dummies = int((r.P > 0) or (r.n_comb ** r.n_robust > 1)) if r.has_dummies is None else int(r.has_dummies)
    H: List[str] = []
    H.append("// GENERATED by do_mpc_b200.solver.codegen — do not edit.")
    H.append("#pragma once")
    H.append("#ifndef DMB_TABLE\n#ifdef DMB_HOSTSIM\n#define DMB_TABLE static\n#else\n#define DMB_TABLE static __device__\n#endif\n#endif")
    dims = dict(NX=r.n_x, NU=r.n_u, NZ=r.n_z, NTVP=r.n_tvp, NPM=r.n_pm, NCOMB=r.n_comb, NEPS=r.n_eps, NNL=n_nl,
                NY=r.n_y, NNODE=n_node, NEDGE=len(r.edge_k), NHOR=r.N, NW=r.n_w, NG=r.n_g, NP=r.n_p,
                NAUX=len(aux_ids), DISCRETE=int(r.discrete), P_X0=r.p_x0, P_TVP=r.p_tvp, P_P=r.p_p, P_UPREV=r.p_uprev,
                NNZ_J=len(jn), NNZ_H=len(hn), NNZ_D=len(dn), HAS_DUMMIES=has_dummies, NPZ=r.Pz, NPC=r.P,
                NSCEN=r.n_comb ** r.n_robust, NTHETA=n_th, NFREE0=r.n_free0, HAS_ARRIVAL=int(r.arrival is not None))
    for k, v in dims.items():
        H.append(f"#define DMB_{k} {v}")
    H.append(f"#define NLAM_ACTIVE {r.n_g}")
    H.append(_ctab("int", "DMB_EDGE_K", r.edge_k)); H.append(_ctab("int", "DMB_EDGE_PARENT", r.edge_parent))
    H.append(_ctab("int", "DMB_EDGE_CHILD", r.edge_child)); H.append(_ctab("int", "DMB_EDGE_PIDX", r.edge_pidx))
    H.append(_ctab("double", "DMB_EDGE_OMEGA", [_dbl(v) for v in r.edge_omega]))
    H.append(_ctab("int", "DMB_NODE_K", r.node_k)); H.append(_ctab("int", "DMB_NODE_FIRST_EDGE", r.node_first_edge))
    H.append(_ctab("int", "DMB_NODE_NEDGE", r.node_n_edges)); H.append(_ctab("int", "DMB_NODE_PARENT_EDGE", r.node_parent_edge))
    H.append(_ctab("double", "DMB_NODE_RW", [_dbl(v) for v in node_rw]))
    # nodes are numbered level by level (k = 0..N): first node of every level, for the level-parallel tree sweeps
    lvl = [0]
    for k in range(r.N + 1):
        lvl.append(lvl[-1] + r.n_scenarios[k])
    assert lvl[-1] == n_node and all(r.node_k[n] == k for k in range(r.N + 1) for n in range(lvl[k], lvl[k + 1]))
    H.append(_ctab("int", "DMB_LEVEL_START", lvl))
    H.append(_ctab("int", "DMB_NODE_X_FLAT", r.node_x_flat)); H.append(_ctab("int", "DMB_NODE_U_FLAT", r.node_u_flat))
    H.append(_ctab("int", "DMB_NODE_EPS_FLAT", r.node_eps_flat))
    H.append(_ctab("int", "DMB_EDGE_YX_FLAT", r.edge_yx_flat)); H.append(_ctab("int", "DMB_EDGE_YZ_FLAT", r.edge_yz_flat))
    H.append(_ctab("int", "DMB_EDGE_G_FLAT", r.edge_g_flat))
    H.append(_ctab("int", "DMB_XF_SEL", r.xf_sel))
    xf_inv = [-1] * r.n_y
    for a_, j_ in enumerate(r.xf_sel): xf_inv[int(j_)] = a_
    H.append(_ctab("int", "DMB_XF_INV", xf_inv))
    S_ = r.n_comb ** r.n_robust
    ks_node = []
    for k in range(r.N):
        nodes_k = [n for n in range(n_node) if r.node_k[n] == k]
        ks_node += [nodes_k[min(s_, len(nodes_k) - 1)] for s_ in range(S_)]
    H.append(_ctab("int", "DMB_KS_NODE", ks_node))
    H.append(_ctab("double", "DMB_INV_XSCALE", [_dbl(1.0 / v) for v in r.x_scaling]))
    H.append(_ctab("double", "DMB_INV_USCALE", [_dbl(1.0 / v) for v in r.u_scaling]))
    H.append(_ctab("double", "DMB_NL_UB", [_dbl(v) for v in r.nl_ub]))
    H.append(_ctab("int", "DMB_J_ROW", jr)); H.append(_ctab("int", "DMB_J_COL", jc))
    H.append(_ctab("int", "DMB_H_ROW", hr)); H.append(_ctab("int", "DMB_H_COL", hc))
    H.append(_ctab("int", "DMB_D_ROW", dr)); H.append(_ctab("int", "DMB_D_COL", dcol))
    H.append(_ctab("int", "DMB_JR_START", jr_start)); H.append(_ctab("int", "DMB_DR_START", dr_start))
    H.append(_ctab("int", "DMB_JC_START", jc_start)); H.append(_ctab("int", "DMB_JC_IDX", jc_idx if jc_idx else [0]))
    H.append(_ctab("int", "DMB_HR_START", hr_start)); H.append(_ctab("int", "DMB_HR_COL", hr_col))
    H.append(_ctab("int", "DMB_HR_IDX", hr_idx))
    H.append(f"#define DMB_NNZ_HR {max(len(hr_col), 1)}")
    # internal <-> flat (reference order) maps
    n_edge = len(r.edge_k)
    w2f = []
    for n in range(n_node):
        w2f += [int(r.node_x_flat[n]) + i for i in range(r.n_x)]
        w2f += [int(r.node_u_flat[n]) + i if r.node_u_flat[n] >= 0 else -1 for i in range(r.n_u)]
        w2f += [int(r.node_eps_flat[n]) + i if r.node_eps_flat[n] >= 0 else -1 for i in range(r.n_eps)]
    for e in range(n_edge):
        if r.discrete:
            w2f += [int(r.edge_yz_flat[e]) + i for i in range(r.n_z)] + [-1] * r.n_x
        else:
            w2f += [int(r.edge_yx_flat[e]) + i for i in range(r.P * r.n_x)]
            w2f += [int(r.edge_yz_flat[e]) + i for i in range(r.P * r.n_z)]
    l2f = list(range(r.n_x))
    for e in range(n_edge):
        g0 = int(r.edge_g_flat[e])
        if r.discrete:
            l2f += [g0 + i for i in range(r.n_z)] + [-1] * r.n_x
            l2f += [g0 + r.n_z + i for i in range(r.n_x)]
            l2f += [g0 + r.n_z + r.n_x + i for i in range(n_nl)]
        else:
            l2f += [g0 + i for i in range(r.n_y + r.n_x + n_nl)]
    if r.w2flat is not None:
        w2f, l2f = [int(v) for v in r.w2flat], [int(v) for v in r.lam2flat]
        assert len(l2f) == r.n_x + n_edge * (r.n_y + r.n_x + n_nl), (len(l2f), r.n_x + n_edge * (r.n_y + r.n_x + n_nl))
    assert len(w2f) == n_node * n_vp + n_edge * r.n_y, (len(w2f), n_node * n_vp + n_edge * r.n_y)
    init_from = [-1] * len(w2f) if r.w_init_from is None else [int(v) for v in r.w_init_from]
    # least-squares multiplier estimate: the copies of a lifted global variable share the unit weight of the single variable of
    # the flat NLP (weight 1/#copies each), which makes the estimate of the reference multipliers identical to the flat one
    grp = {}
    for i_, v in enumerate(init_from):
        if v >= 0: grp.setdefault(v, []).append(i_)
    lsw = [1.0] * len(w2f)
    for root_, members in grp.items():
        for i_ in [root_] + members: lsw[i_] = 1.0 / (1 + len(members))
    H.append(_ctab("double", "DMB_LS_WEIGHT", [_dbl(v) for v in lsw] if grp else [_dbl(1.0)]))
    H.append(f"#define DMB_HAS_INIT_FROM {int(any(v >= 0 for v in init_from))}")
    H.append(_ctab("int", "DMB_W_INIT_FROM", init_from if any(v >= 0 for v in init_from) else [0]))
    H.append("#define DMB_W2FLAT_INIT {" + ", ".join(str(v) for v in w2f) + "}")
    H.append("#define DMB_LAM2FLAT_INIT {" + ", ".join(str(v) for v in l2f) + "}")
    H.append(f"DMB_TABLE const int DMB_W2FLAT[{len(w2f)}] = DMB_W2FLAT_INIT;")
    H.append(f"DMB_TABLE const int DMB_LAM2FLAT[{len(l2f)}] = DMB_LAM2FLAT_INIT;")
    H.append("")
    src = "\n".join(H) + "\n" + "\n".join(out)
    return src


def source_hash(src: str, extra: str = "") -> str:
    return hashlib.sha256((src + extra).encode()).hexdigest()[:16]
