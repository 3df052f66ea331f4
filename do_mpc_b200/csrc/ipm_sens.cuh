// Parametric sensitivities dx*/dp of a solved instance — the batched counterpart of the reference's
// NLPDifferentiator / DoMPCDifferentiator (do_mpc/differentiator/_nlpdifferentiator.py:289-304 sensitivity
// matrices, :345-375 active set, :469-555 reduced linear system), called after the solve like
// `nlp_diff.differentiate()` follows `mpc.make_step()`.
//
// The reference builds  A = d(grad_z L)/dz,  B = d(grad_z L)/dp  symbolically, deletes the rows/columns of the
// inactive constraints and solves  A_red S = -B_red  with a sparse direct solver.  Here the same system is solved
// with the structured elimination of the IPM (edge LU + Riccati sweep over the tree):
//   * active-set reduction == bound / inequality weights: Sigma = BIG on active rows (|x - bound| <= tol), 0 on
//     inactive ones (an active bound pins its variable, an inactive one disappears),
//   * one factorisation at the solution (unscaled NLP: d_f = d_c = 1, multipliers = lam_g as returned),
//   * one right-hand side per parameter column, each solved by the iterative-refinement loop of the IPM started
//     from the zero vector (first pass = the solve, further passes remove the cancellation error of BIG).
// Right-hand sides (column theta of B):  x0_i -> init row; u_prev_i -> rterm cross Hessian at the root;
// tvp_k,i / p_c,i -> generated d/dtheta of (grad_q L_e, G_e, nl_e) on the edges that see that parameter
// (+ rterm / mterm terms).
#pragma once
#include "ipm_solve.cuh"

#define DMB_SENS_OK 0
#define DMB_SENS_INDEFINITE 1       // solved, but the reduced Hessian on the active set is indefinite (not a strict minimiser)
#define DMB_SENS_FAILED 2           // singular collocation block or singular control block (reference: rank-deficient A)
#define DMB_SENS_SKIPPED 3          // instance not solved (status not 0/1): output NaN

// ---------------------------------------------------------------------------------------------------
// load the solution, determine the active set and factorise (L D L' control blocks: the system only has to be regular,
// like the reference's direct solve).  Returns DMB_SENS_OK / _INDEFINITE / _FAILED.
// ---------------------------------------------------------------------------------------------------
DMB_DEV int sens_factor(const Ctx& c, Inst& I, const double* x_flat, const double* lam_g, const double* lbA, const double* ubA,
                        int relaxed_rule, double tol, double big,
                           double* tile, double* Qs, DmbShared* S) {
  I.df = 1.0;
  load_initial_guess(c, x_flat, nullptr, nullptr, nullptr, I.w);
  for (int i = c.tid; i < NLAM; i += c.nthr) {
    const int fl = DMB_LAM2FLAT[i];
    I.dc[i] = 1.0; I.rest[i] = 0.0;
    I.lam[i] = fl >= 0 ? lam_g[fl] : 0.0;
  }
  CTA_SYNC();
#if DMB_DISCRETE
  // the end-of-interval state is an auxiliary unknown here: its defining row carries the continuity multiplier
  for (int idx = c.tid; idx < NEDGE * NX; idx += c.nthr) {
    const int e = idx / NX, i = idx - e * NX;
    I.lam[LAM_EDGE(e) + DMB_NZ + i] = I.lam[LAM_EDGE(e) + NY + i];
  }
#endif
  for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) I.s[i] = 0.0;
  CTA_SYNC();
  double f, theta;
  eval_point(c, I, I.w, I.s, I.res, S->red, &f, &theta, nullptr);      // res[nl rows] = d(x) (s = 0)
  // active set (reference :345-375: primal test |x - bound| <= active_set_tol against the bounds lbA / ubA — the user's
  // bounds like the reference, or the relaxed bounds the IPM worked with) -> Sigma through the multiplier slots
  // (Sigma = zL / (w - lb) with the relaxed lb of I: only "zero or BIG" matters)
  for (int i = c.tid; i < NWI; i += c.nthr) {
    const double lo = I.lb[i], hi = I.ub[i];
    double w = I.w[i], zl = 0.0, zu = 0.0;
    if (lo > -DMB_INF && fabs(w - lbA[i]) <= tol) { if (w == lo) { w = lo + 1e-13 * fmax(1.0, fabs(lo)); I.w[i] = w; } zl = big * (w - lo); }
    if (hi < DMB_INF && fabs(ubA[i] - w) <= tol) { if (w == hi) { w = hi - 1e-13 * fmax(1.0, fabs(hi)); I.w[i] = w; } zu = big * (hi - w); }
    I.zL[i] = zl; I.zU[i] = zu;
  }
#if NNL > 0
  for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) {
    const int e = i / NNL, j = i % NNL;
    const double d = I.res[LAM_EDGE(e) + NY + NX + j];
    I.vU[i] = fabs(d - (relaxed_rule ? nl_ub_relaxed(I, e, j) : DMB_NL_UB[j])) <= tol ? big : 0.0;
    I.s[i] = nl_ub_relaxed(I, e, j) - 1.0;                              // slack distance 1: Sigma_s = vU
  }
#endif
  if (c.tid == 0) S->fail = 0;
  CTA_SYNC();
  eval_derivatives(c, I);
  factor_edges(c, I, I.rest, 0.0, 0.0, 0, tile, &S->fail);
  if (S->fail) return DMB_SENS_FAILED;
  riccati_backward_t<true>(c, I, 0.0, 0.0, 0.0, 0, Qs, &S->flag);
  const int bad = S->flag;
  CTA_SYNC();
  return (bad & 1) ? DMB_SENS_FAILED : ((bad & 2) ? DMB_SENS_INDEFINITE : DMB_SENS_OK);
}

// ---------------------------------------------------------------------------------------------------
// right-hand side of opt_p column pj:  I.go (node rows) / EB_GY (edge unknowns) <- d(grad L)/dtheta,
// I.res <- dc/dtheta.  I.dw, I.lamn, I.ds are reset (the refinement loop starts from zero).
// ---------------------------------------------------------------------------------------------------
DMB_DEV void sens_rhs(const Ctx& c, const Inst& I, int pj) {
  for (int i = c.tid; i < NWI; i += c.nthr) { I.go[i] = 0.0; I.dw[i] = 0.0; }
  for (int i = c.tid; i < NLAM; i += c.nthr) { I.res[i] = 0.0; I.lamn[i] = 0.0; }
  for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) { I.ds[i] = 0.0; I.s[i] = 0.0; }      // I.s: Sigma_s dd/dtheta of this column
  for (int idx = c.tid; idx < NEDGE * NY; idx += c.nthr) I.eb[(size_t)(idx / NY) * EB_SZ + EB_GY + idx % NY] = 0.0;
  for (int idx = c.tid; idx < NEDGE * NVP; idx += c.nthr) I.eb[(size_t)(idx / NVP) * EB_SZ + EB_GCON + idx % NVP] = 0.0;
  CTA_SYNC();
  if (pj >= DMB_P_X0 && pj < DMB_P_X0 + NX) {
    const int i = pj - DMB_P_X0;
    if (c.tid == 0) I.res[i] = -DMB_INV_XSCALE[i];                   // init row x_root - x0/scaling
  } else if (pj >= DMB_P_UPREV && pj < DMB_P_UPREV + NU) {
    const int i = pj - DMB_P_UPREV;
    const double* w0 = I.node;                                        // weighted rterm Hessian of the root, (u, u_prev) block
    for (int a = c.tid; a < NU; a += c.nthr) I.go[NODE_OFF(0) + NX + a] = w0[WN_RH + a * 2 * NU + NU + i] * DMB_INV_USCALE[i];
  }
#if DMB_NTHETA > 0
  else {
    int jth, k_sel = -1, c_sel = -1;
    if (pj >= DMB_P_TVP && pj < DMB_P_TVP + (NHOR + 1) * DMB_NTVP) { k_sel = (pj - DMB_P_TVP) / (DMB_NTVP > 0 ? DMB_NTVP : 1); jth = (pj - DMB_P_TVP) - k_sel * DMB_NTVP; }
    else { c_sel = (pj - DMB_P_P) / (DMB_NPM > 0 ? DMB_NPM : 1); jth = DMB_NTVP + (pj - DMB_P_P) - c_sel * DMB_NPM; }
    for (int e = c.tid; e < NEDGE; e += c.nthr) {
      if (k_sel >= 0 ? EDGE_K_OF(e) != k_sel : EDGE_PIDX_OF(e) != c_sel) continue;
      double q[NQ], lg[NY > 0 ? NY : 1], ln[NNL > 0 ? NNL : 1], dq[NQ], dG[NY > 0 ? NY : 1], dD[NNL > 0 ? NNL : 1];
      gather_q(I.w, e, q);
      const double* le = I.lam + LAM_EDGE(e);
#pragma unroll
      for (int i = 0; i < NY; ++i) lg[i] = le[i];
#pragma unroll
      for (int i = 0; i < NNL; ++i) ln[i] = le[NY + NX + i];
#pragma unroll
      for (int i = 0; i < NQ; ++i) dq[i] = 0.0;
#pragma unroll
      for (int i = 0; i < NY; ++i) dG[i] = 0.0;
#pragma unroll
      for (int i = 0; i < NNL; ++i) dD[i] = 0.0;
      gen_dtheta(jth, q, edge_tvp(I.p, e), edge_pp(I.p, e), lg, EDGE_OMEGA_OF(e), 1.0, ln, dq, dG, dD);
      double* eb = I.eb + (size_t)e * EB_SZ;
#if NNL > 0
      // active inequality rows: condensed  J_d' Sigma_s (J_d dq + dd/dtheta)
      for (int i = 0; i < NNL; ++i) {
        const double v = eb[EB_SIGS + i] * dD[i];
        I.s[e * NNL + i] = v;
        if (v != 0.0) for (int t = DMB_DR_START[i]; t < DMB_DR_START[i + 1]; ++t) dq[DMB_D_COL[t]] += eb[EB_DNZ + t] * v;
      }
#endif
#pragma unroll
      for (int i = 0; i < NVP; ++i) eb[EB_GCON + i] = dq[i];
#pragma unroll
      for (int i = 0; i < NY; ++i) { eb[EB_GY + i] = dq[NVP + i]; I.res[LAM_EDGE(e) + i] = dG[i]; }
    }
    CTA_SYNC();
    for (int idx = c.tid; idx < NNODE * NVP; idx += c.nthr) {
      const int n = idx / NVP, i = idx - n * NVP;
      if (NODE_K_OF(n) == NHOR) continue;
      double s = 0.0;
      for (int j = 0; j < NODE_NEDGE(n); ++j) s += I.eb[(size_t)(NODE_FIRST_EDGE(n) + j) * EB_SZ + EB_GCON + i];
      I.go[idx] = s;
    }
    CTA_SYNC();
    // terminal cost (leaf nodes) and input penalty (time-varying weights) — rare, cheap: thread per node / one thread
    for (int n = c.tid; n < NNODE; n += c.nthr) {
      if (NODE_K_OF(n) != NHOR) continue;
      const int pe = NODE_PARENT_EDGE(n);
      if (k_sel >= 0 ? k_sel != NHOR : EDGE_PIDX_OF(pe) != c_sel) continue;
      double o[NX];
#pragma unroll
      for (int i = 0; i < NX; ++i) o[i] = 0.0;
      gen_mterm_dtheta(jth, I.w + NODE_OFF(n), p_tvp_at(I.p, NHOR), edge_pp(I.p, pe), o);
      for (int i = 0; i < NX; ++i) I.go[NODE_OFF(n) + i] += EDGE_OMEGA_OF(pe) * o[i];
    }
#if DMB_NTVP > 0
    if (k_sel >= 0 && k_sel < NHOR && c.tid == 0) {
      for (int n = 0; n < NNODE; ++n) {
        if (NODE_K_OF(n) != k_sel) continue;
        double up[NU], o[2 * NU];
        for (int a = 0; a < NU; ++a) { up[a] = node_uprev(I, I.w, n, a); o[a] = 0.0; o[NU + a] = 0.0; }
        gen_rterm_dtvp(jth, I.w + NODE_OFF(n) + NX, up, I.p + DMB_P_TVP + k_sel * DMB_NTVP, o);
        const int pe = NODE_PARENT_EDGE(n);
        for (int a = 0; a < NU; ++a) {
          I.go[NODE_OFF(n) + NX + a] += NODE_RW_OF(n) * o[a];
          if (pe >= 0) I.go[NODE_OFF(EDGE_PARENT_OF(pe)) + NX + a] += NODE_RW_OF(n) * o[NU + a];
        }
      }
    }
#endif
  }
#endif
  CTA_SYNC();
}

// ---------------------------------------------------------------------------------------------------
// solve the factorised system for the right-hand side prepared by sens_rhs: I.dw = dx/dtheta (internal layout),
// I.lamn = d lam_g / dtheta.
// ---------------------------------------------------------------------------------------------------
DMB_DEV void sens_solve(const Ctx& c, const Inst& I, int max_refine, double ratio_max, double* tile, double* Qs, DmbShared* S) {
  for (int it = 0; it <= max_refine; ++it) {
    double nr[3];
    kkt_residual(c, I, I.res, 0.0, 0.0, 0.0, S->red, nr, tile);
    if (it > 0) {
      const double ratio = nr[0] / (fmin(nr[1], 1e6 * nr[2]) + nr[2] + 1e-300);
      if (!(ratio > ratio_max)) break;
    }
    for (int i = c.tid; i < NWI; i += c.nthr) I.wt[i] = I.dw[i];
    for (int i = c.tid; i < NLAM; i += c.nthr) I.lamr[i] = I.lamn[i];
    for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) I.st[i] = I.ds[i];
    CTA_SYNC();
    resolve_edges(c, I, 0.0, tile);
    riccati_vector_t<true>(c, I, 0.0, Qs);
    forward_recover(c, I, I.rest, 0.0, 0, 1, tile);
    for (int i = c.tid; i < NWI; i += c.nthr) I.dw[i] += I.wt[i];
    for (int i = c.tid; i < NLAM; i += c.nthr) I.lamn[i] += I.lamr[i];
    for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) I.ds[i] += I.st[i];
    CTA_SYNC();
  }
}

// ---------------------------------------------------------------------------------------------------
// one column -> flat outputs (reference layouts).  dx: [NW] or the first input [NU]; dlam: [NG + NW] (lam_g | lam_x).
// ---------------------------------------------------------------------------------------------------
DMB_DEV void sens_store(const Ctx& c, const Inst& I, int mode, double* dx, double* dlam) {
  if (mode == 1) {
    for (int a = c.tid; a < NU; a += c.nthr) dx[a] = I.dw[NODE_OFF(0) + NX + a];
  } else {
    for (int i = c.tid; i < DMB_NW; i += c.nthr) dx[i] = 0.0;
    if (dlam) for (int i = c.tid; i < DMB_NG + DMB_NW; i += c.nthr) dlam[i] = 0.0;
    CTA_SYNC();
    for (int i = c.tid; i < NWI; i += c.nthr) {
      const int fl = DMB_W2FLAT[i];
      if (fl < 0) continue;
      dx[fl] = I.dw[i];
      if (dlam) {
        double sig = 0.0;
        if (I.zL[i] != 0.0) sig += I.zL[i] / (I.w[i] - I.lb[i]);
        if (I.zU[i] != 0.0) sig += I.zU[i] / (I.ub[i] - I.w[i]);
        dlam[DMB_NG + fl] = sig * I.dw[i];
      }
    }
    if (dlam)
      for (int i = c.tid; i < NLAM; i += c.nthr) {
        const int fl = DMB_LAM2FLAT[i];
        if (fl < 0) continue;
        double v = I.lamn[i];
#if NNL > 0
        if (i >= NX) { const int j = (i - NX) % NLE; if (j >= NY + NX) v += I.s[((i - NX) / NLE) * NNL + (j - NY - NX)]; }
#endif
        dlam[fl] = v;
      }
  }
  CTA_SYNC();
}
