// Iterative refinement of the Newton step (IPOPT: min_refinement_steps / residual_ratio_max, PDFullSpaceSolver).
//
// The structured elimination (edge LU + Riccati sweep) is not backward stable like a pivoted LDL^T: with
// Sigma = z/s ~ 1e12 near the end of badly scaled problems (industrial_poly) the computed step satisfies the
// linear KKT system only to ~1e-8, which stalls the dual infeasibility just above tol.  Here the residual of the
// (condensed) KKT system is evaluated with the original sparse data and the step is corrected by re-solving
// with the existing factors (edge LU, Y, P, K):
//     rho_d = g_phi + (W + Sigma + delta) dw + J' lam+        rho_p = c + J dw
//     (ddw, dlam) = S(rho_d, rho_p) ;  dw -= ... (signs as in the primary solve: S solves  K [dw; lam+] = -[g; c]).
#pragma once
#include "ipm_core.cuh"

// ---------------------------------------------------------------------------------------------------
// residual of the linear system for the current (dw, ds, lamn).  rd -> I.gl (NWI), rp -> I.rest (NLAM).
// Uses the per-edge scratch EB_GCON (NQ) for the edge-local part.  Returns max-norms through out[3]:
// |rho|, |sol|, |rhs|.
// ---------------------------------------------------------------------------------------------------
DMB_PHASE void kkt_residual(const Ctx& c, const Inst& I, const double* res, double mu, double kappa_d, double delta,
                            double* red, double* out, double* tile_base) {
  double* rd = I.gl; double* rp = I.rest;
  double nrm_rho = 0.0, nrm_sol = 0.0, nrm_rhs = 0.0;
  // edge-local part, one lane group per edge (r1e profile: with one thread per edge this section kept a single warp busy for
  // ~4 % of the wall time).  Scratch in the (idle) LU area of the group's tile; rows of H through its CSR view, J dq through
  // the row-major nonzero order, J' lam through the CSC view (DMB_JC_*).
  double* T = tile_base + (size_t)c.grp * TILE_SZ;
  double* dq = T;                               // NQ
  double* ln = dq + NQ;                         // NY + NX : lam_G+, lam_C+
  double* jz = ln + NY + NX;                    // NNZ_J   : row-scaled Jacobian nonzeros
  double* hq = jz + DMB_NNZ_J;                  // NQ
  double* ydv = hq + NQ;                        // NNL     : y_d+ of the condensed inequality rows
  double* dz = ydv + NNL;                       // NNZ_D
  static_assert(2 * NQ + NY + NX + DMB_NNZ_J + NNL + DMB_NNZ_D <= STG_HNZ, "kkt_residual scratch does not fit into the LU area of the tile");
  for (int e0 = 0; e0 < NEDGE; e0 += c.ngrp) {
    const int e = e0 + c.grp;
    const bool act = e < NEDGE;
    double* eb = I.eb + (size_t)(act ? e : 0) * EB_SZ;
    const double* dc = I.dc + LAM_EDGE(act ? e : 0);
    if (act) {
      DMB_COPY_BATCH(DMB_NNZ_H, c.gl, GSZ, eb[EB_HNZ + idx], T[STG_HNZ + idx] = val);
      DMB_COPY_BATCH(DMB_NNZ_J, c.gl, GSZ, eb[EB_JNZ + idx], jz[idx] = dc[DMB_J_ROW[idx]] * val);
      DMB_COPY_BATCH(NY, c.gl, GSZ, eb[EB_GY + idx], T[STG_GY + idx] = val);
      DMB_COPY_BATCH(NY, c.gl, GSZ, eb[EB_SIGY + idx], T[STG_SIGY + idx] = val);
      const double* dvp = I.dw + NODE_OFF(EDGE_PARENT_OF(e));
      const double* dy = I.dw + EDGE_OFF(e);
      for (int i = c.gl; i < NQ; i += GSZ) dq[i] = i < NVP ? dvp[i] : dy[i - NVP];
      for (int i = c.gl; i < NY + NX; i += GSZ) ln[i] = I.lamn[LAM_EDGE(e) + i];
#if NNL > 0
      for (int i = c.gl; i < DMB_NNZ_D; i += GSZ) dz[i] = eb[EB_DNZ + i];
#endif
    }
    WARP_SYNC();
#if NNL > 0
    if (act)
      for (int i = c.gl; i < NNL; i += GSZ) {
        double jd = 0.0;
        for (int t = DMB_DR_START[i]; t < DMB_DR_START[i + 1]; ++t) jd += dz[t] * dq[DMB_D_COL[t]];
        ydv[i] = (eb[EB_SIGS + i] + delta) * jd + eb[EB_RD + i] + delta * res[LAM_EDGE(e) + NY + NX + i];   // = y_d+ of the condensed system
        rp[LAM_EDGE(e) + NY + NX + i] = 0.0;                               // eliminated exactly
      }
    WARP_SYNC();
#endif
    if (act) {
      // rows of (W + Sigma + delta) dq + J' lam+ (+ g for the edge unknowns)
      for (int i = c.gl; i < NQ; i += GSZ) {
        double s = 0.0;
        for (int t = DMB_HR_START[i]; t < DMB_HR_START[i + 1]; ++t) s += T[STG_HNZ + DMB_HR_IDX[t]] * dq[DMB_HR_COL[t]];
        for (int t = DMB_JC_START[i]; t < DMB_JC_START[i + 1]; ++t) { const int k = DMB_JC_IDX[t]; s += jz[k] * ln[DMB_J_ROW[k]]; }
#if NNL > 0
        for (int t = 0; t < DMB_NNZ_D; ++t) if (DMB_D_COL[t] == i) s += dz[t] * ydv[DMB_D_ROW[t]];
#endif
        if (i >= NVP) {
          const int j = i - NVP;
          s += (T[STG_SIGY + j] + delta) * dq[i] + T[STG_GY + j];
#pragma unroll
          for (int a = 0; a < NX; ++a) if (DMB_XF_SEL[a] == j) s += ln[NY + a];    // + E' lam_C+ on the end-of-interval state
        }
        hq[i] = s;
      }
      // rows of c + J dq
      for (int r = c.gl; r < NY; r += GSZ) {
        double jd = 0.0;
        for (int k = DMB_JR_START[r]; k < DMB_JR_START[r + 1]; ++k) jd += jz[k] * dq[DMB_J_COL[k]];
        const double rr = res[LAM_EDGE(e) + r] + jd;
        rp[LAM_EDGE(e) + r] = rr;
        nrm_rho = fmax(nrm_rho, fabs(rr));
        nrm_rhs = fmax(nrm_rhs, fmax(fabs(T[STG_GY + r]), fabs(res[LAM_EDGE(e) + r])));
        nrm_sol = fmax(nrm_sol, fmax(fabs(dq[NVP + r]), fabs(ln[r])));
      }
      const double* dxc = I.dw + NODE_OFF(EDGE_CHILD_OF(e));
      for (int a = c.gl; a < NX; a += GSZ) {
        const double rr = res[LAM_EDGE(e) + NY + a] + dq[NVP + DMB_XF_SEL[a]] - dxc[a];
        rp[LAM_EDGE(e) + NY + a] = rr;
        nrm_rho = fmax(nrm_rho, fabs(rr));
      }
    }
    WARP_SYNC();
    if (act) {
      for (int i = c.gl; i < NQ; i += GSZ) {
        if (i < NVP) eb[EB_GCON + i] = hq[i];
        else { rd[EDGE_OFF(e) + i - NVP] = hq[i]; nrm_rho = fmax(nrm_rho, fabs(hq[i])); }
      }
    }
    WARP_SYNC();
  }
  CTA_SYNC();
  for (int idx = c.tid; idx < NNODE * NVP; idx += c.nthr) {
    const int n = idx / NVP, i = idx - n * NVP;
    const int k = NODE_K_OF(n);
    if (k == NHOR && i >= NX) { rd[idx] = 0.0; continue; }
    const double* wn = I.node + (size_t)n * WN_NODE;
    const double lo = I.lb[idx], hi = I.ub[idx], wx = I.w[idx], d = I.dw[idx];
    const bool hl = lo > -DMB_INF, hu = hi < DMB_INF;
    double sig = delta, g = I.go[idx];
    if (hl) { const double rs = 1.0 / (wx - lo); sig += I.zL[idx] * rs; g -= mu * rs; if (!hu) g += kappa_d * mu; }
    if (hu) { const double rs = 1.0 / (hi - wx); sig += I.zU[idx] * rs; g += mu * rs; if (!hl) g -= kappa_d * mu; }
    double r = g + sig * d;
    nrm_rhs = fmax(nrm_rhs, fabs(g)); nrm_sol = fmax(nrm_sol, fabs(d));
    if (k < NHOR)
      for (int j = 0; j < NODE_NEDGE(n); ++j) r += I.eb[(size_t)(NODE_FIRST_EDGE(n) + j) * EB_SZ + EB_GCON + i];
    const int pe = NODE_PARENT_EDGE(n);
    if (i < NX) {
      if (pe >= 0) {
        r -= I.lamn[LAM_EDGE(pe) + NY + i];
        if (k == NHOR) {
          double hm[NX * NX];
          gen_mterm_hess(I.w + NODE_OFF(n), p_tvp_at(I.p, NHOR), edge_pp(I.p, pe), hm);
          const double wgt = I.df * EDGE_OMEGA_OF(pe);
          for (int b = 0; b < NX; ++b) r += wgt * hm[i * NX + b] * I.dw[NODE_OFF(n) + b];
        }
      } else {
        r += I.lamn[i];
#if DMB_HAS_ARRIVAL
        double ha[NX * NX];
        gen_arrival_hess(I.w + NODE_OFF(0), p_tvp_at(I.p, 0), I.p + DMB_P_P, ha);
        for (int b = 0; b < NX; ++b) r += I.df * ha[i * NX + b] * I.dw[NODE_OFF(0) + b];
#endif
      }
    } else if (i < NX + NU) {
      const int iu = i - NX;
      // own rterm: rows of u ; children's rterm: rows of their u_prev = this u
      const double* dup = pe >= 0 ? I.dw + NODE_OFF(EDGE_PARENT_OF(pe)) + NX : nullptr;
      for (int b = 0; b < NU; ++b) {
        r += wn[WN_RH + iu * 2 * NU + b] * I.dw[NODE_OFF(n) + NX + b];
        if (dup) r += wn[WN_RH + iu * 2 * NU + NU + b] * dup[b];
      }
      for (int j = 0; j < NODE_NEDGE(n); ++j) {
        const int cn = EDGE_CHILD_OF(NODE_FIRST_EDGE(n) + j);
        if (NODE_K_OF(cn) == NHOR) continue;
        const double* wc = I.node + (size_t)cn * WN_NODE;
        for (int b = 0; b < NU; ++b)
          r += wc[WN_RH + (NU + iu) * 2 * NU + b] * I.dw[NODE_OFF(cn) + NX + b] + wc[WN_RH + (NU + iu) * 2 * NU + NU + b] * I.dw[NODE_OFF(n) + NX + b];
      }
    }
    rd[idx] = r;
    nrm_rho = fmax(nrm_rho, fabs(r));
  }
  for (int i = c.tid; i < NX; i += c.nthr) {
    const double r = i < NXF ? res[i] + I.dw[NODE_OFF(0) + i] : 0.0;      // free root components have no initial-condition row
    rp[i] = r;
    nrm_rho = fmax(nrm_rho, fabs(r));
  }
  out[0] = block_max(c, nrm_rho, red);
  out[1] = block_max(c, nrm_sol, red);
  out[2] = block_max(c, nrm_rhs, red);
}

// ---------------------------------------------------------------------------------------------------
// right-hand-side dependent part of the edge elimination for a new rhs (rd = I.gl, rp = I.rest), using the stored
// LU / Y: y0 = -G_y^{-1} rp_G, reduced gradient, Y'y0, b.  One lane group per edge, data staged in the tile.
// Overwrites ER_GH / ER_YTY0 / ER_B, the y0 column of the tile image (EB_IMG) and EB_GY (:= rd_y) — the primary values are no longer
// needed once the primary step has been recovered (the primary g_y in EB_GY is kept for further residual evaluations).
// ---------------------------------------------------------------------------------------------------
DMB_PHASE void resolve_edges(const Ctx& c, const Inst& I, double delta, double* tile_base) {
  const double* rd = I.gl; const double* rp = I.rest;
  double* T = tile_base + (size_t)c.grp * TILE_SZ;
  double* bv = T + STG_DQ;                      // NY: permuted rhs / solution
  double* zq = bv + NY;                         // NQ: H (0; y0)
  for (int e0 = 0; e0 < NEDGE; e0 += c.ngrp) {
    const int e = e0 + c.grp;
    const bool act = e < NEDGE;
    double* eb = I.eb + (size_t)(act ? e : 0) * EB_SZ;
    double* er = I.er + (size_t)(act ? e : 0) * ER_SZ;
    if (act) {
      DMB_COPY_BATCH(DMB_NNZ_H, c.gl, GSZ, eb[EB_HNZ + idx], T[STG_HNZ + idx] = val);
      DMB_COPY_BATCH(NY, c.gl, GSZ, eb[EB_SIGY + idx], T[STG_SIGY + idx] = val);
      for (int m = c.gl; m < NY; m += GSZ) bv[m] = rp[LAM_EDGE(e) + (int)eb[EB_PIV + m]];      // P rp_G
    }
    load_tile_image(c, I, eb, T, act);                 // [LU | Yext] image (one TMA bulk copy)
    WARP_SYNC();
    // L b' = P rp_G (unit lower), U x = b' (reciprocal diagonal): column sweeps, lanes over rows
    for (int k = 0; k < NY; ++k) {
      if (act) { const double bk = bv[k]; for (int r = k + 1 + c.gl; r < NY; r += GSZ) bv[r] -= T[r * TILE_LD + k] * bk; }
      WARP_SYNC();
    }
    for (int k = NY - 1; k >= 0; --k) {
      double xk = 0.0;
      if (act) {
        xk = bv[k] * T[k * TILE_LD + k];
        for (int r = c.gl; r < k; r += GSZ) bv[r] -= T[r * TILE_LD + k] * xk;
      }
      WARP_SYNC();
      if (act && c.gl == 0) bv[k] = -xk;          // y0 = -x
      WARP_SYNC();
    }
    if (act) {
      // zq = H_e (0; y0) + Sigma_y y0 ; staged generated product by one lane (straight-line)
      if (c.gl == 0) {
        double v[NQ], z[NQ];
#pragma unroll
        for (int i = 0; i < NVP; ++i) v[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NY; ++i) v[NVP + i] = bv[i];
        gen_hmul(T + STG_HNZ, v, z);
#pragma unroll
        for (int i = 0; i < NY; ++i) z[NVP + i] += T[STG_SIGY + i] * bv[i];
#if NNL > 0
        for (int i = 0; i < NNL; ++i) {
          double jd = rp[LAM_EDGE(e) + NY + NX + i];          // right-hand side of the condensed row (0 in a refinement solve)
          for (int t = DMB_DR_START[i]; t < DMB_DR_START[i + 1]; ++t) if (DMB_D_COL[t] >= NVP) jd += eb[EB_DNZ + t] * bv[DMB_D_COL[t] - NVP];
          for (int t = DMB_DR_START[i]; t < DMB_DR_START[i + 1]; ++t) z[DMB_D_COL[t]] += eb[EB_DNZ + t] * (eb[EB_SIGS + i] + delta) * jd;
        }
#endif
#pragma unroll
        for (int i = 0; i < NQ; ++i) zq[i] = z[i];
      }
    }
    WARP_SYNC();
    if (act) {
      for (int a = c.gl; a < NVP; a += GSZ) {
        double g = zq[a], yy = 0.0;
#pragma unroll 4
        for (int j = 0; j < NY; ++j) {
          const double yja = T[j * TILE_LD + NY + a];
          g += yja * (rd[EDGE_OFF(e) + j] + zq[NVP + j]);
          yy += yja * bv[j];
        }
        er[ER_GH + a] = g; er[ER_YTY0 + a] = yy;
      }
      for (int i = c.gl; i < NX; i += GSZ) er[ER_B + i] = bv[DMB_XF_SEL[i]] + rp[LAM_EDGE(e) + NY + i];
      for (int j = c.gl; j < NY; j += GSZ) { eb[EB_IMG + j * TILE_LD + NY + NVP] = bv[j]; eb[EB_GY2 + j] = rd[EDGE_OFF(e) + j]; }
    }
    WARP_SYNC();
  }
  CTA_SYNC();
}

// ---------------------------------------------------------------------------------------------------
// vector part of the Riccati sweep for a new right-hand side, re-using Q_xu, Q_uu, P (in place) and K:
// q0 = rd[node vars] + sum_e (GH_e + delta Y'y0_e) ; q += [A B]'(p_c + P_c b) ; k = -Q_uu^{-1} q_u ; p = q_x + Q_xu k.
// ---------------------------------------------------------------------------------------------------
template <bool LDL>
DMB_PHASE void riccati_vector_t(const Ctx& c, const Inst& I, double delta, double* Qs) {
  const double* rd = I.gl;
  for (int idx = c.tid; idx < NNODE * NV; idx += c.nthr) {
    const int n = idx / NV, a = idx - n * NV;
    double* wn = I.node + (size_t)n * WN_NODE;
    const int k = NODE_K_OF(n);
    const int ia = a < NX ? a : (a >= NXI ? a - NU : -1);
    double val = 0.0;
    if (ia >= 0 && !(k == NHOR && ia >= NX)) {
      val = rd[NODE_OFF(n) + ia];
      if (k < NHOR)
        for (int j = 0; j < NODE_NEDGE(n); ++j) {
          const double* er = I.er + (size_t)(NODE_FIRST_EDGE(n) + j) * ER_SZ;
          val += er[ER_GH + ia] + delta * er[ER_YTY0 + ia];
        }
    }
    wn[WN_QV + a] = val;
  }
  CTA_SYNC();
  const bool par = DMB_TREE_PAR && c.nwarp > 1;
  {
    double* pb = Qs + (par ? c.warp : 0) * DMB_QS_WARP;
    for (int lvl = NHOR - 1; lvl >= 0; --lvl) {
     if (par || c.warp == 0)
     for (int n = DMB_LVL(lvl + 1) - 1 - (par ? c.warp : 0); n >= DMB_LVL(lvl); n -= (par ? c.nwarp : 1)) {
      double* wn = I.node + (size_t)n * WN_NODE;
      for (int j = 0; j < NODE_NEDGE(n); ++j) {
        const int e = NODE_FIRST_EDGE(n) + j;
        const double* er = I.er + (size_t)e * ER_SZ;
        const double* wc = I.node + (size_t)EDGE_CHILD_OF(e) * WN_NODE;
        for (int r = c.lane; r < NXI; r += DMB_LANES) {
          double s = wc[WN_QV + r];
#pragma unroll
          for (int t = 0; t < NX; ++t) s += P_AT(wc, r, t) * er[ER_B + t];
          pb[r] = s;
        }
        WARP_SYNC();
        for (int i = c.lane; i < NVP; i += DMB_LANES) {
          double s = (i >= NX && i < NX + NU) ? pb[i] : 0.0;
#pragma unroll
          for (int r = 0; r < NX; ++r) s += er[ER_AB + r * NVP + i] * pb[r];
          wn[WN_QV + VPOS(i)] += s;
        }
        WARP_SYNC();
      }
      // k = -Q_uu^{-1} q_u by every lane redundantly (tiny), then p
      double L[NUT * NUT], kv[NUT];
      if (LDL) quu_ldl(wn + WN_Q + NXI * NV + NXI, L);
      else {
#pragma unroll
      for (int a = 0; a < NUT; ++a)
#pragma unroll
        for (int b = 0; b <= a; ++b) {
          double s = wn[WN_Q + (NXI + a) * NV + NXI + b];
#pragma unroll
          for (int t = 0; t < b; ++t) s -= L[a * NUT + t] * L[b * NUT + t];
          if (a == b) L[a * NUT + a] = rsqrt(s > 0.0 ? s : 1.0);
          else L[a * NUT + b] = s * L[b * NUT + b];
        }
      }
#pragma unroll
      for (int a = 0; a < NUT; ++a) kv[a] = -wn[WN_QV + NXI + a];
      if (LDL) quu_ldl_solve(L, kv);
      else {
#pragma unroll
      for (int a = 0; a < NUT; ++a) { double s = kv[a]; for (int t = 0; t < a; ++t) s -= L[a * NUT + t] * kv[t]; kv[a] = s * L[a * NUT + a]; }
#pragma unroll
      for (int a = NUT - 1; a >= 0; --a) { double s = kv[a]; for (int t = a + 1; t < NUT; ++t) s -= L[t * NUT + a] * kv[t]; kv[a] = s * L[a * NUT + a]; }
      }
      WARP_SYNC();
      for (int a = c.lane; a < NUT; a += DMB_LANES) wn[WN_KV + a] = kv[a];
      for (int a = c.lane; a < NXI; a += DMB_LANES) {
        double s = wn[WN_QV + a];
#pragma unroll
        for (int t = 0; t < NUT; ++t) s += wn[WN_Q + a * NV + NXI + t] * kv[t];
        wn[WN_QV + a] = s;
      }
      WARP_SYNC();
     }
     if (par) CTA_SYNC();
    }
  }
  CTA_SYNC();
}
DMB_PHASE void riccati_vector(const Ctx& c, const Inst& I, double delta, double* Qs) { riccati_vector_t<false>(c, I, delta, Qs); }
