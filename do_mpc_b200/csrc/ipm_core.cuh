// Batched primal-dual interior-point solver for stage-structured collocation NLPs (sm_100a).
//
// Replaces the CasADi nlpsol('ipopt') call of the reference (do_mpc/controller/_mpc.py:1328,
// do_mpc/optimizer.py:770) for a BATCH of problem instances.  One CTA solves one instance at a time
// (persistent CTAs pull instances from an atomic queue); the filter line-search IPM follows
// Waechter & Biegler (2006) with IPOPT's default options, the same algorithm the CPU oracle restates.
// The KKT system of every iteration is solved by a structure-exploiting elimination instead of a
// sparse LDL^T (MUMPS in the reference):
//   per edge  : LU of the square collocation block G_y, sensitivities Y = -G_y^{-1} G_v (null-space
//               projection of the edge Hessian onto the node variables),
//   per node  : Riccati / dynamic-programming sweep over the scenario tree in the augmented state
//               xi = (x, u_prev); positive definiteness of every control block == correct inertia,
//   recovery  : forward roll-out, then multipliers by a transposed LU solve per edge.
// All problem dimensions are compile-time constants from the generated header (problem_gen.h), the
// model functions (residuals, Jacobian / Lagrangian-Hessian nonzeros) are generated straight-line code.
//
// The file compiles for the device (nvcc) and, with DMB_HOSTSIM, as plain single-thread C++ so that
// the numerics can be checked on a machine without a GPU (tests/hostsim; never part of the product).
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef DMB_HOSTSIM
#define DMB_DEV
#define DMB_LANES 1
struct Ctx { int tid, nthr, lane, warp, nwarp; };
#define CTA_SYNC() ((void)0)
#define WARP_SYNC() ((void)0)
#else
#define DMB_DEV __device__ __forceinline__
#define DMB_LANES 32
struct Ctx { int tid, nthr, lane, warp, nwarp; };
#define CTA_SYNC() __syncthreads()
#define WARP_SYNC() __syncwarp()
#endif

// ---------------------------------------------------------------------------------------------------
// options (IPOPT defaults; oracle/ipm.py IpmOptions mirrors this struct)
// ---------------------------------------------------------------------------------------------------
struct DmbOptions {
  double tol, dual_inf_tol, constr_viol_tol, compl_inf_tol;
  double acceptable_tol, acceptable_dual_inf_tol, acceptable_constr_viol_tol, acceptable_compl_inf_tol;
  double s_max, mu_init, mu_min, kappa_mu, theta_mu, kappa_eps, tau_min;
  double bound_push, bound_frac, bound_relax_factor, constr_mult_init_max, bound_mult_init_val;
  double kappa_sigma, kappa_d;
  double alpha_red_factor, eta_phi, theta_max_fact, theta_min_fact, s_phi, s_theta, gamma_phi, gamma_theta;
  double alpha_min_frac, delta, kappa_soc;
  double delta_w_first, delta_w_min, delta_w_max, kappa_w_plus_first, kappa_w_plus, kappa_w_minus;
  double nlp_scaling_max_gradient, nlp_scaling_min_value, kappa_resto;
  int max_iter, acceptable_iter, max_soc, mimic_dummy_reg, resto_max_iter, warm_start;
};

enum DmbStatus {
  DMB_SOLVE_SUCCEEDED = 0, DMB_SOLVED_TO_ACCEPTABLE_LEVEL = 1, DMB_INFEASIBLE_PROBLEM_DETECTED = 2,
  DMB_MAXIMUM_ITERATIONS_EXCEEDED = -1, DMB_RESTORATION_FAILED = -2, DMB_ERROR_IN_STEP_COMPUTATION = -3,
  DMB_INVALID_NUMBER_DETECTED = -13, DMB_NOT_RUN = -100
};

#include "problem_gen.h"   // generated: dimensions, tables, model functions

// ---------------------------------------------------------------------------------------------------
// derived dimensions and internal layouts
// ---------------------------------------------------------------------------------------------------
#define NX DMB_NX
#define NU DMB_NU
#define NEPS DMB_NEPS
#define NY DMB_NY
#define NNL DMB_NNL
#define NVP (NX + NU + NEPS)          // edge-local node variables (x, u, eps)
#define NQ (NVP + NY)                 // edge-local variable vector q = (x, u, eps, y)
#define NXI (NX + NU)                 // augmented Riccati state (x, u_prev)
#define NUT (NU + NEPS)               // stage controls (u, eps)
#define NV (NXI + NUT)                // (x, u_prev, u, eps)
#define NNODE DMB_NNODE
#define NEDGE DMB_NEDGE
#define NHOR DMB_NHOR
#define NWI (NNODE * NVP + NEDGE * NY)               // internal primal vector
#define NODE_OFF(n) ((n) * NVP)
#define EDGE_OFF(e) (NNODE * NVP + (e) * NY)
#define NLE (NY + NX + NNL)                           // multipliers per edge: lamG | lamC | y_d
#define NLAM (NX + NEDGE * NLE)
#define LAM_EDGE(e) (NX + (e) * NLE)
#define TILE_NC (NY + NVP + 1)                        // [G_y | G_vp | rhs] take part in the elimination
#define TILE_M (TILE_NC)                              // then NVP scratch columns for M = H_yy Y
#define TILE_LD ((TILE_NC + NVP) | 1)                 // odd leading dimension: conflict-free 64-bit rows
#define HQ_LD ((NQ) | 1)
#define DMB_INF 1e19
#define MAX_FILTER 64

// per-edge persistent workspace (doubles)
#define WS_LU 0
#define WS_Y (WS_LU + NY * NY)                          // Y (NY x NVP), row-major
#define WS_Y0 (WS_Y + NY * NVP)                         // y0 (NY)
#define WS_HNZ (WS_Y0 + NY)
#define WS_DNZ (WS_HNZ + DMB_NNZ_H)
#define WS_HHAT (WS_DNZ + DMB_NNZ_D)                    // T'HT without regularisation (NVP x NVP)
#define WS_YTY (WS_HHAT + NVP * NVP)                    // Y'Y
#define WS_GHAT (WS_YTY + NVP * NVP)                    // reduced gradient (NVP)
#define WS_YTY0 (WS_GHAT + NVP)                         // Y'y0
#define WS_B (WS_YTY0 + NVP)                            // b (NX)
#define WS_GQ (WS_B + NX)                               // barrier-objective gradient wrt q (NQ)
#define WS_SIGY (WS_GQ + NQ)                            // Sigma on y (NY)
#define WS_SIGS (WS_SIGY + NY)                          // Sigma on inequality slacks (NNL)
#define WS_RD (WS_SIGS + NNL)                           // condensed inequality rhs (NNL)
#define WS_TMP (WS_RD + NNL)                            // scratch (NQ)
#define WS_PIV (WS_TMP + NQ)                            // pivots stored as doubles (NY)
#define WS_EDGE (WS_PIV + NY)
// per-node persistent workspace
#define WN_P 0                                          // P (NXI x NXI)
#define WN_PV (WN_P + NXI * NXI)                        // p (NXI)
#define WN_K (WN_PV + NXI)                              // K (NUT x NXI)
#define WN_KV (WN_K + NUT * NXI)                        // k (NUT)
#define WN_DXI (WN_KV + NUT)                            // d xi (NXI)
#define WN_NODE (WN_DXI + NXI)
// vectors
#define WV_W 0
#define WV_WT (WV_W + NWI)
#define WV_DW (WV_WT + NWI)
#define WV_ZL (WV_DW + NWI)
#define WV_ZU (WV_ZL + NWI)
#define WV_LAM (WV_ZU + NWI)
#define WV_LAMN (WV_LAM + NLAM)
#define WV_RES (WV_LAMN + NLAM)                          // residuals, same layout as lam
#define WV_REST (WV_RES + NLAM)
#define WV_DC (WV_REST + NLAM)                           // row scaling, same layout as lam
#define WV_S (WV_DC + NLAM)                              // inequality slacks (NEDGE*NNL)
#define WV_ST (WV_S + NEDGE * NNL)
#define WV_DS (WV_ST + NEDGE * NNL)
#define WV_VU (WV_DS + NEDGE * NNL)
#define WV_GL (WV_VU + NEDGE * NNL)                      // gradient scratch (NWI)
#define WV_END (WV_GL + NWI)
#define WS_TOTAL (WV_END + NEDGE * WS_EDGE + NNODE * WN_NODE)

struct Inst {
  double *w, *wt, *dw, *zL, *zU, *lam, *lamn, *res, *rest, *dc, *s, *st, *ds, *vU, *gl;
  double *edge, *node;
  const double *p, *lb, *ub;
  double df, relax_fac, relax_max;
};
DMB_DEV const double* p_tvp_at(const double* p, int k) { return p + DMB_P_TVP + k * DMB_NTVP; }
#if DMB_NNL > 0
// relaxed upper bound of the (scaled) inequality row i of edge e (IPOPT bound_relax_factor)
#define DMB_NL_UB_RELAXED(I, e, i) nl_ub_relaxed(I, e, i)
#endif

// ---------------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------------
#ifdef DMB_HOSTSIM
static inline double block_sum(const Ctx&, double v, double*) { return v; }
static inline double block_max(const Ctx&, double v, double*) { return v; }
static inline double block_min(const Ctx&, double v, double*) { return v; }
#else
#define DMB_FULL 0xffffffffu
DMB_DEV double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(DMB_FULL, v, o);
  return v;
}
DMB_DEV double warp_max(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(DMB_FULL, v, o));
  return v;
}
DMB_DEV double warp_min(double v) {
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(DMB_FULL, v, o));
  return v;
}
// scratch: >= 32 doubles of shared memory; result broadcast to all threads
DMB_DEV double block_sum(const Ctx& c, double v, double* scratch) {
  v = warp_sum(v);
  CTA_SYNC();
  if (c.lane == 0) scratch[c.warp] = v;
  CTA_SYNC();
  double r = 0.0;
  for (int i = 0; i < c.nwarp; ++i) r += scratch[i];
  return r;
}
DMB_DEV double block_max(const Ctx& c, double v, double* scratch) {
  v = warp_max(v);
  CTA_SYNC();
  if (c.lane == 0) scratch[c.warp] = v;
  CTA_SYNC();
  double r = scratch[0];
  for (int i = 1; i < c.nwarp; ++i) r = fmax(r, scratch[i]);
  return r;
}
DMB_DEV double block_min(const Ctx& c, double v, double* scratch) {
  v = warp_min(v);
  CTA_SYNC();
  if (c.lane == 0) scratch[c.warp] = v;
  CTA_SYNC();
  double r = scratch[0];
  for (int i = 1; i < c.nwarp; ++i) r = fmin(r, scratch[i]);
  return r;
}
#endif

// ---------------------------------------------------------------------------------------------------
// gather of the edge-local vector q = (x_parent, u_parent, eps_parent, y_e) and its parameters
// ---------------------------------------------------------------------------------------------------
DMB_DEV void gather_q(const double* w, int e, double* q) {
  const double* wn = w + NODE_OFF(DMB_EDGE_PARENT[e]);
#pragma unroll
  for (int i = 0; i < NVP; ++i) q[i] = wn[i];
  const double* wy = w + EDGE_OFF(e);
#pragma unroll
  for (int i = 0; i < NY; ++i) q[NVP + i] = wy[i];
}
DMB_DEV const double* edge_tvp(const double* p, int e) { return p + DMB_P_TVP + DMB_EDGE_K[e] * DMB_NTVP; }
DMB_DEV const double* edge_pp(const double* p, int e) { return p + DMB_P_P + DMB_EDGE_PIDX[e] * DMB_NPM; }

// u_prev (scaled) seen by node n: the parent's input, or the parameter for the root (reference _mpc.py:1273)
DMB_DEV double node_uprev(const Inst& I, const double* w, int n, int i) {
  const int pe = DMB_NODE_PARENT_EDGE[n];
  if (pe < 0) return I.p[DMB_P_UPREV + i] * DMB_INV_USCALE[i];
  return w[NODE_OFF(DMB_EDGE_PARENT[pe]) + NX + i];
}

#if DMB_NNL > 0
DMB_DEV double nl_ub_relaxed(const Inst& I, int e, int i) {
  const double u = I.dc[LAM_EDGE(e) + NY + NX + i] * DMB_NL_UB[i];
  return u + fmin(I.relax_fac * fmax(1.0, fabs(u)), I.relax_max);
}
#endif

// ---------------------------------------------------------------------------------------------------
// residuals + objective at a point (w, s) -> res (scaled), returns objective (scaled by df) through *f_out
// theta = 1-norm of the residual vector.  All threads of the CTA participate.
// ---------------------------------------------------------------------------------------------------
DMB_DEV void eval_point(const Ctx& c, const Inst& I, const double* w, const double* s, double* res,
                        double* scratch, double* f_out, double* theta_out, double* infpr_out) {
  double f = 0.0;
  for (int e = c.tid; e < NEDGE; e += c.nthr) {
    double q[NQ], g[NY > 0 ? NY : 1];
    gather_q(w, e, q);
    const double* tv = edge_tvp(I.p, e);
    const double* pp = edge_pp(I.p, e);
    gen_G(q, tv, pp, g);
    double* r = res + LAM_EDGE(e);
    const double* dc = I.dc + LAM_EDGE(e);
#pragma unroll
    for (int i = 0; i < NY; ++i) r[i] = dc[i] * g[i];
    const double* xc = w + NODE_OFF(DMB_EDGE_CHILD[e]);
#pragma unroll
    for (int i = 0; i < NX; ++i) r[NY + i] = q[NVP + DMB_XF_SEL[i]] - xc[i];
#if NNL > 0
    double d[NNL];
    gen_nl(q, tv, pp, d);
#pragma unroll
    for (int i = 0; i < NNL; ++i) r[NY + NX + i] = dc[NY + NX + i] * d[i] - s[e * NNL + i];
#endif
    double lt, ec;
    gen_cost(q, tv, pp, &lt, &ec);
    f += DMB_EDGE_OMEGA[e] * lt + ec;
    if (DMB_EDGE_K[e] == NHOR - 1) f += DMB_EDGE_OMEGA[e] * gen_mterm(xc, p_tvp_at(I.p, NHOR), pp);
  }
  for (int n = c.tid; n < NNODE; n += c.nthr) {
    if (DMB_NODE_K[n] == NHOR) continue;
    double up[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) up[i] = node_uprev(I, w, n, i);
    f += DMB_NODE_RW[n] * gen_rterm(w + NODE_OFF(n) + NX, up, I.p + DMB_P_TVP + DMB_NODE_K[n] * DMB_NTVP);
  }
  for (int i = c.tid; i < NX; i += c.nthr) res[i] = w[NODE_OFF(0) + i] - I.p[DMB_P_X0 + i] * DMB_INV_XSCALE[i];
  CTA_SYNC();
  double th = 0.0, mx = 0.0;
  for (int i = c.tid; i < NLAM; i += c.nthr) { const double a = fabs(res[i]); th += a; mx = fmax(mx, a); }
  *f_out = I.df * block_sum(c, f, scratch);
  *theta_out = block_sum(c, th, scratch);
  if (infpr_out) *infpr_out = block_max(c, mx, scratch);
}

// ---------------------------------------------------------------------------------------------------
// gradient of  sig*f + lam'c  w.r.t. all primal variables into gl (NWI).  with_lam=0 -> objective only.
// lam are the SCALED multipliers; row scales dc are applied here.  Uses per-edge scratch WS_TMP.
// ---------------------------------------------------------------------------------------------------
DMB_DEV void eval_gradient(const Ctx& c, const Inst& I, const double* w, const double* lam, int with_lam, double* gl) {
  for (int e = c.tid; e < NEDGE; e += c.nthr) {
    double q[NQ], lg[NY > 0 ? NY : 1], ln[NNL > 0 ? NNL : 1], gq[NQ];
    gather_q(w, e, q);
    const double* dc = I.dc + LAM_EDGE(e);
    const double* le = lam + LAM_EDGE(e);
#pragma unroll
    for (int i = 0; i < NY; ++i) lg[i] = with_lam ? dc[i] * le[i] : 0.0;
#pragma unroll
    for (int i = 0; i < NNL; ++i) ln[i] = with_lam ? dc[NY + NX + i] * le[NY + NX + i] : 0.0;
    gen_lag_grad(q, edge_tvp(I.p, e), edge_pp(I.p, e), lg, I.df * DMB_EDGE_OMEGA[e], I.df, ln, gq);
    if (with_lam) {
#pragma unroll
      for (int i = 0; i < NX; ++i) gq[NVP + DMB_XF_SEL[i]] += le[NY + i];
    }
    double* t = I.edge + e * WS_EDGE + WS_TMP;
#pragma unroll
    for (int i = 0; i < NQ; ++i) t[i] = gq[i];
  }
  CTA_SYNC();
  // node variables: sum over child edges (+ continuity of the parent edge, terminal cost, rterm)
  for (int idx = c.tid; idx < NNODE * NVP; idx += c.nthr) {
    const int n = idx / NVP, i = idx - n * NVP;
    const int k = DMB_NODE_K[n];
    double g = 0.0;
    if (k < NHOR) {
      for (int j = 0; j < DMB_NODE_NEDGE[n]; ++j) g += I.edge[(DMB_NODE_FIRST_EDGE[n] + j) * WS_EDGE + WS_TMP + i];
    } else if (i >= NX) { gl[idx] = 0.0; continue; }
    if (i < NX) {
      const int pe = DMB_NODE_PARENT_EDGE[n];
      if (pe >= 0) {
        if (with_lam) g -= lam[LAM_EDGE(pe) + NY + i];
        if (k == NHOR) {
          double gm[NX];
          gen_mterm_grad(w + NODE_OFF(n), p_tvp_at(I.p, NHOR), edge_pp(I.p, pe), gm);
          g += I.df * DMB_EDGE_OMEGA[pe] * gm[i];
        }
      } else if (with_lam) g += lam[i];
    } else if (i < NX + NU) {
      const int iu = i - NX;
      double up[NU], gr[2 * NU];
#pragma unroll
      for (int a = 0; a < NU; ++a) up[a] = node_uprev(I, w, n, a);
      gen_rterm_grad(w + NODE_OFF(n) + NX, up, I.p + DMB_P_TVP + k * DMB_NTVP, gr);
      g += I.df * DMB_NODE_RW[n] * gr[iu];
      // this node's input is u_prev of every child node that still has a control
      for (int j = 0; j < DMB_NODE_NEDGE[n]; ++j) {
        const int cn = DMB_EDGE_CHILD[DMB_NODE_FIRST_EDGE[n] + j];
        if (DMB_NODE_K[cn] == NHOR) continue;
        gen_rterm_grad(w + NODE_OFF(cn) + NX, w + NODE_OFF(n) + NX, I.p + DMB_P_TVP + (k + 1) * DMB_NTVP, gr);
        g += I.df * DMB_NODE_RW[cn] * gr[NU + iu];
      }
    }
    gl[idx] = g;
  }
  for (int idx = c.tid; idx < NEDGE * NY; idx += c.nthr) {
    const int e = idx / NY, i = idx - e * NY;
    gl[NNODE * NVP + idx] = I.edge[e * WS_EDGE + WS_TMP + NVP + i];
  }
  CTA_SYNC();
}

// ---------------------------------------------------------------------------------------------------
// H_e * v for the edge Hessian (Lagrangian nonzeros + inequality condensation + Sigma_y + delta on y)
// serial, one thread.  hv and v have NQ entries.
// ---------------------------------------------------------------------------------------------------
DMB_DEV void edge_hmul(const double* we, double delta, const double* v, double* hv) {
#pragma unroll
  for (int i = 0; i < NQ; ++i) hv[i] = 0.0;
  const double* hnz = we + WS_HNZ;
  for (int k = 0; k < DMB_NNZ_H; ++k) {
    const int r = DMB_H_ROW[k], cc = DMB_H_COL[k];
    const double a = hnz[k];
    hv[r] += a * v[cc];
    if (r != cc) hv[cc] += a * v[r];
  }
#if NNL > 0
  double jd[NNL];
  for (int i = 0; i < NNL; ++i) jd[i] = 0.0;
  for (int k = 0; k < DMB_NNZ_D; ++k) jd[DMB_D_ROW[k]] += we[WS_DNZ + k] * v[DMB_D_COL[k]];
  for (int i = 0; i < NNL; ++i) jd[i] *= we[WS_SIGS + i];
  for (int k = 0; k < DMB_NNZ_D; ++k) hv[DMB_D_COL[k]] += we[WS_DNZ + k] * jd[DMB_D_ROW[k]];
#endif
  for (int i = 0; i < NY; ++i) hv[NVP + i] += (we[WS_SIGY + i] + delta) * v[NVP + i];
}

// solve G_y x = b with the stored LU (row swaps as recorded), in place.  One thread.
DMB_DEV void lu_solve(const double* we, double* b) {
  const double* LU = we + WS_LU;
  for (int k = 0; k < NY; ++k) {
    const int pr = (int)we[WS_PIV + k];
    if (pr != k) { const double t = b[k]; b[k] = b[pr]; b[pr] = t; }
  }
  for (int k = 0; k < NY; ++k) {
    const double bk = b[k];
    for (int r = k + 1; r < NY; ++r) b[r] -= LU[r * NY + k] * bk;
  }
  for (int k = NY - 1; k >= 0; --k) {
    double a = b[k];
    for (int cc = k + 1; cc < NY; ++cc) a -= LU[k * NY + cc] * b[cc];
    b[k] = a / LU[k * NY + k];
  }
}
// solve G_y' x = b, in place
DMB_DEV void lu_solve_t(const double* we, double* b) {
  const double* LU = we + WS_LU;
  for (int k = 0; k < NY; ++k) {            // U' a = b
    double a = b[k];
    for (int r = 0; r < k; ++r) a -= LU[r * NY + k] * b[r];
    b[k] = a / LU[k * NY + k];
  }
  for (int k = NY - 1; k >= 0; --k) {       // L' c = a (unit diagonal)
    double a = b[k];
    for (int r = k + 1; r < NY; ++r) a -= LU[r * NY + k] * b[r];
    b[k] = a;
  }
  for (int k = NY - 1; k >= 0; --k) {
    const int pr = (int)we[WS_PIV + k];
    if (pr != k) { const double t = b[k]; b[k] = b[pr]; b[pr] = t; }
  }
}

// ---------------------------------------------------------------------------------------------------
// factor phase: per edge Jacobian tile -> LU, Y ; Hessian tile -> projected blocks.  One warp per edge.
// tile: NY x TILE_LD, hq: NQ x HQ_LD (warp-private shared memory).
// returns through *fail a nonzero value when a pivot is (numerically) zero.
// ---------------------------------------------------------------------------------------------------
DMB_DEV void factor_edges(const Ctx& c, const Inst& I, double mu, double kappa_d, int ls_mode, double* tile_base, double* hq_base, int* fail) {
  double* T = tile_base + (size_t)c.warp * NY * TILE_LD;
  double* HQ = hq_base + (size_t)c.warp * NQ * HQ_LD;
  for (int e = c.warp; e < NEDGE; e += c.nwarp) {
    double* we = I.edge + (size_t)e * WS_EDGE;
    const double* tv = edge_tvp(I.p, e);
    const double* pp = edge_pp(I.p, e);
    const double* dc = I.dc + LAM_EDGE(e);
    const double* le = I.lam + LAM_EDGE(e);
    // -- zero tiles
    for (int i = c.lane; i < NY * TILE_LD; i += DMB_LANES) T[i] = 0.0;
    for (int i = c.lane; i < NQ * HQ_LD; i += DMB_LANES) HQ[i] = 0.0;
    WARP_SYNC();
    // -- model derivatives: straight-line generated code, one lane
    if (c.lane == 0) {
      double q[NQ];
      gather_q(I.w, e, q);
      {
        double jnz[DMB_NNZ_J > 0 ? DMB_NNZ_J : 1];
        gen_J(q, tv, pp, jnz);
        for (int k = 0; k < DMB_NNZ_J; ++k) {
          const int r = DMB_J_ROW[k], cq = DMB_J_COL[k];
          const int col = cq >= NVP ? cq - NVP : NY + cq;      // tile columns: [G_y | G_vp | rhs]
          T[r * TILE_LD + col] = dc[r] * jnz[k];
        }
      }
      {
        double lg[NY > 0 ? NY : 1], ln[NNL > 0 ? NNL : 1];
        for (int i = 0; i < NY; ++i) lg[i] = 0.0;
        for (int i = 0; i < NNL; ++i) ln[i] = 0.0;
        // edge-local objective gradient (node-level terms are added in the Riccati sweep)
        gen_lag_grad(q, tv, pp, lg, I.df * DMB_EDGE_OMEGA[e], I.df, ln, we + WS_GQ);
        if (ls_mode) {
          for (int k = 0; k < DMB_NNZ_H; ++k) we[WS_HNZ + k] = 0.0;
        } else {
          for (int i = 0; i < NY; ++i) lg[i] = dc[i] * le[i];
          for (int i = 0; i < NNL; ++i) ln[i] = dc[NY + NX + i] * le[NY + NX + i];
          gen_lag_hess(q, tv, pp, lg, I.df * DMB_EDGE_OMEGA[e], I.df, ln, we + WS_HNZ);
          for (int k = 0; k < DMB_NNZ_H; ++k) {
            const int r = DMB_H_ROW[k], cc = DMB_H_COL[k];
            const double a = we[WS_HNZ + k];
            HQ[r * HQ_LD + cc] = a; HQ[cc * HQ_LD + r] = a;
          }
        }
      }
      // barrier terms of the edge unknowns y
      const double* wy = I.w + EDGE_OFF(e);
      const double* lb = I.lb + EDGE_OFF(e);
      const double* ub = I.ub + EDGE_OFF(e);
      const double* zl = I.zL + EDGE_OFF(e);
      const double* zu = I.zU + EDGE_OFF(e);
      for (int i = 0; i < NY; ++i) {
        double sig = 0.0, gb = 0.0;
        const bool hl = lb[i] > -DMB_INF, hu = ub[i] < DMB_INF;
        if (ls_mode) {
          sig = 1.0;
          if (hl) gb -= zl[i];
          if (hu) gb += zu[i];
        } else {
          if (hl) { const double sl = wy[i] - lb[i]; sig += zl[i] / sl; gb -= mu / sl; if (!hu) gb += kappa_d * mu; }
          if (hu) { const double su = ub[i] - wy[i]; sig += zu[i] / su; gb += mu / su; if (!hl) gb -= kappa_d * mu; }
        }
        we[WS_SIGY + i] = sig;
        we[WS_GQ + NVP + i] += gb;
        HQ[(NVP + i) * HQ_LD + NVP + i] += sig;
      }
#if NNL > 0
      {
        double dnz[DMB_NNZ_D > 0 ? DMB_NNZ_D : 1];
        gen_nl_jac(q, tv, pp, dnz);
        for (int k = 0; k < DMB_NNZ_D; ++k) we[WS_DNZ + k] = dc[NY + NX + DMB_D_ROW[k]] * dnz[k];
        for (int i = 0; i < NNL; ++i) {
          const double ubs = DMB_NL_UB_RELAXED(I, e, i);
          const double sl = ubs - I.s[e * NNL + i];
          const double sig = ls_mode ? 1.0 : I.vU[e * NNL + i] / sl;
          we[WS_SIGS + i] = sig;
          // r_d = Sigma_s (d - s) + grad_s phi ; grad_s phi = mu/sl - kappa_d*mu (upper bound only)
          we[WS_RD + i] = ls_mode ? I.vU[e * NNL + i] : sig * I.res[LAM_EDGE(e) + NY + NX + i] + mu / sl - kappa_d * mu;
        }
        for (int k = 0; k < DMB_NNZ_D; ++k) {
          const int r = DMB_D_ROW[k], cq = DMB_D_COL[k];
          we[WS_GQ + cq] += we[WS_DNZ + k] * we[WS_RD + r];
          for (int k2 = 0; k2 < DMB_NNZ_D; ++k2) {
            if (DMB_D_ROW[k2] != r) continue;
            HQ[cq * HQ_LD + DMB_D_COL[k2]] += we[WS_DNZ + k] * we[WS_SIGS + r] * we[WS_DNZ + k2];
          }
        }
      }
#endif
    }
    WARP_SYNC();
    for (int r = c.lane; r < NY; r += DMB_LANES) T[r * TILE_LD + NY + NVP] = ls_mode ? 0.0 : -I.res[LAM_EDGE(e) + r];
    WARP_SYNC();
    // -- LU with partial pivoting on columns [0, NY); the other columns ride along
    for (int k = 0; k < NY; ++k) {
      // pivot search (small: every lane scans; identical result on all lanes)
      int pr = k; double best = fabs(T[k * TILE_LD + k]);
      for (int r = k + 1; r < NY; ++r) { const double a = fabs(T[r * TILE_LD + k]); if (a > best) { best = a; pr = r; } }
      if (!(best > 1e-300)) { *fail = 1; }
      if (pr != k) {
        for (int cc = c.lane; cc < TILE_NC; cc += DMB_LANES) {
          const double t = T[k * TILE_LD + cc]; T[k * TILE_LD + cc] = T[pr * TILE_LD + cc]; T[pr * TILE_LD + cc] = t;
        }
      }
      if (c.lane == 0) we[WS_PIV + k] = (double)pr;
      WARP_SYNC();
      const double piv = 1.0 / T[k * TILE_LD + k];
      WARP_SYNC();
      for (int r = k + 1 + c.lane; r < NY; r += DMB_LANES) T[r * TILE_LD + k] *= piv;
      WARP_SYNC();
      const int nr = NY - k - 1, ncol = TILE_NC - k - 1;
      for (int idx = c.lane; idx < nr * ncol; idx += DMB_LANES) {
        const int r = k + 1 + idx / ncol, cc = k + 1 + idx % ncol;
        T[r * TILE_LD + cc] -= T[r * TILE_LD + k] * T[k * TILE_LD + cc];
      }
      WARP_SYNC();
    }
    // store LU
    for (int idx = c.lane; idx < NY * NY; idx += DMB_LANES) we[WS_LU + idx] = T[(idx / NY) * TILE_LD + idx % NY];
    // -- back substitution on the right-hand-side columns: U X = (L^-1 P rhs)
    for (int k = NY - 1; k >= 0; --k) {
      const double piv = 1.0 / T[k * TILE_LD + k];
      WARP_SYNC();
      for (int cc = NY + c.lane; cc < TILE_NC; cc += DMB_LANES) T[k * TILE_LD + cc] *= piv;
      WARP_SYNC();
      const int ncol = TILE_NC - NY;
      for (int idx = c.lane; idx < k * ncol; idx += DMB_LANES) {
        const int r = idx / ncol, cc = NY + idx % ncol;
        T[r * TILE_LD + cc] -= T[r * TILE_LD + k] * T[k * TILE_LD + cc];
      }
      WARP_SYNC();
    }
    // Y = -X[:, :NVP] ; y0 = X[:, NVP]  (rhs column already holds -res)
    for (int idx = c.lane; idx < NY * NVP; idx += DMB_LANES) {
      const int r = idx / NVP, cc = idx - r * NVP;
      const double y = -T[r * TILE_LD + NY + cc];
      T[r * TILE_LD + NY + cc] = y;
      we[WS_Y + idx] = y;
    }
    for (int r = c.lane; r < NY; r += DMB_LANES) we[WS_Y0 + r] = T[r * TILE_LD + NY + NVP];
    WARP_SYNC();
    // -- projection: M = H_yy Y (into T columns [0,NVP) which are free now) ; Hhat = H_vv + H_vy Y + (H_vy Y)' + Y' M
    for (int idx = c.lane; idx < NY * NVP; idx += DMB_LANES) {
      const int r = idx / NVP, cc = idx - r * NVP;
      double a = 0.0;
      for (int j = 0; j < NY; ++j) a += HQ[(NVP + r) * HQ_LD + NVP + j] * T[j * TILE_LD + NY + cc];
      T[r * TILE_LD + TILE_M + cc] = a;
    }
    WARP_SYNC();
    for (int idx = c.lane; idx < NVP * NVP; idx += DMB_LANES) {
      const int a = idx / NVP, b = idx - a * NVP;
      double hvv = HQ[a * HQ_LD + b], yty = 0.0;
      for (int j = 0; j < NY; ++j) {
        const double yja = T[j * TILE_LD + NY + a], yjb = T[j * TILE_LD + NY + b];
        hvv += HQ[a * HQ_LD + NVP + j] * yjb + yja * HQ[(NVP + j) * HQ_LD + b] + yja * T[j * TILE_LD + TILE_M + b];
        yty += yja * yjb;
      }
      we[WS_HHAT + idx] = hvv;
      we[WS_YTY + idx] = yty;
    }
    WARP_SYNC();
  }
}

// solve phase per edge (one thread per edge): y0 for the given residuals, reduced gradient, offsets
DMB_DEV void solve_edges(const Ctx& c, const Inst& I, const double* res, double delta, int recompute_y0) {
  for (int e = c.tid; e < NEDGE; e += c.nthr) {
    double* we = I.edge + (size_t)e * WS_EDGE;
    double y0[NY > 0 ? NY : 1], v[NQ], hv[NQ];
    if (recompute_y0) {
      for (int i = 0; i < NY; ++i) y0[i] = -res[LAM_EDGE(e) + i];
      lu_solve(we, y0);
      for (int i = 0; i < NY; ++i) we[WS_Y0 + i] = y0[i];
    } else {
      for (int i = 0; i < NY; ++i) y0[i] = we[WS_Y0 + i];
    }
    for (int i = 0; i < NVP; ++i) v[i] = 0.0;
    for (int i = 0; i < NY; ++i) v[NVP + i] = y0[i];
    edge_hmul(we, delta, v, hv);
    const double* gq = we + WS_GQ;
    for (int a = 0; a < NVP; ++a) {
      double g = gq[a] + hv[a], yy = 0.0;
      for (int j = 0; j < NY; ++j) { g += we[WS_Y + j * NVP + a] * (gq[NVP + j] + hv[NVP + j]); yy += we[WS_Y + j * NVP + a] * y0[j]; }
      we[WS_GHAT + a] = g;
      we[WS_YTY0 + a] = yy;
    }
    for (int i = 0; i < NX; ++i) we[WS_B + i] = y0[DMB_XF_SEL[i]] + res[LAM_EDGE(e) + NY + i];
  }
  CTA_SYNC();
}

// ---------------------------------------------------------------------------------------------------
// Riccati sweep over the tree.  Executed by warp 0 (lanes over matrix entries); Qs: shared scratch
// of size NV*NV + NV + NXI*NV.  Returns 0 on success, 1 if a control block is not positive definite.
// v-space ordering: (x [NX], u_prev [NU], u [NU], eps [NEPS]).
// ---------------------------------------------------------------------------------------------------
DMB_DEV int riccati_backward(const Ctx& c, const Inst& I, double mu, double kappa_d, double delta, int ls_mode, double* Qs, int* flag) {
  double* Q = Qs;                 // NV x NV
  double* qv = Qs + NV * NV;      // NV
  double* AB = qv + NV;           // NXI x NV  ([A B] of the current child edge)
  double* PA = AB + NXI * NV;     // NXI x NV  (P_child [A B])
  if (c.tid == 0) *flag = 0;
  CTA_SYNC();
  if (c.warp == 0) {
    for (int n = NNODE - 1; n >= 0; --n) {
      double* wn = I.node + (size_t)n * WN_NODE;
      const int k = DMB_NODE_K[n];
      const double* wx = I.w + NODE_OFF(n);
      const double* lb = I.lb + NODE_OFF(n);
      const double* ub = I.ub + NODE_OFF(n);
      const double* zl = I.zL + NODE_OFF(n);
      const double* zu = I.zU + NODE_OFF(n);
      for (int i = c.lane; i < NV * NV + NV; i += DMB_LANES) Q[i] = 0.0;
      WARP_SYNC();
      // node-level diagonal terms: Sigma + delta and barrier gradient on (x, u, eps); objective node terms
      for (int i = c.lane; i < NVP; i += DMB_LANES) {
        if (k == NHOR && i >= NX) continue;
        const int vi = i < NX ? i : i + NU;           // position in v-space
        double sig = delta, gb = 0.0;
        const bool hl = lb[i] > -DMB_INF, hu = ub[i] < DMB_INF;
        if (ls_mode) {
          sig = 1.0;
          if (hl) gb -= zl[i];
          if (hu) gb += zu[i];
        } else {
          if (hl) { const double sl = wx[i] - lb[i]; sig += zl[i] / sl; gb -= mu / sl; if (!hu) gb += kappa_d * mu; }
          if (hu) { const double su = ub[i] - wx[i]; sig += zu[i] / su; gb += mu / su; if (!hl) gb -= kappa_d * mu; }
        }
        Q[vi * NV + vi] += sig;
        qv[vi] += gb;
      }
      WARP_SYNC();
      if (k == NHOR) {
        // terminal node: mterm
        if (c.lane == 0) {
          const int pe = DMB_NODE_PARENT_EDGE[n];
          double gm[NX], hm[NX * NX];
          gen_mterm_grad(wx, p_tvp_at(I.p, NHOR), edge_pp(I.p, pe), gm);
          gen_mterm_hess(wx, p_tvp_at(I.p, NHOR), edge_pp(I.p, pe), hm);
          const double wgt = I.df * DMB_EDGE_OMEGA[pe];
          for (int a = 0; a < NX; ++a) {
            qv[a] += wgt * gm[a];
            if (!ls_mode) for (int b = 0; b < NX; ++b) Q[a * NV + b] += wgt * hm[a * NX + b];
          }
        }
        WARP_SYNC();
        for (int idx = c.lane; idx < NXI * NXI; idx += DMB_LANES) {
          const int a = idx / NXI, b = idx - a * NXI;
          wn[WN_P + idx] = Q[a * NV + b];
        }
        for (int a = c.lane; a < NXI; a += DMB_LANES) wn[WN_PV + a] = qv[a];
        WARP_SYNC();
        continue;
      }
      // rterm at this node (weights: reference _mpc.py:1263-1275)
      if (c.lane == 0) {
        double up[NU], gr[2 * NU], hr[4 * NU * NU];
        for (int a = 0; a < NU; ++a) up[a] = node_uprev(I, I.w, n, a);
        const double* tv = I.p + DMB_P_TVP + k * DMB_NTVP;
        gen_rterm_grad(wx + NX, up, tv, gr);
        gen_rterm_hess(wx + NX, up, tv, hr);
        const double wgt = I.df * DMB_NODE_RW[n];
        for (int a = 0; a < NU; ++a) {
          qv[NXI + a] += wgt * gr[a];
          qv[NX + a] += wgt * gr[NU + a];
          if (!ls_mode) for (int b = 0; b < NU; ++b) {
            Q[(NXI + a) * NV + NXI + b] += wgt * hr[a * 2 * NU + b];
            Q[(NXI + a) * NV + NX + b] += wgt * hr[a * 2 * NU + NU + b];
            Q[(NX + a) * NV + NXI + b] += wgt * hr[(NU + a) * 2 * NU + b];
            Q[(NX + a) * NV + NX + b] += wgt * hr[(NU + a) * 2 * NU + NU + b];
          }
        }
      }
      WARP_SYNC();
      for (int j = 0; j < DMB_NODE_NEDGE[n]; ++j) {
        const int e = DMB_NODE_FIRST_EDGE[n] + j;
        const double* we = I.edge + (size_t)e * WS_EDGE;
        const double* wc = I.node + (size_t)DMB_EDGE_CHILD[e] * WN_NODE;
        // [A B]: child xi = (x_c, u_prev_c) = (Y[xf rows] dvp + b, du)
        for (int idx = c.lane; idx < NXI * NV; idx += DMB_LANES) {
          const int r = idx / NV, vi = idx - r * NV;
          double a = 0.0;
          if (r < NX) {
            if (vi < NX) a = we[WS_Y + DMB_XF_SEL[r] * NVP + vi];
            else if (vi >= NXI) a = we[WS_Y + DMB_XF_SEL[r] * NVP + (vi - NU)];
          } else if (vi == NXI + (r - NX)) a = 1.0;
          AB[idx] = a;
        }
        WARP_SYNC();
        for (int idx = c.lane; idx < NXI * NV; idx += DMB_LANES) {
          const int r = idx / NV, vi = idx - r * NV;
          double a = 0.0;
          for (int t = 0; t < NXI; ++t) a += wc[WN_P + r * NXI + t] * AB[t * NV + vi];
          PA[idx] = a;
        }
        WARP_SYNC();
        // Q += embed(Hhat + delta*YtY) + AB' PA ; q += embed(ghat + delta*Yty0) + AB'(P b + p)
        for (int idx = c.lane; idx < NV * NV; idx += DMB_LANES) {
          const int a = idx / NV, b = idx - a * NV;
          double s = 0.0;
          for (int t = 0; t < NXI; ++t) s += AB[t * NV + a] * PA[t * NV + b];
          const int ia = a < NX ? a : (a >= NXI ? a - NU : -1), ib = b < NX ? b : (b >= NXI ? b - NU : -1);
          if (ia >= 0 && ib >= 0) s += we[WS_HHAT + ia * NVP + ib] + delta * we[WS_YTY + ia * NVP + ib];
          Q[idx] += s;
        }
        for (int a = c.lane; a < NV; a += DMB_LANES) {
          double s = 0.0;
          for (int t = 0; t < NXI; ++t) {
            double pb = wc[WN_PV + t];
            for (int r = 0; r < NX; ++r) pb += wc[WN_P + t * NXI + r] * we[WS_B + r];
            s += AB[t * NV + a] * pb;
          }
          const int ia = a < NX ? a : (a >= NXI ? a - NU : -1);
          if (ia >= 0) s += we[WS_GHAT + ia] + delta * we[WS_YTY0 + ia];
          qv[a] += s;
        }
        WARP_SYNC();
      }
      // eliminate the controls: Cholesky of Q_uu (NUT x NUT), serial (tiny)
      if (c.lane == 0) {
        double L[NUT * NUT];
        int bad = 0;
        for (int a = 0; a < NUT; ++a)
          for (int b = 0; b <= a; ++b) {
            double s = Q[(NXI + a) * NV + NXI + b];
            for (int t = 0; t < b; ++t) s -= L[a * NUT + t] * L[b * NUT + t];
            if (a == b) { if (!(s > 0.0)) { bad = 1; s = 1.0; } L[a * NUT + a] = sqrt(s); }
            else L[a * NUT + b] = s / L[b * NUT + b];
          }
        if (bad) *flag = 1;
        // K = -Quu^{-1} Qux ; k = -Quu^{-1} qu
        for (int col = 0; col <= NXI; ++col) {
          double rhs[NUT];
          for (int a = 0; a < NUT; ++a) rhs[a] = col < NXI ? -Q[(NXI + a) * NV + col] : -qv[NXI + a];
          for (int a = 0; a < NUT; ++a) { double s = rhs[a]; for (int t = 0; t < a; ++t) s -= L[a * NUT + t] * rhs[t]; rhs[a] = s / L[a * NUT + a]; }
          for (int a = NUT - 1; a >= 0; --a) { double s = rhs[a]; for (int t = a + 1; t < NUT; ++t) s -= L[t * NUT + a] * rhs[t]; rhs[a] = s / L[a * NUT + a]; }
          for (int a = 0; a < NUT; ++a) { if (col < NXI) wn[WN_K + a * NXI + col] = rhs[a]; else wn[WN_KV + a] = rhs[a]; }
        }
      }
      WARP_SYNC();
      for (int idx = c.lane; idx < NXI * NXI; idx += DMB_LANES) {
        const int a = idx / NXI, b = idx - a * NXI;
        double s = Q[a * NV + b];
        for (int t = 0; t < NUT; ++t) s += Q[a * NV + NXI + t] * wn[WN_K + t * NXI + b];
        wn[WN_P + idx] = s;
      }
      for (int a = c.lane; a < NXI; a += DMB_LANES) {
        double s = qv[a];
        for (int t = 0; t < NUT; ++t) s += Q[a * NV + NXI + t] * wn[WN_KV + t];
        wn[WN_PV + a] = s;
      }
      WARP_SYNC();
    }
  }
  CTA_SYNC();
  return *flag;
}

// forward roll-out: dw (all primal steps), new multipliers lamn, inequality steps ds / new y_d
DMB_DEV void forward_recover(const Ctx& c, const Inst& I, const double* res, double delta) {
  if (c.warp == 0) {
    for (int n = 0; n < NNODE; ++n) {
      double* wn = I.node + (size_t)n * WN_NODE;
      const int k = DMB_NODE_K[n];
      if (n == 0) {
        for (int a = c.lane; a < NXI; a += DMB_LANES) wn[WN_DXI + a] = a < NX ? -res[a] : 0.0;
        WARP_SYNC();
      }
      double* dwn = I.dw + NODE_OFF(n);
      for (int a = c.lane; a < NX; a += DMB_LANES) dwn[a] = wn[WN_DXI + a];
      if (k == NHOR) { for (int a = NX + c.lane; a < NVP; a += DMB_LANES) dwn[a] = 0.0; WARP_SYNC(); continue; }
      for (int a = c.lane; a < NUT; a += DMB_LANES) {
        double s = wn[WN_KV + a];
        for (int t = 0; t < NXI; ++t) s += wn[WN_K + a * NXI + t] * wn[WN_DXI + t];
        dwn[NX + a] = s;
      }
      WARP_SYNC();
      for (int j = 0; j < DMB_NODE_NEDGE[n]; ++j) {
        const int e = DMB_NODE_FIRST_EDGE[n] + j;
        const double* we = I.edge + (size_t)e * WS_EDGE;
        double* wc = I.node + (size_t)DMB_EDGE_CHILD[e] * WN_NODE;
        for (int r = c.lane; r < NXI; r += DMB_LANES) {
          double s;
          if (r < NX) {
            s = we[WS_B + r];
            for (int t = 0; t < NVP; ++t) s += we[WS_Y + DMB_XF_SEL[r] * NVP + t] * dwn[t];
          } else s = dwn[NX + (r - NX)];
          wc[WN_DXI + r] = s;
        }
      }
      WARP_SYNC();
    }
  }
  CTA_SYNC();
  // edge unknowns, multipliers
  for (int e = c.tid; e < NEDGE; e += c.nthr) {
    const double* we = I.edge + (size_t)e * WS_EDGE;
    const double* dvp = I.dw + NODE_OFF(DMB_EDGE_PARENT[e]);
    double v[NQ], hv[NQ], r[NY > 0 ? NY : 1];
    for (int i = 0; i < NVP; ++i) v[i] = dvp[i];
    for (int i = 0; i < NY; ++i) {
      double s = we[WS_Y0 + i];
      for (int t = 0; t < NVP; ++t) s += we[WS_Y + i * NVP + t] * dvp[t];
      v[NVP + i] = s;
      I.dw[EDGE_OFF(e) + i] = s;
    }
    // continuity multiplier = gradient of the child's cost-to-go
    const double* wc = I.node + (size_t)DMB_EDGE_CHILD[e] * WN_NODE;
    double lc[NX];
    for (int a = 0; a < NX; ++a) {
      double s = wc[WN_PV + a];
      for (int t = 0; t < NXI; ++t) s += wc[WN_P + a * NXI + t] * wc[WN_DXI + t];
      lc[a] = s;
      I.lamn[LAM_EDGE(e) + NY + a] = s;
    }
    edge_hmul(we, delta, v, hv);
    for (int i = 0; i < NY; ++i) r[i] = -(we[WS_GQ + NVP + i] + hv[NVP + i]);
    for (int a = 0; a < NX; ++a) r[DMB_XF_SEL[a]] -= lc[a];
    lu_solve_t(we, r);
    for (int i = 0; i < NY; ++i) I.lamn[LAM_EDGE(e) + i] = r[i];
#if NNL > 0
    {
      double jd[NNL];
      for (int i = 0; i < NNL; ++i) jd[i] = 0.0;
      for (int kk = 0; kk < DMB_NNZ_D; ++kk) jd[DMB_D_ROW[kk]] += we[WS_DNZ + kk] * v[DMB_D_COL[kk]];
      for (int i = 0; i < NNL; ++i) {
        const double dsi = jd[i] + res[LAM_EDGE(e) + NY + NX + i];
        I.ds[e * NNL + i] = dsi;
        // y_d+ = Sigma_s ds + grad_s phi = Sigma_s*J dq + r_d
        I.lamn[LAM_EDGE(e) + NY + NX + i] = we[WS_SIGS + i] * jd[i] + we[WS_RD + i];
      }
    }
#endif
  }
  {
    const double* w0 = I.node;
    for (int a = c.tid; a < NX; a += c.nthr) {
      double s = w0[WN_PV + a];
      for (int t = 0; t < NXI; ++t) s += w0[WN_P + a * NXI + t] * w0[WN_DXI + t];
      I.lamn[a] = -s;
    }
  }
  CTA_SYNC();
}
