// Batched primal-dual interior-point solver for stage-structured collocation NLPs (sm_100a) — core phases.
//
// Replaces the CasADi nlpsol('ipopt') call of the reference (do_mpc/controller/_mpc.py:1328,
// do_mpc/optimizer.py:770) for a BATCH of problem instances.  One CTA solves one instance at a time
// (persistent CTAs pull instances from an atomic queue); the filter line-search IPM follows
// Waechter & Biegler (2006) with IPOPT's default options, the same algorithm the CPU oracle restates.
// The KKT system of every iteration is solved by a structure-exploiting elimination instead of a
// sparse LDL^T (MUMPS in the reference):
//   per edge  : LU of the square collocation block G_y, sensitivities Y = -G_y^{-1} G_v (null-space
//               projection of the edge Hessian onto the node variables)      -> one lane GROUP per edge,
//   per node  : Riccati / dynamic-programming sweep over the scenario tree in the augmented state
//               xi = (x, u_prev); positive definiteness of every control block == correct inertia,
//   recovery  : forward roll-out, then multipliers by a transposed LU solve per edge.
// All problem dimensions are compile-time constants from the generated header (problem_gen.h); the
// model functions (residuals, Jacobian / Lagrangian-Hessian nonzeros) are generated straight-line code.
//
// Memory plan (v2, after ncu capture r1a showed the kernel bound by dependent L1/L2 round trips in
// single-thread sections): everything that sits on a dependency chain lives in shared memory
// (iterates, per-edge reduced blocks, per-node Riccati data, LU tiles); bulk per-edge data that is
// written once and read once per iteration (LU factors, Y, nonzeros) goes to a per-CTA global slot
// and is only accessed with group-wide coalesced loops.
//
// The file compiles for the device (nvcc) and, with DMB_HOSTSIM, as plain single-thread C++ so that
// the numerics can be checked on a machine without a GPU (tests/hostsim; never part of the product).
#pragma once
#include <math.h>
#include <stdint.h>

struct Ctx { int tid, nthr, lane, warp, nwarp, gl, grp, ngrp; };

#ifdef DMB_HOSTSIM
#define DMB_DEV
#define DMB_PHASE
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
#define DMB_LANES 1
#define CTA_SYNC() ((void)0)
#define WARP_SYNC() ((void)0)
#else
#define DMB_DEV __device__ __forceinline__
#ifdef DMB_NOINLINE_PHASES
#define DMB_PHASE __device__ __noinline__
#else
#define DMB_PHASE __device__ __forceinline__
#endif
#define DMB_LANES 32
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
  double nlp_scaling_max_gradient, nlp_scaling_min_value, kappa_resto, residual_ratio_max;
  double ws_bound_push, ws_bound_frac, ws_mult_bound_push, ws_mult_init_max;   // warm_start_* (used when warm_start != 0)
  int max_iter, acceptable_iter, max_soc, mimic_dummy_reg, resto_max_iter, warm_start, max_refine, pad_;
};

enum DmbStatus {
  DMB_SOLVE_SUCCEEDED = 0, DMB_SOLVED_TO_ACCEPTABLE_LEVEL = 1, DMB_INFEASIBLE_PROBLEM_DETECTED = 2,
  DMB_MAXIMUM_ITERATIONS_EXCEEDED = -1, DMB_RESTORATION_FAILED = -2, DMB_ERROR_IN_STEP_COMPUTATION = -3,
  DMB_INVALID_NUMBER_DETECTED = -13, DMB_NOT_RUN = -100
};

#include "problem_gen.h"   // generated: dimensions, tables, model functions

// tree topology of nodes / edges.  Chains (one scenario) need no table loads — node n is stage n with the single edge n to
// node n + 1 — which takes dependent global loads out of every sweep (r1e capture: 14 % of the sweep's stalls).
#if DMB_NSCEN > 1
#define NODE_NEDGE(n) DMB_NODE_NEDGE[n]
#define NODE_FIRST_EDGE(n) DMB_NODE_FIRST_EDGE[n]
#define EDGE_CHILD_OF(e) DMB_EDGE_CHILD[e]
#define EDGE_PARENT_OF(e) DMB_EDGE_PARENT[e]
#define NODE_PARENT_EDGE(n) DMB_NODE_PARENT_EDGE[n]
#define NODE_K_OF(n) DMB_NODE_K[n]
#define EDGE_K_OF(e) DMB_EDGE_K[e]
#define EDGE_PIDX_OF(e) DMB_EDGE_PIDX[e]
#define EDGE_OMEGA_OF(e) DMB_EDGE_OMEGA[e]
#define NODE_RW_OF(n) DMB_NODE_RW[n]
#else
#define NODE_NEDGE(n) ((n) < DMB_NHOR ? 1 : 0)
#define NODE_FIRST_EDGE(n) (n)
#define EDGE_CHILD_OF(e) ((e) + 1)
#define EDGE_PARENT_OF(e) (e)
#define NODE_PARENT_EDGE(n) ((n) - 1)
#define NODE_K_OF(n) (n)
#define EDGE_K_OF(e) (e)
#define EDGE_PIDX_OF(e) 0
#define EDGE_OMEGA_OF(e) 1.0
#define NODE_RW_OF(n) ((n) < DMB_NHOR ? 1.0 : 0.0)
#endif

// ---------------------------------------------------------------------------------------------------
// derived dimensions and layouts
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
#define NE (NVP + 1)                  // extended right-hand sides [G_vp | rhs]
#define NNODE DMB_NNODE
#define NEDGE DMB_NEDGE
#define NHOR DMB_NHOR
#define NWI (NNODE * NVP + NEDGE * NY)               // internal primal vector
#define NODE_OFF(n) ((n) * NVP)
#define EDGE_OFF(e) (NNODE * NVP + (e) * NY)
#define NLE (NY + NX + NNL)                           // multipliers per edge: lamG | lamC | y_d
#define NLAM (NX + NEDGE * NLE)
#define LAM_EDGE(e) (NX + (e) * NLE)
#define DMB_INF 1e19
#define MAX_FILTER 32
// partially free root (moving-horizon estimation: the whole root state; nl_cons_single_slack: the lifted slack): the LAST
// DMB_NFREE0 components of the root state are decision variables without an initial-condition row; the root may carry its
// own cost term (MHE arrival cost, gen_arrival*).  NXF = number of root components fixed by x0.
#ifndef DMB_NFREE0
#define DMB_NFREE0 0
#endif
#ifndef DMB_HAS_ARRIVAL
#define DMB_HAS_ARRIVAL 0
#endif
#define NXF (NX - DMB_NFREE0)
// weight of an internal variable in the least-squares multiplier estimate (H := I): copies of a lifted global variable share
// the unit weight of the flat NLP's single variable (codegen.py)
#if DMB_HAS_INIT_FROM
#define DMB_LSW(i) DMB_LS_WEIGHT[i]
#else
#define DMB_LSW(i) 1.0
#endif

// lane groups: one group factorises one edge
#ifdef DMB_HOSTSIM
#define GSZ 1
#else
#define GSZ (NY <= 16 ? 16 : 32)
#endif
// LU tile of a group: NY rows x [G_y (NY) | G_vp,rhs (NE) | M_ext (NE)], a compact NVP x NE block N = H_vy*Yext, and a
// staging area for the bulk per-edge data (global loads never sit inside dependent inner loops)
#define TILE_NC (NY + NE)
// Lane-flattened LU update / back substitution of the factor phase (all lanes of a group share the entries of a step instead
// of one lane per row / per right-hand side): fewer dependent steps per lane but more instructions.  Measured on B200
// (profiles/r2_flat_lu_ab.json): DIP N=50 (chain, NY = 27) +8.8 % / +1.5 %, batch reactor tree (NY = 20) -1.4 % / -6 %,
// industrial_poly tree -12 % with both, CSTR (NY = 12, two groups per warp) -0.5 %: it pays where the kernel is latency bound
// (one scenario) and the blocks are large; scenario trees have enough warps in flight and are bound by instruction issue.
#ifndef DMB_LU_FLAT
#define DMB_LU_FLAT (DMB_NSCEN == 1 && DMB_NY > 16)
#endif
#ifndef DMB_BS_FLAT
#define DMB_BS_FLAT (DMB_NSCEN == 1 && DMB_NY > 16)
#endif
#define TILE_M (NY + NE)
#define TILE_LD ((NY + 2 * NE) | 1)                   // odd leading dimension: conflict-free 64-bit rows
#define STG_N (NY * TILE_LD)                           // N block, row-major NVP x NE
#define STG_HNZ (STG_N + NVP * NE)
#define STG_GY (STG_HNZ + DMB_NNZ_H)
#define STG_SIGY (STG_GY + NY)
#define STG_DNZ (STG_SIGY + NY)
#define STG_SIGS (STG_DNZ + DMB_NNZ_D)
#define STG_RD (STG_SIGS + NNL)
#define STG_GOBJ (STG_RD + NNL)                        // factor phase: gobj (NQ), reciprocal pivots (NY), permutation (NY)
#define STG_INVD (STG_GOBJ + NQ)
#define STG_PERM (STG_INVD + NY)                       // perm[m] = original row of the m-th pivot (as doubles)
#define STG_DQ STG_GOBJ                                // recovery phase re-uses that space: dq (NQ), r (NY), jd (NNL)
#define TILE_SZ ((STG_PERM + NY + 1) & ~1)             // even: every tile starts 16-byte aligned (TMA bulk copies)
#define IMG_SZ ((NY * TILE_LD + 1) & ~1)                // tile image [LU | Yext | ..] saved to / loaded from the slot in one bulk copy
static_assert(NY >= NNL, "recovery scratch overlays the factor scratch: needs NY >= NNL");

// per-edge reduced data kept in shared memory (inputs of the Riccati sweep)
#define ER_AB 0                                        // Y[xf rows] : NX x NVP
#define ER_B (ER_AB + NX * NVP)                        // b (NX)
#define ER_HH (ER_B + NX)                              // T'HT without delta (NVP x NVP)
#define ER_YTY (ER_HH + NVP * NVP)                     // Y'Y
#define ER_GH (ER_YTY + NVP * NVP)                     // reduced gradient (NVP)
#define ER_YTY0 (ER_GH + NVP)                          // Y'y0
#define ER_SZ (ER_YTY0 + NVP)
// per-edge bulk data in the per-CTA global slot (coalesced group access only)
#define EB_IMG 0                                       // image of the tile rows [0, NY): LU factors (reciprocal pivots on the
                                                       // diagonal) in columns [0, NY), Yext = [Y | y0] in columns [NY, NY + NE)
#define EB_PIV (EB_IMG + IMG_SZ)
#define EB_JNZ (EB_PIV + NY)
#define EB_HNZ (EB_JNZ + DMB_NNZ_J)
#define EB_DNZ (EB_HNZ + DMB_NNZ_H)                    // scaled inequality Jacobian nonzeros
#define EB_GOBJ (EB_DNZ + DMB_NNZ_D)                   // edge-local objective gradient wrt q (NQ)
#define EB_GCON (EB_GOBJ + NQ)                         // edge-local constraint part of grad L wrt q (NQ)
#define EB_GY (EB_GCON + NQ)                           // total barrier-objective gradient of y (NY)
#define EB_SIGY (EB_GY + NY)
#define EB_SIGS (EB_SIGY + NY)
#define EB_RD (EB_SIGS + NNL)
#define EB_GY2 (EB_RD + NNL)                           // y-part of the refinement right-hand side (NY)
#define EB_SZ ((EB_GY2 + NY + 1) & ~1)                 // even: every edge block starts 16-byte aligned
// per-node data in shared memory.  Q (NV x NV, v-space) is assembled in parallel (node terms + projected edge
// Hessians), completed and reduced in place by the sweep: afterwards its leading NXI x NXI block (stride NV) is P.
#define WN_Q 0                                          // Q / P (NV x NV)
#define WN_QV (WN_Q + NV * NV)                          // q / p (NV)
#define WN_K (WN_QV + NV)                               // K (NUT x NXI)
#define WN_KV (WN_K + NUT * NXI)                        // k (NUT)
#define WN_DXI (WN_KV + NUT)                            // d xi (NXI)
#define WN_RG (WN_DXI + NXI)                            // weighted rterm gradient (u, u_prev)
#define WN_RH (WN_RG + 2 * NU)                          // weighted rterm Hessian (2NU x 2NU)
#define WN_NODE (WN_RH + 4 * NU * NU)
#define P_AT(wn, a, b) (wn)[WN_Q + (a) * NV + (b)]
#define VPOS(i) ((i) < NX ? (i) : (i) + NU)             // position of edge-local node variable i in v-space
// iterate vectors
#define WV_W 0
#define WV_WT (WV_W + NWI)
#define WV_DW (WV_WT + NWI)
#define WV_ZL (WV_DW + NWI)
#define WV_ZU (WV_ZL + NWI)
#define WV_GL (WV_ZU + NWI)                              // grad of Lagrangian without bound multipliers
#define WV_GO (WV_GL + NWI)                              // objective gradient
#define WV_LAM (WV_GO + NWI)
#define WV_LAMN (WV_LAM + NLAM)
#define WV_RES (WV_LAMN + NLAM)                          // residuals, same layout as lam
#define WV_REST (WV_RES + NLAM)
#define WV_LAMR (WV_REST + NLAM)                         // primary multipliers while a refinement solve is running
#define WV_DC (WV_LAMR + NLAM)                           // row scaling, same layout as lam
#define WV_S (WV_DC + NLAM)                              // inequality slacks (NEDGE*NNL)
#define WV_ST (WV_S + NEDGE * NNL)
#define WV_DS (WV_ST + NEDGE * NNL)
#define WV_VU (WV_DS + NEDGE * NNL)
#define WV_S2 (WV_VU + NEDGE * NNL)                       // second-order-correction trial slacks / restoration slack iterate
#define WV_SOC (WV_S2 + NEDGE * NNL)                      // residual of a second-order-correction trial point (NLAM)
#define WV_END (WV_SOC + NLAM)
#define DMB_QS_WARP (NXI * NVP + NXI + 8)                // Riccati scratch of one warp
#if defined(DMB_THREADS) && DMB_NSCEN > 1
#define DMB_QS_SIZE ((DMB_THREADS / 32) * DMB_QS_WARP)    // scenario trees: one scratch per warp, see DMB_TREE_PAR
#else
#define DMB_QS_SIZE DMB_QS_WARP
#endif

#ifdef DMB_HOSTSIM
#define DMB_CLOCK() 0LL
#else
#define DMB_CLOCK() clock64()
#endif
#define DMB_NSUB 8
#ifdef DMB_PROFILE
#define SUB_TICK(I, ph) do { if ((I).sub) { const long long t_ = DMB_CLOCK(); (I).sub[ph] += t_ - tsub_; tsub_ = t_; } } while (0)
#define SUB_START() long long tsub_ = DMB_CLOCK()
#else
#define SUB_TICK(I, ph) ((void)0)
#define SUB_START() ((void)0)
#endif
struct Inst {
  long long* sub;      // optional fine-grained cycle counters (thread 0 only), see SUB_TICK
  double *w, *wt, *dw, *zL, *zU, *gl, *go, *lam, *lamn, *lamr, *res, *rest, *dc, *s, *st, *ds, *vU, *s2, *soc;
  double *er, *node, *eb;
  const double *p, *lb, *ub;
  double df, relax_fac, relax_max;
  // feasibility restoration (ipm_solve.cuh): resto != 0 switches the barrier terms of every primal variable to those of the
  // restoration problem (see resto_terms), zeta = sqrt(mu); the reference point of its proximal term is kept in go / s2
  double zeta;
  int resto;
  unsigned long long* mbar;   // one mbarrier per lane group (TMA loads of the tile image), shared memory
  int* mphase;                // its current phase parity
};
// restoration problem  min zeta/2 |D_R (v - v_R)|^2 - mu log(v - lo) - mu log(hi - v), D_R = 1/max(1,|v_R|): diagonal Hessian
// entry and gradient of one variable (primal barrier: no multipliers involved)
DMB_DEV void resto_terms(double zeta, double mu, double v, double vref, double lo, double hi, bool hl, bool hu, double& sig, double& gb) {
  const double dr = 1.0 / fmax(1.0, fabs(vref));
  sig = zeta * dr * dr; gb = sig * (v - vref);
  if (hl) { const double rs = 1.0 / (v - lo); sig += mu * rs * rs; gb -= mu * rs; }
  if (hu) { const double rs = 1.0 / (hi - v); sig += mu * rs * rs; gb += mu * rs; }
}
DMB_DEV void bind_vectors(Inst& I, double* vec) {
  I.w = vec + WV_W; I.wt = vec + WV_WT; I.dw = vec + WV_DW; I.zL = vec + WV_ZL; I.zU = vec + WV_ZU;
  I.gl = vec + WV_GL; I.go = vec + WV_GO;
  I.lam = vec + WV_LAM; I.lamn = vec + WV_LAMN; I.lamr = vec + WV_LAMR; I.res = vec + WV_RES; I.rest = vec + WV_REST; I.dc = vec + WV_DC;
  I.s = vec + WV_S; I.st = vec + WV_ST; I.ds = vec + WV_DS; I.vU = vec + WV_VU; I.s2 = vec + WV_S2; I.soc = vec + WV_SOC;
  I.zeta = 0.0; I.resto = 0; I.mbar = nullptr; I.mphase = nullptr;
}
DMB_DEV const double* p_tvp_at(const double* p, int k) { return p + DMB_P_TVP + k * DMB_NTVP; }

// ---------------------------------------------------------------------------------------------------
// reductions
// ---------------------------------------------------------------------------------------------------
#define DMB_RED_SUM 0
#define DMB_RED_MAX 1
#define DMB_RED_MIN 2
#ifdef DMB_HOSTSIM
static inline double block_sum(const Ctx&, double v, double*) { return v; }
static inline double block_max(const Ctx&, double v, double*) { return v; }
static inline double block_min(const Ctx&, double v, double*) { return v; }
static inline int group_argmax(const Ctx&, double, int idx) { return idx; }
template <int NRED> static inline void block_reduce(const Ctx&, double (&)[NRED], const int (&)[NRED], double*) {}
#else
#define DMB_FULL 0xffffffffu
DMB_DEV double warp_sum(double v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(DMB_FULL, v, o); return v; }
DMB_DEV double warp_max(double v) { for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(DMB_FULL, v, o)); return v; }
DMB_DEV double warp_min(double v) { for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(DMB_FULL, v, o)); return v; }
// scratch: >= 2*nwarp doubles of shared memory (double-buffered by call parity); result broadcast to all threads
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
// several reductions behind ONE pair of barriers (the IPM loop reduces ~25 scalars per iteration; one block_sum / block_max
// each cost two barriers apiece).  op[k]: 0 = sum, 1 = max, 2 = min.  Same shuffle tree and same order over the warps as
// block_sum / block_max / block_min, so the results are bit-identical to the separate calls.  Result broadcast in v[].
// scratch: >= NV * nwarp doubles.
template <int NRED>
DMB_DEV void block_reduce(const Ctx& c, double (&v)[NRED], const int (&op)[NRED], double* scratch) {
#pragma unroll
  for (int k = 0; k < NRED; ++k) v[k] = op[k] == DMB_RED_SUM ? warp_sum(v[k]) : (op[k] == DMB_RED_MAX ? warp_max(v[k]) : warp_min(v[k]));
  CTA_SYNC();
  if (c.lane == 0) {
#pragma unroll
    for (int k = 0; k < NRED; ++k) scratch[k * c.nwarp + c.warp] = v[k];
  }
  CTA_SYNC();
#pragma unroll
  for (int k = 0; k < NRED; ++k) {
    double r = scratch[k * c.nwarp];
    if (op[k] == DMB_RED_SUM) { r = 0.0; for (int i = 0; i < c.nwarp; ++i) r += scratch[k * c.nwarp + i]; }
    else if (op[k] == DMB_RED_MAX) for (int i = 1; i < c.nwarp; ++i) r = fmax(r, scratch[k * c.nwarp + i]);
    else for (int i = 1; i < c.nwarp; ++i) r = fmin(r, scratch[k * c.nwarp + i]);
    v[k] = r;
  }
}
// index of the largest |val| among the lanes of a group (idx < 64, idx < 0 = no candidate): one REDUX on packed keys
DMB_DEV int group_argmax(const Ctx& c, double val, int idx) {
  unsigned key = 0u;
  if (idx >= 0) key = ((__float_as_uint((float)fabs(val)) & 0x7fffffc0u) | (unsigned)(63 - idx)) + 64u;
  const unsigned mask = GSZ == 32 ? DMB_FULL : (0xffffu << (c.lane & 16));
  const unsigned best = __reduce_max_sync(mask, key);
  return 63 - (int)((best - 64u) & 63u);
}
#endif

#if !defined(DMB_HOSTSIM) && !defined(DMB_NO_DMMA)
// FP64 tensor-core MMA (sm_100a: DMMA.8x8x4): D(8x8) += A(8x4, row) * B(4x8, col).  Fragments per lane (r = lane / 4,
// q = lane % 4): A[r][q], B[q][r], C/D[r][2q], C/D[r][2q + 1].
DMB_DEV void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// element (row, col) of a matrix held in the C-fragment layout, delivered to the calling lane (row / col may differ per lane)
DMB_DEV double frag_get(double c0, double c1, int row, int col) {
  const int src = row * 4 + (col >> 1);
  const double v0 = __shfl_sync(DMB_FULL, c0, src), v1 = __shfl_sync(DMB_FULL, c1, src);
  return (col & 1) ? v1 : v0;
}
// projection products of the factor phase: only with one tile per warp (NY > 16).  Measured on B200 (round 2, r2l): with two
// 16-lane groups per warp the tiles take turns on the tensor pipe and the dense H_d build costs more than the products save
// (cfg #2, NY = 12: 27.65 ms with / 27.24 ms without); NY = 24 / 36 / 30 (cfg #3 / #4 / #5): +2 % / +3 % / +5.5 % throughput
#define DMB_DMMA_FACTOR (DMB_NY > 16)
#define DMB_DMMA_SWEEP (DMB_NX + 2 * DMB_NU + DMB_NEPS <= 8)  // Riccati sweep in the 8 x 8 padded v-space
#else
#define DMB_DMMA_FACTOR 0
#define DMB_DMMA_SWEEP 0
#endif

// ---------------------------------------------------------------------------------------------------
// TMA bulk copies (cp.async.bulk, SASS UBLKCP) between a lane group's tile and its image in the per-CTA slot: one
// instruction by one lane moves the whole [LU | Yext] image (2.6 KB for cfg #2) instead of ~15 LDG + STS per lane, and the
// factor phase keeps computing while the store drains.  One mbarrier per lane group tracks the loads.
// ---------------------------------------------------------------------------------------------------
#ifndef DMB_HOSTSIM
#define DMB_TMA 1
DMB_DEV unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
DMB_DEV void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
DMB_DEV void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  const unsigned m = smem_u32(bar);
  asm volatile("fence.proxy.async;" ::: "memory");                     // earlier generic writes of the image (y0 column) -> async proxy
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(m), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(m) : "memory");
}
DMB_DEV void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile("{\n .reg .pred P1;\n LAB_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n @P1 bra DONE;\n bra LAB_WAIT;\n DONE:\n }"
               :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// smem -> global; returns when the source may be overwritten (the write itself is awaited by tma_store_drain)
DMB_DEV void tma_store_1d(void* gdst, const void* smem_src, unsigned bytes) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");         // generic writes of the tile -> visible to the async proxy
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
DMB_DEV void tma_store_drain() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
#else
#define DMB_TMA 0
#endif

// batched copy: all loads of a lane are issued before the first store (the two pointers may alias for the
// compiler, which otherwise serialises load -> store -> load through the memory latency)
#ifdef DMB_NO_COPY_BATCH
#define DMB_COPY_BATCH(N, GL, STRIDE, LOADEXPR, ...) \
  do { for (int idx = (GL); idx < (N); idx += (STRIDE)) { const double val = (LOADEXPR); __VA_ARGS__; } } while (0)
#else
#define DMB_COPY_BATCH(N, GL, STRIDE, LOADEXPR, ...)                                   \
  do {                                                                                      \
    constexpr int n_it_ = ((N) + (STRIDE) - 1) / (STRIDE);                                  \
    double tmp_[n_it_ > 0 ? n_it_ : 1];                                                                  \
    _Pragma("unroll") for (int it_ = 0; it_ < n_it_; ++it_) {                               \
      const int idx = (GL) + it_ * (STRIDE);                                                \
      tmp_[it_] = idx < (N) ? (LOADEXPR) : 0.0;                                             \
    }                                                                                       \
    _Pragma("unroll") for (int it_ = 0; it_ < n_it_; ++it_) {                               \
      const int idx = (GL) + it_ * (STRIDE);                                                \
      const double val = tmp_[it_];                                                         \
      if (idx < (N)) { __VA_ARGS__; }                                                       \
    }                                                                                       \
  } while (0)
#endif

// ---------------------------------------------------------------------------------------------------
// gather of the edge-local vector q = (x_parent, u_parent, eps_parent, y_e) and its parameters
// ---------------------------------------------------------------------------------------------------
DMB_DEV void gather_q(const double* w, int e, double* q) {
  const double* wn = w + NODE_OFF(EDGE_PARENT_OF(e));
#pragma unroll
  for (int i = 0; i < NVP; ++i) q[i] = wn[i];
  const double* wy = w + EDGE_OFF(e);
#pragma unroll
  for (int i = 0; i < NY; ++i) q[NVP + i] = wy[i];
}
DMB_DEV const double* edge_tvp(const double* p, int e) { return p + DMB_P_TVP + EDGE_K_OF(e) * DMB_NTVP; }
DMB_DEV const double* edge_pp(const double* p, int e) { return p + DMB_P_P + EDGE_PIDX_OF(e) * DMB_NPM; }

// u_prev (scaled) seen by node n: the parent's input, or the parameter for the root (reference _mpc.py:1273)
DMB_DEV double node_uprev(const Inst& I, const double* w, int n, int i) {
  const int pe = NODE_PARENT_EDGE(n);
  if (pe < 0) return I.p[DMB_P_UPREV + i] * DMB_INV_USCALE[i];
  return w[NODE_OFF(EDGE_PARENT_OF(pe)) + NX + i];
}

#if DMB_NNL > 0
// relaxed upper bound of the (scaled) inequality row i of edge e (IPOPT bound_relax_factor)
DMB_DEV double nl_ub_relaxed(const Inst& I, int e, int i) {
  const double u = I.dc[LAM_EDGE(e) + NY + NX + i] * DMB_NL_UB[i];
  return u + fmin(I.relax_fac * fmax(1.0, fabs(u)), I.relax_max);
}
#endif

// ---------------------------------------------------------------------------------------------------
// residuals + objective at a point (w, s) -> res (scaled), objective (scaled by df), theta = ||res||_1.
// One thread per edge for the generated code; all threads for the reductions.
// ---------------------------------------------------------------------------------------------------
// per-thread partial sums of the barrier terms of a point: acc += -mu log(slack) (+ kappa_d mu slack for one-sided bounds),
// bad = 1 outside the bounds.  The logarithms of a thread's slacks are taken as ONE log of their running product (flushed
// before it could leave the normal range): an FP64 log costs ~40 instructions, a multiplication one.
#define DMB_LOGPROD(SLACK)                                                                                   \
  do { prod *= (SLACK); if (prod < 1e-120 || prod > 1e120) { acc -= mu * log(prod); prod = 1.0; } } while (0)
DMB_DEV void barrier_terms(const Ctx& c, const Inst& I, const double* w, const double* s, double mu, double kappa_d, double& acc, double& bad) {
  double prod = 1.0;
  for (int i = c.tid; i < NWI; i += c.nthr) {
    const bool hl = I.lb[i] > -DMB_INF, hu = I.ub[i] < DMB_INF;
    if (hl) { const double sl = w[i] - I.lb[i]; if (sl <= 0.0) bad = 1.0; else { DMB_LOGPROD(sl); if (!hu) acc += kappa_d * mu * sl; } }
    if (hu) { const double su = I.ub[i] - w[i]; if (su <= 0.0) bad = 1.0; else { DMB_LOGPROD(su); if (!hl) acc += kappa_d * mu * su; } }
  }
#if NNL > 0
  for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) {
    const double su = nl_ub_relaxed(I, i / NNL, i % NNL) - s[i];
    if (su <= 0.0) bad = 1.0; else { DMB_LOGPROD(su); acc += kappa_d * mu * su; }
  }
#else
  (void)s;
#endif
  if (prod != 1.0) acc -= mu * log(prod);
}

// phi_out != nullptr (and phi_mu >= 0): also the barrier function phi_mu(w, s) = f + barrier terms, in the same reduction
DMB_PHASE void eval_point(const Ctx& c, const Inst& I, const double* w, const double* s, double* res,
                        double* scratch, double* f_out, double* theta_out, double* infpr_out,
                        double phi_mu = -1.0, double phi_kappa_d = 0.0, double* phi_out = nullptr) {
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
    const double* xc = w + NODE_OFF(EDGE_CHILD_OF(e));
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
    f += EDGE_OMEGA_OF(e) * lt + ec;
    if (EDGE_K_OF(e) == NHOR - 1) f += EDGE_OMEGA_OF(e) * gen_mterm(xc, p_tvp_at(I.p, NHOR), pp);
  }
  for (int n = c.tid; n < NNODE; n += c.nthr) {
    if (NODE_K_OF(n) == NHOR) continue;
    double up[NU];
#pragma unroll
    for (int i = 0; i < NU; ++i) up[i] = node_uprev(I, w, n, i);
    f += NODE_RW_OF(n) * gen_rterm(w + NODE_OFF(n) + NX, up, I.p + DMB_P_TVP + NODE_K_OF(n) * DMB_NTVP);
#if DMB_HAS_ARRIVAL
    if (n == 0) f += gen_arrival(w + NODE_OFF(0), p_tvp_at(I.p, 0), I.p + DMB_P_P);
#endif
  }
  for (int i = c.tid; i < NX; i += c.nthr) res[i] = i < NXF ? w[NODE_OFF(0) + i] - I.p[DMB_P_X0 + i] * DMB_INV_XSCALE[i] : 0.0;
  CTA_SYNC();
  double th = 0.0, mx = 0.0;
  for (int i = c.tid; i < NLAM; i += c.nthr) { const double a = fabs(res[i]); th += a; mx = fmax(mx, a); }
  double v[5] = {f, th, mx, 0.0, 0.0};
  if (phi_mu >= 0.0) barrier_terms(c, I, w, s, phi_mu, phi_kappa_d, v[3], v[4]);
  const int op[5] = {DMB_RED_SUM, DMB_RED_SUM, DMB_RED_MAX, DMB_RED_SUM, DMB_RED_MAX};
  block_reduce<5>(c, v, op, scratch);
  *f_out = I.df * v[0];
  *theta_out = v[1];
  if (infpr_out) *infpr_out = v[2];
  if (phi_out) *phi_out = v[4] > 0.0 ? INFINITY : *f_out + v[3];
}

// ---------------------------------------------------------------------------------------------------
// all derivatives at the current point, once per iteration: per edge (one thread each, fused generated
// code) Jacobian / Hessian nonzeros and gradients into the bulk slot; then the gradients of the
// objective (go) and of the Lagrangian without bound multipliers (gl) for every primal variable.
// lam are the SCALED multipliers; the row scales dc are applied here.
// ---------------------------------------------------------------------------------------------------
// (measured and rejected in round 2: splitting the generated function into 4 parts evaluated by 4 warps — the expensive
//  sub-expressions (exp / division of the Arrhenius terms) are shared by all outputs and were recomputed per part:
//  cfg #2 34.4 ms with the split against 32.3 ms without, gpurun_out/r2e_quick.json / r2f_quick.json)
DMB_PHASE void eval_derivatives(const Ctx& c, const Inst& I) {
  const double* w = I.w;
  for (int e = c.tid; e < NEDGE; e += c.nthr) {
    double q[NQ], lg[NY > 0 ? NY : 1], ln[NNL > 0 ? NNL : 1];
    gather_q(w, e, q);
    const double* dc = I.dc + LAM_EDGE(e);
    const double* le = I.lam + LAM_EDGE(e);
#pragma unroll
    for (int i = 0; i < NY; ++i) lg[i] = dc[i] * le[i];
#pragma unroll
    for (int i = 0; i < NNL; ++i) ln[i] = dc[NY + NX + i] * le[NY + NX + i];
    double* eb = I.eb + (size_t)e * EB_SZ;
    gen_derivs(q, edge_tvp(I.p, e), edge_pp(I.p, e), lg, I.df * EDGE_OMEGA_OF(e), I.df, ln,
               eb + EB_GOBJ, eb + EB_GCON, eb + EB_JNZ, eb + EB_HNZ, eb + EB_DNZ);
  }
  // weighted rterm derivatives per node (reference _mpc.py:1263-1275), used by the gradients below and the sweep
  for (int n = c.nthr - 1 - c.tid; n < NNODE; n += c.nthr) {
    double* wn = I.node + (size_t)n * WN_NODE;
    const int k = NODE_K_OF(n);
    if (k == NHOR) continue;
    double up[NU], gr[2 * NU], hr[4 * NU * NU];
#pragma unroll
    for (int a = 0; a < NU; ++a) up[a] = node_uprev(I, w, n, a);
    const double* tv = I.p + DMB_P_TVP + k * DMB_NTVP;
    gen_rterm_grad(w + NODE_OFF(n) + NX, up, tv, gr);
    gen_rterm_hess(w + NODE_OFF(n) + NX, up, tv, hr);
    const double wgt = I.df * NODE_RW_OF(n);
#pragma unroll
    for (int a = 0; a < 2 * NU; ++a) wn[WN_RG + a] = wgt * gr[a];
#pragma unroll
    for (int a = 0; a < 4 * NU * NU; ++a) wn[WN_RH + a] = wgt * hr[a];
  }
  CTA_SYNC();
  // node variables: sum over child edges (+ continuity of the parent edge, terminal cost, rterm)
  for (int idx = c.tid; idx < NNODE * NVP; idx += c.nthr) {
    const int n = idx / NVP, i = idx - n * NVP;
    const int k = NODE_K_OF(n);
    double go = 0.0, gc = 0.0;
    if (k < NHOR) {
      for (int j = 0; j < NODE_NEDGE(n); ++j) {
        const double* eb = I.eb + (size_t)(NODE_FIRST_EDGE(n) + j) * EB_SZ;
        go += eb[EB_GOBJ + i]; gc += eb[EB_GCON + i];
      }
    } else if (i >= NX) { I.gl[idx] = 0.0; if (!I.resto) I.go[idx] = 0.0; continue; }
    if (i < NX) {
      const int pe = NODE_PARENT_EDGE(n);
      if (pe >= 0) {
        gc -= I.lam[LAM_EDGE(pe) + NY + i];
        if (k == NHOR) {
          double gm[NX];
          gen_mterm_grad(w + NODE_OFF(n), p_tvp_at(I.p, NHOR), edge_pp(I.p, pe), gm);
          go += I.df * EDGE_OMEGA_OF(pe) * gm[i];
        }
      } else {
        gc += I.lam[i];
#if DMB_HAS_ARRIVAL
        double ga[NX];
        gen_arrival_grad(w + NODE_OFF(0), p_tvp_at(I.p, 0), I.p + DMB_P_P, ga);
        go += I.df * ga[i];
#endif
      }
    } else if (i < NX + NU) {
      const int iu = i - NX;
      go += I.node[(size_t)n * WN_NODE + WN_RG + iu];
      // this node's input is u_prev of every child node that still has a control
      for (int j = 0; j < NODE_NEDGE(n); ++j) {
        const int cn = EDGE_CHILD_OF(NODE_FIRST_EDGE(n) + j);
        if (NODE_K_OF(cn) == NHOR) continue;
        go += I.node[(size_t)cn * WN_NODE + WN_RG + NU + iu];
      }
    }
    if (!I.resto) I.go[idx] = go;            // restoration: go holds the reference point of the proximal term
    I.gl[idx] = go + gc;
  }
  for (int idx = c.tid; idx < NEDGE * NY; idx += c.nthr) {
    const int e = idx / NY, i = idx - e * NY;
    const double* eb = I.eb + (size_t)e * EB_SZ;
    const double go = eb[EB_GOBJ + NVP + i];
    if (!I.resto) I.go[NNODE * NVP + idx] = go;
    // constraint part + continuity multiplier on the end-of-interval state (same association as the former in-place update)
    const double gcon = DMB_XF_INV[i] >= 0 ? eb[EB_GCON + NVP + i] + I.lam[LAM_EDGE(e) + NY + DMB_XF_INV[i]] : eb[EB_GCON + NVP + i];
    I.gl[NNODE * NVP + idx] = go + gcon;
  }
#if NNL > 0
  // inequality Jacobian nonzeros: row scaling
  for (int idx = c.tid; idx < NEDGE * DMB_NNZ_D; idx += c.nthr) {
    const int e = idx / DMB_NNZ_D, k = idx - e * DMB_NNZ_D;
    I.eb[(size_t)e * EB_SZ + EB_DNZ + k] *= I.dc[LAM_EDGE(e) + NY + NX + DMB_D_ROW[k]];
  }
#endif
  CTA_SYNC();
}

// ---------------------------------------------------------------------------------------------------
// factor phase.  One lane group per edge: assemble [G_y | G_vp | -res] from the Jacobian nonzeros,
// LU with partial pivoting (REDUX pivot search), Yext = [Y | y0] by back substitution (one lane per
// right-hand side), sparse Hessian products M = (H_yy+Sigma_y) Yext, N = H_vy Yext, and the reduced
// blocks for the Riccati sweep.  `res` supplies the right-hand side (ignored in ls_mode: zero).
// ls_mode = least-squares multiplier estimate: Hessian := identity, gradient := grad f - zL + zU.
// ---------------------------------------------------------------------------------------------------
DMB_PHASE void factor_edges(const Ctx& c, const Inst& I, const double* res, double mu, double kappa_d, int ls_mode,
                               double* tile_base, int* fail) {
  double* T = tile_base + (size_t)c.grp * TILE_SZ;
  SUB_START();
  for (int e0 = 0; e0 < NEDGE; e0 += c.ngrp) {
    const int e = e0 + c.grp;
    const bool act = e < NEDGE;
    double* eb = I.eb + (size_t)(act ? e : 0) * EB_SZ;
    double* er = I.er + (size_t)(act ? e : 0) * ER_SZ;
    const double* dc = I.dc + LAM_EDGE(act ? e : 0);
    if (act) {
      for (int i = c.gl; i < NY * TILE_LD; i += GSZ) T[i] = 0.0;
      // stage the bulk data of this edge (coalesced, independent loads)
      DMB_COPY_BATCH(DMB_NNZ_H, c.gl, GSZ, eb[EB_HNZ + idx], T[STG_HNZ + idx] = ls_mode ? 0.0 : val);
      DMB_COPY_BATCH(NQ, c.gl, GSZ, eb[EB_GOBJ + idx], T[STG_GOBJ + idx] = val);
#if NNL > 0
      for (int i = c.gl; i < DMB_NNZ_D; i += GSZ) T[STG_DNZ + i] = eb[EB_DNZ + i];
#endif
    }
    WARP_SYNC();
    if (act) {
      // Jacobian nonzeros (row scaled) -> tile columns [G_y | G_vp]; right-hand side -res
      DMB_COPY_BATCH(DMB_NNZ_J, c.gl, GSZ, eb[EB_JNZ + idx],
                     const int r = DMB_J_ROW[idx]; const int cq = DMB_J_COL[idx]; T[r * TILE_LD + (cq >= NVP ? cq - NVP : NY + cq)] = dc[r] * val);
      for (int r = c.gl; r < NY; r += GSZ) {
        T[r * TILE_LD + NY + NVP] = ls_mode ? 0.0 : -res[LAM_EDGE(e) + r];
        // barrier terms of the edge unknowns y
        const int i = EDGE_OFF(e) + r;
        const double lo = I.lb[i], hi = I.ub[i], wy = I.w[i];
        const bool hl = lo > -DMB_INF, hu = hi < DMB_INF;
        double sig = 0.0, gb = 0.0;
        if (ls_mode) {
          sig = DMB_LSW(i);
          if (hl) gb -= I.zL[i];
          if (hu) gb += I.zU[i];
        } else if (I.resto) resto_terms(I.zeta, mu, wy, I.go[i], lo, hi, hl, hu, sig, gb);
        else {
          if (hl) { const double rs = 1.0 / (wy - lo); sig += I.zL[i] * rs; gb -= mu * rs; if (!hu) gb += kappa_d * mu; }
          if (hu) { const double rs = 1.0 / (hi - wy); sig += I.zU[i] * rs; gb += mu * rs; if (!hl) gb -= kappa_d * mu; }
        }
        T[STG_SIGY + r] = sig;
        T[STG_GY + r] = T[STG_GOBJ + NVP + r] + gb;
        eb[EB_SIGY + r] = sig;
        eb[EB_GY + r] = T[STG_GY + r];
      }
#if NNL > 0
      for (int i = c.gl; i < NNL; i += GSZ) {
        const double sl = nl_ub_relaxed(I, e, i) - I.s[e * NNL + i];
        double sig = ls_mode ? 1.0 : I.vU[e * NNL + i] / sl;
        // r_d = Sigma_s (d - s) + grad_s phi ; grad_s phi = mu/sl - kappa_d*mu (upper bound only)
        double rd = ls_mode ? I.vU[e * NNL + i] : sig * res[LAM_EDGE(e) + NY + NX + i] + mu / sl - kappa_d * mu;
        if (I.resto && !ls_mode) {
          double gs;
          resto_terms(I.zeta, mu, I.s[e * NNL + i], I.s2[e * NNL + i], 0.0, nl_ub_relaxed(I, e, i), false, true, sig, gs);
          rd = sig * res[LAM_EDGE(e) + NY + NX + i] + gs;
        }
        T[STG_SIGS + i] = sig; T[STG_RD + i] = rd;
        eb[EB_SIGS + i] = sig; eb[EB_RD + i] = rd;
      }
#endif
    }
    WARP_SYNC();
    SUB_TICK(I, 0);
    // ---- LU with partial pivoting; columns [NY, TILE_NC) ride along.  The diagonal keeps the RECIPROCAL pivot. ------
    // (measured and rejected, twice: the lane's row held in registers for the whole elimination, only the pivot row going
    //  through shared memory — one load per multiply-add instead of two loads and a store, but fully unrolled over the pivot
    //  steps.  Round 2, cfg #2 on B200: 24.88 ms against 23.97 ms for this loop, identical iterates; profiles/r2_reg_lu_ab.json)
    for (int k = 0; k < NY; ++k) {
      double best = -1.0; int myr = -1;
      if (act) for (int r = c.gl; r < NY; r += GSZ) if (r >= k) { const double a = fabs(T[r * TILE_LD + k]); if (a > best) { best = a; myr = r; } }
      int pr = group_argmax(c, best, myr);
      if (pr < k || pr >= NY) pr = k;        // a column of NaNs has no candidate: keep row k (the pivot test below fails the step)
      if (act && pr != k) {
        for (int cc = c.gl; cc < TILE_NC; cc += GSZ) {
          const double t = T[k * TILE_LD + cc]; T[k * TILE_LD + cc] = T[pr * TILE_LD + cc]; T[pr * TILE_LD + cc] = t;
        }
      }
      if (act && c.gl == 0) {
        if (k == 0) for (int i = 0; i < NY; ++i) T[STG_PERM + i] = (double)i;
        const double t = T[STG_PERM + k]; T[STG_PERM + k] = T[STG_PERM + pr]; T[STG_PERM + pr] = t;
      }
      WARP_SYNC();
#if DMB_LU_FLAT
      // multipliers first, then the rank-1 update with its (NY-1-k) x (TILE_NC-1-k) entries dealt out over ALL lanes of the
      // group: with one lane per row the lanes of the rows <= k idle (on average half of the NY active ones, and GSZ - NY never
      // work).  Same operation per entry as the row loop (t -= l * u), so the factors are bit-identical.
      if (act) {
        const double pv = T[k * TILE_LD + k];
        if (!(fabs(pv) > 1e-300)) *fail = 1;
        const double inv = 1.0 / pv;
        for (int r = k + 1 + c.gl; r < NY; r += GSZ) T[r * TILE_LD + k] *= inv;
        if (c.gl == 0) T[STG_INVD + k] = inv;
      }
      WARP_SYNC();
      if (act) {
        const int nc = TILE_NC - 1 - k, n = (NY - 1 - k) * nc;
        const float rnc = 1.0f / (float)nc;                 // idx / nc through the reciprocal (exact for these small integers)
        const double* urow = T + k * TILE_LD + k + 1;
#pragma unroll 4
        for (int idx = c.gl; idx < n; idx += GSZ) {
          const int r = (int)(((float)idx + 0.5f) * rnc), cc = idx - r * nc;
          double* row = T + (k + 1 + r) * TILE_LD + k;
          row[1 + cc] -= row[0] * urow[cc];
        }
      }
      WARP_SYNC();
    }
#else
      if (act) {
        const double pv = T[k * TILE_LD + k];
        if (!(fabs(pv) > 1e-300)) *fail = 1;
        const double inv = 1.0 / pv;
        for (int r = c.gl; r < NY; r += GSZ) if (r > k) {
          const double l = T[r * TILE_LD + k] * inv;
          T[r * TILE_LD + k] = l;
#pragma unroll 6
          for (int cc = k + 1; cc < TILE_NC; ++cc) T[r * TILE_LD + cc] -= l * T[k * TILE_LD + cc];
        }
        if (c.gl == 0) T[STG_INVD + k] = inv;
      }
      WARP_SYNC();
    }
#endif
    SUB_TICK(I, 1);
#if NY > 16 && DMB_BS_FLAT
    // ---- back substitution as a column sweep over ALL lanes of the group (one lane per right-hand side keeps NE of the 32
    //      lanes busy): x_k = b_k / u_kk, then b_r -= u_rk x_k for the k * NE entries above; then Yext = [-X | x_rhs] ----------
    if (act) for (int i = c.gl; i < NY; i += GSZ) eb[EB_PIV + i] = T[STG_PERM + i];
    for (int k = NY - 1; k >= 0; --k) {
      if (act) for (int j = c.gl; j < NE; j += GSZ) T[k * TILE_LD + NY + j] *= T[STG_INVD + k];
      WARP_SYNC();
      if (act) {
        const double* xk = T + k * TILE_LD + NY;
#pragma unroll 4
        for (int idx = c.gl; idx < k * NE; idx += GSZ) {
          const int r = idx / NE, j = idx - r * NE;
          T[r * TILE_LD + NY + j] -= T[r * TILE_LD + k] * xk[j];
        }
      }
      WARP_SYNC();
    }
    if (act) for (int idx = c.gl; idx < NY * NVP; idx += GSZ) { const int r = idx / NVP, j = idx - r * NVP; T[r * TILE_LD + NY + j] = -T[r * TILE_LD + NY + j]; }
#else
    if (act) {
      for (int i = c.gl; i < NY; i += GSZ) eb[EB_PIV + i] = T[STG_PERM + i];
      // ---- back substitution, one lane per right-hand side; then Yext = [-X | x_rhs] --------------------------------
      for (int j = c.gl; j < NE; j += GSZ) {
#if NY <= 16
        // the lane's solution column stays in registers (fully unrolled: static indices); same operation order as the loop below
        double xv[NY];
#pragma unroll
        for (int k = NY - 1; k >= 0; --k) {
          double x = T[k * TILE_LD + NY + j];
#pragma unroll
          for (int cc = k + 1; cc < NY; ++cc) x -= T[k * TILE_LD + cc] * xv[cc];
          xv[k] = x * T[STG_INVD + k];
        }
        const double sgn = j < NVP ? -1.0 : 1.0;
#pragma unroll
        for (int k = 0; k < NY; ++k) T[k * TILE_LD + NY + j] = sgn * xv[k];
#else
        for (int k = NY - 1; k >= 0; --k) {
          double x = T[k * TILE_LD + NY + j];
          for (int cc = k + 1; cc < NY; ++cc) x -= T[k * TILE_LD + cc] * T[cc * TILE_LD + NY + j];
          T[k * TILE_LD + NY + j] = x * T[STG_INVD + k];
        }
        if (j < NVP) for (int k = 0; k < NY; ++k) T[k * TILE_LD + NY + j] = -T[k * TILE_LD + NY + j];
#endif
      }
    }
#endif
    WARP_SYNC();
    SUB_TICK(I, 2);
    if (act) {
      // tile image [LU (reciprocal pivots on the diagonal) | Yext] -> slot
      for (int i = c.gl; i < NY; i += GSZ) T[i * TILE_LD + i] = T[STG_INVD + i];
    }
    WARP_SYNC();
    if (act) {
#if DMB_TMA
      if (c.gl == 0) tma_store_1d(eb + EB_IMG, T, IMG_SZ * 8);
#else
      for (int i = c.gl; i < IMG_SZ; i += GSZ) eb[EB_IMG + i] = T[i];
#endif
    }
    WARP_SYNC();
    if (act) {
#if DMB_DMMA_FACTOR
    }
    // ---- projection products on the FP64 tensor cores (DMMA.8x8x4) ---------------------------------------------------------
    // dense H_d = H_yy + Sigma_y in the LU area of the tile (the factors are already saved to the slot), N = H_vy Yext stays
    // sparse / scalar; then all 32 lanes of the warp work on one tile at a time:
    //   M = H_d Yext (+ g_y in the right-hand-side column),   Y'M  and  Y'Yext  -> reduced Hessian / gradient blocks
    if (act) for (int i = c.gl; i < NY * NY; i += GSZ) T[(i / NY) * TILE_LD + i % NY] = 0.0;
    WARP_SYNC();
    if (act) {
      for (int k = c.gl; k < DMB_NNZ_H; k += GSZ) {
        const int r = DMB_H_ROW[k], cq = DMB_H_COL[k];
        if (r >= NVP && cq >= NVP) { const double v = T[STG_HNZ + k]; T[(r - NVP) * TILE_LD + cq - NVP] = v; T[(cq - NVP) * TILE_LD + r - NVP] = v; }
      }
      for (int idx = c.gl; idx < NVP * NE; idx += GSZ) {
        const int rq = idx / NE, a = idx - rq * NE;
        double s = 0.0;
        for (int t = DMB_HR_START[rq]; t < DMB_HR_START[rq + 1]; ++t) {
          const int cq = DMB_HR_COL[t];
          if (cq >= NVP) s += T[STG_HNZ + DMB_HR_IDX[t]] * T[(cq - NVP) * TILE_LD + NY + a];
        }
        T[STG_N + rq * NE + a] = s;
      }
    }
    WARP_SYNC();
    if (act) for (int i = c.gl; i < NY; i += GSZ) T[i * TILE_LD + i] += T[STG_SIGY + i];
    WARP_SYNC();
    SUB_TICK(I, 3);
    {
      constexpr int GPW = 32 / GSZ;                               // lane groups (= tiles) per warp
      const int fr = c.lane >> 2, fq = c.lane & 3;
#pragma unroll 1
      for (int g = 0; g < GPW; ++g) {
        const int gid = (c.grp / GPW) * GPW + g;
        const int eg = e0 + gid;
        if (eg >= NEDGE) continue;                                // uniform over the warp
        double* Tg = tile_base + (size_t)gid * TILE_SZ;
        double* erg = I.er + (size_t)eg * ER_SZ;
        // M = H_d Yext
#pragma unroll 1
        for (int mt = 0; mt < (NY + 7) / 8; ++mt)
#pragma unroll 1
          for (int nt = 0; nt < (NE + 7) / 8; ++nt) {
            double c0 = 0.0, c1 = 0.0;
            const int row = mt * 8 + fr, col = nt * 8 + fr;
#pragma unroll
            for (int kk = 0; kk < (NY + 3) / 4; ++kk) {
              const int k = kk * 4 + fq;
              const double av = (row < NY && k < NY) ? Tg[row * TILE_LD + k] : 0.0;
              const double bv = (k < NY && col < NE) ? Tg[k * TILE_LD + NY + col] : 0.0;
              dmma884(c0, c1, av, bv);
            }
            const int oc = nt * 8 + 2 * fq;
            if (row < NY) {
              if (oc < NE) Tg[row * TILE_LD + TILE_M + oc] = c0 + (oc == NVP ? Tg[STG_GY + row] : 0.0);
              if (oc + 1 < NE) Tg[row * TILE_LD + TILE_M + oc + 1] = c1 + (oc + 1 == NVP ? Tg[STG_GY + row] : 0.0);
            }
          }
        __syncwarp();
        // Y'[M | g_y + M y0]  and  Y'Yext
#pragma unroll 1
        for (int mt = 0; mt < (NVP + 7) / 8; ++mt)
#pragma unroll 1
          for (int nt = 0; nt < (NE + 7) / 8; ++nt) {
            double h0 = 0.0, h1 = 0.0, y0 = 0.0, y1 = 0.0;
            const int m = mt * 8 + fr, n = nt * 8 + fr;
#pragma unroll
            for (int kk = 0; kk < (NY + 3) / 4; ++kk) {
              const int k = kk * 4 + fq;
              const double av = (k < NY && m < NVP) ? Tg[k * TILE_LD + NY + m] : 0.0;
              const double bm = (k < NY && n < NE) ? Tg[k * TILE_LD + TILE_M + n] : 0.0;
              const double by = (k < NY && n < NE) ? Tg[k * TILE_LD + NY + n] : 0.0;
              dmma884(h0, h1, av, bm); dmma884(y0, y1, av, by);
            }
            if (m < NVP) {
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const int b = nt * 8 + 2 * fq + hh;
                const double hv = hh ? h1 : h0, yv = hh ? y1 : y0;
                if (b < NVP) {
                  double h = Tg[STG_N + m * NE + b] + Tg[STG_N + b * NE + m];
                  for (int t = DMB_HR_START[m]; t < DMB_HR_START[m + 1]; ++t) if (DMB_HR_COL[t] == b) h += Tg[STG_HNZ + DMB_HR_IDX[t]];
                  erg[ER_HH + m * NVP + b] = h + hv; erg[ER_YTY + m * NVP + b] = yv;
                } else if (b == NVP) {
                  erg[ER_GH + m] = Tg[STG_GOBJ + m] + Tg[STG_N + m * NE + NVP] + hv; erg[ER_YTY0 + m] = yv;
                }
              }
            }
          }
        __syncwarp();
      }
    }
    if (act) {
#else
      // ---- M = (H_yy + Sigma_y) Yext  (rows NY, cols NE) and N = H_vy Yext (rows NVP) via the CSR rows of H ----------
      // DMB_EMU_FP32_PROJ (experiment, tools/plan_a_experiment.py): the projection products with FP32 operands and FP32
      // accumulation — the accuracy a TF32x3-split tcgen05 product reaches at best — to measure what the FP64 refinement
      // loop can absorb (north_star plan A).  Never defined in product builds.
#ifdef DMB_EMU_FP32_PROJ
#define DMB_PACC float
#define DMB_PMUL(x, y) ((float)(x) * (float)(y))
#else
#define DMB_PACC double
#define DMB_PMUL(x, y) ((x) * (y))
#endif
#if !defined(DMB_EMU_FP32_PROJ)
      // one lane per column of Yext: the column in registers, the sparse rows as generated straight-line code (gen_hy_products)
      for (int a = c.gl; a < NE; a += GSZ) {
        double yc[NY], mc[NY], nc[NVP];
#pragma unroll
        for (int j = 0; j < NY; ++j) yc[j] = T[j * TILE_LD + NY + a];
        gen_hy_products(T + STG_HNZ, T + STG_SIGY, yc, mc, nc);
#pragma unroll
        for (int j = 0; j < NY; ++j) T[j * TILE_LD + TILE_M + a] = mc[j];
#pragma unroll
        for (int i = 0; i < NVP; ++i) T[STG_N + i * NE + a] = nc[i];
      }
#else
      for (int idx = c.gl; idx < NQ * NE; idx += GSZ) {
        const int rq = idx / NE, a = idx - rq * NE;            // rq < NVP: N row, else M row
        DMB_PACC s = 0.0;
        for (int t = DMB_HR_START[rq]; t < DMB_HR_START[rq + 1]; ++t) {
          const int cq = DMB_HR_COL[t];
          if (cq >= NVP) s += DMB_PMUL(T[STG_HNZ + DMB_HR_IDX[t]], T[(cq - NVP) * TILE_LD + NY + a]);
        }
        if (rq >= NVP) {
          s += DMB_PMUL(T[STG_SIGY + rq - NVP], T[(rq - NVP) * TILE_LD + NY + a]);
          T[(rq - NVP) * TILE_LD + TILE_M + a] = s;
        } else T[STG_N + rq * NE + a] = s;
      }
#endif
    }
    WARP_SYNC();
    SUB_TICK(I, 3);
    if (act) {
      // ---- reduced blocks ---------------------------------------------------------------------------------------------
      for (int idx = c.gl; idx < NVP * NVP; idx += GSZ) {
        const int a = idx / NVP, b = idx - a * NVP;
        double h = T[STG_N + a * NE + b] + T[STG_N + b * NE + a], yty = 0.0;
        for (int t = DMB_HR_START[a]; t < DMB_HR_START[a + 1]; ++t) if (DMB_HR_COL[t] == b) h += T[STG_HNZ + DMB_HR_IDX[t]];
        DMB_PACC hp = 0.0;
#pragma unroll 4
        for (int j = 0; j < NY; ++j) {
          const double yja = T[j * TILE_LD + NY + a];
          hp += DMB_PMUL(yja, T[j * TILE_LD + TILE_M + b]);
          yty += yja * T[j * TILE_LD + NY + b];
        }
        h += hp;
        er[ER_HH + idx] = h; er[ER_YTY + idx] = yty;
      }
      for (int a = c.gl; a < NVP; a += GSZ) {
        double g = T[STG_GOBJ + a] + T[STG_N + a * NE + NVP], yy = 0.0;
#pragma unroll 4
        for (int j = 0; j < NY; ++j) {
          const double yja = T[j * TILE_LD + NY + a];
          g += yja * (T[STG_GY + j] + T[j * TILE_LD + TILE_M + NVP]);
          yy += yja * T[j * TILE_LD + NY + NVP];
        }
        er[ER_GH + a] = g; er[ER_YTY0 + a] = yy;
      }
#endif
      for (int idx = c.gl; idx < NX * NVP; idx += GSZ) { const int r = idx / NVP; er[ER_AB + idx] = T[DMB_XF_SEL[r] * TILE_LD + NY + idx - r * NVP]; }
      for (int i = c.gl; i < NX; i += GSZ)
        er[ER_B + i] = T[DMB_XF_SEL[i] * TILE_LD + NY + NVP] + (ls_mode ? 0.0 : res[LAM_EDGE(e) + NY + i]);
    }
    WARP_SYNC();
    SUB_TICK(I, 4);
#if NNL > 0
    if (act) {
      // condensed inequality slacks: H += sum_i sig_i d_i d_i', g += sum_i d_i r_i, projected with t_i = T' d_i
      for (int i = 0; i < NNL; ++i) {
        double* tv = T + STG_DQ;                               // NE entries of scratch
        for (int a = c.gl; a < NE; a += GSZ) {
          double s = 0.0;
          for (int t = DMB_DR_START[i]; t < DMB_DR_START[i + 1]; ++t) {
            const int cq = DMB_D_COL[t];
            if (cq >= NVP) s += T[STG_DNZ + t] * T[(cq - NVP) * TILE_LD + NY + a];
            else if (cq == a) s += T[STG_DNZ + t];
          }
          tv[a] = s;
        }
        WARP_SYNC();
        const double sig = T[STG_SIGS + i], rd = T[STG_RD + i] + sig * tv[NVP];
        // the inertia correction acts on the slacks as well (IPOPT: delta_s = delta_x): (Sigma_s + delta) takes the place of
        // Sigma_s, so the delta-proportional parts join Y'Y / Y'y0, which the node assembly scales with delta
        const double dms = (ls_mode ? 0.0 : res[LAM_EDGE(e) + NY + NX + i]) + tv[NVP];
        for (int idx = c.gl; idx < NVP * NVP; idx += GSZ) {
          const double tt = tv[idx / NVP] * tv[idx % NVP];
          er[ER_HH + idx] += sig * tt; er[ER_YTY + idx] += tt;
        }
        for (int a = c.gl; a < NVP; a += GSZ) { er[ER_GH + a] += tv[a] * rd; er[ER_YTY0 + a] += tv[a] * dms; }
        WARP_SYNC();
      }
    }
    WARP_SYNC();
#endif
  }
#if DMB_TMA
  tma_store_drain();                 // the tile images are in the slot before any group loads them back
#endif
  CTA_SYNC();
}


// ---------------------------------------------------------------------------------------------------
// Riccati sweep over the scenario tree.  v-space ordering: (x [NX], u_prev [NU], u [NU], eps [NEPS]).
//  (1) all threads: node matrices Q0 = node terms (Sigma + delta, barrier gradient, rterm, mterm) + the
//      projected edge Hessians of the children — everything that does not depend on a child's value function;
//  (2) warp 0, backwards over the nodes, everything in shared memory, loops fully unrolled:
//        W = P_c [A B]_aug , Q += [A B]_aug' W , Cholesky of Q_uu , K = -Q_uu^{-1} Q_u. , P = Q_xx + Q_xu K (in place).
// *flag = 1 if a control block is not positive definite (wrong inertia).
// ---------------------------------------------------------------------------------------------------
// Scenario trees: the nodes of one stage are independent of each other in both sweeps, so every warp takes some of them and
// the CTA synchronises once per stage (robust configs: 9 nodes per stage, ~180 nodes — a serial sweep on one warp dominated
// those problems).  Chains (one node per stage) keep the single-warp sweep without barriers.  Compile-time switch.
#if DMB_NSCEN > 1
#define DMB_LVL(l) DMB_LEVEL_START[l]     // first node of stage l
#else
#define DMB_LVL(l) (l)                    // chain: node l is stage l
#endif
#if DMB_NSCEN > 1 && !defined(DMB_HOSTSIM)
#define DMB_TREE_PAR 1
#define DMB_FLAG_OR(flag, bits) atomicOr((flag), (bits))
#else
#define DMB_TREE_PAR 0
#define DMB_FLAG_OR(flag, bits) (*(flag) |= (bits))
#endif

// lower-triangle index -> (row i >= column j)
#define NTRI_VP (NVP * (NVP + 1) / 2)
#define NTRI_XI (NXI * (NXI + 1) / 2)
DMB_DEV void tri_decode(int idx, int& i, int& j) {
  int r = 0;
  while ((r + 1) * (r + 2) / 2 <= idx) ++r;
  i = r; j = idx - r * (r + 1) / 2;
}

// unpivoted L D L' of the control block (row stride NV): strict lower part of L (unit diagonal), diagonal = 1 / D.
// returns bit 0: zero / non-finite pivot (replaced by 1), bit 1: negative pivot.
DMB_DEV int quu_ldl(const double* Quu, double* L) {
  int bad = 0;
#pragma unroll
  for (int j = 0; j < NUT; ++j) {
    double d = Quu[j * NV + j];
#pragma unroll
    for (int t = 0; t < j; ++t) d -= L[j * NUT + t] * L[j * NUT + t] / L[t * NUT + t];
    if (!(fabs(d) > 1e-300) || !(fabs(d) < 1e300)) { bad |= 1; d = 1.0; }
    if (d < 0.0) bad |= 2;
    const double dinv = 1.0 / d;
    L[j * NUT + j] = dinv;
#pragma unroll
    for (int i = j + 1; i < NUT; ++i) {
      double s = Quu[i * NV + j];
#pragma unroll
      for (int t = 0; t < j; ++t) s -= L[i * NUT + t] * L[j * NUT + t] / L[t * NUT + t];
      L[i * NUT + j] = s * dinv;
    }
  }
  return bad;
}
DMB_DEV void quu_ldl_solve(const double* L, double* rhs) {
#pragma unroll
  for (int a = 0; a < NUT; ++a) { double s = rhs[a]; for (int t = 0; t < a; ++t) s -= L[a * NUT + t] * rhs[t]; rhs[a] = s; }
#pragma unroll
  for (int a = 0; a < NUT; ++a) rhs[a] *= L[a * NUT + a];
#pragma unroll
  for (int a = NUT - 1; a >= 0; --a) { double s = rhs[a]; for (int t = a + 1; t < NUT; ++t) s -= L[t * NUT + a] * rhs[t]; rhs[a] = s; }
}

// LDL = false: Cholesky of every control block, *flag = 1 unless all are positive definite (inertia test of the IPM).
// LDL = true : unpivoted L D L' instead, valid for indefinite blocks (sensitivity solve, ipm_sens.cuh: the linear system
//              only has to be regular); *flag bit 0 = singular block, bit 1 = some D < 0 (reduced Hessian indefinite).
//              The "L" returned through the in-register layout then has unit diagonal and carries 1/D on the diagonal.
template <bool LDL>
DMB_PHASE void riccati_backward_t(const Ctx& c, const Inst& I, double mu, double kappa_d, double delta, int ls_mode, double* Qs, int* flag) {
  CTA_SYNC();                       // every thread has consumed the flag of a previous call before it is reset
  if (c.tid == 0) *flag = 0;
  // ---- (1) node matrices ------------------------------------------------------------------------------------------
  // Two kinds of work items, so that the bound terms (five dependent-free global loads, a division each) are evaluated ONCE
  // per variable instead of once for the diagonal and once for the gradient entry:
  //   items [0, NNODE * NV)            : v-position a of node n -> diagonal entry Q[a][a] and gradient entry q[a]
  //   items [NNODE * NV, ... + NV*NV)  : off-diagonal entries
  for (int idx = c.tid; idx < NNODE * (NV + NV * NV); idx += c.nthr) {
    const bool diag = idx < NNODE * NV;
    const int rem = diag ? idx : idx - NNODE * NV;
    const int n = diag ? rem / NV : rem / (NV * NV);
    const int a = diag ? rem - n * NV : (rem - n * NV * NV) / NV;
    const int b = diag ? a : (rem - n * NV * NV) - a * NV;
    if (!diag && a == b) continue;
    double* wn = I.node + (size_t)n * WN_NODE;
    const int k = NODE_K_OF(n);
    // edge-local index of v-space position a / b (u_prev has none)
    const int ia = a < NX ? a : (a >= NXI ? a - NU : -1), ib = b < NX ? b : (b >= NXI ? b - NU : -1);
    double val = 0.0, gq = 0.0;                                  // matrix entry (a, b); gradient entry a (diag items only)
    if (diag && ia >= 0 && !(k == NHOR && ia >= NX)) {
      const int gi = NODE_OFF(n) + ia;
      const double lo = I.lb[gi], hi = I.ub[gi], wx = I.w[gi];
      const bool hl = lo > -DMB_INF, hu = hi < DMB_INF;
      double sig = delta, gb = 0.0;
      if (ls_mode) {
        sig = DMB_LSW(gi);
        if (hl) gb -= I.zL[gi];
        if (hu) gb += I.zU[gi];
      } else if (I.resto) { resto_terms(I.zeta, mu, wx, I.go[gi], lo, hi, hl, hu, sig, gb); sig += delta; }
      else {
        if (hl) { const double sl = wx - lo; sig += I.zL[gi] / sl; gb -= mu / sl; if (!hu) gb += kappa_d * mu; }
        if (hu) { const double su = hi - wx; sig += I.zU[gi] / su; gb += mu / su; if (!hl) gb -= kappa_d * mu; }
      }
      val = sig; gq = gb;
    }
    if (k == NHOR) {
      if (a < NX && b < NX) {
        const int pe = NODE_PARENT_EDGE(n);
        const double wgt = I.df * EDGE_OMEGA_OF(pe);
        if (diag) { double gm[NX]; gen_mterm_grad(I.w + NODE_OFF(n), p_tvp_at(I.p, NHOR), edge_pp(I.p, pe), gm); gq += wgt * gm[a]; }
        if (!ls_mode) { double hm[NX * NX]; gen_mterm_hess(I.w + NODE_OFF(n), p_tvp_at(I.p, NHOR), edge_pp(I.p, pe), hm); val += wgt * hm[a * NX + b]; }
      }
    } else {
#if DMB_HAS_ARRIVAL
      if (n == 0 && a < NX && b < NX) {
        if (diag) { double ga[NX]; gen_arrival_grad(I.w + NODE_OFF(0), p_tvp_at(I.p, 0), I.p + DMB_P_P, ga); gq += I.df * ga[a]; }
        if (!ls_mode) { double ha[NX * NX]; gen_arrival_hess(I.w + NODE_OFF(0), p_tvp_at(I.p, 0), I.p + DMB_P_P, ha); val += I.df * ha[a * NX + b]; }
      }
#endif
      // rterm couples u (v-pos NXI..) and u_prev (v-pos NX..)
      const int ra = (a >= NXI && a < NXI + NU) ? a - NXI : ((a >= NX && a < NXI) ? NU + a - NX : -1);
      if (diag && ra >= 0) gq += wn[WN_RG + ra];
      if (!ls_mode) {
        const int rb = (b >= NXI && b < NXI + NU) ? b - NXI : ((b >= NX && b < NXI) ? NU + b - NX : -1);
        if (ra >= 0 && rb >= 0) val += wn[WN_RH + ra * 2 * NU + rb];
      }
      if (ia >= 0)
        for (int j = 0; j < NODE_NEDGE(n); ++j) {
          const double* eb = I.er + (size_t)(NODE_FIRST_EDGE(n) + j) * ER_SZ;
          if (diag) gq += eb[ER_GH + ia] + delta * eb[ER_YTY0 + ia];
          if (ib >= 0) val += eb[ER_HH + ia * NVP + ib] + delta * eb[ER_YTY + ia * NVP + ib];
        }
    }
    wn[WN_Q + a * NV + b] = val;
    if (diag) wn[WN_QV + a] = gq;
  }
  CTA_SYNC();
  SUB_START();
  // ---- (2) sweep ----------------------------------------------------------------------------------------------------------
  const bool par = DMB_TREE_PAR && c.nwarp > 1;
  {
    double* W = Qs + (par ? c.warp : 0) * DMB_QS_WARP;   // NXI x NVP : P_c [A B]_aug restricted to the (x, u, eps) columns
    double* pb = W + NXI * NVP;                          // NXI       : p_c + P_c[:, x] b
    int tvi, tvj, txi, txj;          // this lane's first lower-triangle entry of the two symmetric updates (loop invariant)
    tri_decode(c.lane, tvi, tvj); tri_decode(c.lane, txi, txj);
    for (int lvl = NHOR - 1; lvl >= 0; --lvl) {
     if (par || c.warp == 0)
     for (int n = DMB_LVL(lvl + 1) - 1 - (par ? c.warp : 0); n >= DMB_LVL(lvl); n -= (par ? c.nwarp : 1)) {
      double* wn = I.node + (size_t)n * WN_NODE;
#if DMB_DMMA_SWEEP
      if (!LDL) {
        // ---- tensor-core sweep step (FP64 DMMA), everything in the padded 8 x 8 v-space (x, u_prev, u, eps) -------------------
        //   F  : child's xi = F v  (rows x_c = [A B] (x,u,eps), rows u_prev_c = u)          T1 = P_c F ,  Q = Q0 + F' T1
        //   L  = Q_xi,u chol(Q_uu)^-T ,  P = Q_xi,xi - L L'  (one DMMA with k = NUT <= 4) ,  K = -chol^-T L'
        const int fr = c.lane >> 2, fq = c.lane & 3;
        double q0 = (fr < NV && 2 * fq < NV) ? wn[WN_Q + fr * NV + 2 * fq] : 0.0;
        double q1 = (fr < NV && 2 * fq + 1 < NV) ? wn[WN_Q + fr * NV + 2 * fq + 1] : 0.0;
        double* pbs = Qs + (par ? c.warp : 0) * DMB_QS_WARP;              // NXI : p_c + P_c[:, x] b
        for (int j = 0; j < NODE_NEDGE(n); ++j) {
          const int e = NODE_FIRST_EDGE(n) + j;
          const double* er = I.er + (size_t)e * ER_SZ;
          const double* wc = I.node + (size_t)EDGE_CHILD_OF(e) * WN_NODE;
          double fa[2], pa[2];
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const int k = kk * 4 + fq, m = fr;                                // F[k][m]
            const int ia = m < NX ? m : ((m >= NXI && m < NV) ? m - NU : -1);
            fa[kk] = k < NX ? (ia >= 0 ? er[ER_AB + k * NVP + ia] : 0.0) : ((k < NXI && m == NXI + k - NX) ? 1.0 : 0.0);
            pa[kk] = (fr < NXI && k < NXI) ? P_AT(wc, fr, k) : 0.0;          // P_c[fr][k]
          }
          if (c.lane < NXI) {
            double sv = wc[WN_QV + c.lane];
#pragma unroll
            for (int t = 0; t < NX; ++t) sv += P_AT(wc, c.lane, t) * er[ER_B + t];
            pbs[c.lane] = sv;
          }
          double t0 = 0.0, t1 = 0.0;
          dmma884(t0, t1, pa[0], fa[0]); dmma884(t0, t1, pa[1], fa[1]);      // T1 = P_c F
          const double tb0 = frag_get(t0, t1, fq, fr), tb1 = frag_get(t0, t1, 4 + fq, fr);   // B fragments: T1[k][fr]
          dmma884(q0, q1, fa[0], tb0); dmma884(q0, q1, fa[1], tb1);          // Q += F' T1
          WARP_SYNC();
          if (c.lane < NV) {                                                  // q += F' (p_c + P_c b)
            const int m = c.lane;
            const int ia = m < NX ? m : (m >= NXI ? m - NU : -1);
            double sv = (m >= NXI && m < NXI + NU) ? pbs[NX + m - NXI] : 0.0;
            if (ia >= 0) {
#pragma unroll
              for (int k = 0; k < NX; ++k) sv += er[ER_AB + k * NVP + ia] * pbs[k];
            }
            wn[WN_QV + m] += sv;
          }
          WARP_SYNC();
        }
        // complete Q: the control columns / rows stay in shared memory for the re-solves (riccati_vector, forward roll-out)
        if (fr < NV) {
          if (2 * fq >= NXI && 2 * fq < NV) wn[WN_Q + fr * NV + 2 * fq] = q0;
          if (2 * fq + 1 >= NXI && 2 * fq + 1 < NV) wn[WN_Q + fr * NV + 2 * fq + 1] = q1;
        }
        // Cholesky of Q_uu, redundantly in every lane (NUT <= 4); diagonal keeps the reciprocal
        double C[NUT * NUT];
        int bad = 0;
#pragma unroll
        for (int a = 0; a < NUT; ++a)
#pragma unroll
          for (int b = 0; b <= a; ++b) {
            double sv = frag_get(q0, q1, NXI + a, NXI + b);
#pragma unroll
            for (int t = 0; t < b; ++t) sv -= C[a * NUT + t] * C[b * NUT + t];
            if (a == b) { if (!(sv > 0.0)) { bad = 1; sv = 1.0; } C[a * NUT + a] = rsqrt(sv); }
            else C[a * NUT + b] = sv * C[b * NUT + b];
          }
        if (bad && c.lane == 0) DMB_FLAG_OR(flag, bad);
        // L[fr][:] = Q[fr][u] chol^-T (forward substitution along the row; all lanes of a row quad compute the same values)
        double Lr[NUT];
#pragma unroll
        for (int a = 0; a < NUT; ++a) {
          double sv = frag_get(q0, q1, fr, NXI + a);
#pragma unroll
          for (int t = 0; t < a; ++t) sv -= Lr[t] * C[a * NUT + t];
          Lr[a] = sv * C[a * NUT + a];
        }
        // vector part: l = chol^-1 q_u ;  k = -chol^-T l ;  p = q_xi - L l
        WARP_SYNC();
        double lv[NUT], kv[NUT];
#pragma unroll
        for (int a = 0; a < NUT; ++a) {
          double sv = wn[WN_QV + NXI + a];
#pragma unroll
          for (int t = 0; t < a; ++t) sv -= lv[t] * C[a * NUT + t];
          lv[a] = sv * C[a * NUT + a];
        }
#pragma unroll
        for (int a = NUT - 1; a >= 0; --a) {
          double sv = lv[a];
#pragma unroll
          for (int t = a + 1; t < NUT; ++t) sv -= C[t * NUT + a] * kv[t];
          kv[a] = sv * C[a * NUT + a];
        }
        // gains K[:, fr] = -chol^-T L[fr][:]' (one lane per row quad writes), k
        if (fq == 0 && fr < NXI) {
          double z[NUT];
#pragma unroll
          for (int a = NUT - 1; a >= 0; --a) {
            double sv = Lr[a];
#pragma unroll
            for (int t = a + 1; t < NUT; ++t) sv -= C[t * NUT + a] * z[t];
            z[a] = sv * C[a * NUT + a];
          }
          double pv = wn[WN_QV + fr];
#pragma unroll
          for (int a = 0; a < NUT; ++a) { wn[WN_K + a * NXI + fr] = -z[a]; pv -= Lr[a] * lv[a]; }
          wn[WN_QV + fr] = pv;
        }
        if (c.lane < NUT) wn[WN_KV + c.lane] = -kv[c.lane];
        // P = Q_xi,xi - L L' : one DMMA (A[fr][fq] = -L[fr][fq], B[fq][fr] = L[fr][fq]); rows / columns beyond xi are not used
        double la = 0.0;
#pragma unroll
        for (int a = 0; a < NUT; ++a) if (fq == a && fr < NXI) la = Lr[a];
        dmma884(q0, q1, -la, la);
        // exact symmetry (the recursion amplifies rounding asymmetries on badly scaled problems): mirror the lower triangle
        {
          const double m0 = frag_get(q0, q1, 2 * fq, fr), m1 = frag_get(q0, q1, (2 * fq + 1) & 7, fr);
          if (2 * fq > fr) q0 = m0;
          if (2 * fq + 1 > fr) q1 = m1;
        }
        if (fr < NXI) {
          if (2 * fq < NXI) wn[WN_Q + fr * NV + 2 * fq] = q0;
          if (2 * fq + 1 < NXI) wn[WN_Q + fr * NV + 2 * fq + 1] = q1;
        }
        WARP_SYNC();
      } else
#endif
      {
      for (int j = 0; j < NODE_NEDGE(n); ++j) {
        const int e = NODE_FIRST_EDGE(n) + j;
        const double* er = I.er + (size_t)e * ER_SZ;
        const double* wc = I.node + (size_t)EDGE_CHILD_OF(e) * WN_NODE;
        for (int idx = c.lane; idx < NXI * NVP + NXI; idx += DMB_LANES) {
          if (idx < NXI * NVP) {
            const int r = idx / NVP, jj = idx - r * NVP;
            double s = (jj >= NX && jj < NX + NU) ? P_AT(wc, r, NX + (jj - NX)) : 0.0;     // u_prev_c = u
#pragma unroll
            for (int t = 0; t < NX; ++t) s += P_AT(wc, r, t) * er[ER_AB + t * NVP + jj];
            W[idx] = s;
          } else {
            const int r = idx - NXI * NVP;
            double s = wc[WN_QV + r];
#pragma unroll
            for (int t = 0; t < NX; ++t) s += P_AT(wc, r, t) * er[ER_B + t];
            pb[r] = s;
          }
        }
        WARP_SYNC();
        // [A B]' W is symmetric: only the lower triangle is formed (one lane pass for NVP <= 7) and mirrored
        for (int idx = c.lane; idx < NTRI_VP + NVP; idx += DMB_LANES) {
          if (idx < NTRI_VP) {
            int i, jj;
            if (idx == c.lane) { i = tvi; jj = tvj; } else tri_decode(idx, i, jj);
            double s = (i >= NX && i < NX + NU) ? W[(NX + i - NX) * NVP + jj] : 0.0;
#pragma unroll
            for (int r = 0; r < NX; ++r) s += er[ER_AB + r * NVP + i] * W[r * NVP + jj];
            wn[WN_Q + VPOS(i) * NV + VPOS(jj)] += s;
            if (i != jj) wn[WN_Q + VPOS(jj) * NV + VPOS(i)] += s;
          } else {
            const int i = idx - NTRI_VP;
            double s = (i >= NX && i < NX + NU) ? pb[NX + i - NX] : 0.0;
#pragma unroll
            for (int r = 0; r < NX; ++r) s += er[ER_AB + r * NVP + i] * pb[r];
            wn[WN_QV + VPOS(i)] += s;
          }
        }
        WARP_SYNC();
      }
      // eliminate the controls: every lane owns one column of [Q_u,xi | q_u] and carries its own Cholesky of Q_uu
      {
        double L[NUT * NUT];              // unit-free Cholesky factor; the diagonal stores 1 / L_aa
        int bad = 0;
        if (LDL) bad = quu_ldl(wn + WN_Q + NXI * NV + NXI, L);
        else {
#pragma unroll
        for (int a = 0; a < NUT; ++a)
#pragma unroll
          for (int b = 0; b <= a; ++b) {
            double s = wn[WN_Q + (NXI + a) * NV + NXI + b];
#pragma unroll
            for (int t = 0; t < b; ++t) s -= L[a * NUT + t] * L[b * NUT + t];
            if (a == b) { if (!(s > 0.0)) { bad = 1; s = 1.0; } L[a * NUT + a] = rsqrt(s); }
            else L[a * NUT + b] = s * L[b * NUT + b];
          }
        }
        if (bad && c.lane == 0) DMB_FLAG_OR(flag, bad);
        for (int col = c.lane; col <= NXI; col += DMB_LANES) {
          double rhs[NUT];
#pragma unroll
          for (int a = 0; a < NUT; ++a) rhs[a] = col < NXI ? -wn[WN_Q + (NXI + a) * NV + col] : -wn[WN_QV + NXI + a];
          if (LDL) quu_ldl_solve(L, rhs);
          else {
#pragma unroll
          for (int a = 0; a < NUT; ++a) { double s = rhs[a]; for (int t = 0; t < a; ++t) s -= L[a * NUT + t] * rhs[t]; rhs[a] = s * L[a * NUT + a]; }
#pragma unroll
          for (int a = NUT - 1; a >= 0; --a) { double s = rhs[a]; for (int t = a + 1; t < NUT; ++t) s -= L[t * NUT + a] * rhs[t]; rhs[a] = s * L[a * NUT + a]; }
          }
#pragma unroll
          for (int a = 0; a < NUT; ++a) { if (col < NXI) wn[WN_K + a * NXI + col] = rhs[a]; else wn[WN_KV + a] = rhs[a]; }
        }
      }
      WARP_SYNC();
      // P = Q_xixi + Q_xiu K is symmetric as well: lower triangle, mirrored
      for (int idx = c.lane; idx < NTRI_XI + NXI; idx += DMB_LANES) {
        if (idx < NTRI_XI) {
          int a, b;
          if (idx == c.lane) { a = txi; b = txj; } else tri_decode(idx, a, b);
          double s = wn[WN_Q + a * NV + b];
#pragma unroll
          for (int t = 0; t < NUT; ++t) s += wn[WN_Q + a * NV + NXI + t] * wn[WN_K + t * NXI + b];
          wn[WN_Q + a * NV + b] = s;
          if (a != b) wn[WN_Q + b * NV + a] = s;
        } else {
          const int a = idx - NTRI_XI;
          double s = wn[WN_QV + a];
#pragma unroll
          for (int t = 0; t < NUT; ++t) s += wn[WN_Q + a * NV + NXI + t] * wn[WN_KV + t];
          wn[WN_QV + a] = s;
        }
      }
      WARP_SYNC();
     }
      }
     if (par) CTA_SYNC();            // the next stage reads the cost-to-go of this one
    }
    SUB_TICK(I, 5);
  }
#if DMB_NFREE0 > 0
  // free root components: in-place Cholesky of P_0[f, f] (lower triangle, reciprocal diagonal); positive definiteness is part
  // of the inertia test.  The roll-outs solve  P_0[f,f] d xi_f = -(p_0[f] + P_0[f, fixed] d xi_fixed)  with it.
  if (!LDL && c.warp == 0) {
    WARP_SYNC();
    if (c.lane == 0) {
      double* w0 = I.node;
      int bad = 0;
      for (int a = 0; a < DMB_NFREE0; ++a)
        for (int b = 0; b <= a; ++b) {
          double sv = P_AT(w0, NXF + a, NXF + b);
          for (int t = 0; t < b; ++t) sv -= P_AT(w0, NXF + a, NXF + t) * P_AT(w0, NXF + b, NXF + t);
          if (a == b) { if (!(sv > 0.0)) { bad = 1; sv = 1.0; } P_AT(w0, NXF + a, NXF + a) = rsqrt(sv); }
          else P_AT(w0, NXF + a, NXF + b) = sv * P_AT(w0, NXF + b, NXF + b);
        }
      if (bad) DMB_FLAG_OR(flag, bad);
    }
    WARP_SYNC();
  }
#endif
  CTA_SYNC();
}

DMB_PHASE void riccati_backward(const Ctx& c, const Inst& I, double mu, double kappa_d, double delta, int ls_mode, double* Qs, int* flag) {
  riccati_backward_t<false>(c, I, mu, kappa_d, delta, ls_mode, Qs, flag);
}

// root state step d xi_0 (NXI entries into wn0[WN_DXI]): fixed components from the initial-condition residual, free components
// (DMB_NFREE0 > 0) from the Cholesky factor left in P_0[f, f] by riccati_backward.  Called by lane 0 of warp 0.
DMB_DEV void root_step(double* w0, const double* res) {
  for (int a = 0; a < NXI; ++a) w0[WN_DXI + a] = a < NXF ? -res[a] : 0.0;
#if DMB_NFREE0 > 0
  double y[DMB_NFREE0];
  for (int a = 0; a < DMB_NFREE0; ++a) {                       // rhs = -(p_f + P[f, fixed] dxi_fixed), then L y = rhs
    double sv = -w0[WN_QV + NXF + a];
    for (int t = 0; t < NXF; ++t) sv -= P_AT(w0, t, NXF + a) * w0[WN_DXI + t];        // P symmetric: P[f][t] = P[t][f]
    for (int t = 0; t < a; ++t) sv -= P_AT(w0, NXF + a, NXF + t) * y[t];
    y[a] = sv * P_AT(w0, NXF + a, NXF + a);
  }
  for (int a = DMB_NFREE0 - 1; a >= 0; --a) {                  // L' x = y
    double sv = y[a];
    for (int t = a + 1; t < DMB_NFREE0; ++t) sv -= P_AT(w0, NXF + t, NXF + a) * w0[WN_DXI + NXF + t];
    w0[WN_DXI + NXF + a] = sv * P_AT(w0, NXF + a, NXF + a);
  }
#endif
}

// tile image of edge block `eb` -> the group's tile T (all lanes of the group call; returns when the data has landed)
DMB_DEV void load_tile_image(const Ctx& c, const Inst& I, const double* eb, double* T, bool act) {
#if DMB_TMA
  if (act && c.gl == 0) tma_load_1d(T, eb + EB_IMG, IMG_SZ * 8, I.mbar + c.grp);
  if (act) mbar_wait(I.mbar + c.grp, (unsigned)I.mphase[c.grp]);
  WARP_SYNC();
  if (act && c.gl == 0) I.mphase[c.grp] ^= 1;
#else
  if (act) for (int i = c.gl; i < IMG_SZ; i += GSZ) T[i] = eb[EB_IMG + i];
#endif
}

// recovery phase, data of one edge that does not depend on the roll-out: the [LU | Yext] tile image (one TMA bulk copy, tracked by
// the group's mbarrier) and the sparse blocks, staged into the group's tile
DMB_DEV void fr_stage_edge(const Ctx& c, const Inst& I, const double* res, double delta, int ls_mode, int refine, int gy_off, int e, bool act, double* T) {
  const double* eb = I.eb + (size_t)(act ? e : 0) * EB_SZ;
#if DMB_TMA
  if (act && c.gl == 0) tma_load_1d(T, eb + EB_IMG, IMG_SZ * 8, I.mbar + c.grp);       // one bulk copy in flight per lane group
#endif
  if (act) {
    // coalesced, independent loads while the image lands
    DMB_COPY_BATCH(DMB_NNZ_H, c.gl, GSZ, eb[EB_HNZ + idx], T[STG_HNZ + idx] = ls_mode ? 0.0 : val);
    DMB_COPY_BATCH(NY, c.gl, GSZ, eb[gy_off + idx], T[STG_GY + idx] = val);
    DMB_COPY_BATCH(NY, c.gl, GSZ, eb[EB_SIGY + idx], T[STG_SIGY + idx] = val);
#if NNL > 0
    for (int i = c.gl; i < DMB_NNZ_D; i += GSZ) T[STG_DNZ + i] = eb[EB_DNZ + i];
    // (Sigma_s + delta) in place of Sigma_s (inertia correction, see factor_edges).  Re-solve with a new right-hand side:
    // r_d = (Sigma_s + delta) rp_d (zero in a refinement solve, where the condensed rows are exact)
    for (int i = c.gl; i < NNL; i += GSZ) {
      const double sg = eb[EB_SIGS + i] + delta, rr = ls_mode ? 0.0 : res[LAM_EDGE(e) + NY + NX + i];
      T[STG_SIGS + i] = sg; T[STG_RD + i] = refine ? sg * rr : eb[EB_RD + i] + delta * rr;
    }
#else
    (void)res; (void)refine; (void)delta;
#endif
  }
}

// ---------------------------------------------------------------------------------------------------
// forward roll-out (warp 0) then, per edge by its lane group: dy = Yext [dv; 1], new multipliers by a
// transposed LU solve in the group tile, inequality steps.
// ---------------------------------------------------------------------------------------------------
DMB_PHASE void forward_recover(const Ctx& c, const Inst& I, const double* res, double delta, int ls_mode, int refine, double* tile_base) {
  const int gy_off = refine ? EB_GY2 : EB_GY;
  SUB_START();
  const bool par = DMB_TREE_PAR && c.nwarp > 1;     // scenario trees: the nodes of a stage roll out on different warps
  // chains: warp 0 rolls out alone — the other warps fetch the data of their first edges meanwhile (the tiles are idle here)
  const bool early = !par && c.nwarp > 1;
  if (early && c.warp != 0) {
    const int e = c.grp;
    fr_stage_edge(c, I, res, delta, ls_mode, refine, gy_off, e, e < NEDGE, tile_base + (size_t)c.grp * TILE_SZ);
  }
  {
#if DMB_NSCEN == 1 && !defined(DMB_HOSTSIM) && (NXI <= 8) && (NVP <= 8) && !defined(DMB_NO_CHAIN_ROLLOUT)
    // chains: the state step d xi travels in registers (lane t owns d xi_t), broadcast by shuffles; nothing on the dependent
    // chain goes through memory (the iterate vectors live in the L2-backed slot for cfg #2: the former roll-out paid a global
    // store -> load round trip per node, ~1.2k cycles; r2 profile gpurun_out/r2d_r1_prof.json: fwd_sweep 7.6 % of the cycles)
    if (c.warp == 0) {
      const int t = c.lane;
#if DMB_NFREE0 > 0
      if (t == 0) root_step(I.node, res);
      WARP_SYNC();
      double dxi = t < NXI ? I.node[WN_DXI + t] : 0.0;
#else
      double dxi = t < NX ? -res[t] : 0.0;
#endif
      for (int n = 0; n <= NHOR; ++n) {
        double* wn = I.node + (size_t)n * WN_NODE;
        double* dwn = I.dw + NODE_OFF(n);
        if (t < NXI) wn[WN_DXI + t] = dxi;                 // kept for the multiplier recovery below
        if (n == NHOR) { if (t < NVP) dwn[t] = t < NX ? dxi : 0.0; break; }
        double xv[NXI], du[NUT];
#pragma unroll
        for (int i = 0; i < NXI; ++i) xv[i] = __shfl_sync(DMB_FULL, dxi, i);
#pragma unroll
        for (int a = 0; a < NUT; ++a) {
          double sv = wn[WN_KV + a];
#pragma unroll
          for (int i = 0; i < NXI; ++i) sv += wn[WN_K + a * NXI + i] * xv[i];
          du[a] = sv;
        }
        double dus = 0.0;                                   // du[t - NX] of this lane
#pragma unroll
        for (int a = 0; a < NUT; ++a) if (t - NX == a) dus = du[a];
        if (t < NVP) dwn[t] = t < NX ? dxi : dus;
        const double* er = I.er + (size_t)n * ER_SZ;
        double sv = 0.0;
        if (t < NX) {
          sv = er[ER_B + t];
#pragma unroll
          for (int i = 0; i < NX; ++i) sv += er[ER_AB + t * NVP + i] * xv[i];
#pragma unroll
          for (int a = 0; a < NUT; ++a) sv += er[ER_AB + t * NVP + NX + a] * du[a];
        } else if (t < NXI) sv = dus;
        dxi = sv;
      }
      WARP_SYNC();
    }
#else
    for (int lvl = 0; lvl <= NHOR; ++lvl) {
     if (par || c.warp == 0)
     for (int n = DMB_LVL(lvl) + (par ? c.warp : 0); n < DMB_LVL(lvl + 1); n += (par ? c.nwarp : 1)) {
      double* wn = I.node + (size_t)n * WN_NODE;
      const int k = lvl;
      if (n == 0) {
#if DMB_NFREE0 > 0
        if (c.lane == 0) root_step(wn, res);
#else
        for (int a = c.lane; a < NXI; a += DMB_LANES) wn[WN_DXI + a] = a < NX ? -res[a] : 0.0;
#endif
        WARP_SYNC();
      }
      double* dwn = I.dw + NODE_OFF(n);
      for (int a = c.lane; a < NX; a += DMB_LANES) dwn[a] = wn[WN_DXI + a];
      if (k == NHOR) { for (int a = NX + c.lane; a < NVP; a += DMB_LANES) dwn[a] = 0.0; WARP_SYNC(); continue; }
      for (int a = c.lane; a < NUT; a += DMB_LANES) {
        double s = wn[WN_KV + a];
#pragma unroll
        for (int t = 0; t < NXI; ++t) s += wn[WN_K + a * NXI + t] * wn[WN_DXI + t];
        dwn[NX + a] = s;
      }
      WARP_SYNC();
      for (int idx = c.lane; idx < NODE_NEDGE(n) * NXI; idx += DMB_LANES) {
        const int j = idx / NXI, r = idx - j * NXI;
        const int e = NODE_FIRST_EDGE(n) + j;
        const double* er = I.er + (size_t)e * ER_SZ;
        double* wc = I.node + (size_t)EDGE_CHILD_OF(e) * WN_NODE;
        double s;
        if (r < NX) {
          s = er[ER_B + r];
#pragma unroll
          for (int t = 0; t < NVP; ++t) s += er[ER_AB + r * NVP + t] * dwn[t];
        } else s = dwn[NX + (r - NX)];
        wc[WN_DXI + r] = s;
      }
      WARP_SYNC();
     }
     if (par) CTA_SYNC();              // the next stage starts from the states written here
    }
#endif
    if (c.warp == 0) {
      const double* w0 = I.node;
      for (int a = c.lane; a < NX; a += DMB_LANES) {
        double s = w0[WN_QV + a];
        for (int t = 0; t < NXI; ++t) s += P_AT(w0, a, t) * w0[WN_DXI + t];      // (rows of fixed components: untouched by the root Cholesky)
        I.lamn[a] = a < NXF ? -s : 0.0;                                          // free root components have no initial-condition row
      }
    }
    SUB_TICK(I, 6);
  }
  CTA_SYNC();
  double* T = tile_base + (size_t)c.grp * TILE_SZ;
  double* dq = T + STG_DQ;                  // dq (NQ), then r (NY), then jd (NNL)
  double* rv = dq + NQ;
  for (int e0 = 0; e0 < NEDGE; e0 += c.ngrp) {
    const int e = e0 + c.grp;
    const bool act = e < NEDGE;
    const double* eb = I.eb + (size_t)(act ? e : 0) * EB_SZ;
    if (!(early && e0 == 0 && c.warp != 0)) fr_stage_edge(c, I, res, delta, ls_mode, refine, gy_off, e, act, T);
    if (act) {
      const double* dvp = I.dw + NODE_OFF(EDGE_PARENT_OF(e));
      for (int i = c.gl; i < NVP; i += GSZ) dq[i] = dvp[i];
      // continuity multiplier = gradient of the child's cost-to-go
      const double* wc = I.node + (size_t)EDGE_CHILD_OF(e) * WN_NODE;
      for (int a = c.gl; a < NX; a += GSZ) {
        double s = wc[WN_QV + a];
#pragma unroll
        for (int t = 0; t < NXI; ++t) s += P_AT(wc, a, t) * wc[WN_DXI + t];
        I.lamn[LAM_EDGE(e) + NY + a] = s;
      }
    }
#if DMB_TMA
    if (act) mbar_wait(I.mbar + c.grp, (unsigned)I.mphase[c.grp]);
    WARP_SYNC();
    if (act && c.gl == 0) I.mphase[c.grp] ^= 1;
#else
    if (act) for (int i = c.gl; i < IMG_SZ; i += GSZ) T[i] = eb[EB_IMG + i];
    WARP_SYNC();
#endif
    if (act) {
      for (int i = c.gl; i < NY; i += GSZ) {
        double s = T[i * TILE_LD + NY + NVP];
#pragma unroll
        for (int t = 0; t < NVP; ++t) s += T[i * TILE_LD + NY + t] * dq[t];
        dq[NVP + i] = s;
        I.dw[EDGE_OFF(e) + i] = s;
      }
    }
    WARP_SYNC();
#if NNL > 0
    double* jd = rv + NY;                    // NNL entries
    if (act) {
      for (int i = c.gl; i < NNL; i += GSZ) {
        double s = 0.0;
        for (int t = DMB_DR_START[i]; t < DMB_DR_START[i + 1]; ++t) s += T[STG_DNZ + t] * dq[DMB_D_COL[t]];
        jd[i] = s;
        I.ds[e * NNL + i] = s + (ls_mode ? 0.0 : res[LAM_EDGE(e) + NY + NX + i]);
        // y_d+ = Sigma_s ds + grad_s phi = Sigma_s J dq + r_d
        I.lamn[LAM_EDGE(e) + NY + NX + i] = T[STG_SIGS + i] * s + T[STG_RD + i];
      }
    }
    WARP_SYNC();
#endif
    if (act) {
      // r_y = -(g_y + [H_e dq]_y) - E' lam_C
      for (int j = c.gl; j < NY; j += GSZ) {
        double s = T[STG_GY + j] + (T[STG_SIGY + j] + delta) * dq[NVP + j];
        for (int t = DMB_HR_START[NVP + j]; t < DMB_HR_START[NVP + j + 1]; ++t) s += T[STG_HNZ + DMB_HR_IDX[t]] * dq[DMB_HR_COL[t]];
#if NNL > 0
        for (int t = 0; t < DMB_NNZ_D; ++t) if (DMB_D_COL[t] == NVP + j) s += T[STG_DNZ + t] * (T[STG_SIGS + DMB_D_ROW[t]] * jd[DMB_D_ROW[t]] + T[STG_RD + DMB_D_ROW[t]]);
#endif
        rv[j] = -s;
      }
    }
    WARP_SYNC();
    if (act) {
      for (int a = c.gl; a < NX; a += GSZ) rv[DMB_XF_SEL[a]] -= I.lamn[LAM_EDGE(e) + NY + a];
    }
    WARP_SYNC();
    // ---- G_y' lam = r with P G_y = L U (diagonal of U stored as reciprocal):  U' a = r ; L' b = a ; lam = P' b ----------
#if !defined(DMB_HOSTSIM) && (NY <= GSZ)
    {
      // lane r of the group carries entry r of the vector; the pivot entry travels by shuffle: no barriers, no shared-memory
      // round trips (same operations in the same order as the loop version below)
      const unsigned gmask = GSZ == 32 ? DMB_FULL : (0xffffu << (c.lane & 16));
      const int gbase = c.lane & ~(GSZ - 1);
      const bool mine = act && c.gl < NY;
      double rr = mine ? rv[c.gl] : 0.0;
#pragma unroll
      for (int k = 0; k < NY; ++k) {
        const double ak = __shfl_sync(gmask, rr, gbase + k) * T[k * TILE_LD + k];
        if (mine && c.gl > k) rr -= T[k * TILE_LD + c.gl] * ak;
        if (c.gl == k) rr = ak;
      }
#pragma unroll
      for (int k = NY - 1; k >= 0; --k) {
        const double bk = __shfl_sync(gmask, rr, gbase + k);
        if (mine && c.gl < k) rr -= T[k * TILE_LD + c.gl] * bk;
      }
      if (mine) I.lamn[LAM_EDGE(e) + (int)eb[EB_PIV + c.gl]] = rr;                                       // lam = P' b
    }
    WARP_SYNC();
#else
    for (int k = 0; k < NY; ++k) {
      double ak = 0.0;
      if (act) {
        ak = rv[k] * T[k * TILE_LD + k];
        for (int r = k + 1 + c.gl; r < NY; r += GSZ) rv[r] -= T[k * TILE_LD + r] * ak;
      }
      WARP_SYNC();
      if (act && c.gl == 0) rv[k] = ak;
      WARP_SYNC();
    }
    for (int k = NY - 1; k >= 0; --k) {
      if (act) { const double bk = rv[k]; for (int r = c.gl; r < k; r += GSZ) rv[r] -= T[k * TILE_LD + r] * bk; }
      WARP_SYNC();
    }
    if (act) for (int j = c.gl; j < NY; j += GSZ) I.lamn[LAM_EDGE(e) + (int)eb[EB_PIV + j]] = rv[j];      // lam = P' b
    WARP_SYNC();
#endif
  }
  CTA_SYNC();
}
