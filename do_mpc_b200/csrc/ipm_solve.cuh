// Interior-point driver for one instance (all threads of a CTA cooperate).  See ipm_core.cuh.
// Algorithm and constants: Waechter & Biegler (2006), IPOPT defaults — identical to oracle/ipm.py.
#pragma once
#include "ipm_core.cuh"
#include "ipm_refine.cuh"

#ifdef DMB_HOSTSIM
#include <stdio.h>
static int dmb_hostsim_verbose = 0;
#endif
struct DmbShared {          // small per-CTA state in shared memory
  double red[64];           // reduction scratch (block_reduce: values x warps)
  double filt_theta[MAX_FILTER], filt_phi[MAX_FILTER];
  int n_filt;
  int flag, fail;
  int mphase[16];                      // phase parity of the lane groups' TMA barriers
  unsigned long long mbar[16];         // one mbarrier per lane group (tile-image loads, see load_tile_image)
};

#define DMB_NPHASE 17
struct DmbResult {
  double f, mu, inf_pr, inf_du, df;
  int status, iters, n_resto, n_resto_it, n_soc, n_refine, n_ls;
  long long phase_cycles[DMB_NPHASE];   // derivs | kkt+mu | factor | riccati | forward+recover | step sizes | line search | accept | init
};
#ifdef DMB_PROFILE
#define DMB_TICK(ph) do { const long long t_ = DMB_CLOCK(); pc[ph] += t_ - t_last; t_last = t_; } while (0)
#else
#define DMB_TICK(ph) ((void)0)
#endif

DMB_DEV double ftb_step(double slack, double dslack, double tau) {   // largest a: slack + a*dslack >= (1-tau)*slack
  return dslack < 0.0 ? -tau * slack / dslack : 1.0;
}

// barrier function value given objective f (scaled); returns +inf outside the bounds
DMB_DEV double barrier_phi(const Ctx& c, const Inst& I, const double* w, const double* s, double mu, double kappa_d,
                           double f, double* red) {
  double v[2] = {0.0, 0.0};
  barrier_terms(c, I, w, s, mu, kappa_d, v[0], v[1]);
  const int op[2] = {DMB_RED_SUM, DMB_RED_MAX};
  block_reduce<2>(c, v, op, red);
  return v[1] > 0.0 ? INFINITY : f + v[0];
}

// optimality error pieces in ONE pass and one reduction: dual infeasibility (gl = grad L without z, then -zL+zU),
// complementarity for the barrier parameter mu AND for mu = 0, the scaling factors s_d, s_c, and the unscaled primal
// infeasibility max |res| / dc
DMB_DEV void kkt_error(const Ctx& c, const Inst& I, double mu, double s_max, double* red,
                       double* dual, double* compl_mu, double* compl_0, double* s_d, double* s_c, double* prim_unscaled) {
  double du = 0.0, cm = 0.0, c0 = 0.0, sz = 0.0, nz = 0.0, sy = 0.0, pu = 0.0;
  for (int i = c.tid; i < NWI; i += c.nthr) {
    const bool hl = I.lb[i] > -DMB_INF, hu = I.ub[i] < DMB_INF;
    double r = I.gl[i];
    if (hl) { const double sz_ = (I.w[i] - I.lb[i]) * I.zL[i]; r -= I.zL[i]; cm = fmax(cm, fabs(sz_ - mu)); c0 = fmax(c0, fabs(sz_)); sz += fabs(I.zL[i]); nz += 1.0; }
    if (hu) { const double sz_ = (I.ub[i] - I.w[i]) * I.zU[i]; r += I.zU[i]; cm = fmax(cm, fabs(sz_ - mu)); c0 = fmax(c0, fabs(sz_)); sz += fabs(I.zU[i]); nz += 1.0; }
    du = fmax(du, fabs(r));
  }
#if NNL > 0
  for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) {
    const int e = i / NNL, j = i % NNL;
    const double yd = I.lam[LAM_EDGE(e) + NY + NX + j];
    du = fmax(du, fabs(-yd + I.vU[i]));
    const double sv = (nl_ub_relaxed(I, e, j) - I.s[i]) * I.vU[i];
    cm = fmax(cm, fabs(sv - mu)); c0 = fmax(c0, fabs(sv));
    sz += fabs(I.vU[i]); nz += 1.0;
  }
#endif
  for (int i = c.tid; i < NLAM; i += c.nthr) { sy += fabs(I.lam[i]); pu = fmax(pu, fabs(I.res[i]) / I.dc[i]); }
  double v[7] = {du, cm, c0, sz, nz, sy, pu};
  const int op[7] = {DMB_RED_MAX, DMB_RED_MAX, DMB_RED_MAX, DMB_RED_SUM, DMB_RED_SUM, DMB_RED_SUM, DMB_RED_MAX};
  block_reduce<7>(c, v, op, red);
  *dual = v[0]; *compl_mu = v[1]; *compl_0 = v[2]; *prim_unscaled = v[6];
  const double m_tot = (double)NLAM_ACTIVE;
  *s_d = fmax(s_max, (v[5] + v[3]) / fmax(m_tot + v[4], 1.0)) / s_max;
  *s_c = fmax(s_max, v[3] / fmax(v[4], 1.0)) / s_max;
}

// ---- cold-path helpers: ONE out-of-line copy each, shared by the starting point, the second-order correction and the
// restoration phase (the hot loop keeps its own inlined copies; DESIGN.md §4 on the instruction footprint) ------------------
#ifdef DMB_HOSTSIM
#define DMB_COLD static
#else
#define DMB_COLD __device__ __noinline__
#endif
// (arguments BY VALUE: the caller's Ctx / Inst never have their address taken, so they stay in registers in the hot loop)
DMB_COLD void eval_point_cold(const Ctx c, const Inst I, const double* w, const double* s, double* res, double* scratch,
                              double* f_out, double* theta_out, double* infpr_out) {
  eval_point(c, I, w, s, res, scratch, f_out, theta_out, infpr_out);
}
// K [dw; lam+; ds] = -[rd; rp] with the factors of the current iteration: rd = I.gl, rp = I.rest  ->  I.dw, I.lamn, I.ds
DMB_DEV void resolve_system(const Ctx& c, const Inst& I, double delta_w, double* tile, double* Qs) {
  resolve_edges(c, I, delta_w, tile);
  riccati_vector(c, I, delta_w, Qs);
  forward_recover(c, I, I.rest, delta_w, 0, 1, tile);
}
DMB_COLD void resolve_system_cold(const Ctx c, const Inst I, double delta_w, double* tile, double* Qs) { resolve_system(c, I, delta_w, tile, Qs); }

// filter acceptance test of a trial point (Waechter & Biegler 2006, Sec. 2.3): 0 = rejected, 1 = accepted (the filter gets
// augmented), 2 = accepted under the Armijo condition
DMB_DEV int trial_acceptable(const DmbOptions& o, const DmbShared* S, double theta_t, double phi_t, double theta, double phi,
                             double dphi, double alpha, double theta_max, double theta_min) {
  const double eps_r = 10.0 * 2.220446049250313e-16;
  if (!(phi_t == phi_t && phi_t < INFINITY && theta_t <= theta_max)) return 0;
  const bool switching = (dphi < 0.0) && (alpha * pow(-dphi, o.s_phi) > o.delta * pow(theta, o.s_theta));
  bool ok; int kind;
  if (theta <= theta_min && switching) { ok = (phi_t - phi - eps_r * fabs(phi) <= o.eta_phi * alpha * dphi); kind = 2; }
  else {
    ok = (theta_t - (1.0 - o.gamma_theta) * theta <= eps_r * fabs(theta)) || (phi_t - (phi - o.gamma_phi * theta) <= eps_r * fabs(phi));
    kind = 1;
  }
  if (ok) for (int k = 0; k < S->n_filt; ++k) if (!(theta_t < S->filt_theta[k] || phi_t < S->filt_phi[k])) { ok = false; break; }
  return ok ? kind : 0;
}

// largest step in [0,1] that keeps every bounded primal variable and slack a fraction tau inside its bounds
DMB_DEV double primal_ftb(const Ctx& c, const Inst& I, const double* dw, const double* ds, double tau, double* red) {
  double a = 1.0;
  for (int i = c.tid; i < NWI; i += c.nthr) {
    const double d = dw[i];
    if (I.lb[i] > -DMB_INF) a = fmin(a, ftb_step(I.w[i] - I.lb[i], d, tau));
    if (I.ub[i] < DMB_INF) a = fmin(a, ftb_step(I.ub[i] - I.w[i], -d, tau));
  }
#if NNL > 0
  for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) a = fmin(a, ftb_step(nl_ub_relaxed(I, i / NNL, i % NNL) - I.s[i], -ds[i], tau));
#else
  (void)ds;
#endif
  return block_min(c, a, red);
}

// dual step size: fraction-to-the-boundary for the bound multipliers along dz = mu/s - z -+ z/s * d (d from dw / ds)
DMB_DEV double dual_ftb(const Ctx& c, const Inst& I, const double* dw, const double* ds, double mu, double tau, double* red) {
  double a_z = 1.0;
  for (int i = c.tid; i < NWI; i += c.nthr) {
    const double d = dw[i];
    if (I.lb[i] > -DMB_INF) { const double sl = I.w[i] - I.lb[i]; a_z = fmin(a_z, ftb_step(I.zL[i], mu / sl - I.zL[i] - I.zL[i] / sl * d, tau)); }
    if (I.ub[i] < DMB_INF) { const double su = I.ub[i] - I.w[i]; a_z = fmin(a_z, ftb_step(I.zU[i], mu / su - I.zU[i] + I.zU[i] / su * d, tau)); }
  }
#if NNL > 0
  for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) {
    const double su = nl_ub_relaxed(I, i / NNL, i % NNL) - I.s[i];
    a_z = fmin(a_z, ftb_step(I.vU[i], mu / su - I.vU[i] + I.vU[i] / su * ds[i], tau));
  }
#else
  (void)ds;
#endif
  return block_min(c, a_z, red);
}

// lam_g0 / lam_x0 (flat reference layouts, may be NULL): the multipliers of a previous solution; used only with
// o.warm_start (IPOPT warm_start_init_point=yes: the point is pushed by warm_start_bound_push/frac, the multipliers are taken
// over — bound multipliers at least warm_start_mult_bound_push —, no least-squares estimate, mu starts at mu_init).
//
// One loop, four modes, so that every phase of the KKT solve is instantiated ONCE in the kernel:
//   MODE_SCALE  first pass: gradient-based scaling from the derivatives at the user's point, bound push, slack start
//   MODE_LS     least-squares multiplier estimate (cold start, and after a restoration): H := I, g := grad f - zL + zU
//   MODE_NORMAL an interior-point iteration (convergence test, mu update, step, refinement, filter line search with
//               second-order correction)
//   MODE_RESTO  an iteration of the feasibility restoration substitute (oracle/ipm.py::_restoration): same KKT solve with the
//               barrier terms of the restoration problem (resto_terms); objective and multipliers off through df = 0, lam = 0
enum { MODE_NORMAL = 0, MODE_LS = 1, MODE_RESTO = 2, MODE_SCALE = 3 };
DMB_DEV void solve_instance(const Ctx& c, const Inst& I_in, const DmbOptions& o, const double* x0_internal,
                            double* tile, double* Qs, DmbShared* S, DmbResult* out,
                            const double* lam_g0 = nullptr, const double* lam_x0 = nullptr) {
  Inst I = I_in;
  double* red = S->red;
#ifdef DMB_PROFILE
  long long pc[DMB_NPHASE];
  for (int i = 0; i < DMB_NPHASE; ++i) pc[i] = 0;
  I.sub = c.tid == 0 ? pc + 9 : nullptr;
  long long t_last = DMB_CLOCK();
#else
  I.sub = nullptr;
#endif
  for (int i = c.tid; i < NWI; i += c.nthr) { I.w[i] = x0_internal[i]; I.zL[i] = 0.0; I.zU[i] = 0.0; }
  for (int i = c.tid; i < NLAM; i += c.nthr) { I.dc[i] = 1.0; I.lam[i] = 0.0; }
  I.df = 1.0;
  if (c.tid == 0) { S->fail = 0; S->n_filt = 0; }
  CTA_SYNC();
  const bool ws = o.warm_start != 0 && lam_g0 != nullptr && lam_x0 != nullptr;
  double f = 0.0, theta = 0.0, inf_pr = 0.0;
  double mu = o.mu_init, tau = fmax(o.tau_min, 1.0 - mu);
  double theta_max = 0.0, theta_min = 0.0;
  double delta_w_last = 0.0;
  int it = 0, status = DMB_MAXIMUM_ITERATIONS_EXCEEDED, acceptable_cnt = 0, n_resto = 0, n_resto_it = 0, n_soc = 0, n_refine = 0, n_ls = 0;
  const int degenerate_reg = o.mimic_dummy_reg && DMB_HAS_DUMMIES;
  bool acceptable_now = false;
  double dual0 = 0.0;
  int mode = MODE_SCALE;
  // restoration state
  double theta_ref = 0.0, df_save = 1.0;
  int resto_it = 0;
  bool after_resto = false;

  while (true) {
    eval_derivatives(c, I);
    if (mode == MODE_SCALE) {
      // ---- gradient based scaling at the user's starting point (before the bound push) ----------------------------
      {
        double g = 0.0;
        for (int i = c.tid; i < NWI; i += c.nthr) {
          double gi = I.go[i];
#if DMB_HAS_INIT_FROM
          // the objective gradient w.r.t. a lifted global variable is the sum over its copies (what the flat NLP's scaling sees)
          if (DMB_W_INIT_FROM[i] >= 0) gi = 0.0;
          else if (DMB_LS_WEIGHT[i] < 1.0) for (int j = 0; j < NWI; ++j) if (DMB_W_INIT_FROM[j] == i) gi += I.go[j];
#endif
          g = fmax(g, fabs(gi));
        }
        g = block_max(c, g, red);
        if (g > o.nlp_scaling_max_gradient) I.df = fmax(o.nlp_scaling_max_gradient / g, o.nlp_scaling_min_value);
      }
      for (int e = c.tid; e < NEDGE; e += c.nthr) {
        // Jacobian nonzeros of this edge as left by eval_derivatives (dc = 1 so far: unscaled)
        const double* eb = I.eb + (size_t)e * EB_SZ;
        double mx[NY > 0 ? NY : 1];
        for (int i = 0; i < NY; ++i) mx[i] = 0.0;
        for (int k = 0; k < DMB_NNZ_J; ++k) mx[DMB_J_ROW[k]] = fmax(mx[DMB_J_ROW[k]], fabs(eb[EB_JNZ + k]));
        for (int i = 0; i < NY; ++i)
          I.dc[LAM_EDGE(e) + i] = mx[i] > o.nlp_scaling_max_gradient ? fmax(o.nlp_scaling_max_gradient / mx[i], o.nlp_scaling_min_value) : 1.0;
#if NNL > 0
        double md[NNL];
        for (int i = 0; i < NNL; ++i) md[i] = 0.0;
        for (int k = 0; k < DMB_NNZ_D; ++k) md[DMB_D_ROW[k]] = fmax(md[DMB_D_ROW[k]], fabs(eb[EB_DNZ + k]));
        for (int i = 0; i < NNL; ++i)
          I.dc[LAM_EDGE(e) + NY + NX + i] = md[i] > o.nlp_scaling_max_gradient ? fmax(o.nlp_scaling_max_gradient / md[i], o.nlp_scaling_min_value) : 1.0;
#endif
      }
      CTA_SYNC();
      // ---- starting point: push into the (relaxed) bounds, slacks, bound multipliers ----------------------------
      const double bpush = ws ? o.ws_bound_push : o.bound_push, bfrac = ws ? o.ws_bound_frac : o.bound_frac;
      for (int i = c.tid; i < NWI; i += c.nthr) {
        const double lo = I.lb[i], hi = I.ub[i];
        const bool hl = lo > -DMB_INF, hu = hi < DMB_INF;
        double v = I.w[i];
        if (hl && hu) {
          const double pl = fmin(bpush * fmax(1.0, fabs(lo)), bfrac * (hi - lo));
          const double pu = fmin(bpush * fmax(1.0, fabs(hi)), bfrac * (hi - lo));
          v = fmin(fmax(v, lo + pl), hi - pu);
        } else if (hl) v = fmax(v, lo + bpush * fmax(1.0, fabs(lo)));
        else if (hu) v = fmin(v, hi - bpush * fmax(1.0, fabs(hi)));
        I.w[i] = v;
        double zl = o.bound_mult_init_val, zu = o.bound_mult_init_val;
        if (ws) {
          const int fl = DMB_W2FLAT[i];
          const double lx = fl >= 0 ? I.df * lam_x0[fl] : 0.0;             // lam_x = zU - zL (unscaled objective)
          zl = fmin(fmax(-lx, o.ws_mult_bound_push), o.ws_mult_init_max);
          zu = fmin(fmax(lx, o.ws_mult_bound_push), o.ws_mult_init_max);
        }
        I.zL[i] = hl ? zl : 0.0;
        I.zU[i] = hu ? zu : 0.0;
      }
#if DMB_HAS_INIT_FROM
      // copies of a lifted global variable follow the pushed root copy: the linear consistency rows hold at the starting point, so
      // all copies take the same Newton steps as the single variable of the flat NLP
      CTA_SYNC();
      for (int i = c.tid; i < NWI; i += c.nthr) if (DMB_W_INIT_FROM[i] >= 0) I.w[i] = I.w[DMB_W_INIT_FROM[i]];
#endif
      if (ws) {
        for (int i = c.tid; i < NLAM; i += c.nthr) {
          const int fl = DMB_LAM2FLAT[i];
          const double y = fl >= 0 ? I.df * lam_g0[fl] / I.dc[i] : 0.0;
          I.lam[i] = fmin(fmax(y, -o.ws_mult_init_max), o.ws_mult_init_max);
        }
        CTA_SYNC();
#if DMB_DISCRETE
        for (int idx = c.tid; idx < NEDGE * NX; idx += c.nthr) {           // auxiliary end-of-interval rows: lam_G = lam_C
          const int e = idx / NX, i = idx - e * NX;
          I.lam[LAM_EDGE(e) + DMB_NZ + i] = I.lam[LAM_EDGE(e) + NY + i];
        }
#endif
      }
      for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) {
        I.s[i] = 0.0;
        I.vU[i] = ws ? fmin(fmax(I.lam[LAM_EDGE(i / (NNL > 0 ? NNL : 1)) + NY + NX + i % (NNL > 0 ? NNL : 1)], o.ws_mult_bound_push), o.ws_mult_init_max)
                     : o.bound_mult_init_val;
      }
      CTA_SYNC();
      eval_point_cold(c, I, I.w, I.s, I.res, red, &f, &theta, &inf_pr);
#if NNL > 0
      // slack init: s = d(x) pushed below the (relaxed) upper bound; residual d - s follows
      for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) {
        const int e = i / NNL, j = i % NNL;
        const double d = I.res[LAM_EDGE(e) + NY + NX + j];      // s was 0
        const double hi = nl_ub_relaxed(I, e, j);
        I.s[i] = fmin(d, hi - bpush * fmax(1.0, fabs(hi)));
      }
      CTA_SYNC();
      eval_point_cold(c, I, I.w, I.s, I.res, red, &f, &theta, &inf_pr);
#endif
      theta_max = o.theta_max_fact * fmax(1.0, theta); theta_min = o.theta_min_fact * fmax(1.0, theta);
      mode = ws ? MODE_NORMAL : MODE_LS;
      DMB_TICK(8);
      continue;
    }
    DMB_TICK(0);
    if (mode == MODE_NORMAL) {
      // ---- convergence tests ----------------------------------------------------------------------------------------------
      double dual, cmpl_mu, cmpl, s_d, s_c, pu;
      kkt_error(c, I, mu, o.s_max, red, &dual, &cmpl_mu, &cmpl, &s_d, &s_c, &pu);
      dual0 = dual;
      const double E0 = fmax(fmax(dual / s_d, inf_pr), cmpl / s_c);
#ifdef DMB_HOSTSIM
      if (dmb_hostsim_verbose) printf("%4d f=% .10e inf_pr=%.2e inf_du=%.2e cmpl=%.2e lg(mu)=%5.1f E0=%.2e theta=%.3e nfilt=%d\n", it, f / I.df, inf_pr, dual, cmpl, log10(mu), E0, theta, S->n_filt);
#endif
      {
        // unscaled secondary tolerances
        const double du_u = dual / I.df, cm_u = cmpl / I.df;
        if (!(E0 == E0)) { status = DMB_INVALID_NUMBER_DETECTED; break; }
        if (E0 <= o.tol && du_u <= o.dual_inf_tol && pu <= o.constr_viol_tol && cm_u <= o.compl_inf_tol) { status = DMB_SOLVE_SUCCEEDED; break; }
        acceptable_now = (E0 <= o.acceptable_tol && du_u <= o.acceptable_dual_inf_tol && pu <= o.acceptable_constr_viol_tol && cm_u <= o.acceptable_compl_inf_tol);
        if (acceptable_now) {
          if (++acceptable_cnt >= o.acceptable_iter) { status = DMB_SOLVED_TO_ACCEPTABLE_LEVEL; break; }
        } else acceptable_cnt = 0;
        if (it >= o.max_iter) { status = DMB_MAXIMUM_ITERATIONS_EXCEEDED; break; }
      }
      // ---- barrier parameter (monotone, repeated decrease allowed); the first test re-uses the pass above ----------------
      while (true) {
        const double Emu = fmax(fmax(dual / s_d, inf_pr), cmpl_mu / s_c);
        if (Emu <= o.kappa_eps * mu && mu > o.mu_min) {
          double nm = fmax(o.tol / 10.0, fmin(o.kappa_mu * mu, pow(mu, o.theta_mu)));
          nm = fmax(nm, o.mu_min);
          if (nm >= mu) break;
          mu = nm; tau = fmax(o.tau_min, 1.0 - mu);
          if (c.tid == 0) S->n_filt = 0;
          CTA_SYNC();
          double c0_;
          kkt_error(c, I, mu, o.s_max, red, &dual, &cmpl_mu, &c0_, &s_d, &s_c, &pu);
        } else break;
      }
    }
    DMB_TICK(1);
    // ---- search direction: one KKT solve in the mode's setting ---------------------------------------------------------------
    const int lsm = mode == MODE_LS;
    const double mu_k = lsm ? 0.0 : mu, kd_k = lsm ? 0.0 : o.kappa_d;
    if (lsm) for (int i = c.tid; i < NLAM; i += c.nthr) I.rest[i] = 0.0;
    if (c.tid == 0) S->fail = 0;
    CTA_SYNC();
    const double* rhs_c = lsm ? I.rest : I.res;
    factor_edges(c, I, rhs_c, mu_k, kd_k, lsm, tile, &S->fail);
    DMB_TICK(2);
    double delta_w = 0.0;
    if (mode == MODE_NORMAL && degenerate_reg) delta_w = delta_w_last == 0.0 ? o.delta_w_first : fmax(o.delta_w_min, o.kappa_w_minus * delta_w_last);
    bool step_fail = S->fail != 0;
    while (!step_fail) {
      riccati_backward(c, I, mu_k, kd_k, delta_w, lsm, Qs, &S->flag);
      const int wrong_inertia = S->flag;      // (re-set only behind the barrier at the top of the next call)
      if (!wrong_inertia) break;
      if (mode != MODE_NORMAL) { step_fail = true; break; }      // LS / restoration: the Hessian is a positive diagonal
      if (delta_w == 0.0) delta_w = delta_w_last == 0.0 ? o.delta_w_first : fmax(o.delta_w_min, o.kappa_w_minus * delta_w_last);
      else delta_w *= (delta_w_last == 0.0 ? o.kappa_w_plus_first : o.kappa_w_plus);
      if (delta_w > o.delta_w_max) step_fail = true;
    }
    if (mode == MODE_NORMAL) {
      // IPOPT: a failure at a point that satisfies the acceptable tolerances terminates as "acceptable"
      if (step_fail) { status = acceptable_now ? DMB_SOLVED_TO_ACCEPTABLE_LEVEL : DMB_ERROR_IN_STEP_COMPUTATION; break; }
      if (delta_w > 0.0) delta_w_last = delta_w;
    }
    DMB_TICK(3);
    if (!step_fail) forward_recover(c, I, rhs_c, delta_w, lsm, 0, tile);

    if (mode == MODE_LS) {
      // ---- least-squares multiplier estimate --------------------------------------------------------------------------------
      double m = 0.0;
      if (!step_fail) for (int i = c.tid; i < NLAM; i += c.nthr) m = fmax(m, fabs(I.lamn[i]));
      m = block_max(c, m, red);
      const bool ok = !step_fail && (m <= o.constr_mult_init_max) && (m == m);
      for (int i = c.tid; i < NLAM; i += c.nthr) I.lam[i] = ok ? I.lamn[i] : 0.0;
      CTA_SYNC();
      mode = MODE_NORMAL;
      if (after_resto) { after_resto = false; ++it; }
      DMB_TICK(8);
      continue;
    }

    if (mode == MODE_RESTO) {
      // ---- one step of the feasibility restoration: fraction-to-the-boundary, Armijo backtracking on theta --------------------
      bool leave = false, success = false;
      if (step_fail) leave = true;
      else {
        double a = primal_ftb(c, I, I.dw, I.ds, tau, red);
        double f_t = 0.0, theta_t = theta, infpr_t = inf_pr;
        bool improved = false;
        while (a > 1e-10) {
          for (int i = c.tid; i < NWI; i += c.nthr) I.wt[i] = I.w[i] + a * I.dw[i];
          for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) I.st[i] = I.s[i] + a * I.ds[i];
          CTA_SYNC();
          eval_point_cold(c, I, I.wt, I.st, I.rest, red, &f_t, &theta_t, &infpr_t);
          if (theta_t < (1.0 - 1e-4 * a) * theta) { improved = true; break; }
          a *= 0.5;
        }
#ifdef DMB_HOSTSIM
        if (dmb_hostsim_verbose) printf("     resto %d: theta %.6e -> %.6e alpha=%.3e improved=%d\n", resto_it, theta, theta_t, a, (int)improved);
#endif
        if (!improved) leave = true;
        else {
          for (int i = c.tid; i < NWI; i += c.nthr) I.w[i] = I.wt[i];
          for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) I.s[i] = I.st[i];
          for (int i = c.tid; i < NLAM; i += c.nthr) I.res[i] = I.rest[i];
          theta = theta_t; inf_pr = infpr_t;
          CTA_SYNC();
          if (theta <= o.kappa_resto * theta_ref && theta <= theta_max) {
            // acceptable to the (already augmented) filter of the outer problem?  needs the real objective
            I.df = df_save;
            eval_point_cold(c, I, I.w, I.s, I.rest, red, &f_t, &theta_t, &infpr_t);
            I.df = 0.0;
            const double phi_t = barrier_phi(c, I, I.w, I.s, mu, o.kappa_d, f_t, red);
            bool okf = phi_t == phi_t && phi_t < INFINITY;
            if (okf) for (int k = 0; k < S->n_filt; ++k) if (!(theta < S->filt_theta[k] || phi_t < S->filt_phi[k])) { okf = false; break; }
            if (okf) { leave = true; success = true; f = f_t; }
          }
          ++n_resto_it;
          if (!leave && ++resto_it >= o.resto_max_iter) leave = true;
        }
      }
      if (!leave) continue;
      // ---- leave the restoration; on failure the outer iterate comes back -----------------------------------------------------------
      if (!success) {
        for (int i = c.tid; i < NWI; i += c.nthr) I.w[i] = I.go[i];
        for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) I.s[i] = I.s2[i];
      }
      CTA_SYNC();
      I.df = df_save; I.resto = 0;
      if (!success) {
        status = theta_ref > 1e-6 ? DMB_INFEASIBLE_PROBLEM_DETECTED : DMB_RESTORATION_FAILED;
        theta = theta_ref;
        break;
      }
      // bound multipliers capped, equality multipliers from a least-squares estimate at the restored point
      for (int i = c.tid; i < NWI; i += c.nthr) { I.zL[i] = fmin(I.zL[i], 1e3); I.zU[i] = fmin(I.zU[i], 1e3); }
      for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) I.vU[i] = fmin(I.vU[i], 1e3);
      CTA_SYNC();
      mode = MODE_LS; after_resto = true;
      continue;
    }

    // ---- iterative refinement of the step (IPOPT: residual_ratio_max 1e-10) ------------------------------------------
    for (int ref_it = 0; ref_it < o.max_refine; ++ref_it) {
      double nr[3];
      kkt_residual(c, I, I.res, mu, o.kappa_d, delta_w, red, nr, tile);
      const double ratio = nr[0] / (fmin(nr[1], 1e6 * nr[2]) + nr[2] + 1e-300);
#ifdef DMB_HOSTSIM
      if (dmb_hostsim_verbose) printf("     refine %d: |rho|=%.3e |sol|=%.3e |rhs|=%.3e ratio=%.3e\n", ref_it, nr[0], nr[1], nr[2], ratio);
#endif
      if (!(ratio > o.residual_ratio_max)) break;
      for (int i = c.tid; i < NWI; i += c.nthr) I.wt[i] = I.dw[i];
      for (int i = c.tid; i < NLAM; i += c.nthr) I.lamr[i] = I.lamn[i];
      for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) I.st[i] = I.ds[i];
      CTA_SYNC();
      ++n_refine;
      resolve_system(c, I, delta_w, tile, Qs);
      for (int i = c.tid; i < NWI; i += c.nthr) I.dw[i] += I.wt[i];
      for (int i = c.tid; i < NLAM; i += c.nthr) I.lamn[i] += I.lamr[i];
      for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) I.ds[i] += I.st[i];
      CTA_SYNC();
    }
    DMB_TICK(4);

    // ---- step sizes, directional derivative (I.go = objective gradient, scaled by df) ------------------------------
    double a_max = 1.0, a_z = 1.0, dphi = 0.0;
    for (int i = c.tid; i < NWI; i += c.nthr) {
      const bool hl = I.lb[i] > -DMB_INF, hu = I.ub[i] < DMB_INF;
      double g = I.go[i];
      const double d = I.dw[i];
      if (hl) {
        const double sl = I.w[i] - I.lb[i];
        a_max = fmin(a_max, ftb_step(sl, d, tau));
        const double dz = mu / sl - I.zL[i] - I.zL[i] / sl * d;
        a_z = fmin(a_z, ftb_step(I.zL[i], dz, tau));
        g -= mu / sl; if (!hu) g += o.kappa_d * mu;
      }
      if (hu) {
        const double su = I.ub[i] - I.w[i];
        a_max = fmin(a_max, ftb_step(su, -d, tau));
        const double dz = mu / su - I.zU[i] + I.zU[i] / su * d;
        a_z = fmin(a_z, ftb_step(I.zU[i], dz, tau));
        g += mu / su; if (!hl) g -= o.kappa_d * mu;
      }
      dphi += g * d;
    }
#if NNL > 0
    for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) {
      const double su = nl_ub_relaxed(I, i / NNL, i % NNL) - I.s[i];
      const double d = I.ds[i];
      a_max = fmin(a_max, ftb_step(su, -d, tau));
      const double dv = mu / su - I.vU[i] + I.vU[i] / su * d;
      a_z = fmin(a_z, ftb_step(I.vU[i], dv, tau));
      dphi += (mu / su - o.kappa_d * mu) * d;
    }
#endif
    double phi;
    {
      double v[5] = {a_max, a_z, dphi, 0.0, 0.0};
      barrier_terms(c, I, I.w, I.s, mu, o.kappa_d, v[3], v[4]);
      const int op[5] = {DMB_RED_MIN, DMB_RED_MIN, DMB_RED_SUM, DMB_RED_SUM, DMB_RED_MAX};
      block_reduce<5>(c, v, op, red);
      a_max = v[0]; a_z = v[1]; dphi = v[2];
      phi = v[4] > 0.0 ? INFINITY : f + v[3];
    }
    DMB_TICK(5);

    // ---- filter line search -------------------------------------------------------------------------------------------------
    double a_min;
    if (dphi < 0.0 && theta <= theta_min)
      a_min = fmin(fmin(o.gamma_theta, o.gamma_phi * theta / (-dphi)), o.delta * pow(theta, o.s_theta) / pow(-dphi, o.s_phi));
    else if (dphi < 0.0) a_min = fmin(o.gamma_theta, o.gamma_phi * theta / (-dphi));
    else a_min = o.gamma_theta;
    a_min *= o.alpha_min_frac;
    double alpha = a_max, a_used = a_max;
    bool accepted = false, armijo = false;
    double f_t = f, theta_t = theta, infpr_t = inf_pr;
    int ls = 0;
    while (alpha >= a_min || ls == 0) {
      ++ls; ++n_ls;
      for (int i = c.tid; i < NWI; i += c.nthr) I.wt[i] = I.w[i] + alpha * I.dw[i];
      for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) I.st[i] = I.s[i] + alpha * I.ds[i];
      CTA_SYNC();
      double phi_t;
      eval_point(c, I, I.wt, I.st, I.rest, red, &f_t, &theta_t, &infpr_t, mu, o.kappa_d, &phi_t);
      const int acc = trial_acceptable(o, S, theta_t, phi_t, theta, phi, dphi, alpha, theta_max, theta_min);
      if (acc) { accepted = true; armijo = acc == 2; a_used = alpha; break; }
      // ---- second-order correction on the first trial when the infeasibility did not decrease (W&B 2006, Sec. 2.4) -------------
      if (ls == 1 && theta_t >= theta && o.max_soc > 0) {
        // the corrected step solves the same system with c_soc in place of c: it is the primary step plus the solution for
        // the right-hand side (0, c_soc - c).  Primary step kept in wt / lamr / st (the trial point is no longer needed).
        for (int i = c.tid; i < NWI; i += c.nthr) { I.wt[i] = I.dw[i]; I.gl[i] = 0.0; }
        for (int i = c.tid; i < NLAM; i += c.nthr) { I.lamr[i] = I.lamn[i]; I.rest[i] = (alpha - 1.0) * I.res[i] + I.rest[i]; }
        for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) I.st[i] = I.ds[i];
        CTA_SYNC();
        double theta_old = theta_t;
        for (int soc_it = 0; soc_it < o.max_soc; ++soc_it) {
          ++n_soc;
          resolve_system_cold(c, I, delta_w, tile, Qs);
          for (int i = c.tid; i < NWI; i += c.nthr) I.dw[i] += I.wt[i];
          for (int i = c.tid; i < NLAM; i += c.nthr) I.lamn[i] += I.lamr[i];
          for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) I.ds[i] += I.st[i];
          CTA_SYNC();
          const double a_soc = primal_ftb(c, I, I.dw, I.ds, tau, red);
          for (int i = c.tid; i < NWI; i += c.nthr) I.gl[i] = I.w[i] + a_soc * I.dw[i];
          for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) I.s2[i] = I.s[i] + a_soc * I.ds[i];
          CTA_SYNC();
          double f2, th2, ip2;
          eval_point_cold(c, I, I.gl, I.s2, I.soc, red, &f2, &th2, &ip2);
          const double phi2 = barrier_phi(c, I, I.gl, I.s2, mu, o.kappa_d, f2, red);
          const int acc2 = trial_acceptable(o, S, th2, phi2, theta, phi, dphi, alpha, theta_max, theta_min);
#ifdef DMB_HOSTSIM
          if (dmb_hostsim_verbose) printf("     soc %d: a_soc=%.3e theta %.6e -> %.6e acc=%d\n", soc_it, a_soc, theta_t, th2, acc2);
#endif
          if (acc2) {
            accepted = true; armijo = acc2 == 2; a_used = a_soc;
            for (int i = c.tid; i < NWI; i += c.nthr) I.wt[i] = I.gl[i];
            for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) I.st[i] = I.s2[i];
            for (int i = c.tid; i < NLAM; i += c.nthr) I.rest[i] = I.soc[i];
            f_t = f2; theta_t = th2; infpr_t = ip2;
            CTA_SYNC();
            a_z = dual_ftb(c, I, I.dw, I.ds, mu, tau, red);          // the whole direction is replaced, multipliers included
            break;
          }
          if (th2 > o.kappa_soc * theta_old || soc_it + 1 == o.max_soc) break;
          theta_old = th2;
          // next correction: c_soc <- a_soc c_soc + c(x_soc);  right-hand side (0, c_soc - c)
          for (int i = c.tid; i < NWI; i += c.nthr) I.gl[i] = 0.0;
          for (int i = c.tid; i < NLAM; i += c.nthr) I.rest[i] = a_soc * (I.rest[i] + I.res[i]) + I.soc[i] - I.res[i];
          CTA_SYNC();
        }
        if (accepted) break;
        for (int i = c.tid; i < NWI; i += c.nthr) I.dw[i] = I.wt[i];     // back to the primary step
        for (int i = c.tid; i < NLAM; i += c.nthr) I.lamn[i] = I.lamr[i];
        for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) I.ds[i] = I.st[i];
        CTA_SYNC();
      }
      alpha *= o.alpha_red_factor;
    }
    CTA_SYNC();                     // all threads are done reading the filter before thread 0 augments it
    DMB_TICK(6);
#ifdef DMB_HOSTSIM
    if (dmb_hostsim_verbose) printf("     alpha=%.3e a_max=%.3e a_z=%.3e ls=%d dphi=%.3e phi=%.12e a_min=%.3e accepted=%d armijo=%d delta_w=%.1e\n", a_used, a_max, a_z, ls, dphi, phi, a_min, (int)accepted, (int)armijo, delta_w);
#endif
    if (!accepted) {
      // ---- enter the feasibility restoration (IPOPT: called at an acceptable point -> terminate as acceptable) -------------
      ++n_resto;
      if (acceptable_now) { status = DMB_SOLVED_TO_ACCEPTABLE_LEVEL; break; }
      if (o.resto_max_iter <= 0) { status = theta > 1e-6 ? DMB_INFEASIBLE_PROBLEM_DETECTED : DMB_RESTORATION_FAILED; break; }
      if (c.tid == 0) {
        int k = S->n_filt < MAX_FILTER ? S->n_filt++ : MAX_FILTER - 1;
        S->filt_theta[k] = (1.0 - o.gamma_theta) * theta; S->filt_phi[k] = phi - o.gamma_phi * theta;
      }
      // the restoration iterates in place (w, s); the point it starts from — the reference x_R of the proximal term, restored
      // if the phase fails — is kept in the storage of `go` (no objective gradient is needed while df = 0) and `s2`
      theta_ref = theta; resto_it = 0; df_save = I.df;
      for (int i = c.tid; i < NWI; i += c.nthr) I.go[i] = I.w[i];
      for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) I.s2[i] = I.s[i];
      for (int i = c.tid; i < NLAM; i += c.nthr) I.lam[i] = 0.0;
      CTA_SYNC();
      I.df = 0.0; I.zeta = sqrt(mu); I.resto = 1;
      mode = MODE_RESTO;
      continue;
    }
    // ---- accept ------------------------------------------------------------------------------------------------------------------
    if (!armijo && c.tid == 0) {
      int k = S->n_filt < MAX_FILTER ? S->n_filt++ : MAX_FILTER - 1;
      S->filt_theta[k] = (1.0 - o.gamma_theta) * theta; S->filt_phi[k] = phi - o.gamma_phi * theta;
    }
    for (int i = c.tid; i < NWI; i += c.nthr) {
      const bool hl = I.lb[i] > -DMB_INF, hu = I.ub[i] < DMB_INF;
      const double d = I.dw[i], wn = I.wt[i];
      if (hl) {
        const double sl = I.w[i] - I.lb[i];
        double z = I.zL[i] + a_z * (mu / sl - I.zL[i] - I.zL[i] / sl * d);
        const double sn = wn - I.lb[i];
        z = fmax(fmin(z, o.kappa_sigma * mu / sn), mu / (o.kappa_sigma * sn));
        I.zL[i] = z;
      }
      if (hu) {
        const double su = I.ub[i] - I.w[i];
        double z = I.zU[i] + a_z * (mu / su - I.zU[i] + I.zU[i] / su * d);
        const double sn = I.ub[i] - wn;
        z = fmax(fmin(z, o.kappa_sigma * mu / sn), mu / (o.kappa_sigma * sn));
        I.zU[i] = z;
      }
      I.w[i] = wn;
    }
#if NNL > 0
    for (int i = c.tid; i < NEDGE * NNL; i += c.nthr) {
      const double hi = nl_ub_relaxed(I, i / NNL, i % NNL);
      const double su = hi - I.s[i];
      double v = I.vU[i] + a_z * (mu / su - I.vU[i] + I.vU[i] / su * I.ds[i]);
      const double sn = hi - I.st[i];
      v = fmax(fmin(v, o.kappa_sigma * mu / sn), mu / (o.kappa_sigma * sn));
      I.vU[i] = v; I.s[i] = I.st[i];
    }
#endif
    for (int i = c.tid; i < NLAM; i += c.nthr) { I.lam[i] += a_used * (I.lamn[i] - I.lam[i]); I.res[i] = I.rest[i]; }
    f = f_t; theta = theta_t; inf_pr = infpr_t;
    CTA_SYNC();
    DMB_TICK(7);
    ++it;
  }
  if (c.tid == 0) {
    out->f = f / I.df; out->mu = mu; out->inf_pr = inf_pr; out->inf_du = dual0;
    out->status = status; out->iters = it; out->n_resto = n_resto; out->df = I.df;
    out->n_resto_it = n_resto_it; out->n_soc = n_soc; out->n_refine = n_refine; out->n_ls = n_ls;
#ifdef DMB_PROFILE
    for (int i = 0; i < DMB_NPHASE; ++i) out->phase_cycles[i] = pc[i];
#else
    for (int i = 0; i < DMB_NPHASE; ++i) out->phase_cycles[i] = 0;
#endif
  }
  CTA_SYNC();
}

// ---------------------------------------------------------------------------------------------------
// instance I/O in the reference's flat layouts
// ---------------------------------------------------------------------------------------------------
// initial guess -> internal layout (into dst).  Either a flat opt_x vector (scaled variables), or the
// broadcast of (x0, u0, z0) that MPC.set_initial_guess performs (reference _mpc.py:969-971).
DMB_DEV void load_initial_guess(const Ctx& c, const double* x_flat, const double* x0s, const double* u0s,
                                const double* z0s, double* dst) {
  for (int i = c.tid; i < NWI; i += c.nthr) {
    double v = 0.0;
    const int fl = DMB_W2FLAT[i];
    if (x_flat) {
      if (fl >= 0) v = x_flat[fl];
      else if (i >= NNODE * NVP) {
        // auxiliary end-of-interval state of a discrete edge: start from the child's node state
        const int e = (i - NNODE * NVP) / NY, j = (i - NNODE * NVP) % NY;
        v = x_flat[DMB_NODE_X_FLAT[EDGE_CHILD_OF(e)] + (j - DMB_NZ)];
      }
    } else {
      if (i < NNODE * NVP) {
        const int j = i % NVP;
        if (j < NX) v = x0s[j]; else if (j < NX + NU) v = u0s[j - NX];
      } else {
        const int j = (i - NNODE * NVP) % NY;
#if DMB_DISCRETE
        v = j < DMB_NZ ? z0s[j] : x0s[j - DMB_NZ];
#else
        v = j < DMB_NPC * NX ? x0s[j % NX] : z0s[(j - DMB_NPC * NX) % (DMB_NZ > 0 ? DMB_NZ : 1)];
#endif
      }
    }
    dst[i] = v;
  }
  CTA_SYNC();
#if DMB_HAS_INIT_FROM
  // lifted copies of a global variable (MHE: the estimated parameters at every node) start from the value of the original
  for (int i = c.tid; i < NWI; i += c.nthr) if (DMB_W_INIT_FROM[i] >= 0) dst[i] = dst[DMB_W_INIT_FROM[i]];
  CTA_SYNC();
#endif
}

// solution -> flat arrays.  x_out must already hold the initial guess (unused reference variables keep it).
DMB_DEV void store_solution(const Ctx& c, const Inst& I, double df, double* x_out, double* lam_g_out, double* lam_x_out,
                            double* g_out) {
  for (int i = c.tid; i < NWI; i += c.nthr) {
    const int fl = DMB_W2FLAT[i];
    if (fl < 0) continue;
    if (x_out) x_out[fl] = I.w[i];
    if (lam_x_out) {
      double l = 0.0;
      if (I.lb[i] > -DMB_INF) l -= I.zL[i];
      if (I.ub[i] < DMB_INF) l += I.zU[i];
      lam_x_out[fl] = l / df;
    }
  }
  for (int i = c.tid; i < NLAM; i += c.nthr) {
    const int fl = DMB_LAM2FLAT[i];
    if (fl < 0) continue;
    if (lam_g_out) lam_g_out[fl] = I.dc[i] * I.lam[i] / df;
    if (g_out) {
      double r = I.res[i];
#if NNL > 0
      if (i >= NX) { const int j = (i - NX) % NLE; if (j >= NY + NX) r += I.s[((i - NX) / NLE) * NNL + (j - NY - NX)]; }
#endif
      g_out[fl] = r / I.dc[i];
    }
  }
  CTA_SYNC();
}

