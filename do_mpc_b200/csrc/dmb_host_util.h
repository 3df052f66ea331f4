// Host-side helpers shared by the CUDA C-ABI (dmb_api.cu) and the CPU emulation used in tests.
#pragma once
#include <math.h>
#include <string.h>

static const int DMB_W2FLAT_HOST[] = DMB_W2FLAT_INIT;

static inline void dmb_fill_default_options(DmbOptions* o) {
  memset(o, 0, sizeof(*o));
  o->tol = 1e-8; o->dual_inf_tol = 1.0; o->constr_viol_tol = 1e-4; o->compl_inf_tol = 1e-4;
  o->acceptable_tol = 1e-6; o->acceptable_dual_inf_tol = 1e10; o->acceptable_constr_viol_tol = 1e-2;
  o->acceptable_compl_inf_tol = 1e-2;
  o->s_max = 100.0; o->mu_init = 0.1; o->mu_min = 1e-11; o->kappa_mu = 0.2; o->theta_mu = 1.5; o->kappa_eps = 10.0;
  o->tau_min = 0.99; o->bound_push = 1e-2; o->bound_frac = 1e-2; o->bound_relax_factor = 1e-8;
  o->constr_mult_init_max = 1e3; o->bound_mult_init_val = 1.0; o->kappa_sigma = 1e10; o->kappa_d = 1e-5;
  o->alpha_red_factor = 0.5; o->eta_phi = 1e-8; o->theta_max_fact = 1e4; o->theta_min_fact = 1e-4;
  o->s_phi = 2.3; o->s_theta = 1.1; o->gamma_phi = 1e-8; o->gamma_theta = 1e-5; o->alpha_min_frac = 0.05;
  o->delta = 1.0; o->kappa_soc = 0.99;
  o->delta_w_first = 1e-4; o->delta_w_min = 1e-20; o->delta_w_max = 1e20; o->kappa_w_plus_first = 100.0;
  o->kappa_w_plus = 8.0; o->kappa_w_minus = 1.0 / 3.0;
  o->nlp_scaling_max_gradient = 100.0; o->nlp_scaling_min_value = 1e-8; o->kappa_resto = 0.9;
  o->max_iter = 3000; o->acceptable_iter = 15; o->max_soc = 4; o->mimic_dummy_reg = 1; o->resto_max_iter = 60;
  o->warm_start = 0; o->residual_ratio_max = 1e-10; o->max_refine = 4; o->pad_ = 0;
  o->ws_bound_push = 1e-3; o->ws_bound_frac = 1e-3; o->ws_mult_bound_push = 1e-3; o->ws_mult_init_max = 1e6;   // IPOPT warm_start_* defaults
}

// flat bounds (scaled variables, reference order) -> internal layout with IPOPT's bound relaxation.
// entries without a flat counterpart (terminal-node inputs, auxiliary states) are unbounded.
static inline void dmb_internal_bounds(const double* lbx, const double* ubx, const DmbOptions* o, double* lbI, double* ubI) {
  const int n = (int)(sizeof(DMB_W2FLAT_HOST) / sizeof(int));
  for (int i = 0; i < n; ++i) {
    const int fl = DMB_W2FLAT_HOST[i];
    double lo = -1e20, hi = 1e20;
    if (fl >= 0) {
      if (lbx[fl] > -DMB_INF) lo = lbx[fl] - fmin(o->bound_relax_factor * fmax(1.0, fabs(lbx[fl])), o->constr_viol_tol);
      if (ubx[fl] < DMB_INF) hi = ubx[fl] + fmin(o->bound_relax_factor * fmax(1.0, fabs(ubx[fl])), o->constr_viol_tol);
    }
    lbI[i] = lo; ubI[i] = hi;
  }
}
