// CPU emulation of the CUDA interior-point kernel (one "thread", DMB_LANES=1) — TEST TOOL ONLY.
// It compiles the very same algorithm source (do_mpc_b200/csrc/ipm_core.cuh, ipm_solve.cuh) with g++ so that the
// numerics can be compared with the oracle on machines without a GPU.  The product never loads this.
#define DMB_HOSTSIM 1
#include <stdlib.h>
#include <vector>
#include "ipm_solve.cuh"
#include "ipm_sens.cuh"
#include "dmb_host_util.h"

extern "C" {
void hostsim_dims(int* d) {
  d[0] = DMB_NW; d[1] = DMB_NG; d[2] = DMB_NP; d[3] = DMB_NX; d[4] = DMB_NU; d[5] = DMB_NZ; d[6] = DMB_NAUX;
  d[7] = NWI; d[8] = NLAM; d[9] = (int)sizeof(DmbOptions);
}
void hostsim_default_options(DmbOptions* o) { dmb_fill_default_options(o); }
void hostsim_set_verbose(int v) { dmb_hostsim_verbose = v; }

// x_init: [B, n_w] flat scaled initial guess or NULL (then x0s/u0s/z0s broadcast); outputs may be NULL.
int hostsim_solve_ws(int B, const double* x_init, const double* lam_g0, const double* lam_x0, const double* x0s, const double* u0s,
                     const double* z0s, const double* p, const double* lbx, const double* ubx, const DmbOptions* opts,
                     double* x, double* lam_g, double* lam_x, double* g, double* f, int* status, int* iters, double* u0);
int hostsim_solve(int B, const double* x_init, const double* x0s, const double* u0s, const double* z0s, const double* p,
                  const double* lbx, const double* ubx, const DmbOptions* opts,
                  double* x, double* lam_g, double* lam_x, double* g, double* f, int* status, int* iters, double* u0) {
  return hostsim_solve_ws(B, x_init, nullptr, nullptr, x0s, u0s, z0s, p, lbx, ubx, opts, x, lam_g, lam_x, g, f, status, iters, u0);
}
int hostsim_solve_ws(int B, const double* x_init, const double* lam_g0, const double* lam_x0, const double* x0s, const double* u0s,
                     const double* z0s, const double* p, const double* lbx, const double* ubx, const DmbOptions* opts,
                     double* x, double* lam_g, double* lam_x, double* g, double* f, int* status, int* iters, double* u0) {
  std::vector<double> lbI(NWI), ubI(NWI), vec(WV_END), er((size_t)NEDGE * ER_SZ), node((size_t)NNODE * WN_NODE),
      eb((size_t)NEDGE * EB_SZ);
  std::vector<double> tile((size_t)TILE_SZ + 1), Qs(DMB_QS_SIZE), x0i(NWI);
  dmb_internal_bounds(lbx, ubx, opts, lbI.data(), ubI.data());
  Ctx c{0, 1, 0, 0, 1, 0, 0, 1};
  DmbShared S;
  for (int b = 0; b < B; ++b) {
    Inst I;
    bind_vectors(I, vec.data());
    I.er = er.data(); I.node = node.data(); I.eb = eb.data();
    I.p = p + (size_t)b * DMB_NP; I.lb = lbI.data(); I.ub = ubI.data(); I.df = 1.0; I.sub = nullptr;
    I.relax_fac = opts->bound_relax_factor; I.relax_max = opts->constr_viol_tol;
    load_initial_guess(c, x_init ? x_init + (size_t)b * DMB_NW : nullptr, x0s ? x0s + (size_t)b * DMB_NX : nullptr,
                       u0s ? u0s + (size_t)b * DMB_NU : nullptr, z0s ? z0s + (size_t)b * DMB_NZ : nullptr, x0i.data());
    DmbResult r;
    solve_instance(c, I, *opts, x0i.data(), tile.data(), Qs.data(), &S, &r, lam_g0 ? lam_g0 + (size_t)b * DMB_NG : nullptr,
                   lam_x0 ? lam_x0 + (size_t)b * DMB_NW : nullptr);
    I.df = r.df;
    if (x && x_init) for (int i = 0; i < DMB_NW; ++i) x[(size_t)b * DMB_NW + i] = x_init[(size_t)b * DMB_NW + i];
    store_solution(c, I, r.df, x ? x + (size_t)b * DMB_NW : nullptr, lam_g ? lam_g + (size_t)b * DMB_NG : nullptr,
                   lam_x ? lam_x + (size_t)b * DMB_NW : nullptr, g ? g + (size_t)b * DMB_NG : nullptr);
    if (f) f[b] = r.f;
    if (status) status[b] = r.status;
    if (iters) iters[b] = r.iters;
    if (u0) for (int i = 0; i < NX * 0 + NU; ++i) u0[(size_t)b * NU + i] = I.w[NODE_OFF(0) + NX + i];
  }
  return 0;
}

// sensitivities of solved instances (same phases as dmb_sens_kernel).  dx [B, ncols, n_w], dlam [B, ncols, n_g+n_w] or NULL
int hostsim_sens(int B, const double* x, const double* lam_g, const double* p, const double* lbx, const double* ubx,
                 const DmbOptions* opts, double tol, double big, int rule, const int* cols, int ncols, double* dx, double* dlam, int* sstat) {
  std::vector<double> lbI(NWI), ubI(NWI), vec(WV_END), er((size_t)NEDGE * ER_SZ), node((size_t)NNODE * WN_NODE),
      eb((size_t)NEDGE * EB_SZ);
  std::vector<double> tile((size_t)TILE_SZ + 1), Qs(DMB_QS_SIZE), lb0(NWI), ub0(NWI);
  dmb_internal_bounds(lbx, ubx, opts, lbI.data(), ubI.data());
  DmbOptions o0 = *opts; o0.bound_relax_factor = 0.0;
  dmb_internal_bounds(lbx, ubx, &o0, lb0.data(), ub0.data());
  Ctx c{0, 1, 0, 0, 1, 0, 0, 1};
  DmbShared S;
  for (int b = 0; b < B; ++b) {
    Inst I;
    bind_vectors(I, vec.data());
    I.er = er.data(); I.node = node.data(); I.eb = eb.data();
    I.p = p + (size_t)b * DMB_NP; I.lb = lbI.data(); I.ub = ubI.data(); I.df = 1.0; I.sub = nullptr;
    I.relax_fac = opts->bound_relax_factor; I.relax_max = opts->constr_viol_tol;
    sstat[b] = sens_factor(c, I, x + (size_t)b * DMB_NW, lam_g + (size_t)b * DMB_NG, rule ? lbI.data() : lb0.data(),
                           rule ? ubI.data() : ub0.data(), rule, tol, big, tile.data(), Qs.data(), &S);
    if (sstat[b] >= DMB_SENS_FAILED) continue;
    for (int jc = 0; jc < ncols; ++jc) {
      sens_rhs(c, I, cols[jc]);
      sens_solve(c, I, opts->max_refine, opts->residual_ratio_max, tile.data(), Qs.data(), &S);
      sens_store(c, I, 0, dx + ((size_t)b * ncols + jc) * DMB_NW, dlam ? dlam + ((size_t)b * ncols + jc) * (DMB_NG + DMB_NW) : nullptr);
    }
  }
  return 0;
}
}
