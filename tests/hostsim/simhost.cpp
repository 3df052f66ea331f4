// CPU build of the plant-simulator integrator source (do_mpc_b200/csrc/sim_core.cuh) — TEST TOOL ONLY.
#define DMS_HOSTSIM 1
#include "sim_core.cuh"
extern "C" {
void simhost_dims(int* d) { d[0] = SNX; d[1] = DMS_NU; d[2] = SNZ; d[3] = DMS_NTVP; d[4] = DMS_NPM; d[5] = DMS_NAUX; }
int simhost_step(int B, const double* x, const double* u, const double* z, const double* tvp, const double* pp, double t_step,
                 int n_sub, double tol, int newton_max, int mode, double* x_next, double* z_next, int* status) {
  for (int b = 0; b < B; ++b) {
    double xl[SNX], zl[SNZ > 0 ? SNZ : 1];
    for (int i = 0; i < SNX; ++i) xl[i] = x[b * SNX + i];
    for (int i = 0; i < SNZ; ++i) zl[i] = z[b * SNZ + i];
    const double* ub = u ? u + b * DMS_NU : nullptr; const double* tb = tvp ? tvp + b * DMS_NTVP : nullptr; const double* pb = pp ? pp + b * DMS_NPM : nullptr;
    status[b] = mode == 1 ? dms_init_alg(xl, zl, ub, tb, pb, tol, newton_max) : dms_integrate(xl, zl, ub, tb, pb, t_step, n_sub, tol, newton_max);
    for (int i = 0; i < SNX; ++i) x_next[b * SNX + i] = xl[i];
    for (int i = 0; i < SNZ; ++i) z_next[b * SNZ + i] = zl[i];
  }
  return 0;
}
}
