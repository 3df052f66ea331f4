// Batched plant simulator: one thread integrates one instance over t_step with n_sub steps of the 3-stage Radau IIA
// collocation method (order 5, stiffly accurate; same collocation constants as the controller's discretisation) and
// a full Newton iteration per step; index-1 DAEs are handled by carrying the algebraic equations at every stage.
// Replaces the SUNDIALS CVODES / IDAS integrators the reference calls in do_mpc/simulator.py:690-755 (`simulate`).
// Compiles for the device and, with DMS_HOSTSIM, for the CPU (tests only).
#pragma once
#include <math.h>
#ifdef DMS_HOSTSIM
#define DMS_DEV static inline
#define DMS_TABLE static
#else
#define DMS_DEV __device__ __forceinline__
#define DMS_TABLE static __device__
#endif
#include "sim_gen.h"

#define SNX DMS_NX
#define SNZ DMS_NZ
#define SN (SNX + SNZ)                 // unknowns per stage
#define SNS (DMS_DEG * SN)             // unknowns per step
#define SLD (SNS + 1)

// solve A x = b (A: n x n in a[n][ld], b in column n) by Gaussian elimination with partial pivoting; returns 0 if singular
DMS_DEV int dms_gauss(double* a, int n, int ld) {
  for (int k = 0; k < n; ++k) {
    int p = k; double best = fabs(a[k * ld + k]);
    for (int r = k + 1; r < n; ++r) { const double v = fabs(a[r * ld + k]); if (v > best) { best = v; p = r; } }
    if (!(best > 1e-300)) return 0;
    if (p != k) for (int c = k; c <= n; ++c) { const double t = a[k * ld + c]; a[k * ld + c] = a[p * ld + c]; a[p * ld + c] = t; }
    const double inv = 1.0 / a[k * ld + k];
    for (int r = k + 1; r < n; ++r) {
      const double l = a[r * ld + k] * inv;
      if (l != 0.0) for (int c = k + 1; c <= n; ++c) a[r * ld + c] -= l * a[k * ld + c];
    }
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = a[k * ld + n];
    for (int c = k + 1; c < n; ++c) s -= a[k * ld + c] * a[c * ld + n];
    a[k * ld + n] = s / a[k * ld + k];
  }
  return 1;
}

// one instance: x,z in/out (physical units). returns 0 ok, 1 Newton failure
DMS_DEV int dms_integrate(double* x, double* z, const double* u, const double* tvp, const double* pp, double t_step, int n_sub,
                          double newton_tol, int newton_max) {
#if DMS_DISCRETE
  double fa[SN > 0 ? SN : 1], jac[SN * SN > 0 ? SN * SN : 1];
#if SNZ > 0
  int converged = 0;
  for (int it = 0; it < newton_max; ++it) {            // algebraic equations 0 = g(x, u, z) for z
    gen_sim_fa(x, u, z, tvp, pp, fa);
    gen_sim_jac(x, u, z, tvp, pp, jac);
    double A[SNZ * (SNZ + 1)];
    for (int i = 0; i < SNZ; ++i) { for (int j = 0; j < SNZ; ++j) A[i * (SNZ + 1) + j] = jac[(SNX + i) * SN + SNX + j]; A[i * (SNZ + 1) + SNZ] = -fa[SNX + i]; }
    if (!dms_gauss(A, SNZ, SNZ + 1)) return 1;
    double nrm = 0.0;
    for (int i = 0; i < SNZ; ++i) { z[i] += A[i * (SNZ + 1) + SNZ]; nrm = fmax(nrm, fabs(A[i * (SNZ + 1) + SNZ]) / (1.0 + fabs(z[i]))); }
    if (!(nrm == nrm)) return 1;
    if (nrm < newton_tol) { converged = 1; break; }
  }
  if (!converged) return 1;                            // an unconverged z must not be accepted (as in the continuous branch)
#endif
  gen_sim_fa(x, u, z, tvp, pp, fa);
  for (int i = 0; i < SNX; ++i) x[i] = fa[i];
  (void)jac; (void)t_step; (void)n_sub;
  return 0;
#else
  const double h = t_step / n_sub;
  double Y[SNS], A[SNS * SLD], fa[SN], jac[SN * SN];
  for (int s = 0; s < n_sub; ++s) {
    for (int j = 0; j < DMS_DEG; ++j) { for (int i = 0; i < SNX; ++i) Y[j * SN + i] = x[i]; for (int i = 0; i < SNZ; ++i) Y[j * SN + SNX + i] = z[i]; }
    int converged = 0;
    for (int it = 0; it < newton_max; ++it) {
      for (int i = 0; i < SNS * SLD; ++i) A[i] = 0.0;
      for (int j = 0; j < DMS_DEG; ++j) {
        const double* Xj = Y + j * SN; const double* Zj = Xj + SNX;
        gen_sim_fa(Xj, u, Zj, tvp, pp, fa);
        gen_sim_jac(Xj, u, Zj, tvp, pp, jac);
        for (int i = 0; i < SNX; ++i) {
          // residual: h f(X_j, Z_j) - sum_r C[r][j+1] X_r   (X_0 = x)
          double r = h * fa[i] - DMS_C[0 * (DMS_DEG + 1) + (j + 1)] * x[i];
          for (int q = 0; q < DMS_DEG; ++q) r -= DMS_C[(q + 1) * (DMS_DEG + 1) + (j + 1)] * Y[q * SN + i];
          double* row = A + (j * SN + i) * SLD;
          for (int c = 0; c < SN; ++c) row[j * SN + c] = h * jac[i * SN + c];
          for (int q = 0; q < DMS_DEG; ++q) row[q * SN + i] -= DMS_C[(q + 1) * (DMS_DEG + 1) + (j + 1)];
          row[SNS] = -r;
        }
        for (int i = 0; i < SNZ; ++i) {
          double* row = A + (j * SN + SNX + i) * SLD;
          for (int c = 0; c < SN; ++c) row[j * SN + c] = jac[(SNX + i) * SN + c];
          row[SNS] = -fa[SNX + i];
        }
      }
      if (!dms_gauss(A, SNS, SLD)) return 1;
      double nrm = 0.0;
      for (int i = 0; i < SNS; ++i) { const double d = A[i * SLD + SNS]; Y[i] += d; nrm = fmax(nrm, fabs(d) / (1.0 + fabs(Y[i]))); }
      if (!(nrm == nrm)) return 1;
      if (nrm < newton_tol) { converged = 1; break; }
    }
    if (!converged) return 1;
    for (int i = 0; i < SNX; ++i) x[i] = Y[(DMS_DEG - 1) * SN + i];
    for (int i = 0; i < SNZ; ++i) z[i] = Y[(DMS_DEG - 1) * SN + SNX + i];
  }
  return 0;
#endif
}

// consistent algebraic variables for given x, u: Newton on g(x, u, z) = 0
DMS_DEV int dms_init_alg(const double* x, double* z, const double* u, const double* tvp, const double* pp, double newton_tol, int newton_max) {
#if SNZ > 0
  double fa[SN], jac[SN * SN], A[SNZ * (SNZ + 1)];
  for (int it = 0; it < newton_max; ++it) {
    gen_sim_fa(x, u, z, tvp, pp, fa);
    gen_sim_jac(x, u, z, tvp, pp, jac);
    for (int i = 0; i < SNZ; ++i) { for (int j = 0; j < SNZ; ++j) A[i * (SNZ + 1) + j] = jac[(SNX + i) * SN + SNX + j]; A[i * (SNZ + 1) + SNZ] = -fa[SNX + i]; }
    if (!dms_gauss(A, SNZ, SNZ + 1)) return 1;
    double nrm = 0.0;
    for (int i = 0; i < SNZ; ++i) { z[i] += A[i * (SNZ + 1) + SNZ]; nrm = fmax(nrm, fabs(A[i * (SNZ + 1) + SNZ]) / (1.0 + fabs(z[i]))); }
    if (nrm < newton_tol) return 0;
  }
  return 1;
#else
  (void)x; (void)z; (void)u; (void)tvp; (void)pp; (void)newton_tol; (void)newton_max;
  return 0;
#endif
}
