// C ABI of the batched plant simulator (include/dompc_b200_sim.h).  One generated library per model.
#include <cuda_runtime.h>
#include <string>
#include <new>
#include "sim_core.cuh"
#include "dompc_b200_sim.h"

__global__ void dms_step_kernel(int B, const double* x, const double* u, const double* z, const double* tvp, const double* pp,
                                double t_step, int n_sub, double tol, int newton_max, int mode,
                                double* x_next, double* z_next, double* aux, int* status) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double xl[SNX], zl[SNZ > 0 ? SNZ : 1], ul[DMS_NU > 0 ? DMS_NU : 1], tl[DMS_NTVP > 0 ? DMS_NTVP : 1], pl[DMS_NPM > 0 ? DMS_NPM : 1];
  for (int i = 0; i < SNX; ++i) xl[i] = x[(size_t)b * SNX + i];
  for (int i = 0; i < SNZ; ++i) zl[i] = z[(size_t)b * SNZ + i];
  for (int i = 0; i < DMS_NU; ++i) ul[i] = u[(size_t)b * DMS_NU + i];
  for (int i = 0; i < DMS_NTVP; ++i) tl[i] = tvp[(size_t)b * DMS_NTVP + i];
  for (int i = 0; i < DMS_NPM; ++i) pl[i] = pp[(size_t)b * DMS_NPM + i];
  if (aux) { double al[DMS_NAUX > 0 ? DMS_NAUX : 1]; gen_sim_aux(xl, ul, zl, tl, pl, al); for (int i = 0; i < DMS_NAUX; ++i) aux[(size_t)b * DMS_NAUX + i] = al[i]; }
  int st = mode == 1 ? dms_init_alg(xl, zl, ul, tl, pl, tol, newton_max) : dms_integrate(xl, zl, ul, tl, pl, t_step, n_sub, tol, newton_max);
  for (int i = 0; i < SNX; ++i) x_next[(size_t)b * SNX + i] = xl[i];
  for (int i = 0; i < SNZ; ++i) z_next[(size_t)b * SNZ + i] = zl[i];
  status[b] = st;
}

struct dms_sim { int device = 0; double* d = nullptr; size_t cap = 0; std::string err; long long n_launches = 0; };
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { s->err = std::string(#call) + ": " + cudaGetErrorString(e_); return -1; } } while (0)

extern "C" {
void dms_get_dims(int* d) { d[0] = SNX; d[1] = DMS_NU; d[2] = SNZ; d[3] = DMS_NTVP; d[4] = DMS_NPM; d[5] = DMS_NAUX; d[6] = DMS_DISCRETE; d[7] = DMS_DEG; }
dms_sim* dms_create(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device >= n) return nullptr;
  dms_sim* s = new (std::nothrow) dms_sim();
  if (s) s->device = device;
  return s;
}
void dms_destroy(dms_sim* s) { if (!s) return; cudaSetDevice(s->device); cudaFree(s->d); delete s; }
const char* dms_last_error(dms_sim* s) { return s ? s->err.c_str() : "null handle"; }
long long dms_launch_count(dms_sim* s) { return s ? s->n_launches : 0; }

int dms_step_device(dms_sim* s, int B, const double* d_x, const double* d_u, const double* d_z, const double* d_tvp, const double* d_p,
                    double t_step, int n_sub, double newton_tol, int newton_max, int mode,
                    double* d_x_next, double* d_z_next, double* d_aux, int* d_status, void* stream_) {
  if (!s || B <= 0 || !d_x || !d_x_next || !d_status || (DMS_NU && !d_u) || (SNZ && (!d_z || !d_z_next)) || (DMS_NTVP && !d_tvp) ||
      (DMS_NPM && !d_p)) return -2;
  CK(cudaSetDevice(s->device));
  dms_step_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream_>>>(B, d_x, d_u, d_z, d_tvp, d_p, t_step, n_sub, newton_tol, newton_max, mode,
                                                                  d_x_next, d_z_next, d_aux, d_status);
  CK(cudaGetLastError());
  s->n_launches++;
  return 0;
}

int dms_step_host(dms_sim* s, int B, const double* x, const double* u, const double* z, const double* tvp, const double* pp,
                  double t_step, int n_sub, double newton_tol, int newton_max, int mode,
                  double* x_next, double* z_next, double* aux, int* status) {
  if (!s || B <= 0 || !x || !x_next || !status) return -2;
  CK(cudaSetDevice(s->device));
  const size_t nB = B;
  const size_t n_in = nB * (SNX + DMS_NU + SNZ + DMS_NTVP + DMS_NPM), n_out = nB * (SNX + SNZ + DMS_NAUX + 1);
  if (n_in + n_out + 16 > s->cap) { cudaFree(s->d); s->d = nullptr; s->cap = 0; CK(cudaMalloc(&s->d, (n_in + n_out + 16) * sizeof(double))); s->cap = n_in + n_out + 16; }
  double* dx = s->d; double* du = dx + nB * SNX; double* dz = du + nB * DMS_NU; double* dt = dz + nB * SNZ; double* dp = dt + nB * DMS_NTVP;
  double* ox = dp + nB * DMS_NPM; double* oz = ox + nB * SNX; double* oa = oz + nB * SNZ; int* os = (int*)(oa + nB * DMS_NAUX);
  CK(cudaMemcpy(dx, x, nB * SNX * sizeof(double), cudaMemcpyHostToDevice));
  if (DMS_NU) CK(cudaMemcpy(du, u, nB * DMS_NU * sizeof(double), cudaMemcpyHostToDevice));
  if (SNZ) CK(cudaMemcpy(dz, z, nB * SNZ * sizeof(double), cudaMemcpyHostToDevice));
  if (DMS_NTVP) CK(cudaMemcpy(dt, tvp, nB * DMS_NTVP * sizeof(double), cudaMemcpyHostToDevice));
  if (DMS_NPM) CK(cudaMemcpy(dp, pp, nB * DMS_NPM * sizeof(double), cudaMemcpyHostToDevice));
  dms_step_kernel<<<(B + 63) / 64, 64>>>(B, dx, du, dz, dt, dp, t_step, n_sub, newton_tol, newton_max, mode, ox, oz, aux ? oa : nullptr, os);
  CK(cudaGetLastError());
  s->n_launches++;
  CK(cudaMemcpy(x_next, ox, nB * SNX * sizeof(double), cudaMemcpyDeviceToHost));
  if (SNZ && z_next) CK(cudaMemcpy(z_next, oz, nB * SNZ * sizeof(double), cudaMemcpyDeviceToHost));
  if (aux && DMS_NAUX) CK(cudaMemcpy(aux, oa, nB * DMS_NAUX * sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(status, os, nB * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}
}
