// C-ABI + persistent solve kernel of the batched NLP solver (see include/dompc_b200.h).
// Compiled once per generated problem (problem_gen.h on the include path) for sm_100a.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <new>

#include "ipm_solve.cuh"
#include "ipm_sens.cuh"
#include "dmb_host_util.h"
#include "dompc_b200.h"

#ifndef DMB_THREADS
#define DMB_THREADS 128
#endif
#ifndef DMB_VEC_SMEM
#define DMB_VEC_SMEM 1
#endif
#ifndef DMB_ERN_SMEM
#define DMB_ERN_SMEM 1
#endif
#define DMB_NWARP (DMB_THREADS / 32)
#define DMB_NGRP (DMB_THREADS / GSZ)
#ifndef DMB_MIN_CTAS
#define DMB_MIN_CTAS 1
#endif

// ---- shared memory carve-up (doubles first, then the small struct) ----------------------------------------
// tiles of the lane groups | Riccati scratch | per-edge reduced blocks | per-node data | iterate vectors (optional)
#define SM_TILE 0
#define SM_QS (SM_TILE + DMB_NGRP * TILE_SZ)
#define SM_ER (SM_QS + DMB_QS_SIZE)
#define ERN_DOUBLES (NEDGE * ER_SZ + NNODE * WN_NODE)
#define SM_NODE (SM_ER + NEDGE * ER_SZ)
#define SM_VEC (SM_ER + (DMB_ERN_SMEM ? ERN_DOUBLES : 0))
#define SM_DOUBLES ((SM_VEC + (DMB_VEC_SMEM ? WV_END : 0) + 1) & ~1)
#define SM_BYTES (SM_DOUBLES * sizeof(double) + sizeof(DmbShared) + 16)
// per-CTA global workspace slot: bulk per-edge data (+ the iterate vectors when they do not fit on chip)
#define SLOT_VEC 0
#define SLOT_ERN (SLOT_VEC + (DMB_VEC_SMEM ? 0 : WV_END))
#define SLOT_EB ((SLOT_ERN + (DMB_ERN_SMEM ? 0 : ERN_DOUBLES) + 1) & ~1)      // even: edge blocks 16-byte aligned (TMA)
#define SLOT_DOUBLES (SLOT_EB + NEDGE * EB_SZ)
static_assert(DMB_NGRP <= 16, "DmbShared holds 16 TMA barriers");
static_assert(7 * DMB_NWARP <= 64, "DmbShared::red holds the 7 values of the largest fused reduction for at most 9 warps");
static_assert(SLOT_DOUBLES % 2 == 0 && EB_SZ % 2 == 0 && TILE_SZ % 2 == 0, "TMA bulk copies need 16-byte aligned tiles and edge blocks");

struct KernelArgs {
  int B;
  const double *x_init, *x0s, *u0s, *z0s, *p, *lbI, *ubI, *lam_g0, *lam_x0;
  double *x, *lam_g, *lam_x, *g, *f, *u0, *stats, *ws;
  int *status, *iters, *counter;
  DmbOptions o;
};

__global__ void __launch_bounds__(DMB_THREADS, DMB_MIN_CTAS) dmb_solve_kernel(const KernelArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  DmbShared* S = reinterpret_cast<DmbShared*>(sm + SM_DOUBLES);
  __shared__ int cur;
  __shared__ DmbResult result;
  // the block size is the compile-time DMB_THREADS (dmb_launch); the thread index is read ONCE: under register pressure the
  // compiler otherwise re-reads %tid.x (S2R, ~6 % of the stall samples in the r2y profile) wherever c.tid / c.lane / c.gl is used
  unsigned tid_;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_));
  Ctx c{(int)tid_, (int)DMB_THREADS, (int)(tid_ & 31), (int)(tid_ >> 5), (int)(DMB_THREADS >> 5),
        (int)(tid_ % GSZ), (int)(tid_ / GSZ), (int)(DMB_THREADS / GSZ)};
  unsigned bid_;
  asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bid_));
  double* slot = a.ws + (size_t)bid_ * SLOT_DOUBLES;
  Inst I;
  bind_vectors(I, DMB_VEC_SMEM ? sm + SM_VEC : slot + SLOT_VEC);
  I.er = DMB_ERN_SMEM ? sm + SM_ER : slot + SLOT_ERN;
  I.node = I.er + (size_t)NEDGE * ER_SZ;
  I.eb = slot + SLOT_EB;
  I.lb = a.lbI; I.ub = a.ubI; I.df = 1.0; I.sub = nullptr;
  I.relax_fac = a.o.bound_relax_factor; I.relax_max = a.o.constr_viol_tol;
  I.mbar = S->mbar; I.mphase = S->mphase;
  if (c.tid == 0) {
    for (int i = 0; i < DMB_NGRP; ++i) { mbar_init(S->mbar + i, 1); S->mphase[i] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  while (true) {
    if (c.tid == 0) cur = atomicAdd(a.counter, 1);
    __syncthreads();
    const int b = cur;
    __syncthreads();
    if (b >= a.B) break;
    I.p = a.p + (size_t)b * DMB_NP;
    load_initial_guess(c, a.x_init ? a.x_init + (size_t)b * DMB_NW : nullptr, a.x0s ? a.x0s + (size_t)b * NX : nullptr,
                       a.u0s ? a.u0s + (size_t)b * NU : nullptr, a.z0s ? a.z0s + (size_t)b * DMB_NZ : nullptr, I.wt);
    solve_instance(c, I, a.o, I.wt, sm + SM_TILE, sm + SM_QS, S, &result, a.lam_g0 ? a.lam_g0 + (size_t)b * DMB_NG : nullptr,
                   a.lam_x0 ? a.lam_x0 + (size_t)b * DMB_NW : nullptr);
    const double df = result.df;
    if (a.x && a.x_init)
      for (int i = c.tid; i < DMB_NW; i += c.nthr) a.x[(size_t)b * DMB_NW + i] = a.x_init[(size_t)b * DMB_NW + i];
    __syncthreads();
    store_solution(c, I, df, a.x ? a.x + (size_t)b * DMB_NW : nullptr, a.lam_g ? a.lam_g + (size_t)b * DMB_NG : nullptr,
                   a.lam_x ? a.lam_x + (size_t)b * DMB_NW : nullptr, a.g ? a.g + (size_t)b * DMB_NG : nullptr);
    for (int i = c.tid; i < NU; i += c.nthr) a.u0[(size_t)b * NU + i] = I.w[NODE_OFF(0) + NX + i];
    if (c.tid == 0) {
      a.f[b] = result.f; a.status[b] = result.status; a.iters[b] = result.iters;
      if (a.stats) {
        a.stats[(size_t)b * DMB_NSTAT + DMB_STAT_MU] = result.mu;
        a.stats[(size_t)b * DMB_NSTAT + DMB_STAT_INF_PR] = result.inf_pr;
        a.stats[(size_t)b * DMB_NSTAT + DMB_STAT_INF_DU] = result.inf_du;
        a.stats[(size_t)b * DMB_NSTAT + DMB_STAT_OBJ_SCALE] = result.df;
        for (int i = 0; i < DMB_NPHASE; ++i) a.stats[(size_t)b * DMB_NSTAT + DMB_STAT_PHASE0 + i] = (double)result.phase_cycles[i];
        a.stats[(size_t)b * DMB_NSTAT + DMB_STAT_N_RESTO] = result.n_resto; a.stats[(size_t)b * DMB_NSTAT + DMB_STAT_N_RESTO_IT] = result.n_resto_it;
        a.stats[(size_t)b * DMB_NSTAT + DMB_STAT_N_SOC] = result.n_soc; a.stats[(size_t)b * DMB_NSTAT + DMB_STAT_N_REFINE] = result.n_refine;
        a.stats[(size_t)b * DMB_NSTAT + DMB_STAT_N_LS] = result.n_ls;
      }
    }
    __syncthreads();
  }
}

// ---- parametric sensitivities of solved instances (see ipm_sens.cuh) --------------------------------------------
struct SensArgs {
  int B, mode, ncols;
  const double *x, *lam_g, *p, *lbI, *ubI, *lb0, *ub0;
  const int *cols, *status;
  int rule;
  double *dx, *dlam, *ws;
  int *sstat, *counter;
  double tol, big, ratio_max;
  int max_refine;
  double relax_fac, relax_max;
};

__global__ void __launch_bounds__(DMB_THREADS, DMB_MIN_CTAS) dmb_sens_kernel(const SensArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  DmbShared* S = reinterpret_cast<DmbShared*>(sm + SM_DOUBLES);
  __shared__ int cur;
  // the block size is the compile-time DMB_THREADS (dmb_launch); the thread index is read ONCE: under register pressure the
  // compiler otherwise re-reads %tid.x (S2R, ~6 % of the stall samples in the r2y profile) wherever c.tid / c.lane / c.gl is used
  unsigned tid_;
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_));
  Ctx c{(int)tid_, (int)DMB_THREADS, (int)(tid_ & 31), (int)(tid_ >> 5), (int)(DMB_THREADS >> 5),
        (int)(tid_ % GSZ), (int)(tid_ / GSZ), (int)(DMB_THREADS / GSZ)};
  unsigned bid_;
  asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bid_));
  double* slot = a.ws + (size_t)bid_ * SLOT_DOUBLES;
  Inst I;
  bind_vectors(I, DMB_VEC_SMEM ? sm + SM_VEC : slot + SLOT_VEC);
  I.er = DMB_ERN_SMEM ? sm + SM_ER : slot + SLOT_ERN;
  I.node = I.er + (size_t)NEDGE * ER_SZ;
  I.eb = slot + SLOT_EB;
  I.lb = a.lbI; I.ub = a.ubI; I.df = 1.0; I.sub = nullptr;
  I.relax_fac = a.relax_fac; I.relax_max = a.relax_max;
  I.mbar = S->mbar; I.mphase = S->mphase;
  if (c.tid == 0) {
    for (int i = 0; i < DMB_NGRP; ++i) { mbar_init(S->mbar + i, 1); S->mphase[i] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t nrow = a.mode == 1 ? NU : DMB_NW;
  while (true) {
    if (c.tid == 0) cur = atomicAdd(a.counter, 1);
    __syncthreads();
    const int b = cur;
    __syncthreads();
    if (b >= a.B) break;
    I.p = a.p + (size_t)b * DMB_NP;
    int st;
    if (a.status && a.status[b] != DMB_SOLVE_SUCCEEDED && a.status[b] != DMB_SOLVED_TO_ACCEPTABLE_LEVEL) st = DMB_SENS_SKIPPED;
    else st = sens_factor(c, I, a.x + (size_t)b * DMB_NW, a.lam_g + (size_t)b * DMB_NG, a.rule ? a.lbI : a.lb0, a.rule ? a.ubI : a.ub0, a.rule, a.tol, a.big, sm + SM_TILE, sm + SM_QS, S);
    for (int jc = 0; jc < a.ncols; ++jc) {
      double* dx = a.dx + ((size_t)b * a.ncols + jc) * nrow;
      double* dl = (a.dlam && a.mode != 1) ? a.dlam + ((size_t)b * a.ncols + jc) * (DMB_NG + DMB_NW) : nullptr;
      if (st >= DMB_SENS_FAILED) {
        for (int i = c.tid; i < (int)nrow; i += c.nthr) dx[i] = nan("");
        if (dl) for (int i = c.tid; i < DMB_NG + DMB_NW; i += c.nthr) dl[i] = nan("");
        continue;
      }
      sens_rhs(c, I, a.cols ? a.cols[jc] : jc);
      sens_solve(c, I, a.max_refine, a.ratio_max, sm + SM_TILE, sm + SM_QS, S);
      sens_store(c, I, a.mode, dx, dl);
    }
    if (c.tid == 0 && a.sstat) a.sstat[b] = st;
    __syncthreads();
  }
}

// broadcast-free fill of x for the cold-start path when the caller wants the full primal vector:
// unused reference variables keep the broadcast initial guess (reference _mpc.py:969-971)
__global__ void dmb_fill_guess_kernel(int B, const double* x0s, const double* u0s, const double* z0s, double* x) {
  const int S = DMB_NW;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)B * S; idx += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / S), i = (int)(idx % S);
    const int nxall = (DMB_NHOR + 1) * DMB_NSCEN * (DMB_NPC + 1) * NX;
    const int nzall = DMB_NHOR * DMB_NSCEN * DMB_NPZ * DMB_NZ;
    const int nuall = DMB_NHOR * DMB_NSCEN * NU;
    double v = 0.0;
    if (i < nxall) v = x0s[(size_t)b * NX + i % NX];
    else if (i < nxall + nzall) v = z0s ? z0s[(size_t)b * DMB_NZ + (i - nxall) % (DMB_NZ > 0 ? DMB_NZ : 1)] : 0.0;
    else if (i < nxall + nzall + nuall) v = u0s[(size_t)b * NU + (i - nxall - nzall) % NU];
    x[idx] = v;
  }
}

// aux expressions for every (k, s): thread per (instance, k, s)
__global__ void dmb_aux_kernel(int B, const double* x, const double* p, double* aux) {
  const int per = DMB_NHOR * DMB_NSCEN;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)B * per; idx += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / per), ks = (int)(idx % per);
    // node (k, s); scenario slots beyond the populated ones repeat the last one (reference _mpc.py:1282-1284)
    const int n = DMB_KS_NODE[ks];
    const int e = NODE_FIRST_EDGE(n) + NODE_NEDGE(n) - 1;
    const double* xb = x + (size_t)b * DMB_NW;
    const double* pb = p + (size_t)b * DMB_NP;
    double q[NQ], a[DMB_NAUX > 0 ? DMB_NAUX : 1];
    for (int i = 0; i < NVP; ++i) { const int fl = DMB_W2FLAT[NODE_OFF(n) + i]; q[i] = fl >= 0 ? xb[fl] : 0.0; }
    for (int i = 0; i < NY; ++i) { const int fl = DMB_W2FLAT[EDGE_OFF(e) + i]; q[NVP + i] = fl >= 0 ? xb[fl] : 0.0; }
    gen_aux(q, edge_tvp(pb, e), edge_pp(pb, e), a);
    for (int i = 0; i < DMB_NAUX; ++i) aux[((size_t)b * per + ks) * DMB_NAUX + i] = a[i];
  }
}

// ---------------------------------------------------------------------------------------------------
struct dmb_solver {
  int device = 0, n_sm = 0, grid_max = 0, have_bounds = 0;
  DmbOptions opts;
  double *d_lbI = nullptr, *d_ubI = nullptr, *d_lb0 = nullptr, *d_ub0 = nullptr, *d_ws = nullptr;
  int* d_counter = nullptr;
  int* d_cols = nullptr; int cols_cap = 0;
  double sens_tol = 1e-6, sens_big = 1e10; int sens_rule = 0;
  size_t ws_slots = 0;
  // staging for the host entry point
  double* d_stage = nullptr; size_t stage_doubles = 0;
  double* h_stage = nullptr; size_t h_stage_doubles = 0;
  int last_grid = 0; long long n_launches = 0; float last_ms = 0.f;
  size_t l2_window_bytes = 0; float l2_hit_ratio = 1.f;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { s->err = std::string(#call) + ": " + cudaGetErrorString(e_); return DMB_E_CUDA; } } while (0)

extern "C" {

int dmb_abi_version(void) { return DMB_ABI_VERSION; }

void dmb_get_dims(int* d) {
  d[DMB_DIM_NW] = DMB_NW; d[DMB_DIM_NG] = DMB_NG; d[DMB_DIM_NP] = DMB_NP; d[DMB_DIM_NX] = NX; d[DMB_DIM_NU] = NU;
  d[DMB_DIM_NZ] = DMB_NZ; d[DMB_DIM_NAUX] = DMB_NAUX; d[DMB_DIM_NHOR] = DMB_NHOR; d[DMB_DIM_NSCEN] = DMB_NSCEN;
  d[DMB_DIM_NEDGE] = NEDGE; d[DMB_DIM_NNODE] = NNODE; d[DMB_DIM_OPTSIZE] = (int)sizeof(DmbOptions);
  d[DMB_DIM_THREADS] = DMB_THREADS; d[DMB_DIM_SMEM] = (int)SM_BYTES; d[DMB_DIM_WS_DOUBLES] = (int)SLOT_DOUBLES;
}

void dmb_default_options(DmbOptions* o) { dmb_fill_default_options(o); }

dmb_solver* dmb_create(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device >= n) return nullptr;
  dmb_solver* s = new (std::nothrow) dmb_solver();
  if (!s) return nullptr;
  s->device = device;
  dmb_fill_default_options(&s->opts);
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete s; return nullptr; }
  s->n_sm = prop.multiProcessorCount;
  if (cudaFuncSetAttribute(dmb_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_BYTES) != cudaSuccess) { delete s; return nullptr; }
  if (cudaFuncSetAttribute(dmb_sens_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_BYTES) != cudaSuccess) { delete s; return nullptr; }
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dmb_solve_kernel, DMB_THREADS, SM_BYTES) != cudaSuccess || per_sm < 1) { delete s; return nullptr; }
  s->grid_max = per_sm * s->n_sm;
  bool ok = cudaMalloc(&s->d_lbI, sizeof(double) * NWI) == cudaSuccess && cudaMalloc(&s->d_ubI, sizeof(double) * NWI) == cudaSuccess &&
            cudaMalloc(&s->d_lb0, sizeof(double) * NWI) == cudaSuccess && cudaMalloc(&s->d_ub0, sizeof(double) * NWI) == cudaSuccess &&
            cudaMalloc(&s->d_counter, sizeof(int)) == cudaSuccess &&
            cudaMalloc(&s->d_ws, sizeof(double) * (size_t)SLOT_DOUBLES * s->grid_max) == cudaSuccess &&
            cudaEventCreate(&s->ev0) == cudaSuccess && cudaEventCreate(&s->ev1) == cudaSuccess;
  if (!ok) { dmb_destroy(s); return nullptr; }
  s->ws_slots = s->grid_max;
  {
    // optional L2 persistence window over the workspace slots (DMB_L2_PERSIST=1).  Measured on B200, cfg #2: same kernel time
    // (27.2 ms) and MORE DRAM write-back with it (4.49 GB vs 2.27 GB per launch, gpurun round 2) -> off by default
    const char* env = getenv("DMB_L2_PERSIST");
    const size_t ws_bytes = sizeof(double) * (size_t)SLOT_DOUBLES * s->grid_max;
    if ((env && env[0] == '1') && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
      const size_t carve = ws_bytes < (size_t)prop.persistingL2CacheMaxSize ? ws_bytes : (size_t)prop.persistingL2CacheMaxSize;
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
        s->l2_window_bytes = ws_bytes < (size_t)prop.accessPolicyMaxWindowSize ? ws_bytes : (size_t)prop.accessPolicyMaxWindowSize;
        s->l2_hit_ratio = ws_bytes <= carve ? 1.f : (float)((double)carve / (double)ws_bytes);
      } else (void)cudaGetLastError();
    }
  }
  return s;
}

void dmb_destroy(dmb_solver* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  cudaFree(s->d_lbI); cudaFree(s->d_ubI); cudaFree(s->d_lb0); cudaFree(s->d_ub0); cudaFree(s->d_counter); cudaFree(s->d_ws); cudaFree(s->d_stage); cudaFree(s->d_cols);
  if (s->h_stage) cudaFreeHost(s->h_stage);
  if (s->ev0) cudaEventDestroy(s->ev0);
  if (s->ev1) cudaEventDestroy(s->ev1);
  delete s;
}

const char* dmb_last_error(dmb_solver* s) { return s ? s->err.c_str() : "null solver handle"; }

int dmb_set_options(dmb_solver* s, const DmbOptions* o, int size) {
  if (!s || !o || size != (int)sizeof(DmbOptions)) return DMB_E_ARG;
  s->opts = *o;
  s->have_bounds = 0;     // relaxed bounds depend on the options
  return DMB_E_OK;
}

int dmb_set_bounds(dmb_solver* s, const double* lbx, const double* ubx) {
  if (!s || !lbx || !ubx) return DMB_E_ARG;
  CK(cudaSetDevice(s->device));
  static thread_local double lbI[NWI], ubI[NWI];
  dmb_internal_bounds(lbx, ubx, &s->opts, lbI, ubI);
  CK(cudaMemcpy(s->d_lbI, lbI, sizeof(double) * NWI, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(s->d_ubI, ubI, sizeof(double) * NWI, cudaMemcpyHostToDevice));
  DmbOptions o0 = s->opts; o0.bound_relax_factor = 0.0;          // the user's bounds, for the reference's active-set test
  dmb_internal_bounds(lbx, ubx, &o0, lbI, ubI);
  CK(cudaMemcpy(s->d_lb0, lbI, sizeof(double) * NWI, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(s->d_ub0, ubI, sizeof(double) * NWI, cudaMemcpyHostToDevice));
  s->have_bounds = 1;
  return DMB_E_OK;
}

int dmb_solve_device_ws(dmb_solver* s, int B, const double* d_x_init, const double* d_lam_g0, const double* d_lam_x0, const double* d_x0s,
                        const double* d_u0s, const double* d_z0s, const double* d_p, double* d_x, double* d_lam_g, double* d_lam_x,
                        double* d_g, double* d_f, int* d_status, int* d_iters, double* d_u0, double* d_stats, void* stream_) {
  if (!s) return DMB_E_ARG;
  if ((d_lam_g0 || d_lam_x0) && (!d_lam_g0 || !d_lam_x0 || !d_x_init)) { s->err = "warm start needs x_init, lam_g0 and lam_x0 together"; return DMB_E_ARG; }
  if ((d_lam_g0 && d_lam_g0 == d_lam_g) || (d_lam_x0 && d_lam_x0 == d_lam_x)) { s->err = "warm-start multipliers must not alias the outputs"; return DMB_E_ARG; }
  if (B <= 0 || !d_p || !d_f || !d_status || !d_iters || !d_u0 || (!d_x_init && (!d_x0s || !d_u0s)) ||
      (!d_x_init && DMB_NZ > 0 && !d_z0s)) { s->err = "invalid argument"; return DMB_E_ARG; }
  if (!s->have_bounds) { s->err = "dmb_set_bounds must be called before solving"; return DMB_E_STATE; }
  cudaStream_t st = (cudaStream_t)stream_;
  CK(cudaSetDevice(s->device));
  CK(cudaMemsetAsync(s->d_counter, 0, sizeof(int), st));
  KernelArgs a;
  a.lam_g0 = d_lam_g0; a.lam_x0 = d_lam_x0;
  a.B = B; a.x_init = d_x_init; a.x0s = d_x0s; a.u0s = d_u0s; a.z0s = d_z0s; a.p = d_p; a.lbI = s->d_lbI; a.ubI = s->d_ubI;
  a.x = d_x; a.lam_g = d_lam_g; a.lam_x = d_lam_x; a.g = d_g; a.f = d_f; a.u0 = d_u0; a.stats = d_stats; a.ws = s->d_ws;
  a.status = d_status; a.iters = d_iters; a.counter = s->d_counter; a.o = s->opts;
  const int grid = B < s->grid_max ? B : s->grid_max;
  if (d_x && !d_x_init) {
    dmb_fill_guess_kernel<<<s->n_sm * 4, 256, 0, st>>>(B, d_x0s, d_u0s, d_z0s, d_x);
    a.x_init = nullptr;
    s->n_launches++;
  }
  CK(cudaEventRecord(s->ev0, st));
  if (s->l2_window_bytes > 0) {
    // the per-CTA workspace slots (L2-resident scratch, rewritten every IPM iteration) are marked persisting for this launch so
    // that the L2 keeps the dirty lines instead of writing them back to DRAM between iterations (profiles/README.md: traffic)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(DMB_THREADS); cfg.dynamicSmemBytes = SM_BYTES; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeAccessPolicyWindow;
    at[0].val.accessPolicyWindow.base_ptr = s->d_ws;
    at[0].val.accessPolicyWindow.num_bytes = s->l2_window_bytes;
    at[0].val.accessPolicyWindow.hitRatio = s->l2_hit_ratio;
    at[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    at[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cfg.attrs = at; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, dmb_solve_kernel, a));
  } else
    dmb_solve_kernel<<<grid, DMB_THREADS, SM_BYTES, st>>>(a);
  CK(cudaEventRecord(s->ev1, st));
  CK(cudaGetLastError());
  s->last_grid = grid; s->n_launches++;
  return DMB_E_OK;
}

int dmb_solve_device(dmb_solver* s, int B, const double* d_x_init, const double* d_x0s, const double* d_u0s,
                     const double* d_z0s, const double* d_p, double* d_x, double* d_lam_g, double* d_lam_x, double* d_g,
                     double* d_f, int* d_status, int* d_iters, double* d_u0, double* d_stats, void* stream_) {
  return dmb_solve_device_ws(s, B, d_x_init, nullptr, nullptr, d_x0s, d_u0s, d_z0s, d_p, d_x, d_lam_g, d_lam_x, d_g, d_f, d_status, d_iters,
                             d_u0, d_stats, stream_);
}

static int ensure_stage(dmb_solver* s, size_t doubles) {
  if (doubles > s->stage_doubles) {
    cudaFree(s->d_stage); s->d_stage = nullptr; s->stage_doubles = 0;
    CK(cudaMalloc(&s->d_stage, doubles * sizeof(double)));
    s->stage_doubles = doubles;
  }
  return DMB_E_OK;
}

int dmb_solve_host(dmb_solver* s, int B, const double* x_init, const double* x0s, const double* u0s, const double* z0s,
                   const double* p, double* x, double* lam_g, double* lam_x, double* g, double* f, int* status,
                   int* iters, double* u0, double* stats) {
  return dmb_solve_host_ws(s, B, x_init, nullptr, nullptr, x0s, u0s, z0s, p, x, lam_g, lam_x, g, f, status, iters, u0, stats);
}

int dmb_solve_host_ws(dmb_solver* s, int B, const double* x_init, const double* lam_g0, const double* lam_x0, const double* x0s,
                      const double* u0s, const double* z0s, const double* p, double* x, double* lam_g, double* lam_x, double* g,
                      double* f, int* status, int* iters, double* u0, double* stats) {
  if (!s) return DMB_E_ARG;
  if (B <= 0 || !p || !f || !status || !iters || !u0) { s->err = "invalid argument"; return DMB_E_ARG; }
  CK(cudaSetDevice(s->device));
  const size_t nB = (size_t)B;
  const bool have_z = z0s && DMB_NZ > 0;
  // one staging image, identical on the host (pinned) and on the device: inputs first, outputs behind them, so that a call
  // costs ONE host->device and ONE device->host transfer from / to page-locked memory
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 1) & ~(size_t)1; return o; };
  const size_t o_p = take(nB * DMB_NP), o_xi = take(x_init ? nB * DMB_NW : 0), o_lg0 = take(lam_g0 ? nB * DMB_NG : 0),
               o_lx0 = take(lam_x0 ? nB * DMB_NW : 0), o_x0 = take(x0s ? nB * NX : 0), o_u0s = take(u0s ? nB * NU : 0),
               o_z0 = take(have_z ? nB * DMB_NZ : 0);
  const size_t n_in = off;
  const size_t o_x = take(x ? nB * DMB_NW : 0), o_lg = take(lam_g ? nB * DMB_NG : 0), o_lx = take(lam_x ? nB * DMB_NW : 0),
               o_g = take(g ? nB * DMB_NG : 0), o_f = take(nB), o_u = take(nB * NU), o_st = take(stats ? nB * DMB_NSTAT : 0),
               o_i1 = take(nB), o_i2 = take(nB);
  int rc = ensure_stage(s, off);
  if (rc) return rc;
  if (off > s->h_stage_doubles) {
    if (s->h_stage) cudaFreeHost(s->h_stage);
    s->h_stage = nullptr; s->h_stage_doubles = 0;
    CK(cudaMallocHost(&s->h_stage, off * sizeof(double)));
    s->h_stage_doubles = off;
  }
  double* d = s->d_stage; double* h = s->h_stage;
  auto put = [&](size_t o, const double* src, size_t n) { if (src && n) memcpy(h + o, src, n * sizeof(double)); };
  put(o_p, p, nB * DMB_NP); put(o_xi, x_init, nB * DMB_NW); put(o_lg0, lam_g0, nB * DMB_NG); put(o_lx0, lam_x0, nB * DMB_NW);
  put(o_x0, x0s, nB * NX); put(o_u0s, u0s, nB * NU); if (have_z) put(o_z0, z0s, nB * DMB_NZ);
  CK(cudaMemcpyAsync(d, h, n_in * sizeof(double), cudaMemcpyHostToDevice, 0));
  rc = dmb_solve_device_ws(s, B, x_init ? d + o_xi : nullptr, lam_g0 ? d + o_lg0 : nullptr, lam_x0 ? d + o_lx0 : nullptr, x0s ? d + o_x0 : nullptr, u0s ? d + o_u0s : nullptr,
                        have_z ? d + o_z0 : nullptr, d + o_p, x ? d + o_x : nullptr, lam_g ? d + o_lg : nullptr,
                        lam_x ? d + o_lx : nullptr, g ? d + o_g : nullptr, d + o_f, (int*)(d + o_i1), (int*)(d + o_i2),
                        d + o_u, stats ? d + o_st : nullptr, nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h + n_in, d + n_in, (off - n_in) * sizeof(double), cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  auto get = [&](double* dst, size_t o, size_t n) { if (dst && n) memcpy(dst, h + o, n * sizeof(double)); };
  get(x, o_x, nB * DMB_NW); get(lam_g, o_lg, nB * DMB_NG); get(lam_x, o_lx, nB * DMB_NW); get(g, o_g, nB * DMB_NG);
  get(f, o_f, nB); get(u0, o_u, nB * NU); get(stats, o_st, nB * DMB_NSTAT);
  memcpy(status, h + o_i1, nB * sizeof(int)); memcpy(iters, h + o_i2, nB * sizeof(int));
  cudaEventElapsedTime(&s->last_ms, s->ev0, s->ev1);
  return DMB_E_OK;
}

int dmb_aux_host(dmb_solver* s, int B, const double* x, const double* p, double* aux) {
  if (!s || B <= 0 || !x || !p || !aux) return DMB_E_ARG;
  CK(cudaSetDevice(s->device));
  const size_t nB = (size_t)B, na = nB * DMB_NHOR * DMB_NSCEN * (DMB_NAUX > 0 ? DMB_NAUX : 1);
  int rc = ensure_stage(s, nB * DMB_NW + nB * DMB_NP + na + 8);
  if (rc) return rc;
  double* d = s->d_stage;
  CK(cudaMemcpy(d, x, nB * DMB_NW * sizeof(double), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d + nB * DMB_NW, p, nB * DMB_NP * sizeof(double), cudaMemcpyHostToDevice));
  dmb_aux_kernel<<<(int)((nB * DMB_NHOR * DMB_NSCEN + 127) / 128), 128>>>(B, d, d + nB * DMB_NW, d + nB * DMB_NW + nB * DMB_NP);
  CK(cudaGetLastError());
  s->n_launches++;
  CK(cudaMemcpy(aux, d + nB * DMB_NW + nB * DMB_NP, nB * DMB_NHOR * DMB_NSCEN * DMB_NAUX * sizeof(double), cudaMemcpyDeviceToHost));
  return DMB_E_OK;
}

int dmb_set_sens_options(dmb_solver* s, double active_set_tol, double active_weight, int relaxed_bounds_rule) {
  if (!s || !(active_set_tol >= 0.0) || !(active_weight > 0.0)) return DMB_E_ARG;
  s->sens_tol = active_set_tol; s->sens_big = active_weight; s->sens_rule = relaxed_bounds_rule ? 1 : 0;
  return DMB_E_OK;
}

int dmb_sens_device(dmb_solver* s, int B, const double* d_x, const double* d_lam_g, const double* d_p, const int* d_status,
                    int mode, const int* cols, int ncols, double* d_dx, double* d_dlam, int* d_sens_status, void* stream_) {
  if (!s) return DMB_E_ARG;
  if (B <= 0 || !d_x || !d_lam_g || !d_p || !d_dx || (mode != DMB_SENS_FULL && mode != DMB_SENS_U0) || ncols < 0 ||
      (!cols && ncols != 0 && ncols != DMB_NP)) { s->err = "invalid argument"; return DMB_E_ARG; }
  if (!s->have_bounds) { s->err = "dmb_set_bounds must be called before dmb_sens_*"; return DMB_E_STATE; }
  cudaStream_t st = (cudaStream_t)stream_;
  CK(cudaSetDevice(s->device));
  if (!cols) ncols = DMB_NP;
  if (cols) {
    for (int i = 0; i < ncols; ++i) if (cols[i] < 0 || cols[i] >= DMB_NP) { s->err = "parameter column out of range"; return DMB_E_ARG; }
    if (ncols > s->cols_cap) { cudaFree(s->d_cols); s->d_cols = nullptr; s->cols_cap = 0; CK(cudaMalloc(&s->d_cols, sizeof(int) * ncols)); s->cols_cap = ncols; }
    CK(cudaMemcpyAsync(s->d_cols, cols, sizeof(int) * ncols, cudaMemcpyHostToDevice, st));
  }
  CK(cudaMemsetAsync(s->d_counter, 0, sizeof(int), st));
  SensArgs a;
  a.B = B; a.mode = mode; a.ncols = ncols; a.x = d_x; a.lam_g = d_lam_g; a.p = d_p; a.lbI = s->d_lbI; a.ubI = s->d_ubI;
  a.lb0 = s->d_lb0; a.ub0 = s->d_ub0; a.rule = s->sens_rule;
  a.cols = cols ? s->d_cols : nullptr; a.status = d_status; a.dx = d_dx; a.dlam = d_dlam; a.ws = s->d_ws; a.sstat = d_sens_status;
  a.counter = s->d_counter; a.tol = s->sens_tol; a.big = s->sens_big; a.ratio_max = s->opts.residual_ratio_max;
  a.max_refine = s->opts.max_refine; a.relax_fac = s->opts.bound_relax_factor; a.relax_max = s->opts.constr_viol_tol;
  const int grid = B < s->grid_max ? B : s->grid_max;
  CK(cudaEventRecord(s->ev0, st));
  dmb_sens_kernel<<<grid, DMB_THREADS, SM_BYTES, st>>>(a);
  CK(cudaEventRecord(s->ev1, st));
  CK(cudaGetLastError());
  s->last_grid = grid; s->n_launches++;
  return DMB_E_OK;
}

int dmb_sens_host(dmb_solver* s, int B, const double* x, const double* lam_g, const double* p, const int* status, int mode,
                  const int* cols, int ncols, double* dx, double* dlam, int* sens_status) {
  if (!s) return DMB_E_ARG;
  if (B <= 0 || !x || !lam_g || !p || !dx) { s->err = "invalid argument"; return DMB_E_ARG; }
  CK(cudaSetDevice(s->device));
  const size_t nB = (size_t)B, nc = cols ? (size_t)ncols : (size_t)DMB_NP, nrow = mode == DMB_SENS_U0 ? NU : DMB_NW;
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += (n + 1) & ~(size_t)1; return o; };
  const size_t o_x = take(nB * DMB_NW), o_l = take(nB * DMB_NG), o_p = take(nB * DMB_NP), o_dx = take(nB * nc * nrow),
               o_dl = take((dlam && mode != DMB_SENS_U0) ? nB * nc * (DMB_NG + DMB_NW) : 0), o_st = take(status ? nB : 0), o_ss = take(nB);
  int rc = ensure_stage(s, off);
  if (rc) return rc;
  double* d = s->d_stage;
  CK(cudaMemcpyAsync(d + o_x, x, nB * DMB_NW * sizeof(double), cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(d + o_l, lam_g, nB * DMB_NG * sizeof(double), cudaMemcpyHostToDevice, 0));
  CK(cudaMemcpyAsync(d + o_p, p, nB * DMB_NP * sizeof(double), cudaMemcpyHostToDevice, 0));
  if (status) CK(cudaMemcpyAsync(d + o_st, status, nB * sizeof(int), cudaMemcpyHostToDevice, 0));
  const bool want_dl = dlam && mode != DMB_SENS_U0;
  rc = dmb_sens_device(s, B, d + o_x, d + o_l, d + o_p, status ? (const int*)(d + o_st) : nullptr, mode, cols, ncols, d + o_dx,
                       want_dl ? d + o_dl : nullptr, (int*)(d + o_ss), nullptr);
  if (rc) return rc;
  CK(cudaMemcpyAsync(dx, d + o_dx, nB * nc * nrow * sizeof(double), cudaMemcpyDeviceToHost, 0));
  if (want_dl) CK(cudaMemcpyAsync(dlam, d + o_dl, nB * nc * (DMB_NG + DMB_NW) * sizeof(double), cudaMemcpyDeviceToHost, 0));
  if (sens_status) CK(cudaMemcpyAsync(sens_status, d + o_ss, nB * sizeof(int), cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  cudaEventElapsedTime(&s->last_ms, s->ev0, s->ev1);
  return DMB_E_OK;
}

int dmb_get_stats(dmb_solver* s, int* grid, int* block, int* smem_bytes, long long* n_launches, float* last_kernel_ms) {
  if (!s) return DMB_E_ARG;
  if (grid) *grid = s->last_grid;
  if (block) *block = DMB_THREADS;
  if (smem_bytes) *smem_bytes = (int)SM_BYTES;
  if (n_launches) *n_launches = s->n_launches;
  if (last_kernel_ms) {
    // events that were never recorded make cudaEventElapsedTime fail with cudaErrorInvalidResourceHandle and that error would be
    // picked up by the cudaGetLastError() check of the NEXT launch: only query after a launch, and never leave an error behind
    if (s->n_launches > 0 && cudaEventSynchronize(s->ev1) == cudaSuccess && cudaEventElapsedTime(&s->last_ms, s->ev0, s->ev1) == cudaSuccess) *last_kernel_ms = s->last_ms;
    else { (void)cudaGetLastError(); *last_kernel_ms = 0.f; }
  }
  return DMB_E_OK;
}

}  // extern "C"
