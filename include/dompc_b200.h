/* dompc_b200.h — C ABI of the B200 batched NLP solver that stands in for CasADi's nlpsol('ipopt') Function
 * on do-mpc's controller path.
 *
 * Every generated problem library (one shared object per MPC configuration, built by
 * do_mpc_b200/solver/build.py from csrc/dmb_api.cu + the generated problem_gen.h) exports exactly
 * these symbols.  Plain pointers and sizes only; no torch / CasADi types.
 *
 * Reference interfaces replaced (file:line in the do-mpc tree):
 *   dmb_create / dmb_destroy     <-> castools.nlpsol('S','ipopt',nlp,opts)      do_mpc/controller/_mpc.py:1326-1328
 *   dmb_set_options              <-> settings.nlpsol_opts passed to nlpsol       do_mpc/controller/_mpc.py:1328
 *   dmb_set_bounds               <-> lbx/ubx arguments of S(...)                 do_mpc/optimizer.py:756-757
 *   dmb_solve_host               <-> r = self.S(x0=,lbx=,ubx=,lbg=,ubg=,p=)      do_mpc/optimizer.py:754-777
 *                                    (B instances per call; x/g/lam_g/lam_x/f as in the returned dict)
 *   dmb_solve_device             <-> same call with device-resident batches (no reference counterpart:
 *                                    the reference is single-instance, host only)
 *   dmb_get_stats                <-> S.stats()                                    do_mpc/optimizer.py:778
 *   dmb_aux_host                 <-> opt_aux_expression_fun(opt_x, opt_p)         do_mpc/controller/_mpc.py:1331
 *   dmb_sens_host / _device      <-> NLPDifferentiator.differentiate(nlp_sol, p)  do_mpc/differentiator/_nlpdifferentiator.py:682-741
 *                                    (d x_opt / d p and d lam / d p of solved instances; active_set_tol as settings.active_set_tol)
 *
 * Layouts: all vectors are the reference's flat vectors: opt_x (n_w, scaled
 * variables), opt_p (n_p), g / lam_g (n_g), row-major over the batch: element i of instance b at [b*n + i].
 * Equality rows have lbg = ubg = 0 and inequality rows -inf <= g <= ub exactly as the MPC builds them, so
 * lbg/ubg are compiled into the library and not passed.
 *
 * Errors: functions return 0 on success, a negative DMB_E_* code otherwise (no exceptions, no aborts).
 * A failed NLP solve is NOT an error: it is reported per instance in `status` (IPOPT ApplicationReturnStatus
 * numbering), like the reference, which only exposes it through stats()['success'] (optimizer.py:778).
 * Threading: one solver handle may be used by one host thread at a time; calls are blocking unless a
 * stream is given to dmb_solve_device.
 */
#ifndef DOMPC_B200_H
#define DOMPC_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define DMB_ABI_VERSION 1

#define DMB_E_OK 0
#define DMB_E_CUDA (-1)        /* a CUDA runtime call failed; see dmb_last_error */
#define DMB_E_ARG (-2)         /* invalid argument (null pointer, batch <= 0, wrong option struct size) */
#define DMB_E_NO_DEVICE (-3)   /* no usable sm_100 device */
#define DMB_E_STATE (-4)       /* bounds not set before solve */

/* indices into the array filled by dmb_get_dims */
enum { DMB_DIM_NW = 0, DMB_DIM_NG, DMB_DIM_NP, DMB_DIM_NX, DMB_DIM_NU, DMB_DIM_NZ, DMB_DIM_NAUX, DMB_DIM_NHOR,
       DMB_DIM_NSCEN, DMB_DIM_NEDGE, DMB_DIM_NNODE, DMB_DIM_OPTSIZE, DMB_DIM_THREADS, DMB_DIM_SMEM, DMB_DIM_WS_DOUBLES,
       DMB_DIM_COUNT };

/* per-instance statistics written by the solve calls: stats[b*DMB_NSTAT + k] */
enum { DMB_STAT_MU = 0, DMB_STAT_INF_PR, DMB_STAT_INF_DU, DMB_STAT_OBJ_SCALE,
       DMB_STAT_PHASE0 /* 9 + 8 SM-clock cycle counters of the solver (sub)phases, see DmbResult */,
       DMB_STAT_N_RESTO = DMB_STAT_PHASE0 + 17 /* restoration phases entered */, DMB_STAT_N_RESTO_IT /* restoration iterations */,
       DMB_STAT_N_SOC /* second-order-correction solves */, DMB_STAT_N_REFINE /* iterative-refinement solves */,
       DMB_STAT_N_LS /* line-search trial points */, DMB_NSTAT };

typedef struct dmb_solver dmb_solver;
struct DmbOptions;   /* layout: do_mpc_b200/csrc/ipm_core.cuh, mirrored by do_mpc_b200/solver/options.py */

int dmb_abi_version(void);
void dmb_get_dims(int* dims /* [DMB_DIM_COUNT] */);
void dmb_default_options(struct DmbOptions* opts);

dmb_solver* dmb_create(int device);
void dmb_destroy(dmb_solver* s);
const char* dmb_last_error(dmb_solver* s);

int dmb_set_options(dmb_solver* s, const struct DmbOptions* opts, int opts_size);
int dmb_set_bounds(dmb_solver* s, const double* lbx, const double* ubx /* host, n_w each, +-inf allowed */);

/* Initial guess: either x_init [B, n_w] (scaled opt_x, as MPC.opt_x_num) or, when x_init is NULL, the
 * broadcast of x0s [B,n_x], u0s [B,n_u], z0s [B,n_z] (scaled) that MPC.set_initial_guess() performs
 * (do_mpc/controller/_mpc.py:969-971).  Outputs x, lam_g, lam_x, g, stats may be NULL. */
int dmb_solve_device(dmb_solver* s, int B, const double* d_x_init, const double* d_x0s, const double* d_u0s,
                     const double* d_z0s, const double* d_p, double* d_x, double* d_lam_g, double* d_lam_x, double* d_g,
                     double* d_f, int* d_status, int* d_iters, double* d_u0, double* d_stats, void* cuda_stream);

int dmb_solve_host(dmb_solver* s, int B, const double* x_init, const double* x0s, const double* u0s, const double* z0s,
                   const double* p, double* x, double* lam_g, double* lam_x, double* g, double* f, int* status,
                   int* iters, double* u0, double* stats);

/* The same calls with the multipliers of a previous solution (`lam_g0` [B,n_g], `lam_x0` [B,n_w], reference sign convention)
 * — what optimizer.solve() passes as lam_g0= / lam_x0= (do_mpc/optimizer.py:760-768).  They are used only when the option
 * warm_start is set (IPOPT: warm_start_init_point yes, with warm_start_bound_push / _frac / _mult_bound_push / _mult_init_max
 * and mu_init), otherwise ignored exactly like IPOPT ignores them.  Require x_init. */
int dmb_solve_device_ws(dmb_solver* s, int B, const double* d_x_init, const double* d_lam_g0, const double* d_lam_x0, const double* d_x0s,
                        const double* d_u0s, const double* d_z0s, const double* d_p, double* d_x, double* d_lam_g, double* d_lam_x,
                        double* d_g, double* d_f, int* d_status, int* d_iters, double* d_u0, double* d_stats, void* cuda_stream);
int dmb_solve_host_ws(dmb_solver* s, int B, const double* x_init, const double* lam_g0, const double* lam_x0, const double* x0s,
                      const double* u0s, const double* z0s, const double* p, double* x, double* lam_g, double* lam_x, double* g,
                      double* f, int* status, int* iters, double* u0, double* stats);

/* aux expressions at every (k, scenario): aux [B, n_horizon * n_scenarios * n_aux] from x [B,n_w], p [B,n_p] (host) */
int dmb_aux_host(dmb_solver* s, int B, const double* x, const double* p, double* aux);

/* Parametric sensitivities of solved instances (call after a solve, with its x / lam_g / p / status).
 * mode DMB_SENS_FULL: dx [B, ncols, n_w] = d opt_x(scaled) / d opt_p[col], dlam [B, ncols, n_g + n_w] (lam_g | lam_x) or NULL;
 * mode DMB_SENS_U0  : dx [B, ncols, n_u] = d u_0(scaled) / d opt_p[col], dlam ignored.
 * cols: HOST array of opt_p indices (NULL = all n_p columns).  status: solver status per instance or NULL; instances that
 * were not solved get NaN and sens_status DMB_SENS_SKIPPED.  sens_status: 0 ok, 1 ok but the reduced Hessian
 * on the active set is indefinite, 2 failed (singular system), 3 skipped. */
#define DMB_SENS_FULL 0
#define DMB_SENS_U0 1
/* relaxed_bounds_rule = 0 (default): a bound is active when |x - bound| <= active_set_tol for the USER's bound, exactly the
 * reference's test (_nlpdifferentiator.py:345-375); 1: measured against the relaxed bound the interior-point method
 * worked with (bound_relax_factor) — differs when |bound| * 1e-8 > active_set_tol, i.e. for unscaled bounds > 100. */
int dmb_set_sens_options(dmb_solver* s, double active_set_tol /* 1e-6 */, double active_weight /* 1e10 */, int relaxed_bounds_rule);
int dmb_sens_device(dmb_solver* s, int B, const double* d_x, const double* d_lam_g, const double* d_p, const int* d_status,
                    int mode, const int* cols, int ncols, double* d_dx, double* d_dlam, int* d_sens_status, void* cuda_stream);
int dmb_sens_host(dmb_solver* s, int B, const double* x, const double* lam_g, const double* p, const int* status, int mode,
                  const int* cols, int ncols, double* dx, double* dlam, int* sens_status);

/* launch geometry of the last solve and number of kernel launches issued by this handle so far */
int dmb_get_stats(dmb_solver* s, int* grid, int* block, int* smem_bytes, long long* n_launches, float* last_kernel_ms);

#ifdef __cplusplus
}
#endif
#endif
