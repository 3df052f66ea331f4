/* dompc_b200_sim.h — C ABI of the batched plant simulator (first "next" row of SURVEY.md §8f).
 *
 * Stands in for the CasADi integrator / rootfinder objects of the reference's simulator:
 *   dms_create / dms_destroy <-> castools.integrator('simulator','cvodes'|'idas',dae,opts)   do_mpc/simulator.py:404-416
 *   dms_step_host (mode 0)   <-> Simulator.simulate(): one step of length t_step                do_mpc/simulator.py:690-755
 *   dms_step_host (mode 1)   <-> Simulator.init_algebraic_variables(): consistent z for (x,u)   do_mpc/simulator.py:621-688
 * Physical (unscaled) variables, instance-major rows: x [B,n_x], u [B,n_u], z [B,n_z], tvp [B,n_tvp], p [B,n_p].
 * Returns 0 or a negative error code; per-instance status 0 = ok, 1 = Newton iteration failed.
 */
#ifndef DOMPC_B200_SIM_H
#define DOMPC_B200_SIM_H
#ifdef __cplusplus
extern "C" {
#endif
typedef struct dms_sim dms_sim;
void dms_get_dims(int* dims /* [8]: n_x,n_u,n_z,n_tvp,n_p,n_aux,discrete,collocation degree */);
dms_sim* dms_create(int device);
void dms_destroy(dms_sim* s);
const char* dms_last_error(dms_sim* s);
long long dms_launch_count(dms_sim* s);
int dms_step_host(dms_sim* s, int B, const double* x, const double* u, const double* z, const double* tvp, const double* p,
                  double t_step, int n_sub, double newton_tol, int newton_max, int mode,
                  double* x_next, double* z_next, double* aux /* may be NULL */, int* status);
/* same step with device-resident arrays, queued on cuda_stream (no reference counterpart: the reference is host only) */
int dms_step_device(dms_sim* s, int B, const double* d_x, const double* d_u, const double* d_z, const double* d_tvp, const double* d_p,
                    double t_step, int n_sub, double newton_tol, int newton_max, int mode,
                    double* d_x_next, double* d_z_next, double* d_aux /* may be NULL */, int* d_status, void* cuda_stream);
#ifdef __cplusplus
}
#endif
#endif
