/*
 * cgmres_oracle.h -- CPU oracle for the batched C/GMRES hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the shipped product path
 * (cgmres_cpp_b200/, include/) may include, link or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker / the reported CPU baseline.
 *
 * Two libraries export this same interface:
 *   oracle/_build/libcgmres_oracle.so   prefix "oracle_"  plain-C restatement
 *                                       (cgmres_oracle.c), kind "port"
 *   oracle/_build/libcgmres_oracle_ptrig.so  prefix "oraclept_"  the same restatement with the arm model's
 *                                       sin/cos replaced by oracle/portable_trig.h, kind "port_ptrig": the
 *                                       bit-exact checker of the GPU's exact modes on the arm model
 *   oracle/_ref/libcgmres_ref.so        prefix "ref_"     the UNMODIFIED reference
 *                                       headers under /root/reference compiled by
 *                                       ref_harness.cpp, kind "reference"
 * so the tests can pin the restatement against the reference itself bit for bit.
 *
 * Layout of every array argument is the reference's own (instance-major, then
 * the reference's stage-major AoS): x[n][dim_x], u[n][dim_u], U[n][dv*dim_u],
 * ptau[n][(dv+1)*dim_p].
 */
#ifndef CGMRES_ORACLE_H
#define CGMRES_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef ORACLE_PREFIX
#define ORACLE_PREFIX oracle_
#endif
#define ORACLE_CAT2(a, b) a##b
#define ORACLE_CAT(a, b) ORACLE_CAT2(a, b)
#define ORACLE_FN(name) ORACLE_CAT(ORACLE_PREFIX, name)

/* model ids (same numbering as include/cgmres_b200.h) */
enum { ORACLE_MODEL_MSD = 0, ORACLE_MODEL_ARM = 1, ORACLE_MODEL_SEMIACTIVE = 2 };

/* exit path of the last gmres() call: matches include/cgmres_b200.h status codes */
enum {
  ORACLE_EXIT_FULL = 0,      /* all k_max iterations                      */
  ORACLE_EXIT_CONVERGED = 1, /* |rho[k+1]| < tol break (k columns used)   */
  ORACLE_EXIT_RHO0 = 2,      /* rho0 < tol silent return                  */
  ORACLE_EXIT_BREAKDOWN = 3  /* |h_{k+1,k}| < DBL_EPSILON "Breakdown"     */
};

/* dims[0..5] = dim_x, dim_u, dim_p, dv, k_max, n_control_inputs; returns 0 / -1 */
int ORACLE_FN(model_dims)(int model, int* dims);
/* par[0..5] = dt, h, zeta, Tf, alpha, tol */
int ORACLE_FN(model_params)(int model, double* par);

/* ---- single controller object (mirrors Cgmres<Model>, include/cgmres.hpp:8-207) ---- */
void* ORACLE_FN(create)(int model);
void ORACLE_FN(destroy)(void* ctl);
void ORACLE_FN(set_ptau)(void* ctl, const double* ptau_buf);
void ORACLE_FN(set_ptau_repeat)(void* ctl, const double* p_buf);
void ORACLE_FN(init_u0)(void* ctl, const double* u0);
void ORACLE_FN(init_u0_newton)(void* ctl, double* u0, const double* x0, const double* p0, int n_loop);
void ORACLE_FN(control)(void* ctl, double* u, const double* x);
double ORACLE_FN(get_dtau)(void* ctl, double t);
/* state access (checkpoint / teacher forcing): any pointer may be NULL */
void ORACLE_FN(get_state)(void* ctl, double* t, double* U, double* dUdt);
void ORACLE_FN(set_state)(void* ctl, const double* t, const double* U, const double* dUdt);
/* exit path of the last control(): status = exit code | (columns used << 8) */
int ORACLE_FN(last_status)(void* ctl);

/* the sin/cos this oracle build evaluates the arm model with (glibc, or the portable one in the ptrig build);
 * not exported by the reference harness */
void ORACLE_FN(sincos)(double x, double* s, double* c);

/* x <- x + Simulator::dxdt(x,u)*dt   (main.cpp:74-76 of each example) */
void ORACLE_FN(plant_step)(int model, double* x, const double* u);

/*
 * Batch closed loop, the shape of <example>/main.cpp run over n independent instances
 * on n_threads host threads (one live controller per thread at a time):
 *   per instance: set_ptau_repeat(p) [or set_ptau when p_full], init_u0(u0),
 *   init_u0_newton(u0,x0,p0,newton_iters), then n_steps x { control(u,x); plant_step }.
 * x0[n][dim_x], p[n][dim_p] (or [n][(dv+1)*dim_p] when p_full), u0[n][dim_u] (not mutated).
 * Outputs (any may be NULL):
 *   x_traj[(n_steps/rec_stride)][n][dim_x], u_traj[same][n][dim_u]  state/input AFTER steps
 *        rec_stride, 2*rec_stride, ... (rec_stride<=0: nothing recorded)
 *   x_fin[n][dim_x], u_fin[n][dim_u], U_fin[n][L], dUdt_fin[n][L]
 *   exit_hist[n][4]  per-instance count of each gmres exit path
 *   ctl_seconds[n]   accumulated wall time inside control() per instance
 * returns 0, or -1 on bad arguments.
 */
int ORACLE_FN(run_closed_loop)(int model, int64_t n, const double* x0, const double* p, int p_full,
                               const double* u0, int newton_iters, int n_steps, int rec_stride,
                               double* x_traj, double* u_traj, double* x_fin, double* u_fin,
                               double* U_fin, double* dUdt_fin, int32_t* exit_hist,
                               double* ctl_seconds, int n_threads);
/* max over worker threads of the wall time inside the step loops of the last run_closed_loop */
double ORACLE_FN(last_loop_seconds)(void);

#ifdef __cplusplus
}
#endif
#endif
