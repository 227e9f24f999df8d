/*
 * cgmres_b200.h -- C ABI of the B200-native batched C/GMRES controller.
 *
 * The reference (blockahead/CGMRES_cpp) has no FFI layer: its boundary is the
 * header-only C++ class `Cgmres<Model>` (include/cgmres.hpp:8-207), one object per
 * controller.  This library is that object with a leading batch dimension: one
 * handle owns n independent controller instances of one model on one GPU, and
 * every entry point below is the batched form of one reference method (cited).
 * include/cgmres.hpp of this repository wraps these calls back into a class
 * template with the reference's names.
 *
 * Array layout at this boundary is the reference's own, one instance after another
 * (instance-major): x[n][dim_x], u[n][dim_u], U[n][dv*dim_u] with U[i*dim_u+j] inside an
 * instance (cgmres.hpp:55-58), ptau[n][(dv+1)*dim_p] (cgmres.hpp:36-39).  Plain
 * pointers and sizes only; "host" pointers are ordinary (preferably pinned) host
 * memory, "_dev" variants take device pointers in the same instance-major layout.
 *
 * Every function returns 0 on success or a negative CGMRES_B200_E* code;
 * cgmres_b200_last_error() gives the message of the calling thread's last failure.
 * There is no CPU fallback: creation fails when no CUDA device is usable.
 *
 * Numerical exit paths of the reference's gmres() that only printf or return
 * silently (include/gmres.hpp:39-41, 63-65, 93-95) are reported per instance in a
 * status word instead; the numbers produced on those paths are the reference's.
 */
#ifndef CGMRES_B200_H
#define CGMRES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cgmres_b200_controller* cgmres_b200_handle;

/* model ids: the reference's example directories */
enum {
  CGMRES_B200_MODEL_MASS_SPRING_DAMPER = 0,         /* mass_spring_damper/model.hpp, multiple_controller/model1.hpp */
  CGMRES_B200_MODEL_ARM_TYPE_INVERTED_PENDULUM = 1, /* arm_type_inverted_pendulum/model.hpp, multiple_controller/model2.hpp */
  CGMRES_B200_MODEL_SEMIACTIVE_DAMPER = 2           /* semiactive_damper/model.hpp */
};

/* build modes */
enum {
  CGMRES_B200_MODE_EXACT = 0, /* reference operation order, no FMA: bit-identical U and x (no-libm models) */
  CGMRES_B200_MODE_FAST = 1,  /* persistent on-chip kernel (instance groups resident in shared / tensor memory,
                                 serial-recursion and vector warps pipelined), FMA + shuffle reductions: within
                                 the north-star tolerances                                                  */
  CGMRES_B200_MODE_ONCHIP_EXACT = 2, /* on-chip kernel (one CTA per instance group) with the reference's sequential
                                 sums and no FMA: bit-identical like MODE_EXACT                             */
  CGMRES_B200_MODE_PIPELINED_EXACT = 3 /* verification build of the MODE_FAST kernel: same persistent pipelined
                                 kernel, sequential sums, no FMA: bit-identical (slower than mode 2)          */
};

/* status word of an instance's last update: exit code in bits 0..7, Krylov columns used in bits 8..15 */
enum {
  CGMRES_B200_EXIT_FULL = 0,      /* k_max iterations                               (gmres.hpp:46)    */
  CGMRES_B200_EXIT_CONVERGED = 1, /* |rho[k+1]| < tol, k columns used               (gmres.hpp:93-95) */
  CGMRES_B200_EXIT_RHO0 = 2,      /* ||r0|| < tol, dUdt left unchanged              (gmres.hpp:39-41) */
  CGMRES_B200_EXIT_BREAKDOWN = 3  /* |h(k+1,k)| < DBL_EPSILON, dUdt left unchanged  (gmres.hpp:63-65) */
};

enum {
  CGMRES_B200_OK = 0,
  CGMRES_B200_EINVAL = -1,  /* bad argument / unknown model or mode */
  CGMRES_B200_ECUDA = -2,   /* CUDA runtime error (message in last_error) */
  CGMRES_B200_ENOMEM = -3,
  CGMRES_B200_ENOTIMPL = -4 /* combination not built (e.g. fast mode for a model without one) */
};

const char* cgmres_b200_last_error(void);
/* CUDA devices usable by this process (0 when there is none: the library then cannot create handles) */
int cgmres_b200_device_count(void);

/* Static problem description == Model's static constexpr members (<example>/model.hpp:7-34):
 * dims[0..5] = dim_x, dim_u, dim_p, dv, k_max, control_input; params[0..5] = dt, h, zeta, Tf, alpha, tol */
int cgmres_b200_model_dims(int model, int* dims);
int cgmres_b200_model_params(int model, double* params);
const char* cgmres_b200_model_name(int model);

/* Cgmres<Model>() x n on `device` (cgmres.hpp:11-20): t=0, dUdt=0 (the de-facto contract),
 * U/ptau/x zero until set.  All work of the handle is ordered on one CUDA stream. */
int cgmres_b200_create(int model, int64_t n_instances, int device, int mode, cgmres_b200_handle* out);
/* ~Cgmres (cgmres.hpp:22-30) */
int cgmres_b200_destroy(cgmres_b200_handle h);

int64_t cgmres_b200_size(cgmres_b200_handle h);
int cgmres_b200_model(cgmres_b200_handle h);
int cgmres_b200_mode(cgmres_b200_handle h);
/* Use `stream` (a cudaStream_t) for all subsequent work of the handle; NULL = the handle's own stream. */
int cgmres_b200_set_stream(cgmres_b200_handle h, void* stream);
void* cgmres_b200_get_stream(cgmres_b200_handle h);
int cgmres_b200_synchronize(cgmres_b200_handle h);

/* get_dtau(t) (cgmres.hpp:32-34): horizon step Tf(1-exp(-alpha t))/dv, evaluated with the host libm */
double cgmres_b200_get_dtau(cgmres_b200_handle h, double t);

/* set_ptau (cgmres.hpp:36-39): ptau[n][(dv+1)*dim_p] host */
int cgmres_b200_set_ptau(cgmres_b200_handle h, const double* ptau);
/* set_ptau_repeat (cgmres.hpp:41-49): p[n][dim_p] host, broadcast over the horizon */
int cgmres_b200_set_ptau_repeat(cgmres_b200_handle h, const double* p);
/* init_u0 (cgmres.hpp:51-59): u0[n][dim_u] host */
int cgmres_b200_init_u0(cgmres_b200_handle h, const double* u0);
/* init_u0_newton (cgmres.hpp:61-76): u0[n][dim_u] in/out (mutated like the reference), x0[n][dim_x],
 * p0[n][dim_p]; n_loop Newton steps with the reference's pivoted linsolve (matrix.hpp:166-224), on the device */
int cgmres_b200_init_u0_newton(cgmres_b200_handle h, double* u0, const double* x0, const double* p0, int n_loop);

/* control(u, x) (cgmres.hpp:78-110) for all n instances: copies x[n][dim_x] host->device, runs one update,
 * copies u[n][dim_u] device->host, returns when u is valid.  The plant state held by the handle becomes x. */
int cgmres_b200_control(cgmres_b200_handle h, double* u, const double* x);
/* same with device pointers (instance-major); asynchronous on the handle's stream */
int cgmres_b200_control_dev(cgmres_b200_handle h, double* u_dev, const double* x_dev);

/* plant state resident on the device (north star item 5: the closed loop never round-trips to the host) */
int cgmres_b200_set_x(cgmres_b200_handle h, const double* x);
int cgmres_b200_get_x(cgmres_b200_handle h, double* x);
/* last u = U[0:dim_u] per instance (cgmres.hpp:109) */
int cgmres_b200_get_u(cgmres_b200_handle h, double* u);

/* n_steps x { control(u,x); x += Simulator::dxdt(x,u)*dt } entirely on the device: the loop body of
 * <example>/main.cpp:66-77 (forward Euler).  Asynchronous on the handle's stream. */
int cgmres_b200_step_closed_loop(cgmres_b200_handle h, int n_steps);
/* (MODE_FAST / MODE_PIPELINED_EXACT run n_steps > 1 as multi-step launches of the persistent kernel: the same resident
 *  instances advance up to 256 steps per launch, per-step horizon ramps from a host-evaluated table.  In the bit-exact
 *  mode the results are those of n_steps single-step calls bit for bit; in MODE_FAST likewise for batches larger than
 *  16 instances per SM -- smaller batches take a shorter-latency kernel for single-step calls, whose FMA contraction
 *  may differ in the last bit.) */

/* The same loop with the trajectory recorded ON THE DEVICE and copied out chunk-wise: x_log[n_steps][n][dim_x] = the
 * plant state after every step, u_log[n_steps][n][dim_u] = the input every control update returned -- the rows the
 * reference's mains fprintf() to <example>_{x,u}.txt (mass_spring_damper/main.cpp:78-87), without a host round trip
 * per step.  Host arrays; returns when they are complete. */
int cgmres_b200_step_closed_loop_log(cgmres_b200_handle h, int n_steps, double* x_log, double* u_log);

/* Per-instance controller clocks (controllers started at different times): t[n] host.  From then on
 * every instance evaluates its own horizon step get_dtau(t_i), get_dtau(t_i + h) on the device (CUDA exp: within
 * 1 ulp of the host libm, so results follow the reference to the tolerance bars rather than bit for bit) and
 * advances t_i by dt per update.  t = NULL returns to the batch-uniform clock of get_state/set_state, whose horizon
 * steps are computed on the host and are bit-identical to the reference's. */
int cgmres_b200_set_t(cgmres_b200_handle h, const double* t);
int cgmres_b200_get_t(cgmres_b200_handle h, double* t);

/* Plant integrator used by step_closed_loop: EULER (default; what every reference example does) or
 * classical RK4 on the same Simulator::dxdt with u held over the step (the north star's wording; the reference has
 * no RK4 to compare with, so this option is excluded from the parity claims). */
enum { CGMRES_B200_PLANT_EULER = 1, CGMRES_B200_PLANT_RK4 = 2 };
int cgmres_b200_set_plant_integrator(cgmres_b200_handle h, int integrator);

/* checkpoint / teacher forcing: the complete controller state {t, U, dUdt} (cgmres.hpp:195-197).
 * Any array pointer may be NULL.  U, dUdt: [n][dv*dim_u] host. */
int cgmres_b200_get_state(cgmres_b200_handle h, double* t, double* U, double* dUdt);
int cgmres_b200_set_state(cgmres_b200_handle h, const double* t, const double* U, const double* dUdt);
/* status[n] of the last update (see CGMRES_B200_EXIT_*) */
int cgmres_b200_get_status(cgmres_b200_handle h, int32_t* status);

/* Host-side plant of the reference's example loop, batched: x[n][dim_x] += Simulator::dxdt(x, u)*dt with
 * u[n][dim_u] (<example>/main.cpp:74-76, <example>/simulator.hpp).  Plain host code on the caller's thread: the
 * plant belongs to the user's program (in the reference it lives in main.cpp, not in the controller); this is the
 * same Simulator functor the device epilogue inlines, for callers that drive control() from a non-C++ host. */
int cgmres_b200_plant_step_host(int model, int64_t n, double* x, const double* u);

/* sin and cos as the arm_type_inverted_pendulum functors evaluate them (include/cgmres_b200/portable_trig.hpp:
 * +,-,* only, bit-identical on host and device in the exact build modes, <= 1 ulp from glibc); host code */
void cgmres_b200_portable_sincos(double x, double* s, double* c);

/* Debug aid for the on-chip kernel: 64 clock64() phase timestamps of one warp of CTA 0 from the last launch
 * (all zero unless the library was built with -DCG_FAST_TIMING; tools/phase_times.py decodes them). */
int cgmres_b200_debug_phase_times(cgmres_b200_handle h, int64_t* out64);

/* kernels launched by this library in this process so far (for the benchmark's launch accounting) */
int64_t cgmres_b200_launch_count(void);

/* FP64 vector-pipe peak of `device`, measured with a register-resident chain microbenchmark: use_fma=1 counts
 * DFMA as 2 flop, use_fma=0 issues separate DMUL+DADD (the ceiling of the exact mode).  The roofline denominator
 * of the benchmark (MEASURED_PEAKS.json has no FP64 entry). sm_clock_mhz (may be NULL) = the device's nominal clock. */
int cgmres_b200_measure_fp64_peak(int device, int use_fma, double* tflops, double* sm_clock_mhz);
/* dependent-issue latency of the FP64 pipe in SM cycles: op 0 = DFMA, 1 = DADD, 2 = DMUL (one warp, one chain).  The
 * on-chip kernel is bound by dv-serial recursions, i.e. by this number times their expression depth. */
int cgmres_b200_measure_fp64_latency(int device, int op, double* cycles_per_op);

#ifdef __cplusplus
}
#endif
#endif
