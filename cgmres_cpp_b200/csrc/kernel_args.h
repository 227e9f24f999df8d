// kernel_args.h -- internal interface between the C-ABI host code (capi.cu) and the
// kernel translation units (exact_kernels.cu is built with -fmad=false, fast_kernels.cu
// with FMA contraction on; they must not share inlined device code at link time, so
// everything crosses this plain-struct boundary).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cgmres_b200 {

enum { MODEL_MSD = 0, MODEL_ARM = 1, MODEL_SEMIACTIVE = 2, MODEL_COUNT = 3 };

// exit path of the GMRES solve inside one update (status word = code | columns_used << 8)
enum { EXIT_FULL = 0, EXIT_CONVERGED = 1, EXIT_RHO0 = 2, EXIT_BREAKDOWN = 3 };

struct ModelInfo {
  int dim_x, dim_u, dim_p, dv, k_max, n_ctrl;
  double dt, h, zeta, Tf, alpha, tol;
  double plant_dt;
  const char* name;
  int L() const { return dim_u * dv; }
};
const ModelInfo* model_info(int model);

// Structure-of-arrays device state of the exact mode: every matrix is [rows][ld],
// instance index fastest (ld = instance count rounded up to 32).
struct ExactArgs {
  int64_t n, ld;
  double* x;           // [dim_x][ld]      plant state
  double* U;           // [L][ld]          input trajectory     (cgmres.hpp:196)
  double* dUdt;        // [L][ld]          its time derivative  (cgmres.hpp:197)
  const double* ptau;  // [dim_p][ld] (repeat) or [(dv+1)*dim_p][ld] (full)   (cgmres.hpp:200)
  double* F1;          // [L][ld]          F(U, x+dx*h, t+h)    (cgmres.hpp:202)
  double* V;           // [(k_max+1)*L][ld] un-normalised Krylov basis (gmres.hpp:11)
  double* xtau;        // [3][dim_x*(dv-1)][ld] rollout scratch, one plane per fused trajectory (cgmres.hpp:116)
  double* u_out;       // [dim_u][ld]      u = U[0:dim_u]       (cgmres.hpp:109)
  int32_t* status;     // [ld]
  double dtau_t, dtau_th;  // get_dtau(t), get_dtau(t+h) evaluated on the host (cgmres.hpp:32-34)
  double* t_inst;          // null: all instances share the handle's clock; else per-instance t [n], advanced by dt
  int plant;               // 0: no plant step; 1: Euler (reference); 2: RK4 (include/cgmres_b200/plant.hpp)
};

// Instance-major device state of the on-chip ("fast") kernels: the layout of the C ABI itself, every instance's
// vectors contiguous (U[n][L], x[n][dim_x], ...), so a warp loads its instance with coalesced 256-byte requests.
struct FastArgs {
  int64_t n;
  double* x;           // [n][dim_x]
  double* U;           // [n][L]
  double* dUdt;        // [n][L]
  const double* ptau;  // [n][dim_p] (repeat) or [n][(dv+1)*dim_p] (full)
  double* u_out;       // [n][dim_u]
  int32_t* status;     // [n]
  double dtau_t, dtau_th;
  double* t_inst;      // null: lock step (dtau from the host); else per-instance clocks [n], dtau evaluated on device
  int plant;
  long long* dbg;      // phase timestamps (debug builds with -DCG_FAST_TIMING), else unused / null
  double* scratch;     // pipelined fast kernel: per-CTA spill area (fast_scratch_doubles()), one region per
                       // concurrently running launch; unused by the other on-chip kernels
  // third-generation kernel (pipe2_update.cuh) only; the other kernels advance exactly one step per launch:
  int n_steps = 1;                   // closed-loop steps of the same resident instances inside this launch
  const double* dtau_tab = nullptr;  // n_steps > 1, lock step: {get_dtau(t_s), get_dtau(t_s + h)} per step, host-evaluated
  double* x_log = nullptr;           // optional trajectory log [n_steps][n][dim_x] (x after each plant step)
  double* u_log = nullptr;           // optional [n_steps][n][dim_u] (u returned by each control update)
};

// on-chip kernels: fast_kernels.cu (FMA, shuffle reductions) and their sequential-sum, no-FMA twin compiled in
// exact_kernels.cu (bit-identical to the reference; verification build of the same kernel)
cudaError_t fast_launch_control(int model, bool ptau_full, const FastArgs& a, cudaStream_t s);
cudaError_t onchip_exact_launch_control(int model, bool ptau_full, const FastArgs& a, cudaStream_t s);
// the fast mode's persistent pipelined kernel built with sequential sums and no FMA (onchip_exact_kernels.cu)
cudaError_t pipelined_exact_launch_control(int model, bool ptau_full, const FastArgs& a, cudaStream_t s);
// third generation (pipe2_update.cuh): FMA build (pipe2_fast_kernels.cu) and bit-exact build (pipe2_exact_kernels.cu)
cudaError_t pipe2_fast_launch_control(int model, bool ptau_full, const FastArgs& a, cudaStream_t s);
cudaError_t pipe2_exact_launch_control(int model, bool ptau_full, const FastArgs& a, cudaStream_t s);
size_t pipe2_fast_scratch_doubles(int model, int device, int64_t n);
size_t pipe2_exact_scratch_doubles(int model, int device, int64_t n);
int pipe2_fast_instances_per_cta(int model);   // (the two builds may hold different numbers of groups per SM)
int pipe2_exact_instances_per_cta(int model);
inline size_t pipe2_scratch_doubles(int model, int device, int64_t n, bool exact) {
  return exact ? pipe2_exact_scratch_doubles(model, device, n) : pipe2_fast_scratch_doubles(model, device, n);
}
inline int pipe2_instances_per_cta(int model, bool exact) {
  return exact ? pipe2_exact_instances_per_cta(model) : pipe2_fast_instances_per_cta(model);
}
int fast_instances_per_cta(int model);
int onchip_exact_instances_per_cta(int model);
inline int onchip_instances_per_cta(int model, int mode) {  // mode 1 = fast, 2 = onchip_exact, 3 = pipelined exact
  return mode == 2 ? onchip_exact_instances_per_cta(model) : fast_instances_per_cta(model);
}
// doubles of global scratch one launch of the fast kernel over (up to) n instances needs on `device` (0: none)
size_t fast_scratch_doubles(int model, int device, int64_t n);
size_t pipelined_exact_scratch_doubles(int model, int device, int64_t n);

// exact mode (exact_kernels.cu)
cudaError_t exact_launch_control(int model, bool ptau_full, const ExactArgs& a, cudaStream_t s);
// U element e of instance i is written to U[e*elem_stride + i*inst_stride] (SoA: ld,1; instance-major: 1,L)
cudaError_t exact_launch_newton(int model, int64_t n, int64_t elem_stride, int64_t inst_stride, double* u0_aos,
                                const double* x0_aos, const double* p0_aos, int p_stride, int n_loop, double* U,
                                cudaStream_t s);
size_t exact_control_smem_bytes(int model, int block);

// layout kernels (layout.cu): instance-major host layout <-> structure of arrays
//   aos[n][rows]  <->  soa[rows][ld]
cudaError_t launch_aos_to_soa(const double* aos, double* soa, int64_t n, int rows, int64_t ld, cudaStream_t s);
cudaError_t launch_soa_to_aos(const double* soa, double* aos, int64_t n, int rows, int64_t ld, cudaStream_t s);
// soa[(i*rows_per + j)][n] = aos[n][j] for i in 0..reps-1 (init_u0 / set_ptau_repeat broadcasts)
cudaError_t launch_broadcast_rows(const double* aos, double* soa, int64_t n, int rows_per, int reps, int64_t ld,
                                  cudaStream_t s);
// dst[n][i*rows_per + j] = aos[n][j] for i in 0..reps-1 (instance-major destination)
cudaError_t launch_broadcast_inst(const double* aos, double* dst, int64_t n, int rows_per, int reps, cudaStream_t s);

}  // namespace cgmres_b200
