// exact_kernels.cu -- instantiations of the exact (bit-reproducible) mode.
// MUST be compiled with -fmad=false: the reference is built without FMA
// contraction (CMakeLists.txt:40, x86-64 baseline ISA) and the closed loop
// amplifies a single contracted rounding to ~3e-7 in x over 1000 steps.
#include "exact_update.cuh"

namespace cgmres_b200 {

namespace {
constexpr int kBlock = 64;  // 65,536 instances -> 1024 CTAs = 6.92 per SM (tail 1.2 %)

template <class M, class Sim>
cudaError_t launch_control_t(bool pfull, const ExactArgs& a, cudaStream_t s) {
  const unsigned grid = (unsigned)((a.n + kBlock - 1) / kBlock);
  const size_t smem = sizeof(double) * exact::Ws<M>::COUNT * kBlock;
  if (grid == 0) return cudaSuccess;
  // Shared-memory carve-out is left to the driver on purpose: 6 resident CTAs need 157 KB for the per-thread
  // GMRES scalars and the remaining ~70 KB of L1 serves the rollout-scratch re-reads; forcing the maximum
  // carve-out measured 13 % slower (profiles/README.md).
  if (pfull)
    exact::control_kernel<M, Sim, true><<<grid, kBlock, smem, s>>>(a);
  else
    exact::control_kernel<M, Sim, false><<<grid, kBlock, smem, s>>>(a);
  return cudaGetLastError();
}

template <class M>
cudaError_t launch_newton_t(int64_t n, int64_t es, int64_t is, double* u0, const double* x0, const double* p0,
                            int p_stride, int n_loop, double* U, cudaStream_t s) {
  const unsigned grid = (unsigned)((n + 127) / 128);
  if (grid == 0) return cudaSuccess;
  exact::newton_init_kernel<M><<<grid, 128, 0, s>>>(n, es, is, u0, x0, p0, p_stride, n_loop, U);
  return cudaGetLastError();
}

}  // namespace

cudaError_t exact_launch_control(int model, bool ptau_full, const ExactArgs& a, cudaStream_t s) {
  switch (model) {
    case MODEL_MSD: return launch_control_t<MassSpringDamperModel, MassSpringDamperSimulator>(ptau_full, a, s);
    case MODEL_ARM: return launch_control_t<ArmPendulumModel, ArmPendulumSimulator>(ptau_full, a, s);
    case MODEL_SEMIACTIVE: return launch_control_t<SemiactiveDamperModel, SemiactiveDamperSimulator>(ptau_full, a, s);
  }
  return cudaErrorInvalidValue;
}

cudaError_t exact_launch_newton(int model, int64_t n, int64_t es, int64_t is, double* u0, const double* x0,
                                const double* p0, int p_stride, int n_loop, double* U, cudaStream_t s) {
  switch (model) {
    case MODEL_MSD: return launch_newton_t<MassSpringDamperModel>(n, es, is, u0, x0, p0, p_stride, n_loop, U, s);
    case MODEL_ARM: return launch_newton_t<ArmPendulumModel>(n, es, is, u0, x0, p0, p_stride, n_loop, U, s);
    case MODEL_SEMIACTIVE:
      return launch_newton_t<SemiactiveDamperModel>(n, es, is, u0, x0, p0, p_stride, n_loop, U, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace cgmres_b200
