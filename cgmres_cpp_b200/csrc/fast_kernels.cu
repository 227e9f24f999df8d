// fast_kernels.cu -- instantiations of the on-chip control-update kernel with FMA contraction and shuffle
// reductions (default nvcc floating-point flags; NOT -fmad=false).
#include "fast_update.cuh"

// experiment knob (tools/sweep_build.sh): build the "fast" entry point with the reference's sequential sums
#ifndef CG_FAST_SEQ_SUMS
#define CG_FAST_SEQ_SUMS false
#endif

namespace cgmres_b200 {
namespace {
template <class M, class Sim>
cudaError_t launch_t(bool pfull, const FastArgs& a, cudaStream_t s) {
  using Y = fast::Lay<M>;
  if (a.n == 0) return cudaSuccess;
  const unsigned grid = (unsigned)((a.n + Y::G - 1) / Y::G);
  cudaError_t e;
  if (pfull) {
    e = cudaFuncSetAttribute(fast::control_kernel<M, Sim, true, CG_FAST_SEQ_SUMS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Y::smem_bytes);
    if (e != cudaSuccess) return e;
    fast::control_kernel<M, Sim, true, CG_FAST_SEQ_SUMS><<<grid, Y::threads, Y::smem_bytes, s>>>(a);
  } else {
    e = cudaFuncSetAttribute(fast::control_kernel<M, Sim, false, CG_FAST_SEQ_SUMS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Y::smem_bytes);
    if (e != cudaSuccess) return e;
    fast::control_kernel<M, Sim, false, CG_FAST_SEQ_SUMS><<<grid, Y::threads, Y::smem_bytes, s>>>(a);
  }
  return cudaGetLastError();
}
}  // namespace

cudaError_t fast_launch_control(int model, bool ptau_full, const FastArgs& a, cudaStream_t s) {
  switch (model) {
    case MODEL_MSD: return launch_t<MassSpringDamperModel, MassSpringDamperSimulator>(ptau_full, a, s);
    case MODEL_ARM: return launch_t<ArmPendulumModel, ArmPendulumSimulator>(ptau_full, a, s);
    case MODEL_SEMIACTIVE: return launch_t<SemiactiveDamperModel, SemiactiveDamperSimulator>(ptau_full, a, s);
  }
  return cudaErrorInvalidValue;
}

int fast_instances_per_cta(int model) {
  switch (model) {
    case MODEL_MSD: return fast::Lay<MassSpringDamperModel>::G;
    case MODEL_ARM: return fast::Lay<ArmPendulumModel>::G;
    case MODEL_SEMIACTIVE: return fast::Lay<SemiactiveDamperModel>::G;
  }
  return 0;
}
}  // namespace cgmres_b200
