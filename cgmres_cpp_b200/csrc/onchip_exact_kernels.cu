// onchip_exact_kernels.cu -- the on-chip control-update kernel (fast_update.cuh) with the reference's sequential
// sums, compiled WITHOUT FMA contraction (-fmad=false): bit-identical to the reference.  A translation unit of its
// own so that it builds in parallel with the streaming exact kernels.
#include "fast_update.cuh"

namespace cgmres_b200 {
namespace {
// the on-chip kernel with sequential sums, compiled here WITHOUT FMA contraction: bit-identical to the reference
template <class M, class Sim>
cudaError_t launch_onchip_exact_t(bool pfull, const FastArgs& a, cudaStream_t s) {
  using Y = fast::Lay<M>;
  if (a.n == 0) return cudaSuccess;
  const unsigned grid = (unsigned)((a.n + Y::G - 1) / Y::G);
  cudaError_t e;
  if (pfull) {
    e = cudaFuncSetAttribute(fast::control_kernel<M, Sim, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Y::smem_bytes);
    if (e != cudaSuccess) return e;
    fast::control_kernel<M, Sim, true, true><<<grid, Y::threads, Y::smem_bytes, s>>>(a);
  } else {
    e = cudaFuncSetAttribute(fast::control_kernel<M, Sim, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Y::smem_bytes);
    if (e != cudaSuccess) return e;
    fast::control_kernel<M, Sim, false, true><<<grid, Y::threads, Y::smem_bytes, s>>>(a);
  }
  return cudaGetLastError();
}
}  // namespace

cudaError_t onchip_exact_launch_control(int model, bool ptau_full, const FastArgs& a, cudaStream_t s) {
  switch (model) {
    case MODEL_MSD: return launch_onchip_exact_t<MassSpringDamperModel, MassSpringDamperSimulator>(ptau_full, a, s);
    case MODEL_ARM: return launch_onchip_exact_t<ArmPendulumModel, ArmPendulumSimulator>(ptau_full, a, s);
    case MODEL_SEMIACTIVE:
      return launch_onchip_exact_t<SemiactiveDamperModel, SemiactiveDamperSimulator>(ptau_full, a, s);
  }
  return cudaErrorInvalidValue;
}

int onchip_exact_instances_per_cta(int model) {
  switch (model) {
    case MODEL_MSD: return fast::Lay<MassSpringDamperModel>::G;
    case MODEL_ARM: return fast::Lay<ArmPendulumModel>::G;
    case MODEL_SEMIACTIVE: return fast::Lay<SemiactiveDamperModel>::G;
  }
  return 1;
}

}  // namespace cgmres_b200
