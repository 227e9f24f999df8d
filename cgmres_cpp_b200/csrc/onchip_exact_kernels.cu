// onchip_exact_kernels.cu -- the on-chip control-update kernel (fast_update.cuh) with the reference's sequential
// sums, compiled WITHOUT FMA contraction (-fmad=false): bit-identical to the reference.  A translation unit of its
// own so that it builds in parallel with the streaming exact kernels.
#include "fast_update.cuh"
#include "pipe_launch.cuh"

// Two bit-identical on-chip kernels are built here: MODE_ONCHIP_EXACT = the first-generation one-round-per-CTA kernel
// (fast_update.cuh, EXACT_SUMS = true; 2.9e7 msd updates/s), MODE_PIPELINED_EXACT = the persistent pipelined kernel
// of the fast mode with EXACT = true (2.7e7: its sequential sums sit on every vector warp's critical path), kept as
// the verification build of the fast kernel.

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

cudaError_t pipelined_exact_launch_control(int model, bool ptau_full, const FastArgs& a, cudaStream_t s) {
  switch (model) {
    case MODEL_MSD: return pipe::launch<MassSpringDamperModel, MassSpringDamperSimulator, true>(ptau_full, a, s);
    case MODEL_ARM: return pipe::launch<ArmPendulumModel, ArmPendulumSimulator, true>(ptau_full, a, s);
    case MODEL_SEMIACTIVE: return pipe::launch<SemiactiveDamperModel, SemiactiveDamperSimulator, true>(ptau_full, a, s);
  }
  return cudaErrorInvalidValue;
}

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

size_t pipelined_exact_scratch_doubles(int model, int device, int64_t n) {
  switch (model) {
    case MODEL_MSD: return pipe::scratch_for<MassSpringDamperModel>(device, n);
    case MODEL_ARM: return pipe::scratch_for<ArmPendulumModel>(device, n);
    case MODEL_SEMIACTIVE: return pipe::scratch_for<SemiactiveDamperModel>(device, n);
  }
  return 0;
}

}  // namespace cgmres_b200
