// pipe2_fast_kernels.cu -- the third-generation persistent kernel (pipe2_update.cuh) with FMA contraction and
// butterfly sums (default nvcc floating-point flags): serves MODE_FAST for batches larger than one first-generation
// CTA per SM, and every multi-step launch of that mode.
#ifndef CG_SWEEP_UNROLL
#define CG_SWEEP_UNROLL 10
#endif
#include "pipe2_launch.cuh"

namespace cgmres_b200 {

cudaError_t pipe2_fast_launch_control(int model, bool ptau_full, const FastArgs& a, cudaStream_t s) {
  switch (model) {
    case MODEL_MSD: return pipe2::launch<MassSpringDamperModel, MassSpringDamperSimulator, false>(ptau_full, a, s);
#ifndef CG_ONLY_MSD  // (tuning builds instantiate the msd kernels only: tools/quick_msd_build.sh)
    case MODEL_ARM: return pipe2::launch<ArmPendulumModel, ArmPendulumSimulator, false>(ptau_full, a, s);
    case MODEL_SEMIACTIVE:
      return pipe2::launch<SemiactiveDamperModel, SemiactiveDamperSimulator, false>(ptau_full, a, s);
#endif
  }
  return cudaErrorInvalidValue;
}

size_t pipe2_fast_scratch_doubles(int model, int device, int64_t n) {
  switch (model) {
    case MODEL_MSD: return pipe2::scratch_for<MassSpringDamperModel, false>(device, n);
    case MODEL_ARM: return pipe2::scratch_for<ArmPendulumModel, false>(device, n);
    case MODEL_SEMIACTIVE: return pipe2::scratch_for<SemiactiveDamperModel, false>(device, n);
  }
  return 0;
}

int pipe2_fast_instances_per_cta(int model) {
  switch (model) {
    case MODEL_MSD: return pipe2::Lay<MassSpringDamperModel, false>::NI;
    case MODEL_ARM: return pipe2::Lay<ArmPendulumModel, false>::NI;
    case MODEL_SEMIACTIVE: return pipe2::Lay<SemiactiveDamperModel, false>::NI;
  }
  return 1;
}

}  // namespace cgmres_b200
