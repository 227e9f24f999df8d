// pipe2_exact_kernels.cu -- the third-generation persistent kernel (pipe2_update.cuh) built bit-exact: compiled
// WITHOUT FMA contraction (-fmad=false), EXACT = true (the reference's operation order; its 21 sequential sums per
// update run lane-per-instance on the serial warps).  Serves MODE_PIPELINED_EXACT.
#ifndef CG_SWEEP_UNROLL
#define CG_SWEEP_UNROLL 10
#endif
#include "pipe2_launch.cuh"

namespace cgmres_b200 {

cudaError_t pipe2_exact_launch_control(int model, bool ptau_full, const FastArgs& a, cudaStream_t s) {
  switch (model) {
    case MODEL_MSD: return pipe2::launch<MassSpringDamperModel, MassSpringDamperSimulator, true>(ptau_full, a, s);
#ifndef CG_ONLY_MSD  // (tuning builds instantiate the msd kernels only: tools/quick_msd_build.sh)
    case MODEL_ARM: return pipe2::launch<ArmPendulumModel, ArmPendulumSimulator, true>(ptau_full, a, s);
    case MODEL_SEMIACTIVE:
      return pipe2::launch<SemiactiveDamperModel, SemiactiveDamperSimulator, true>(ptau_full, a, s);
#endif
  }
  return cudaErrorInvalidValue;
}

size_t pipe2_exact_scratch_doubles(int model, int device, int64_t n) {
  switch (model) {
    case MODEL_MSD: return pipe2::scratch_for<MassSpringDamperModel, true>(device, n);
    case MODEL_ARM: return pipe2::scratch_for<ArmPendulumModel, true>(device, n);
    case MODEL_SEMIACTIVE: return pipe2::scratch_for<SemiactiveDamperModel, true>(device, n);
  }
  return 0;
}

int pipe2_exact_instances_per_cta(int model) {
  switch (model) {
    case MODEL_MSD: return pipe2::Lay<MassSpringDamperModel, true>::NI;
    case MODEL_ARM: return pipe2::Lay<ArmPendulumModel, true>::NI;
    case MODEL_SEMIACTIVE: return pipe2::Lay<SemiactiveDamperModel, true>::NI;
  }
  return 1;
}

}  // namespace cgmres_b200
