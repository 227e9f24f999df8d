// fast_kernels.cu -- instantiations of the on-chip control-update kernel with FMA contraction and shuffle
// reductions (default nvcc floating-point flags; NOT -fmad=false).
// The serial recursions are unrolled 10 stages deep in this translation unit (5 in the verification twin): in
// the pipelined kernel the serial warp competes with 16 busy vector warps for the shared-memory pipeline, and
// exposing the loads of 10 stages at once hides that latency better (GPU sweep: 2 -> 5.3e7, 5 -> 6.1e7,
// 10 -> 6.3e7, 25 -> 6.0e7 with spills).
#ifndef CG_SWEEP_UNROLL
#define CG_SWEEP_UNROLL 10
#endif
#include "fast_update.cuh"
#include "pipe_launch.cuh"

// (EXACT_SUMS = true instantiations of fast::control_kernel belong to onchip_exact_kernels.cu alone: that unit is
//  compiled with -fmad=false, and two units instantiating the same specialisation with different floating-point
//  flags would collide at link time and silently pick one.)

// 1: the persistent warp-specialised kernel (pipe_update.cuh) serves the fast mode; 0: the first-generation
// one-round-per-CTA kernel (fast_update.cuh).  Per model: the pipelined kernel wins where the vector work and the
// serial recursion are of similar length (msd, semiactive).
#ifndef CG_FAST_PIPE_MSD
#define CG_FAST_PIPE_MSD 1
#endif
#ifndef CG_FAST_PIPE_ARM
#define CG_FAST_PIPE_ARM 1
#endif
#ifndef CG_FAST_PIPE_SEMI
#define CG_FAST_PIPE_SEMI 1
#endif

namespace cgmres_b200 {
namespace {
template <class M, class Sim>
cudaError_t launch_pipe(bool pfull, const FastArgs& a, cudaStream_t s) {
  return pipe::launch<M, Sim, false>(pfull, a, s);
}

template <class M, class Sim>
cudaError_t launch_t(bool pfull, const FastArgs& a, cudaStream_t s) {
  using Y = fast::Lay<M>;
  if (a.n == 0) return cudaSuccess;
  const unsigned grid = (unsigned)((a.n + Y::G - 1) / Y::G);
  cudaError_t e;
  if (pfull) {
    e = cudaFuncSetAttribute(fast::control_kernel<M, Sim, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Y::smem_bytes);
    if (e != cudaSuccess) return e;
    fast::control_kernel<M, Sim, true, false><<<grid, Y::threads, Y::smem_bytes, s>>>(a);
  } else {
    e = cudaFuncSetAttribute(fast::control_kernel<M, Sim, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Y::smem_bytes);
    if (e != cudaSuccess) return e;
    fast::control_kernel<M, Sim, false, false><<<grid, Y::threads, Y::smem_bytes, s>>>(a);
  }
  return cudaGetLastError();
}
}  // namespace

// Small batches (at most one first-generation CTA per SM) have nothing for the pipelined kernel to overlap, and
// the first-generation kernel's dependent chain per update is shorter (45 us vs 70 us for msd): it serves them.
// Both kernels perform the same arithmetic, so which one ran cannot be seen in the results
// (tests/test_gpu_onchip.py::test_fast_full_size_batch_shard_invariance compares them bit for bit).
template <class M, class Sim>
cudaError_t launch_auto(bool use_pipe, bool pfull, const FastArgs& a, cudaStream_t s) {
  if (use_pipe) {
    int device = 0;
    cudaError_t e = cudaGetDevice(&device);
    if (e != cudaSuccess) return e;
    if (a.n > (int64_t)fast::Lay<M>::G * pipe::sm_count(device)) return launch_pipe<M, Sim>(pfull, a, s);
  }
  return launch_t<M, Sim>(pfull, a, s);
}

cudaError_t fast_launch_control(int model, bool ptau_full, const FastArgs& a, cudaStream_t s) {
  switch (model) {
    case MODEL_MSD:
      return launch_auto<MassSpringDamperModel, MassSpringDamperSimulator>(CG_FAST_PIPE_MSD, ptau_full, a, s);
    case MODEL_ARM: return launch_auto<ArmPendulumModel, ArmPendulumSimulator>(CG_FAST_PIPE_ARM, ptau_full, a, s);
    case MODEL_SEMIACTIVE:
      return launch_auto<SemiactiveDamperModel, SemiactiveDamperSimulator>(CG_FAST_PIPE_SEMI, ptau_full, a, s);
  }
  return cudaErrorInvalidValue;
}

size_t fast_scratch_doubles(int model, int device, int64_t n) {
  switch (model) {
    case MODEL_MSD: return CG_FAST_PIPE_MSD ? pipe::scratch_for<MassSpringDamperModel>(device, n) : 0;
    case MODEL_ARM: return CG_FAST_PIPE_ARM ? pipe::scratch_for<ArmPendulumModel>(device, n) : 0;
    case MODEL_SEMIACTIVE: return CG_FAST_PIPE_SEMI ? pipe::scratch_for<SemiactiveDamperModel>(device, n) : 0;
  }
  return 0;
}

int fast_instances_per_cta(int model) {
  switch (model) {
    case MODEL_MSD: return CG_FAST_PIPE_MSD ? pipe::Lay<MassSpringDamperModel>::NI : fast::Lay<MassSpringDamperModel>::G;
    case MODEL_ARM: return CG_FAST_PIPE_ARM ? pipe::Lay<ArmPendulumModel>::NI : fast::Lay<ArmPendulumModel>::G;
    case MODEL_SEMIACTIVE:
      return CG_FAST_PIPE_SEMI ? pipe::Lay<SemiactiveDamperModel>::NI : fast::Lay<SemiactiveDamperModel>::G;
  }
  return 0;
}
}  // namespace cgmres_b200
