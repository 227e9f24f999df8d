// pipe2_launch.cuh -- host-side launcher of the third-generation persistent kernel (pipe2_update.cuh), shared by
// the two translation units that instantiate it: pipe2_fast_kernels.cu (FMA, shuffle sums) and
// pipe2_exact_kernels.cu (-fmad=false, EXACT = true: sequential sums on the serial warps, bit-identical results).
#pragma once
#include "pipe2_update.cuh"
#include "pipe_launch.cuh"

namespace cgmres_b200 {
namespace pipe2 {

template <class M, class Sim, bool EXACT>
cudaError_t launch(bool pfull, const FastArgs& a, cudaStream_t s) {
  using Y = Lay<M, EXACT>;
  if (a.n == 0) return cudaSuccess;
  if (a.scratch == nullptr) return cudaErrorInvalidValue;
  if (a.n_steps > 1 && a.dtau_tab == nullptr && a.t_inst == nullptr) return cudaErrorInvalidValue;
  int device = 0;
  cudaError_t e = cudaGetDevice(&device);
  if (e != cudaSuccess) return e;
  const int64_t rounds = (a.n + Y::NI - 1) / Y::NI;
  const int sms = pipe::sm_count(device);
  const unsigned grid = (unsigned)(rounds < (int64_t)sms ? rounds : (int64_t)sms);  // persistent: one CTA per SM
  if (pfull) {
    e = cudaFuncSetAttribute(control_kernel<M, Sim, true, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Y::smem_bytes);
    if (e != cudaSuccess) return e;
    control_kernel<M, Sim, true, EXACT><<<grid, Y::threads, Y::smem_bytes, s>>>(a);
  } else {
    e = cudaFuncSetAttribute(control_kernel<M, Sim, false, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Y::smem_bytes);
    if (e != cudaSuccess) return e;
    control_kernel<M, Sim, false, EXACT><<<grid, Y::threads, Y::smem_bytes, s>>>(a);
  }
  return cudaGetLastError();
}

template <class M, bool EXACT>
size_t scratch_for(int device, int64_t n) {
  using Y = Lay<M, EXACT>;
  const int64_t rounds = (n + Y::NI - 1) / Y::NI;
  const int64_t ctas = rounds < (int64_t)pipe::sm_count(device) ? rounds : (int64_t)pipe::sm_count(device);
  return (size_t)(ctas > 0 ? ctas : 1) * Y::scratch_doubles_per_cta;
}

}  // namespace pipe2
}  // namespace cgmres_b200
