// pipe_launch.cuh -- host-side launcher of the persistent pipelined kernel (pipe_update.cuh), shared by the two
// translation units that instantiate it: fast_kernels.cu (FMA, shuffle sums) and onchip_exact_kernels.cu
// (-fmad=false, EXACT = true: sequential sums, bit-identical to the reference).
#pragma once
#include "pipe_update.cuh"

namespace cgmres_b200 {
namespace pipe {

static inline int sm_count(int device) {
  static int cached[64] = {0};
  if (device < 0 || device >= 64) device = 0;
  if (cached[device] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
    cached[device] = v;
  }
  return cached[device];
}

template <class M, class Sim, bool EXACT>
cudaError_t launch(bool pfull, const FastArgs& a, cudaStream_t s) {
  using Y = Lay<M>;
  if (a.n == 0) return cudaSuccess;
  if (a.scratch == nullptr) return cudaErrorInvalidValue;
  int device = 0;
  cudaError_t e = cudaGetDevice(&device);
  if (e != cudaSuccess) return e;
  const int64_t rounds = (a.n + Y::NI - 1) / Y::NI;
  const int sms = sm_count(device);
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

// doubles of global scratch one launch over (up to) n instances needs
template <class M>
size_t scratch_for(int device, int64_t n) {
  using Y = Lay<M>;
  const int64_t rounds = (n + Y::NI - 1) / Y::NI;
  const int64_t ctas = rounds < (int64_t)sm_count(device) ? rounds : (int64_t)sm_count(device);
  return (size_t)(ctas > 0 ? ctas : 1) * Y::scratch_doubles_per_cta;
}

}  // namespace pipe
}  // namespace cgmres_b200
