// peak.cu -- FP64 vector-pipe peak microbenchmarks.  MEASURED_PEAKS.json has no FP64 entry,
// so the benchmark measures the roofline denominator itself, on the same box and in the same run:
//   use_fma=1 : 8 independent DFMA chains per thread           (2 flop per issue slot)
//   use_fma=0 : the same chains as separate DMUL + DADD        (1 flop per issue slot: the exact mode's ceiling)
#include "cgmres_b200.h"
#include "kernel_args.h"

namespace cgmres_b200 {
namespace {

template <bool FMA>
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
  double r[8];
#pragma unroll
  for (int i = 0; i < 8; i++) r[i] = (double)(threadIdx.x + i) * 1e-3;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (FMA) {
        r[i] = __fma_rn(r[i], a, b);
      } else {
        r[i] = __dadd_rn(__dmul_rn(r[i], a), b);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += r[i];
  if (s == 12345.678) out[0] = s;  // keeps the chains alive, practically never stores
}

// dependent-issue latency of the FP64 pipe: one warp, one chain
template <int OP>
__global__ void fp64_latency_kernel(double* out, long long* cycles, int iters, double a, double b) {
  double r = (double)threadIdx.x * 1e-3 + 1.0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (OP == 0)
        r = __fma_rn(r, a, b);
      else if (OP == 1)
        r = __dadd_rn(r, b);
      else
        r = __dmul_rn(r, a);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[0] = t1 - t0;
  if (r == 12345.678) out[0] = r;
}

}  // namespace
}  // namespace cgmres_b200

// cycles per dependent FP64 instruction for op = 0 (DFMA), 1 (DADD), 2 (DMUL), measured with one warp on one SM
extern "C" int cgmres_b200_measure_fp64_latency(int device, int op, double* cycles_per_op) {
  using namespace cgmres_b200;
  if (!cycles_per_op || op < 0 || op > 2) return CGMRES_B200_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return CGMRES_B200_ECUDA;
  double* d = nullptr;
  long long* c = nullptr;
  if (cudaMalloc(&d, 64) != cudaSuccess || cudaMalloc(&c, 64) != cudaSuccess) return CGMRES_B200_ECUDA;
  const int iters = 4096;
  for (int rep = 0; rep < 2; rep++) {
    if (op == 0)
      fp64_latency_kernel<0><<<1, 32>>>(d, c, iters, 0.999999, 1e-9);
    else if (op == 1)
      fp64_latency_kernel<1><<<1, 32>>>(d, c, iters, 0.999999, 1e-9);
    else
      fp64_latency_kernel<2><<<1, 32>>>(d, c, iters, 0.999999, 1e-9);
  }
  long long h = 0;
  if (cudaMemcpy(&h, c, sizeof(h), cudaMemcpyDeviceToHost) != cudaSuccess) return CGMRES_B200_ECUDA;
  cudaFree(d);
  cudaFree(c);
  *cycles_per_op = (double)h / ((double)iters * 16.0);
  return 0;
}

extern "C" int cgmres_b200_measure_fp64_peak(int device, int use_fma, double* tflops, double* sm_clock_mhz_hint) {
  using namespace cgmres_b200;
  if (!tflops) return CGMRES_B200_EINVAL;
  if (cudaSetDevice(device) != cudaSuccess) return CGMRES_B200_ECUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return CGMRES_B200_ECUDA;
  double* d = nullptr;
  if (cudaMalloc(&d, 64) != cudaSuccess) return CGMRES_B200_ECUDA;
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(e0);
    if (use_fma)
      fp64_peak_kernel<true><<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    else
      fp64_peak_kernel<false><<<blocks, threads>>>(d, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) {
      cudaFree(d);
      return CGMRES_B200_ECUDA;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  const double flop = (double)blocks * threads * (double)iters * 8.0 * 2.0;
  *tflops = flop / (best * 1e-3) / 1e12;
  if (sm_clock_mhz_hint) *sm_clock_mhz_hint = prop.clockRate / 1000.0;
  return 0;
}
