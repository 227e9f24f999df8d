// fp64_halfwarp.cu -- FP64 issue cost of a warp instruction with 16 vs 32 active lanes (one warp per sub-partition,
// 8 independent chains per thread): does the 16-lane-wide FP64 pipe skip the empty half of a half-active warp?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int active, long long* out, double* sink, int iters) {
  const int lane = threadIdx.x & 31;
  double a[8];
  for (int i = 0; i < 8; i++) a[i] = 1.0 + lane + i;
  const double m = 1.0 + 1e-9 * lane;
  long long t0 = clock64();
  if (lane < active) {
    for (int i = 0; i < iters; i++) {
#pragma unroll
      for (int j = 0; j < 8; j++) a[j] = __dadd_rn(a[j], m);
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  double s = 0;
  for (int i = 0; i < 8; i++) s += a[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  long long* out; double* sink;
  cudaMalloc(&out, 64); cudaMalloc(&sink, 8 * 148 * 128);
  const int iters = 20000;
  for (int warps = 1; warps <= 4; warps *= 4)
    for (int act = 8; act <= 32; act *= 2) {
      long long h = 0;
      for (int r = 0; r < 2; r++) { k<<<148, 32 * warps>>>(act, out, sink, iters); cudaDeviceSynchronize(); }
      cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      printf("%d warp(s)/CTA (one per sub-partition), %2d active lanes: %.2f cycles per DADD warp-instruction\n", warps, act, (double)h / (iters * 8.0));
    }
  return 0;
}
