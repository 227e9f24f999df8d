// fp64_interference.cu -- does a dependent FP64 chain on one warp slow down when OTHER warps (same or other SM
// sub-partitions) keep the FP64 pipes / the shared-memory pipe busy?  (Why the serial warps of the persistent
// kernel run 1.5-1.7x slower next to the vector warps than alone.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_interference fp64_interference.cu && ./fp64_interference
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(640, 1) k(int mode, int chain_kind, long long* out, double* sink, int iters) {
  extern __shared__ double sm[];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + i * 1e-9;
  __syncthreads();
  if (wid == 0) {
    // the measured warp: 16 active lanes, dependent chain
    if (lane < 16) {
      double acc = sm[lane];
      const double inc = sm[lane + 32];
      long long t0 = clock64();
      if (chain_kind == 0) {
#pragma unroll 16
        for (int i = 0; i < iters; i++) acc = __dadd_rn(acc, inc);
      } else {  // DADD chain fed by shared-memory loads issued 10 ahead (like the sequential sums)
        const double* p = sm + lane * 65;
        for (int b = 0; b < iters / 10; b++) {
          double v[10];
#pragma unroll
          for (int q = 0; q < 10; q++) v[q] = p[(b * 10 + q) & 63];
#pragma unroll
          for (int q = 0; q < 10; q++) acc = __dadd_rn(acc, v[q]);
        }
      }
      long long t1 = clock64();
      if (lane == 0 && blockIdx.x == 0) out[0] = t1 - t0;
      sink[blockIdx.x * 32 + lane] = acc;
    }
  } else {
    const bool same_smsp = (wid & 3) == 0;
    const bool fp = (mode & 1) && !same_smsp || (mode & 2) && same_smsp;
    const bool ld = (mode & 4) && !same_smsp;
    if (fp) {
      double a0 = sm[lane], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, m = sm[lane + 1];
      for (int i = 0; i < iters * 2; i++) {
        a0 = fma(a0, m, m); a1 = fma(a1, m, m); a2 = fma(a2, m, m); a3 = fma(a3, m, m);
      }
      sink[4096 + blockIdx.x * 640 + threadIdx.x] = a0 + a1 + a2 + a3;
    } else if (ld) {
      double s = 0;
      for (int i = 0; i < iters * 2; i++) {
        s += sm[(lane * 6 + i) & 4095];  // strided (conflicting) shared loads
        sm[(lane * 6 + i * 7 + 2048) & 4095] = s;
      }
      sink[4096 + blockIdx.x * 640 + threadIdx.x] = s;
    }
  }
}

int main() {
  long long* out;
  double* sink;
  cudaMalloc(&out, 64);
  cudaMalloc(&sink, sizeof(double) * (4096 + 148 * 640 + 64));
  const int iters = 20000;
  const char* names[] = {"alone", "FP64 streams on the other 3 sub-partitions (15 warps)", "FP64 stream on the SAME sub-partition (4 warps)",
                         "FP64 on all", "shared-memory traffic on the other sub-partitions", "FP64 others + smem others"};
  const int modes[] = {0, 1, 2, 3, 4, 5};
  for (int ck = 0; ck < 2; ck++)
    for (int m = 0; m < 6; m++) {
      long long h = 0;
      for (int rep = 0; rep < 2; rep++) {
        k<<<148, 640, 4096 * 8>>>(modes[m], ck, out, sink, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      printf("%s chain, %-60s: %.2f cycles per dependent DADD\n", ck ? "LDS-fed" : "register", names[m], (double)h / iters);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
