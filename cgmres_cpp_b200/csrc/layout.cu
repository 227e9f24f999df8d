// layout.cu -- instance-major (the reference's per-object arrays, one object after
// another) <-> structure-of-arrays [row][instance] used by the kernels.
// Tiled through shared memory so both the global read and the global write are coalesced.
#include "kernel_args.h"

namespace cgmres_b200 {
namespace {
constexpr int TI = 32;  // instances per tile
constexpr int TR = 32;  // rows per tile

__global__ void aos_to_soa_kernel(const double* __restrict__ aos, double* __restrict__ soa, int64_t n, int rows,
                                  int64_t ld) {
  __shared__ double tile[TI][TR + 1];
  const int64_t n0 = (int64_t)blockIdx.x * TI;
  const int r0 = blockIdx.y * TR;
  // read: consecutive threads walk along a row-chunk of one instance (contiguous in aos)
  for (int i = threadIdx.y; i < TI; i += blockDim.y) {
    const int64_t nn = n0 + i;
    const int r = r0 + threadIdx.x;
    if (nn < n && r < rows) tile[i][threadIdx.x] = aos[nn * rows + r];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < TR; r += blockDim.y) {
    const int64_t nn = n0 + threadIdx.x;
    const int rr = r0 + r;
    if (nn < n && rr < rows) soa[(int64_t)rr * ld + nn] = tile[threadIdx.x][r];
  }
}

__global__ void soa_to_aos_kernel(const double* __restrict__ soa, double* __restrict__ aos, int64_t n, int rows,
                                  int64_t ld) {
  __shared__ double tile[TI][TR + 1];
  const int64_t n0 = (int64_t)blockIdx.x * TI;
  const int r0 = blockIdx.y * TR;
  for (int r = threadIdx.y; r < TR; r += blockDim.y) {
    const int64_t nn = n0 + threadIdx.x;
    const int rr = r0 + r;
    if (nn < n && rr < rows) tile[threadIdx.x][r] = soa[(int64_t)rr * ld + nn];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < TI; i += blockDim.y) {
    const int64_t nn = n0 + i;
    const int r = r0 + threadIdx.x;
    if (nn < n && r < rows) aos[nn * rows + r] = tile[i][threadIdx.x];
  }
}

__global__ void broadcast_rows_kernel(const double* __restrict__ aos, double* __restrict__ soa, int64_t n,
                                      int rows_per, int reps, int64_t ld) {
  const int64_t nn = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (nn >= n) return;
  for (int j = 0; j < rows_per; j++) {
    const double v = aos[nn * rows_per + j];
    for (int i = 0; i < reps; i++) soa[(int64_t)(i * rows_per + j) * ld + nn] = v;
  }
}
__global__ void broadcast_inst_kernel(const double* __restrict__ aos, double* __restrict__ dst, int64_t n,
                                      int rows_per, int reps) {
  const int64_t total = n * (int64_t)rows_per * reps;
  const int64_t per = (int64_t)rows_per * reps;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t nn = t / per;
    const int j = (int)((t % per) % rows_per);
    dst[t] = aos[nn * rows_per + j];
  }
}
}  // namespace

cudaError_t launch_broadcast_inst(const double* aos, double* dst, int64_t n, int rows_per, int reps, cudaStream_t s) {
  if (n == 0 || rows_per == 0 || reps == 0) return cudaSuccess;
  const int64_t total = n * (int64_t)rows_per * reps;
  const unsigned grid = (unsigned)((total + 255) / 256 > 65535 * 16 ? 65535 * 16 : (total + 255) / 256);
  broadcast_inst_kernel<<<grid, 256, 0, s>>>(aos, dst, n, rows_per, reps);
  return cudaGetLastError();
}

cudaError_t launch_aos_to_soa(const double* aos, double* soa, int64_t n, int rows, int64_t ld, cudaStream_t s) {
  if (n == 0 || rows == 0) return cudaSuccess;
  dim3 grid((unsigned)((n + TI - 1) / TI), (unsigned)((rows + TR - 1) / TR));
  aos_to_soa_kernel<<<grid, dim3(32, 8), 0, s>>>(aos, soa, n, rows, ld);
  return cudaGetLastError();
}
cudaError_t launch_soa_to_aos(const double* soa, double* aos, int64_t n, int rows, int64_t ld, cudaStream_t s) {
  if (n == 0 || rows == 0) return cudaSuccess;
  dim3 grid((unsigned)((n + TI - 1) / TI), (unsigned)((rows + TR - 1) / TR));
  soa_to_aos_kernel<<<grid, dim3(32, 8), 0, s>>>(soa, aos, n, rows, ld);
  return cudaGetLastError();
}
cudaError_t launch_broadcast_rows(const double* aos, double* soa, int64_t n, int rows_per, int reps, int64_t ld,
                                  cudaStream_t s) {
  if (n == 0 || rows_per == 0 || reps == 0) return cudaSuccess;
  broadcast_rows_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(aos, soa, n, rows_per, reps, ld);
  return cudaGetLastError();
}

}  // namespace cgmres_b200
