// Microbenchmark: dependent-chain latency and warp-level throughput of fp64 ops on sm_100a.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o fp64_lat fp64_lat.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double* out, long long* cyc, double a, double b, int iters) {
  double x = a + threadIdx.x * 1e-9, y = b;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (OP == 0) x = __dadd_rn(x, y);
      else if (OP == 1) x = __dmul_rn(x, y);
      else if (OP == 2) x = __fma_rn(x, y, y);
      else if (OP == 3) x = 1.0 / x;                       // full IEEE reciprocal
      else if (OP == 4) x = 1.0 / (y - x * 0.25);          // Thomas-like step: mul, add, rcp
      else if (OP == 5) x = y / x;                         // full division
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
// ILP: R independent chains per thread
template <int R>
__global__ void chain_ilp(double* out, long long* cyc, double a, double b, int iters) {
  double x[R];
#pragma unroll
  for (int r = 0; r < R; ++r) x[r] = a + threadIdx.x * 1e-9 + r * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int r = 0; r < R; ++r) x[r] = 1.0 / (b - x[r] * 0.25);
    }
  }
  long long t1 = clock64();
  double s = 0; for (int r = 0; r < R; ++r) s += x[r];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 8);
  const char* names[] = {"DADD", "DMUL", "DFMA", "rcp 1/x", "1/(y-x*c)", "div y/x"};
  long long h;
  const int iters = 256;
#define RUN(OP, blocks, threads) { chain<OP><<<blocks, threads>>>(out, cyc, 1.5, 1.25, iters); chain<OP><<<blocks, threads>>>(out, cyc, 1.5, 1.25, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("%-10s blocks %4d threads %4d : %.1f cycles per op (per warp-chain)\n", names[OP], blocks, threads, (double)h / (iters * 16)); }
  RUN(0, 1, 32) RUN(1, 1, 32) RUN(2, 1, 32) RUN(3, 1, 32) RUN(4, 1, 32) RUN(5, 1, 32)
  // throughput: many warps per SM (cycles per op per warp shrink until the pipe saturates)
  RUN(2, 148, 128) RUN(2, 148, 256) RUN(2, 148, 512) RUN(2, 148, 1024)
  RUN(0, 148, 1024) RUN(4, 148, 128) RUN(4, 148, 512) RUN(4, 148, 1024)
#define RUNI(R, blocks, threads) { chain_ilp<R><<<blocks, threads>>>(out, cyc, 1.5, 1.25, iters); chain_ilp<R><<<blocks, threads>>>(out, cyc, 1.5, 1.25, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("thomas-step ILP %d blocks %4d threads %4d : %.1f cycles per step per chain-warp, %.1f per step\n", R, blocks, threads, (double)h / (iters * 4), (double)h / (iters * 4 * R)); }
  RUNI(1, 148, 64) RUNI(2, 148, 64) RUNI(4, 148, 64) RUNI(8, 148, 64) RUNI(4, 148, 128) RUNI(4, 148, 256)
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
