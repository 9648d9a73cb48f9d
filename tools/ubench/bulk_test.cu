// Validates the cp.async.bulk (1-D TMA bulk copy) + mbarrier sequence used by k_step3d_t6.cu on sm_100a.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bulk_test bulk_test.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
constexpr int ROW = 34, NARR = 10, NSTG = 2;
__global__ void k(const double* __restrict__ in, double* __restrict__ out, int nrows, int pitch) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  double* stg = (double*)smraw + (size_t)w * NSTG * NARR * ROW;
  uint64_t* bars = (uint64_t*)((double*)smraw + (size_t)nw * NSTG * NARR * ROW) + w * NSTG;
  if (lane == 0) { for (int s = 0; s < NSTG; ++s) mbar_init(&bars[s], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  auto issue = [&](int r) {
    const int s = r % NSTG;
    if (lane == 0) {
      mbar_expect_tx(&bars[s], NARR * ROW * 8);
      for (int a = 0; a < NARR; ++a) bulk_g2s(stg + (s * NARR + a) * ROW, in + ((size_t)(blockIdx.x * nw + w) * nrows + r) * pitch + a * ROW, ROW * 8, &bars[s]);
    }
  };
  issue(0);
  for (int r = 0; r < nrows; ++r) {
    if (r + 1 < nrows) issue(r + 1);
    const int s = r % NSTG;
    mbar_wait(&bars[s], (r / NSTG) & 1);
    double acc = 0;
    for (int a = 0; a < NARR; ++a) acc += stg[(s * NARR + a) * ROW + lane + 1];
    out[((size_t)(blockIdx.x * nw + w) * nrows + r) * 32 + lane] = acc;
    __syncwarp();          // all lanes done with stage s before it is refilled (next-next iteration)
  }
}
int main() {
  const int nblk = 148, nw = 16, nrows = 64, pitch = NARR * ROW;   // pitch*8 = 2720 B (multiple of 16)
  const size_t n = (size_t)nblk * nw * nrows * pitch;
  std::vector<double> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (double)(i % 1000) * 0.5;
  double *din, *dout; cudaMalloc(&din, n * 8); cudaMalloc(&dout, (size_t)nblk * nw * nrows * 32 * 8);
  cudaMemcpy(din, h.data(), n * 8, cudaMemcpyHostToDevice);
  const size_t smem = (size_t)nw * NSTG * NARR * ROW * 8 + nw * NSTG * 8;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<nblk, nw * 32, smem>>>(din, dout, nrows, pitch);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<double> o((size_t)nblk * nw * nrows * 32);
  cudaMemcpy(o.data(), dout, o.size() * 8, cudaMemcpyDeviceToHost);
  size_t bad = 0;
  for (size_t q = 0; q < o.size(); ++q) {
    const size_t rowid = q / 32; const int lane = q % 32;
    double acc = 0; for (int a = 0; a < NARR; ++a) acc += h[rowid * pitch + a * ROW + lane + 1];
    if (acc != o[q]) ++bad;
  }
  printf("mismatches: %zu of %zu\n", bad, o.size());
  return bad != 0;
}
