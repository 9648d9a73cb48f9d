// tools/ubench/tma3d_test.cu -- isolates what the TMA tensor-map path (cp.async.bulk.tensor.3d, fp64 volumes) accepts on B200.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma3d_test tma3d_test.cu
// usage: tma3d_test MODE     (one mode per process: a faulting kernel poisons the context)
//   0  3-D FLOAT64 map, box {20,1,30}, even start column, map passed as a __grid_constant__ kernel parameter
//   1  same, odd start column (box start only 8-byte aligned)
//   2  map nested in a __grid_constant__ struct (what k_step3d_t8.cu does)
//   3  FLOAT32-typed map over the same doubles (dims, box, coordinate doubled in i)
//   4  2-D map (rows = j*k flattened), one load per level
//   5  box {16,1,30} (128-byte rows), even start
//   6  box {20,1,30}, start column -2 and last rows (out-of-bound fill)
//   7  UINT64-typed map, box {20,1,30}, odd start
//   8  map in global memory (cudaMemcpy'd), box {20,1,30}, odd start
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  for (long it = 0; it < 20000000; ++it) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void tma3d(void* dst, const void* m, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(s32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tma2d(void* dst, const void* m, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               ::"r"(s32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(s32(bar)) : "memory");
}

struct Args { CUtensorMap m; int c0, c1, bw, nk, mode; double* out; int* flag; };

// one CTA per row j: loads the box (bw columns starting at c0, row c1+blockIdx.x, all nk levels) and writes it out densely
template <int VIA>   // 0: direct map parameter, 1: struct, 2: global pointer
__global__ void k(const __grid_constant__ CUtensorMap m, const __grid_constant__ Args a, const CUtensorMap* gm) {
  extern __shared__ __align__(128) unsigned char smraw[];
  unsigned char* sm = (unsigned char*)(((uintptr_t)smraw + 127) & ~(uintptr_t)127);
  uint64_t* bar = (uint64_t*)sm;
  double* tile = (double*)(sm + 128);
  const void* mp = VIA == 0 ? (const void*)&m : (VIA == 1 ? (const void*)&a.m : (const void*)gm);
  const int bytes = a.bw * a.nk * 8;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, bytes);
    if (a.mode == 3) tma3d(tile, mp, 2 * a.c0, a.c1 + blockIdx.x, 0, bar);
    else tma3d(tile, mp, a.c0, a.c1 + blockIdx.x, 0, bar);
  }
  __syncthreads();
  if (!mbar_wait(bar, 0)) { if (threadIdx.x == 0) atomicAdd(a.flag, 1); return; }
  for (int q = threadIdx.x; q < a.bw * a.nk; q += blockDim.x) a.out[(size_t)blockIdx.x * a.bw * a.nk + q] = tile[q];
}
// mode 4 variant with the level stride folded into the row coordinate
__global__ void k2d(const __grid_constant__ CUtensorMap m, const __grid_constant__ Args a, int nj) {
  extern __shared__ __align__(128) unsigned char smraw[];
  unsigned char* sm = (unsigned char*)(((uintptr_t)smraw + 127) & ~(uintptr_t)127);
  uint64_t* bar = (uint64_t*)sm;
  double* tile = (double*)(sm + 128);
  const int bytes = a.bw * a.nk * 8;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar, bytes);
    for (int k = 0; k < a.nk; ++k) tma2d(tile + k * a.bw, &m, a.c0, (a.c1 + (int)blockIdx.x) + k * nj, bar);
  }
  __syncthreads();
  if (!mbar_wait(bar, 0)) { if (threadIdx.x == 0) atomicAdd(a.flag, 1); return; }
  for (int q = threadIdx.x; q < a.bw * a.nk; q += blockDim.x) a.out[(size_t)blockIdx.x * a.bw * a.nk + q] = tile[q];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int ni = 518, nj = 67, nk = 30, nrows = 8;
  std::vector<double> h((size_t)ni * nj * nk);
  for (size_t q = 0; q < h.size(); ++q) h[q] = 1.0 + (double)q * 0.25;
  double* d; cudaMalloc(&d, h.size() * 8); cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  void* p = nullptr; cudaDriverEntryPointQueryResult qr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) { printf("mode %d: no encoder\n", mode); return 2; }
  EncodeTiledFn enc = (EncodeTiledFn)p;
  int bw = (mode == 5) ? 16 : 20, c0 = (mode == 0 || mode == 5) ? 2 : (mode == 6 ? -2 : 1), c1 = (mode == 6) ? nj - 3 : 5;
  CUtensorMap m; memset(&m, 0, sizeof(m));
  CUresult r;
  const cuuint32_t es[3] = {1, 1, 1};
  if (mode == 4) {
    const cuuint64_t dim[2] = {(cuuint64_t)ni, (cuuint64_t)nj * nk}; const cuuint64_t str[1] = {(cuuint64_t)ni * 8}; const cuuint32_t box[2] = {(cuuint32_t)bw, 1};
    r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (mode == 3) {
    const cuuint64_t dim[3] = {(cuuint64_t)ni * 2, (cuuint64_t)nj, (cuuint64_t)nk}; const cuuint64_t str[2] = {(cuuint64_t)ni * 8, (cuuint64_t)ni * nj * 8};
    const cuuint32_t box[3] = {(cuuint32_t)bw * 2, 1, (cuuint32_t)nk};
    r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t dim[3] = {(cuuint64_t)ni, (cuuint64_t)nj, (cuuint64_t)nk}; const cuuint64_t str[2] = {(cuuint64_t)ni * 8, (cuuint64_t)ni * nj * 8};
    const cuuint32_t box[3] = {(cuuint32_t)bw, 1, (cuuint32_t)nk};
    r = enc(&m, mode == 7 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  printf("mode %d: encode -> %d\n", mode, (int)r);
  if (r != CUDA_SUCCESS) return 3;
  Args a; memset(&a, 0, sizeof(a)); a.m = m; a.c0 = c0; a.c1 = c1; a.bw = bw; a.nk = nk; a.mode = mode;
  cudaMalloc(&a.out, (size_t)nrows * bw * nk * 8); cudaMemset(a.out, 0, (size_t)nrows * bw * nk * 8);
  cudaMalloc(&a.flag, 4); cudaMemset(a.flag, 0, 4);
  CUtensorMap* gm; cudaMalloc(&gm, sizeof(m)); cudaMemcpy(gm, &m, sizeof(m), cudaMemcpyHostToDevice);
  const size_t smem = 128 + 128 + (size_t)bw * nk * 8;
  if (mode == 4) k2d<<<nrows, 128, smem>>>(m, a, nj);
  else if (mode == 2) k<1><<<nrows, 128, smem>>>(m, a, gm);
  else if (mode == 8) k<2><<<nrows, 128, smem>>>(m, a, gm);
  else k<0><<<nrows, 128, smem>>>(m, a, gm);
  cudaError_t e = cudaDeviceSynchronize();
  printf("mode %d: kernel -> %s\n", mode, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  int flag = 0; cudaMemcpy(&flag, a.flag, 4, cudaMemcpyDeviceToHost);
  std::vector<double> o((size_t)nrows * bw * nk);
  cudaMemcpy(o.data(), a.out, o.size() * 8, cudaMemcpyDeviceToHost);
  size_t bad = 0;
  for (int rr = 0; rr < nrows; ++rr) for (int k = 0; k < nk; ++k) for (int x = 0; x < bw; ++x) {
    const int i = c0 + x, j = c1 + rr;
    const double ref = (i >= 0 && i < ni && j >= 0 && j < nj) ? h[(size_t)i + (size_t)ni * (j + (size_t)nj * k)] : 0.0;
    if (o[((size_t)rr * nk + k) * bw + x] != ref) ++bad;
  }
  printf("mode %d: timeouts %d, mismatches %zu of %zu -> %s\n", mode, flag, bad, o.size(), (bad == 0 && flag == 0) ? "OK" : "FAIL");
  return bad != 0 || flag != 0;
}
