// tests/emu/include/cuda_runtime.h -- TEST INFRASTRUCTURE ONLY.
//
// A minimal stand-in for the CUDA runtime and device language so that the kernel SOURCES of roms_b200/csrc can be built
// with g++ and run on the host: the container this repository is developed in has no GPU, and a B200 box costs minutes per
// call, so the per-point arithmetic of a kernel (loop bounds, stencil indices, operation order) is checked here bit-for-bit
// against the oracle before it goes to the GPU.  What this does NOT check: anything about performance, races between
// threads of different blocks (blocks run one after the other), and the halo transport (k_halo.cu is not built; emu_rt.cpp
// replaces it by copies between the mirrors of the ranks of one process).
// The product library libroms_b200.so never includes this header; `roms_b200` never loads the emulation library.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <functional>

#define ROMS_B200_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __constant__
#define __shared__ static
#define __grid_constant__
#define __align__(n)

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
extern thread_local uint3 threadIdx;
extern uint3 blockIdx;
extern dim3 blockDim, gridDim;

namespace emu {
void run_grid(dim3 g, dim3 b, size_t smem, const char* kernel, const std::function<void()>& body);
void run_grid_coop(dim3 g, dim3 b, size_t smem, const char* kernel, const std::function<void()>& body);   // all blocks at once
bool coop_supported();
void barrier();
void* dyn_smem();
void named_barrier(int id, int nthreads, bool wait);     // PTX bar.sync / bar.arrive id, nthreads
bool warp_any(bool pred);                                // __any_sync over the 32 lanes of the calling thread's warp
void yield();
}
#define EMU_LAUNCH(kernel, g, b, smem, ...) emu::run_grid((g), (b), (smem), #kernel, [&]() { kernel(__VA_ARGS__); })
inline void __syncthreads() { emu::barrier(); }
namespace emu { void sync_warp(); double warp_shfl(double v, int src_lane); int lane_id(); }
inline double __shfl_sync(unsigned, double v, int src) { return emu::warp_shfl(v, src); }
inline double __shfl_up_sync(unsigned, double v, unsigned d) { const int l = emu::lane_id(); return emu::warp_shfl(v, l - (int)d >= 0 ? l - (int)d : l); }
inline double __shfl_down_sync(unsigned, double v, unsigned d) { const int l = emu::lane_id(); return emu::warp_shfl(v, l + (int)d <= 31 ? l + (int)d : l); }
inline void __syncwarp() { emu::sync_warp(); }              // rendezvous of the 32 lanes (team kernels), no-op in loop mode
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline bool __any_sync(unsigned, bool pred) { return emu::warp_any(pred); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline long long clock64() { return 0; }
inline int atomicOr(int* p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }

// ---- runtime API subset used by roms_b200.cu / k_*.cu (synchronous, host memory)
typedef int cudaError_t;
enum { cudaSuccess = 0 };
struct emuStream; struct emuEvent { double t; };
typedef emuStream* cudaStream_t; typedef emuEvent* cudaEvent_t; typedef void* cudaGraph_t; typedef void* cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaStreamCaptureModeThreadLocal = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
inline const char* cudaGetErrorString(cudaError_t) { return "emulation"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrMaxSharedMemoryPerBlockOptin = 97, cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {     // a B200; EMU_SM_COUNT changes the CTA count of persistent-style launches
  if (a == cudaDevAttrMaxSharedMemoryPerBlockOptin) *v = 232448;
  else { const char* e = getenv("EMU_SM_COUNT"); *v = e ? atoi(e) : 148; }
  return cudaSuccess;
}
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < h; ++r) memcpy((char*)d + r * dp, (const char*)s + r * sp, w);
  return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
double emu_now_ms();
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent{0.0}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = emu_now_ms(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
// graphs and programmatic launches are never used by the emulation build (ROMS_B200_NO_GRAPH / ROMS_B200_NO_PDL are forced)
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { return 3; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t*) { return 3; }
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, unsigned) { return 3; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return 3; }
struct cudaLaunchAttribute { int id; struct { int programmaticStreamSerializationAllowed; } val; };
enum { cudaLaunchAttributeProgrammaticStreamSerialization = 1 };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes; cudaStream_t stream; cudaLaunchAttribute* attrs; int numAttrs; };
template <class F, class... A> inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t*, F, A...) { return 3; }
