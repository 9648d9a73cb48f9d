// roms_b200/csrc/k_step3d_t8.cu -- step3d_t_tile as a TMA-fed, mbarrier-pipelined 2.5-D blocked sweep (production layout, sm_100a).
//
// Reference: Nonlinear/step3d_t.F:393-399 (1/Hz), :641-916 (U3 horizontal advection of t(3)), :1150-1365 (C4 vertical advection),
// :1672-1721 (spline implicit vertical diffusion), t3dbc_im.F:334-341,415-422 + exchange_3d.F (wall rows, E-W periodic images).
//
// A persistent CTA (one per SM) takes work items = (i-stripe of 16 water columns) x (chunk of JCH rows) and marches along j.
// NO thread of the compute warps ever loads from global memory: every operand is staged into shared memory by TMA
// (cp.async.bulk.tensor.3d, j x k tiles of the 3-D volumes: box = {columns, 1 row, all N levels}) issued by one elected
// thread a full row or more ahead, completion is signalled on mbarriers (expect_tx / complete_tx):
//   * t(3) rows land in a 4-deep ring (rows j, j+1, j+2 in use, row j+3 in flight): the 5-point x/eta stencil and the 4-point
//     vertical stencil read it with LDS; the eta-direction is rolled through registers (the north-face flux of row j is the
//     south-face flux of row j+1), so each t(3) row is fetched once per stripe;
//   * the read-once operands of a row (t(nnew) x2, Akt x2, Hz, Hvom(j+1), Huon, W) land in one of S "slots"; the slot is then
//     worked on IN PLACE: producer warps (level-parallel, a warp = 16 columns x 2 levels, both tracers per thread) overwrite
//     t(nnew) by q = (t(nnew) - dt*pm*pn*div F)/Hz and Hvom by 1/Hz; a consumer warp (lane = column x tracer) then runs the
//     spline tridiagonal (Thomas, software-pipelined, operation order of the reference) out of the slot, parks CF/DC in the
//     slot's dead Huon/W areas and writes t(nnew) with its periodic images and wall rows.
// Slots and ring rows are recycled through full/ready/empty mbarriers; roles never meet at a CTA-wide barrier.
// Per-point arithmetic (operation order, no FMA contraction, rcp_ieee == 1.0/x) is that of k_step3d_t6.cu, which is
// bit-identical to the oracle; tests/emu runs this file on the CPU (mbarriers and TMA boxes emulated) before it goes to a GPU.
#include "common.cuh"
#include <cstdint>
#include <cstring>
#include <vector>
#ifndef ROMS_B200_EMU
#include <cuda.h>       // CUtensorMap and its enums (types only: the encoder is fetched with cudaGetDriverEntryPoint)
#else
#include <mutex>
#endif

namespace {
constexpr int BI = 16;            // water columns per stripe (a half-warp)
constexpr int TW = BI + 4;        // t(3) ring row: columns i0-2 .. i0+17
constexpr int HUW = BI + 2;       // Huon row: columns i0 .. i0+17 (i0+16 is used; 18 keeps the box rows 16-byte multiples)
constexpr int RING = 4;           // t(3) rows resident per stripe
constexpr int MAXS = 8;           // most slots
constexpr int LS = BI;            // doubles per level of a slot array
constexpr int BARD = 48;          // doubles reserved for the mbarriers (32) and the tensor-memory base word
constexpr int RPAD = 48;          // doubles in front of the ring: the vertical stencil of level 1 reaches two (unused) levels below a row
__host__ __device__ constexpr int pad16(int n) { return (n + 15) & ~15; }

#ifdef ROMS_B200_EMU
struct TMap { const double* base; long dim[3]; long stride[3]; int box[3]; };   // strides in elements
#else
typedef CUtensorMap TMap;
#endif

struct alignas(64) A8 {
  TMap t3[2], tw[2], ak[2], hz, hv, hu, w;      // tensor maps (3-D: i, j, k) of the volumes this launch reads
  double* out[2];                               // t(:,:,:,nnew,itrc) volumes (stores)
  const double *pm, *pn;
  int* err;
  double dt;
  int N, ni, sk, LBi, LBj, Lm, backoff, pfd;
  long long* prof;                              // -DS3T_PROF builds only: cycles per role and phase (tools/prof_step3d_t.py)
  int i0, i1, ib, j0, j1, Jstr, Jend, wallS, wallN, wrapEW;   // ib: first column of stripe 0 (<= i0, see k_step3d_t_v8)
  int nstripes, JCH, nitems, S, NC, NP;
};

// shared-memory layout of one slot (offsets in doubles; every TMA destination is a multiple of 16 doubles = 128 bytes)
template <int NTR, int TM> struct Lay {
  int N;
  __host__ __device__ int q(int c) const { return c * N * LS; }                               // t(nnew) -> q           [N][16]
  __host__ __device__ int ak(int c) const { return NTR * N * LS + c * (N + 1) * LS; }         // Akt, levels 0..N       [N+1][16]
  __host__ __device__ int hz() const { return NTR * N * LS + NTR * (N + 1) * LS; }            // Hz                     [N][16]
  __host__ __device__ int hv() const { return hz() + N * LS; }                                // Hvom(j+1) -> 1/Hz      [N][16]
  __host__ __device__ int hu() const { return hv() + N * LS; }                                // Huon -> CF of tracer 0 [N][18]
  __host__ __device__ int w() const { return hu() + pad16(N * HUW); }                         // W 0..N -> CF of tracer 1 [N+1][16]
  __host__ __device__ int dc(int c) const { return w() + (N + 1) * LS + c * N * LS; }         // DC                     [N][16]
  __host__ __device__ int slot() const { return TM ? w() + (N + 1) * LS : dc(0) + NTR * N * LS; }   // TM != 0: no DC arrays
  __host__ __device__ int cf(int c) const { return c == 0 ? hu() : w(); }
  __host__ __device__ int ring_tr() const { return pad16(N * TW); }                           // one tracer of a ring row [N][20]
  __host__ __device__ int ring_row() const { return NTR * ring_tr(); }
  __host__ __device__ size_t bytes(int S) const { return 8 * BARD + 128 + sizeof(double) * (RPAD + (size_t)RING * ring_row() + (size_t)S * slot()); }
  __host__ __device__ unsigned slot_tx() const {                                              // bytes TMA delivers into a full slot
    return 8u * (unsigned)(NTR * N * LS + NTR * (N + 1) * LS + N * LS + N * LS + N * HUW + (N + 1) * LS);
  }
};

#ifdef ROMS_B200_EMU
// ---- host emulation of mbarriers and TMA box loads (tests/emu): the threads of a block are cooperative fibers, a wait yields.
// word: bit 63 phase, bits 48..62 arrival count per phase, bits 32..47 pending arrivals, bits 0..31 pending transaction bytes (signed)
struct EmuBar { int32_t tx; uint16_t pending; uint16_t init_phase; };
static std::mutex emu_bar_mu;                    // the AddressSanitizer build runs the threads of a block as OS threads
__device__ __forceinline__ EmuBar* eb(uint64_t* b) { return (EmuBar*)b; }
__device__ __forceinline__ void eb_check(EmuBar* e) { if (e->pending == 0 && e->tx == 0) { e->pending = e->init_phase & 0x7fff; e->init_phase ^= 0x8000; } }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) { std::lock_guard<std::mutex> lk(emu_bar_mu); EmuBar* e = eb(b); e->tx = 0; e->pending = (uint16_t)count; e->init_phase = (uint16_t)count; }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { std::lock_guard<std::mutex> lk(emu_bar_mu); EmuBar* e = eb(b); if (e->pending == 0) abort(); --e->pending; eb_check(e); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) { std::lock_guard<std::mutex> lk(emu_bar_mu); EmuBar* e = eb(b); e->tx += (int32_t)bytes; if (e->pending == 0) abort(); --e->pending; eb_check(e); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity, int*, unsigned = 0) {
  for (;;) {
    { std::lock_guard<std::mutex> lk(emu_bar_mu); if (((eb(b)->init_phase >> 15) & 1u) != parity) return; }
    emu::yield();
  }
}
__device__ __forceinline__ void tma3d(double* dst, const TMap* m, int c0, int c1, int c2, uint64_t* bar) {
  if ((c0 & 1) || ((uintptr_t)dst & 127)) { fprintf(stderr, "emu: TMA box start not 16-byte aligned in global memory (column %d) or destination not 128-byte aligned\n", c0); abort(); }
  for (int z = 0; z < m->box[2]; ++z) for (int y = 0; y < m->box[1]; ++y) for (int x = 0; x < m->box[0]; ++x) {
    const long X = c0 + x, Y = c1 + y, Z = c2 + z;
    const bool in = X >= 0 && X < m->dim[0] && Y >= 0 && Y < m->dim[1] && Z >= 0 && Z < m->dim[2];
    dst[((size_t)z * m->box[1] + y) * m->box[0] + x] = in ? m->base[X + m->stride[1] * Y + m->stride[2] * Z] : 0.0;
  }
  std::lock_guard<std::mutex> lk(emu_bar_mu);
  EmuBar* e = eb(bar); e->tx -= 8 * m->box[0] * m->box[1] * m->box[2]; eb_check(e);
}
__device__ __forceinline__ void fence_barrier_init() {}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ int __double2hiint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32); }
__device__ __forceinline__ int __double2loint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u & 0xffffffffu); }
__device__ __forceinline__ double __hiloint2double(int hi, int lo) { uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double x; memcpy(&x, &u, 8); return x; }
// tensor memory of the running CTA: 128 lanes x 512 32-bit columns; a warp reaches the 32 lanes of its quadrant (warp % 4)
static uint32_t emu_tmem[128][512];
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) { if (ncols > 512) abort(); *dst = 0; }
__device__ __forceinline__ void tmem_dealloc(uint32_t, uint32_t) {}
__device__ __forceinline__ void tmem_fence_before() {}
__device__ __forceinline__ void tmem_fence_after() {}
__device__ __forceinline__ void tmem_wait_ld() {}
__device__ __forceinline__ void tmem_wait_st() {}
__device__ __forceinline__ void tmem_st2d(uint32_t taddr, double x, double y) {
  const int ln = (int)(taddr >> 16) + (int)(threadIdx.x & 31), col = (int)(taddr & 0xffffu);
  if (ln < 0 || ln > 127 || col < 0 || col + 4 > 512 || ((taddr >> 16) != 32u * ((threadIdx.x >> 5) & 3u))) { fprintf(stderr, "emu: tensor-memory access outside the warp's quadrant or the columns\n"); abort(); }
  memcpy(&emu_tmem[ln][col], &x, 8); memcpy(&emu_tmem[ln][col + 2], &y, 8);
}
__device__ __forceinline__ void tmem_ld2d(uint32_t taddr, double& x, double& y) {
  const int ln = (int)(taddr >> 16) + (int)(threadIdx.x & 31), col = (int)(taddr & 0xffffu);
  if (ln < 0 || ln > 127 || col < 0 || col + 4 > 512 || ((taddr >> 16) != 32u * ((threadIdx.x >> 5) & 3u))) { fprintf(stderr, "emu: tensor-memory access outside the warp's quadrant or the columns\n"); abort(); }
  memcpy(&x, &emu_tmem[ln][col], 8); memcpy(&y, &emu_tmem[ln][col + 2], 8);
}
__device__ __forceinline__ void tmem_st4d(uint32_t taddr, double x, double y, double z, double w) { tmem_st2d(taddr, x, y); tmem_st2d(taddr + 4, z, w); }
__device__ __forceinline__ void tmem_ld4d(uint32_t taddr, double& x, double& y, double& z, double& w) { tmem_ld2d(taddr, x, y); tmem_ld2d(taddr + 4, z, w); }
__device__ __forceinline__ void tmem_st1d(uint32_t taddr, double x) {
  const int ln = (int)(taddr >> 16) + (int)(threadIdx.x & 31), col = (int)(taddr & 0xffffu);
  if (ln < 0 || ln > 127 || col < 0 || col + 2 > 512) { fprintf(stderr, "emu: tensor-memory access outside the columns\n"); abort(); }
  memcpy(&emu_tmem[ln][col], &x, 8);
}
__device__ __forceinline__ void tmem_ld1d(uint32_t taddr, double& x) {
  const int ln = (int)(taddr >> 16) + (int)(threadIdx.x & 31), col = (int)(taddr & 0xffffu);
  if (ln < 0 || ln > 127 || col < 0 || col + 2 > 512) { fprintf(stderr, "emu: tensor-memory access outside the columns\n"); abort(); }
  memcpy(&x, &emu_tmem[ln][col], 8);
}
__device__ __forceinline__ void tma3d_prefetch(const TMap*, int, int, int) {}
__device__ __forceinline__ double* align128(double* p) { return (double*)(((uintptr_t)p + 127) & ~(uintptr_t)127); }
#else
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
// try_wait suspends the warp in hardware until the phase completes or the time hint (ns) passes, so a waiting warp takes no
// issue slots from the working warps of its scheduler; a wait that lasts longer than ~4 s (a lost arrival: a bug) raises the
// device error word and traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity, int* err, unsigned backoff = 0) {
  uint32_t ok; unsigned spins = 0; unsigned long long t0 = 0;
  for (;;) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok) : "r"(s32(b)), "r"(parity), "r"(20000u) : "memory");
    if (ok) return;
    if (backoff) __nanosleep(backoff);
    if ((++spins & 255u) == 0) {
      unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (!t0) t0 = t; else if (t - t0 > 4000000000ull) { atomicOr(err, 4); __threadfence(); asm volatile("trap;"); }
    }
  }
}
__device__ __forceinline__ void tma3d(double* dst, const TMap* m, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(s32(dst)), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// ---- tensor memory (TMEM, 256 KB per SM: 128 lanes x 512 32-bit columns) as per-thread scratch of the consumer warps: a warp
// reaches the 32 lanes of its quadrant (warp % 4), thread i its lane i, columns are addressed dynamically.  The Thomas
// coefficients CF(k), DC(k) of a (column, tracer) live in 4 consecutive columns per level: one tcgen05.st per level in the forward
// sweep, one tcgen05.ld per level in the backward sweep, no shared-memory space or bandwidth.
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t ncols) {                 // one full warp; ncols: power of two >= 32
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) { asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory"); }
__device__ __forceinline__ void tmem_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st2d(uint32_t taddr, double x, double y) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
               ::"r"(taddr), "r"(__double2loint(x)), "r"(__double2hiint(x)), "r"(__double2loint(y)), "r"(__double2hiint(y)) : "memory");
}
__device__ __forceinline__ void tmem_ld2d(uint32_t taddr, double& x, double& y) {            // values valid after tmem_wait_ld()
  int a, b, c, d;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr) : "memory");
  x = __hiloint2double(b, a); y = __hiloint2double(d, c);
}
__device__ __forceinline__ void tmem_st4d(uint32_t taddr, double x, double y, double z, double w) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(__double2loint(x)), "r"(__double2hiint(x)), "r"(__double2loint(y)), "r"(__double2hiint(y)),
                 "r"(__double2loint(z)), "r"(__double2hiint(z)), "r"(__double2loint(w)), "r"(__double2hiint(w)) : "memory");
}
__device__ __forceinline__ void tmem_ld4d(uint32_t taddr, double& x, double& y, double& z, double& w) {
  int a, b, c, d, e, f, g, h;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "r"(taddr) : "memory");
  x = __hiloint2double(b, a); y = __hiloint2double(d, c); z = __hiloint2double(f, e); w = __hiloint2double(h, g);
}
__device__ __forceinline__ void tmem_st1d(uint32_t taddr, double x) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__double2loint(x)), "r"(__double2hiint(x)) : "memory");
}
__device__ __forceinline__ void tmem_ld1d(uint32_t taddr, double& x) {
  int a, b;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
  x = __hiloint2double(b, a);
}
// L2 prefetch of a TMA box (the loader requests the operands of a later row so that the real copy finds them in L2)
__device__ __forceinline__ void tma3d_prefetch(const TMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(m), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// pointer arithmetic on the shared array (a round trip through uintptr_t would lose the address space: generic LD/ST instead of LDS/STS)
__device__ __forceinline__ double* align128(double* p) { return p + (((128u - (s32(p) & 127u)) & 127u) >> 3); }
#endif

// max(h,0) and min(h,0) by masking with the sign: 5 integer instructions for both (the compare+select form is turned into
// DSETP.MAX/MIN + NaN fix-ups by the compiler, ~6 each).  Same values as max/min up to the sign of a zero, which contributes
// nothing to a flux (as in k_step3d_t6.cu).
__device__ __forceinline__ void posneg(double h, double& hp, double& hn) {
  const int hi = __double2hiint(h), lo = __double2loint(h);
  const int m = hi >> 31;                       // all ones for a negative value
  hp = __hiloint2double(hi & ~m, lo & ~m);
  hn = __hiloint2double(hi & m, lo & m);
}

// C4 vertical flux at w-level k from t(k-1),t(k),t(k+1),t(k+2) (step3d_t.F:1150-1185)
__device__ __forceinline__ double vflux8(int k, int N, double tm1, double t0, double tp1, double tp2, double w) {
  if (k <= 0 || k >= N) return 0.0;
  if (k == 1) return w * (0.5 * t0 + (7.0 / 12.0) * tp1 - (1.0 / 12.0) * tp2);
  if (k == N - 1) return w * (0.5 * tp1 + (7.0 / 12.0) * t0 - (1.0 / 12.0) * tm1);
  return w * ((7.0 / 12.0) * (t0 + tp1) - (1.0 / 12.0) * (tm1 + tp2));
}
}  // namespace

// ring of n entries walked in order: index and phase parity without integer division
#ifdef S3T_PROF
#define PROF_T(v) const long long v = clock64()
#define PROF_ADD(acc, t0) acc += clock64() - (t0)
#else
#define PROF_T(v)
#define PROF_ADD(acc, t0)
#endif
struct Ring8 {
  int i, n; unsigned ph;
  __device__ __forceinline__ Ring8(int n_) : i(0), n(n_), ph(0) {}
  __device__ __forceinline__ void next() { if (++i == n) { i = 0; ph ^= 1u; } }
};

// KP: level-pair batches per producer warp (NP * KP >= ceil(N/2))
// TM: 0 = CF/DC of the tridiagonal solve in the slot; 1 = CF/DC in tensor memory (smaller slots -> more of them, more consumer
// warps); 2 = everything the back substitution needs (CF, DC, q, Akt, dt/Hz per level) in tensor memory: the consumer hands the
// slot back to the loader right after its forward sweep, the backward sweep and the stores run out of tensor memory
// MAXT: launch bound (384: up to 12 warps with up to 168 registers per thread; 512: up to 16 warps with 128; 640: up to 20 with 96)
template <int NTR, int KP, int TM, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) step3d_t_v8_kernel(const __grid_constant__ A8 a) {
  constexpr int TMC = (TM == 2) ? 10 : 4;         // tensor-memory columns per level
  extern __shared__ __align__(128) double sm_raw[];
  double* sm = align128(sm_raw);
  const int N = a.N, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int S = a.S, NC = a.NC, NP = a.NP;
  uint64_t* ring_full = (uint64_t*)sm;          // [RING]  TMA has filled ring row
  uint64_t* ring_empty = ring_full + RING;      // [RING]  every producer warp is done with the row
  uint64_t* slot_full = ring_empty + RING;      // [MAXS]  TMA has filled the slot
  uint64_t* slot_ready = slot_full + MAXS;      // [MAXS]  every producer warp has written q, 1/Hz
  uint64_t* slot_empty = slot_ready + MAXS;     // [MAXS]  the consumer warp is done with the slot
  const Lay<NTR, TM> L{N};
  const int trD = L.ring_tr(), rowD = L.ring_row(), slotD = L.slot();
  uint32_t* tmem_word = (uint32_t*)(sm + 32);   // base address of the tensor-memory allocation (TM)
  double* ring = sm + BARD + RPAD;
  double* slots = ring + RING * rowD;
  if (threadIdx.x == 0) {
    for (int q = 0; q < RING; ++q) { mbar_init(&ring_full[q], 1); mbar_init(&ring_empty[q], NP); }
    for (int q = 0; q < MAXS; ++q) { mbar_init(&slot_full[q], 1); mbar_init(&slot_ready[q], NP); mbar_init(&slot_empty[q], 1); }
    fence_barrier_init();
  }
  uint32_t tm_cols = 32;                        // 4 (CF, DC) or 10 (CF, DC, q, Akt, dt/Hz) 32-bit columns per level, a power of two
  while (tm_cols < (TM == 2 ? 10u : 4u) * (unsigned)N) tm_cols *= 2;
  if (TM && warp == 1) tmem_alloc(tmem_word, tm_cols);
  if (TM) tmem_fence_before();
  __syncthreads();
  if (TM) tmem_fence_after();
  const uint32_t tm_base = TM ? *tmem_word : 0u;
  const double dt = a.dt, c16 = 1.0 / 6.0;
  int bad = 0;

  if (warp == 0) {
    // ======================= loader: one thread issues every TMA copy of the CTA =======================
    if (lane == 0) {
      Ring8 g(RING), q(S);                      // ring rows / slots issued so far
#ifdef S3T_PROF
      long long p_ring = 0, p_slot = 0; const long long p_t0 = clock64();
#endif
      for (int it = blockIdx.x; it < a.nitems; it += gridDim.x) {
        const int stripe = it % a.nstripes, chunk = it / a.nstripes;
        const int i0s = a.ib + stripe * BI, ja = a.j0 + chunk * a.JCH, jb = min(ja + a.JCH - 1, a.j1);
        const int nrows = jb - ja + 1, ci = i0s - a.LBi;
        for (int r = 0; r < nrows + 4; ++r) {
          const int n = ja - 2 + r;             // t(3) row that arrives at this step
          {
            { PROF_T(t_); mbar_wait(&ring_empty[g.i], g.ph ^ 1u, a.err); PROF_ADD(p_ring, t_); }
            mbar_arrive_expect_tx(&ring_full[g.i], 8u * (unsigned)(NTR * N * TW));
#pragma unroll
            for (int c = 0; c < NTR; ++c) tma3d(ring + g.i * rowD + c * trD, &a.t3[c], ci - 2, n - a.LBj, 0, &ring_full[g.i]);
            g.next();
          }
          if (r >= 3) {                         // operands of row j = n-2 (r == 3: the row below the chunk, only its Hvom(j+1) is needed)
            const int cj = n - 2 - a.LBj, sl = q.i;
            { PROF_T(t_); mbar_wait(&slot_empty[sl], q.ph ^ 1u, a.err); PROF_ADD(p_slot, t_); }
            double* sp = slots + sl * slotD;
            if (r == 3) {
              mbar_arrive_expect_tx(&slot_full[sl], 8u * (unsigned)(N * LS));
              tma3d(sp + L.hv(), &a.hv, ci, cj + 1, 0, &slot_full[sl]);
            } else {
              mbar_arrive_expect_tx(&slot_full[sl], L.slot_tx());
#pragma unroll
              for (int c = 0; c < NTR; ++c) {
                tma3d(sp + L.q(c), &a.tw[c], ci, cj, 0, &slot_full[sl]);
                tma3d(sp + L.ak(c), &a.ak[c], ci, cj, 0, &slot_full[sl]);
              }
              tma3d(sp + L.hz(), &a.hz, ci, cj, 0, &slot_full[sl]);
              tma3d(sp + L.hv(), &a.hv, ci, cj + 1, 0, &slot_full[sl]);
              tma3d(sp + L.hu(), &a.hu, ci, cj, 0, &slot_full[sl]);
              tma3d(sp + L.w(), &a.w, ci, cj, 0, &slot_full[sl]);
            }
            q.next();
          }
          if (a.pfd > 0) {                      // L2 prefetch of what the step `pfd` steps ahead will copy
            const int rp = r + a.pfd;
            if (rp < nrows + 4) {
              const int np_ = ja - 2 + rp;
#pragma unroll
              for (int c = 0; c < NTR; ++c) tma3d_prefetch(&a.t3[c], ci - 2, np_ - a.LBj, 0);
              if (rp >= 4) {
                const int cjp = np_ - 2 - a.LBj;
#pragma unroll
                for (int c = 0; c < NTR; ++c) { tma3d_prefetch(&a.tw[c], ci, cjp, 0); tma3d_prefetch(&a.ak[c], ci, cjp, 0); }
                tma3d_prefetch(&a.hz, ci, cjp, 0); tma3d_prefetch(&a.hv, ci, cjp + 1, 0); tma3d_prefetch(&a.hu, ci, cjp, 0); tma3d_prefetch(&a.w, ci, cjp, 0);
              }
            }
          }
        }
      }
#ifdef S3T_PROF
      if (a.prof) { atomicAdd((unsigned long long*)&a.prof[0], (unsigned long long)(clock64() - p_t0)); atomicAdd((unsigned long long*)&a.prof[1], (unsigned long long)p_ring); atomicAdd((unsigned long long*)&a.prof[2], (unsigned long long)p_slot); }
#endif
    }
  } else if (warp > NC) {
    // ======================= producers: advection, level-parallel =======================
    const int p = warp - 1 - NC, h = lane >> 4, col = lane & 15;
    double Cj[KP][NTR], FEs[KP][NTR];           // eta-direction carry per (batch, tracer): curv(j), FE(j) (south face of the next row)
    // A thread owns KP CONSECUTIVE levels of its column (half-warp h of producer warp p: levels KP*(2p+h)+1 ...), one per batch m,
    // so the vertical stencil walks through registers: t(3) at k-1, k, k+1 and the flux through the lower face of level k are
    // those of the batch before (one new t(3) level and one vertical flux per level instead of five and two).
    // per batch: level k; element offset of (level k, this column) inside one tracer of a ring row / inside a slot array / inside
    // the Huon array.  The vertical neighbours k-2..k+2 sit at fixed distances; at k <= 2 and k >= N-1 they fall outside the
    // tracer's levels (padding, the neighbouring tracer or row): those values are not used by the C4 flux at these levels.
    int kk[KP], o0[KP], so[KP], sh[KP];
#pragma unroll
    for (int m = 0; m < KP; ++m) {
      const int kl = KP * 2 * p + m + 1;        // level of the lower half-warp in this batch (warp-uniform)
      kk[m] = (kl <= N) ? min(KP * (2 * p + h) + m + 1, N + 1) : 0;   // 0: no such batch; N+1: shadow of level N, not stored
      const int kc = min(max(kk[m], 1), N);
      o0[m] = (kc - 1) * TW + col + 2;
      so[m] = (kc - 1) * LS + col;
      sh[m] = (kc - 1) * HUW + col;
    }
    const int oq = L.q(0), dq_ = L.q(1) - L.q(0), ohz_ = L.hz(), ohv = L.hv(), ohu = L.hu(), ow = L.w();
    int gi = 0;                                 // ring buffer of the row that arrives at this step
    unsigned gph = 0;
#ifdef S3T_PROF
    long long p_ring = 0, p_slot = 0, p_work = 0; const long long p_t0 = clock64();
#endif
    Ring8 q(S);
    for (int it = blockIdx.x; it < a.nitems; it += gridDim.x) {
      const int stripe = it % a.nstripes, chunk = it / a.nstripes;
      const int i0s = a.ib + stripe * BI, ja = a.j0 + chunk * a.JCH, jb = min(ja + a.JCH - 1, a.j1);
      const int nrows = jb - ja + 1;
      const bool act = (i0s + col >= a.i0) && (i0s + col <= a.i1);
      const int ic = min(max(i0s + col, a.i0), a.i1);   // idle lanes shadow an active column for the 2-D metric loads
      const bool southw = a.wallS && ja == a.Jstr;
      int o2 = (ic - a.LBi) + a.ni * (ja - a.LBj);
      double cffn = dt * __ldg(a.pm + o2) * __ldg(a.pn + o2);         // dt*pm*pn of the chunk's first row (later rows: one row ahead)
      for (int r = 0; r < nrows + 4; ++r) {
        { PROF_T(t_); mbar_wait(&ring_full[gi], gph, a.err, a.backoff); PROF_ADD(p_ring, t_); }
        const int g0 = gi;                                            // row n
        const int g1 = (gi + RING - 1) & (RING - 1), g2 = (gi + RING - 2) & (RING - 1);   // rows n-1, n-2
        if (++gi == RING) { gi = 0; gph ^= 1u; }
        if (r < 2) continue;
        const double* R0 = ring + g2 * rowD;    // row n-2
        const double* R1 = ring + g1 * rowD;    // row n-1
        const double* R2 = ring + g0 * rowD;    // row n
        if (r == 2) {
          // start of a chunk, part 1 (rows ja-2, ja-1, ja): cm1 = curv(ja-1), e0 = FE-gradient at ja  (step3d_t.F:697-724)
#pragma unroll
          for (int m = 0; m < KP; ++m) {
            if (kk[m] != 0) {
#pragma unroll
            for (int c = 0; c < NTR; ++c) {
              const double tm2 = R0[c * trD + o0[m]], tm1 = R1[c * trD + o0[m]], tA = R2[c * trD + o0[m]];
              const double e0 = tA - tm1;
              const double em1 = southw ? e0 : (tm1 - tm2);           // FE(i,Jstr-1)=FE(i,Jstr) on the southern wall
              Cj[m][c] = e0 - em1; FEs[m][c] = e0;
            }
            }
          }
        } else if (r == 3) {
          // part 2 (rows ja-1, ja, ja+1; Hvom(ja) in the slot of the row below the chunk): FE(ja), curv(ja)
          const int sl = q.i;
          mbar_wait(&slot_full[sl], q.ph, a.err, a.backoff);
          const double* sp = slots + sl * slotD;
#pragma unroll
          for (int m = 0; m < KP; ++m) {
            if (kk[m] != 0) {
            const double hv = sp[ohv + so[m]];
            double hvx, hvn; posneg(hv, hvx, hvn);
            const double hvh = hv * 0.5;
#pragma unroll
            for (int c = 0; c < NTR; ++c) {
              const double tm1 = R0[c * trD + o0[m]], tA = R1[c * trD + o0[m]], tB = R2[c * trD + o0[m]];
              const double e1 = tB - tA, c0 = e1 - FEs[m][c];
              FEs[m][c] = hvh * (tm1 + tA) - c16 * (Cj[m][c] * hvx + c0 * hvn);
              Cj[m][c] = c0;
            }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&slot_ready[sl]);
          q.next();
        } else {
          // ---- row j = n-2: q = (t(nnew) - dt*pm*pn*(div_h F + d_k FC)) / Hz  into the slot, in place
          const int j = ja + r - 4, sl = q.i;
          const bool lastN = a.wallN && (j == a.Jend);                // FE(i,Jend+2)=FE(i,Jend+1) (step3d_t.F:718-724)
          const double cff = cffn;
          o2 += a.ni;
          const bool more = (r < nrows + 3);
          double pmn = 0.0, pnn = 0.0;
          if (more) { pmn = __ldg(a.pm + o2); pnn = __ldg(a.pn + o2); }    // metrics of the next row: requested now, used after the batches
          { PROF_T(t_); mbar_wait(&slot_full[sl], q.ph, a.err, a.backoff); PROF_ADD(p_slot, t_); }
          PROF_T(tw_);
          double* sp = slots + sl * slotD;
          double cA[NTR], cM1[NTR], cP1[NTR], cP2[NTR], cFC[NTR], cW = 0.0;    // t(3) at k, k-1, k+1, k+2, FC(k), W(k) of the batch before
#pragma unroll
          for (int m = 0; m < KP; ++m) {
            if (kk[m] != 0) {
            const int k = kk[m];
            const bool valid = (k <= N);
            // ---- load phase (shared memory only)
            const int sm_ = so[m];
            const double hu = sp[ohu + sh[m]], hup = sp[ohu + sh[m] + 1];
            const double hvn_ = sp[ohv + sm_], hz = sp[ohz_ + sm_];
            const double wk = sp[ow + sm_ + LS];                        // W: level index 0..N
            const double wkm = (m == 0) ? sp[ow + sm_] : cW;
            double qm2[NTR], qm1[NTR], A[NTR], qp1[NTR], qp2[NTR], Bv[NTR], T2[NTR], tm2[NTR], tm1[NTR], tp1[NTR], tp2[NTR], twv[NTR];
#pragma unroll
            for (int c = 0; c < NTR; ++c) {
              const double* pr = R0 + c * trD + o0[m];
              qm2[c] = pr[-2]; qm1[c] = pr[-1]; qp1[c] = pr[1]; qp2[c] = pr[2];
              if (m == 0) { A[c] = pr[0]; tm2[c] = pr[-2 * TW]; tm1[c] = pr[-TW]; tp1[c] = pr[TW]; }
              else { A[c] = cP1[c]; tm2[c] = cM1[c]; tm1[c] = cA[c]; tp1[c] = cP2[c]; }
              tp2[c] = pr[2 * TW];
              Bv[c] = R1[c * trD + o0[m]]; T2[c] = R2[c * trD + o0[m]];
              twv[c] = sp[oq + c * dq_ + sm_];
            }
            // ---- compute phase
            double hux, hun, hpx, hpn, hvx, hvm;
            posneg(hu, hux, hun); posneg(hup, hpx, hpn); posneg(hvn_, hvx, hvm);
            const double huh = hu * 0.5, hph = hup * 0.5, hvh = hvn_ * 0.5;
            int badl = 0;
            const double ohz = rcp_ieee(hz, badl);
            if (act && valid) bad |= badl;
#pragma unroll
            for (int c = 0; c < NTR; ++c) {
              const double d0 = qm1[c] - qm2[c], d1 = A[c] - qm1[c], d2 = qp1[c] - A[c], d3 = qp2[c] - qp1[c];
              const double cvm = d1 - d0, cv0 = d2 - d1, cvp = d3 - d2;
              const double FXi = huh * (qm1[c] + A[c]) - c16 * (cvm * hux + cv0 * hun);
              const double FXp = hph * (A[c] + qp1[c]) - c16 * (cv0 * hpx + cvp * hpn);
              const double e1 = Bv[c] - A[c];
              const double e2 = lastN ? e1 : (T2[c] - Bv[c]);
              const double c1 = e2 - e1;
              const double FEn = hvh * (A[c] + Bv[c]) - c16 * (Cj[m][c] * hvx + c1 * hvm);
              const double x1 = cff * (FXp - FXi), x2 = cff * (FEn - FEs[m][c]), x3 = x1 + x2;
              double tv = twv[c] - x3;
              const double FCm = (m == 0) ? vflux8(k - 1, N, tm2[c], tm1[c], A[c], tp1[c], wkm) : cFC[c];
              const double FCk = vflux8(k, N, tm1[c], A[c], tp1[c], tp2[c], wk);
              const double cv = cff * (FCk - FCm);
              tv = tv - cv;
              if (valid) sp[oq + c * dq_ + sm_] = tv * ohz;
              Cj[m][c] = c1; FEs[m][c] = FEn;
              cA[c] = A[c]; cM1[c] = tm1[c]; cP1[c] = tp1[c]; cP2[c] = tp2[c]; cFC[c] = FCk;
            }
            cW = wk;
            if (valid) sp[ohv + sm_] = ohz;
            }
          }
          PROF_ADD(p_work, tw_);
          __syncwarp();
          if (lane == 0) mbar_arrive(&slot_ready[sl]);
          q.next();
          if (more) cffn = dt * pmn * pnn;
        }
        // ---- the oldest row of the ring is no longer needed (the last step of a chunk frees all three)
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&ring_empty[g2]);
          if (r == nrows + 3) { mbar_arrive(&ring_empty[g1]); mbar_arrive(&ring_empty[g0]); }
        }
      }
    }
#ifdef S3T_PROF
    if (a.prof && lane == 0) { atomicAdd((unsigned long long*)&a.prof[4], (unsigned long long)(clock64() - p_t0)); atomicAdd((unsigned long long*)&a.prof[5], (unsigned long long)p_ring); atomicAdd((unsigned long long*)&a.prof[6], (unsigned long long)p_slot); atomicAdd((unsigned long long*)&a.prof[7], (unsigned long long)p_work); }
#endif
  } else {
    // ======================= consumers: spline tridiagonal per (column, tracer) =======================
    const int cw = warp - 1, col = lane & 15, c = (NTR == 2) ? (lane >> 4) : 0;
    const bool dup = (NTR == 1) && lane >= 16;                        // one tracer: the upper half-warp only shadows the lower one
    const double c13 = 1.0 / 3.0;
    const int ni = a.ni, sk = a.sk, Lm = a.Lm;
    const int oq = L.q(c) + col, oa = L.ak(c) + col, oh = L.hz() + col, oo = L.hv() + col, ocf = TM ? 0 : L.cf(c) + col, odc = TM ? 0 : L.dc(c) + col;
    const uint32_t tm0 = tm_base + ((32u * (unsigned)(warp & 3)) << 16);      // this warp's lanes, level 1
    Ring8 q(S);
    int turn = 0;                                                     // consumer warp that owns the next slot
#ifdef S3T_PROF
    long long p_wait = 0, p_fwd = 0, p_bwd = 0; const long long p_t0 = clock64();
#endif
    for (int it = blockIdx.x; it < a.nitems; it += gridDim.x) {
      const int stripe = it % a.nstripes, chunk = it / a.nstripes;
      const int i0s = a.ib + stripe * BI, ja = a.j0 + chunk * a.JCH, jb = min(ja + a.JCH - 1, a.j1);
      const int nrows = jb - ja + 1, i = i0s + col;
      const bool act = (i >= a.i0) && (i <= a.i1) && !dup;
      const bool wE = a.wrapEW && i >= 1 && i <= 2, wW = a.wrapEW && i >= Lm - 2 && i <= Lm;
      const bool edge_stripe = __any_sync(0xffffffffu, wE || wW);
      for (int rr = -1; rr < nrows; ++rr, q.next(), turn = (turn + 1 == NC) ? 0 : turn + 1) {
        if (turn != cw) continue;
        const int sl = q.i;
        { PROF_T(t_); mbar_wait(&slot_ready[sl], q.ph, a.err); PROF_ADD(p_wait, t_); }
        if (rr >= 0) {
          PROF_T(tf_);
          const int j = ja + rr;
          double* sp = slots + sl * slotD;
          const double* pq = sp + oq;                    // q(k)    at pq[(k-1)*LS]
          const double* pa = sp + oa;                    // Akt(k)  at pa[k*LS], k = 0..N
          const double* ph = sp + oh;                    // Hz(k)   at ph[(k-1)*LS]
          const double* po = sp + oo;                    // 1/Hz(k) at po[(k-1)*LS]
          double* pcf = sp + ocf;                        // CF(k)   at pcf[(k-1)*LS]
          double* pdc = sp + odc;                        // DC(k)   at pdc[(k-1)*LS]
          // Software-pipelined forward elimination (k_step3d_t6.cu): while the recurrence of level k runs (mul, add, rcp, mul:
          // ~90 cycles of dependent latency), the coefficients FC,CF,BC,dq of level k+1 are formed and the operands of level
          // k+2 are fetched.  On the last pass "level N+1" reads the first level of the array that follows in the slot (every
          // array here is followed by another one): the values are not used.
          double dtakK, hzN, ohzN, c16N, dtakN, qN, akN;
          double FC, CF, BC, dq;
          {
            const double hz1 = ph[0], ohz1 = po[0], ak1 = pa[LS], q1 = pq[0], ak0 = pa[0];
            hzN = ph[LS]; ohzN = po[LS]; akN = pa[2 * LS]; qN = pq[LS];
            dtakK = dt * ak1; c16N = c16 * hzN; dtakN = dt * akN;
            FC = c16 * hz1 - dt * ak0 * ohz1;                            // 1/6*Hz(k)   - dt*Akt(k-1)*oHz(k)
            CF = c16N - dtakN * ohzN;                                    // 1/6*Hz(k+1) - dt*Akt(k+1)*oHz(k+1)
            BC = c13 * (hz1 + hzN) + dtakK * (ohz1 + ohzN);
            dq = qN - q1;
          }
          pq += 2 * LS; pa += 3 * LS; ph += 2 * LS; po += 2 * LS;        // -> level 3
          double cf_prev = 0.0, dc_prev = 0.0;
          double ak_top = akN, q_top = qN, ohz_top = ohzN;
          double ak_lvl = pa[-2 * LS], q_lvl = pq[-2 * LS], ohz_lvl = po[-2 * LS];   // level k of the coming pass (TM == 2)
          int badl = 0;
          uint32_t tmc = tm0;                                            // tensor-memory address of (CF, DC)(k)
#pragma unroll 4
          for (int k = 1; k <= N - 1; ++k) {
            ak_top = akN; q_top = qN; ohz_top = ohzN;                    // level k+1 (== N on the last pass)
            const double hzL = *ph, ohzL = *po, akL = *pa, qL = *pq;     // level k+2
            ph += LS; po += LS; pa += LS; pq += LS;
            const double cf = rcp_ieee(BC - FC * cf_prev, badl);
            cf_prev = cf * CF;
            dc_prev = cf * (dq - FC * dc_prev);
            if (TM == 2) { tmem_st4d(tmc, cf_prev, dc_prev, q_lvl, ak_lvl); tmem_st1d(tmc + 8, dt * ohz_lvl); tmc += TMC; ak_lvl = akN; q_lvl = qN; ohz_lvl = ohzN; }
            else if (TM == 1) { tmem_st2d(tmc, cf_prev, dc_prev); tmc += TMC; }
            else { *pcf = cf_prev; *pdc = dc_prev; pcf += LS; pdc += LS; }
            const double c16L = c16 * hzL, dtakL = dt * akL;
            FC = c16N - dtakK * ohzN;
            CF = c16L - dtakL * ohzL;
            BC = c13 * (hzN + hzL) + dtakN * (ohzN + ohzL);
            dq = qL - qN;
            dtakK = dtakN; hzN = hzL; ohzN = ohzL; c16N = c16L; dtakN = dtakL; qN = qL; akN = akL;
          }
          if (act) bad |= badl;
          PROF_ADD(p_fwd, tf_);
          PROF_T(tb_);
          // back substitution + final update, level N first.  pcf/pdc point one past level N-1; pq/po/pa at level N+2.
          const bool south = a.wallS && j == a.Jstr, north = a.wallN && j == a.Jend;
          double* tw = a.out[c] + ((i - a.LBi) + (size_t)ni * (j - a.LBj)) + (size_t)sk * (N - 1);   // t(nnew)(i,j,N)
          double dc_next = 0.0;                                          // DC(N)
          double a_next = dc_next * ak_top;                              // DC(N)*Akt(N)
          double q_next = q_top, dtohz_next = dt * ohz_top;
          pcf -= LS; pdc -= LS; pq -= 3 * LS; po -= 3 * LS; pa -= 3 * LS;    // level N-1
          double Xk, Yk, akk, qk, dtk;                                   // CF, DC, Akt, q, dt/Hz of level k
          if (TM == 2) {
            tmem_wait_st();
            fence_proxy_async();                                         // the slot goes back to the loader before the backward sweep
            __syncwarp();
            if (lane == 0) mbar_arrive(&slot_empty[sl]);
            tmc -= TMC; tmem_ld4d(tmc, Xk, Yk, qk, akk); tmem_ld1d(tmc + 8, dtk); tmem_wait_ld();
          } else {
            if (TM == 1) { tmem_wait_st(); tmc -= TMC; tmem_ld2d(tmc, Xk, Yk); tmem_wait_ld(); }
            else { Xk = *pcf; Yk = *pdc; }
            akk = *pa; qk = *pq; dtk = dt * *po;
          }
          // operands of level k-1 are fetched one level ahead; for k = 1 that is "level 0": tensor memory re-reads level 1, the
          // shared-memory path reads the array space just below (neither is used)
          auto level_below = [&](int k, double& Xm, double& Ym, double& akm, double& qm, double& dtm) {
            if (TM == 2) { if (k > 1) tmc -= TMC; tmem_ld4d(tmc, Xm, Ym, qm, akm); tmem_ld1d(tmc + 8, dtm); }
            else {
              pq -= LS; po -= LS; pa -= LS;
              if (TM == 1) { if (k > 1) tmc -= TMC; tmem_ld2d(tmc, Xm, Ym); }
              else { pcf -= LS; pdc -= LS; Xm = *pcf; Ym = *pdc; }
              akm = *pa; qm = *pq; dtm = dt * *po;
            }
          };
          if (!(edge_stripe || south || north)) {
            // interior stripe and row: one store per level
#pragma unroll 4
            for (int k = N - 1; k >= 1; --k) {
              double Xm, Ym, akm, qm, dtm;
              level_below(k, Xm, Ym, akm, qm, dtm);
              const double dc_k = Yk - Xk * dc_next;
              const double a_k = dc_k * akk;
              if (act) *tw = q_next + dtohz_next * (a_next - a_k);       // level k+1
              tw -= sk;
              dc_next = dc_k; a_next = a_k; q_next = qk; dtohz_next = dtk;
              if (TM) tmem_wait_ld();
              Xk = Xm; Yk = Ym; akk = akm; qk = qm; dtk = dtm;
            }
            if (act) *tw = q_next + dtohz_next * (a_next - 0.0);         // level 1; DC(0)=0 is not scaled by Akt
          } else {
            auto put = [&](double out) {                                 // st() + t3dbc wall rows (t3dbc_im.F:334-341,415-422)
              if (act) {
                tw[0] = out;
                if (wE) tw[Lm] = out;
                if (wW) tw[-Lm] = out;
                if (south) { tw[-ni] = out; if (wE) tw[Lm - ni] = out; if (wW) tw[-Lm - ni] = out; }
                if (north) { tw[ni] = out; if (wE) tw[Lm + ni] = out; if (wW) tw[-Lm + ni] = out; }
              }
              tw -= sk;
            };
#pragma unroll 1
            for (int k = N - 1; k >= 1; --k) {
              double Xm, Ym, akm, qm, dtm;
              level_below(k, Xm, Ym, akm, qm, dtm);
              const double dc_k = Yk - Xk * dc_next;
              const double a_k = dc_k * akk;
              put(q_next + dtohz_next * (a_next - a_k));
              dc_next = dc_k; a_next = a_k; q_next = qk; dtohz_next = dtk;
              if (TM) tmem_wait_ld();
              Xk = Xm; Yk = Ym; akk = akm; qk = qm; dtk = dtm;
            }
            put(q_next + dtohz_next * (a_next - 0.0));
          }
          PROF_ADD(p_bwd, tb_);
          if (TM == 2) continue;                                         // slot already released
        }
        fence_proxy_async();                     // generic-proxy accesses of the slot are ordered before the TMA refill
        __syncwarp();
        if (lane == 0) mbar_arrive(&slot_empty[sl]);
      }
    }
#ifdef S3T_PROF
    if (a.prof && lane == 0) { atomicAdd((unsigned long long*)&a.prof[8], (unsigned long long)(clock64() - p_t0)); atomicAdd((unsigned long long*)&a.prof[9], (unsigned long long)p_wait); atomicAdd((unsigned long long*)&a.prof[10], (unsigned long long)p_fwd); atomicAdd((unsigned long long*)&a.prof[11], (unsigned long long)p_bwd); }
#endif
  }
  if (bad) atomicOr(a.err, 1);
  if (TM) {
    tmem_fence_before();
    __syncthreads();                              // every consumer warp is done with its tensor-memory columns
    if (warp == 1) tmem_dealloc(tm_base, tm_cols);
  }
}

namespace {
#ifndef ROMS_B200_EMU
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr; static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr; cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && qr == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}
#endif
// tensor map of a (ni, nj, nk) fp64 volume, box = (bw, 1, bk)
int make_map(TMap* m, const double* base, int ni, int nj, int nk, int bw, int bk) {
#ifdef ROMS_B200_EMU
  m->base = base; m->dim[0] = ni; m->dim[1] = nj; m->dim[2] = nk; m->stride[0] = 1; m->stride[1] = ni; m->stride[2] = (long)ni * nj;
  m->box[0] = bw; m->box[1] = 1; m->box[2] = bk;
  return 0;
#else
  EncodeTiledFn fn = encode_fn();
  if (!fn) return 2;
  const cuuint64_t dim[3] = {(cuuint64_t)ni, (cuuint64_t)nj, (cuuint64_t)nk};
  const cuuint64_t str[2] = {(cuuint64_t)ni * 8, (cuuint64_t)ni * nj * 8};
  const cuuint32_t box[3] = {(cuuint32_t)bw, 1, (cuuint32_t)bk}, es[3] = {1, 1, 1};
  const CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)base, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "roms_b200: cuTensorMapEncodeTiled failed (%d) for a %dx%dx%d volume, box %dx1x%d\n", (int)r, ni, nj, nk, bw, bk); return 2; }
  return 0;
#endif
}

template <int NTR, int KP, int TM, int MAXT>
int launch_v8_t(roms_b200_ctx* c, const A8& a, int grid, size_t smem) {
  static AttrOnce set;
  if (set.need(smem)) CUDA_OK(cudaFuncSetAttribute(step3d_t_v8_kernel<NTR, KP, TM, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  step3d_t_v8_kernel<NTR, KP, TM, MAXT><<<dim3(grid), dim3(32 * (1 + a.NC + a.NP)), smem, c->stream>>>(a);
  return 0;
}
template <int NTR, int KP, int TM>
int launch_v8(roms_b200_ctx* c, const A8& a, int grid, size_t smem) {
  const int nt = 32 * (1 + a.NC + a.NP);
  if (nt <= 384) return launch_v8_t<NTR, KP, TM, 384>(c, a, grid, smem);
  if (nt <= 512) return launch_v8_t<NTR, KP, TM, 512>(c, a, grid, smem);
  return launch_v8_t<NTR, KP, TM, 640>(c, a, grid, smem);
}
template <int NTR, int TM>
int launch_v8_kp(roms_b200_ctx* c, const A8& a, int grid, size_t smem, int KP) {
  if (KP == 1) return launch_v8<NTR, 1, TM>(c, a, grid, smem);
  if (KP == 2) return launch_v8<NTR, 2, TM>(c, a, grid, smem);
  return launch_v8<NTR, 4, TM>(c, a, grid, smem);
}
size_t smem_v8(int ntr, int tm, int N, int S) {
  if (ntr == 2) return tm ? Lay<2, 1>{N}.bytes(S) : Lay<2, 0>{N}.bytes(S);
  return tm ? Lay<1, 1>{N}.bytes(S) : Lay<1, 0>{N}.bytes(S);
}
}  // namespace

// returns 0 on success, 2 if this layout does not apply (caller falls back to k_step3d_t6.cu / k_step3d_t4.cu)
int k_step3d_t_v8(roms_b200_ctx* c, int nnew) {
  const Dev& D = c->D; const roms_b200_bounds& b = D.b;
  const int N = b.N;
  if (N < 4 || N > 254) return 2;
  if (!b.EWperiodic && (b.Western_Edge || b.Eastern_Edge)) return 2;     // closed W/E walls: FX edge copies not implemented here
  if (D.nij * (size_t)(N + 1) >= (size_t)1 << 31) return 2;              // 32-bit element offsets inside one volume
  if (D.ni & 1) return 2;                                                // TMA: global strides must be multiples of 16 bytes
  static int max_smem = -1, nsm = 148;
  if (max_smem < 0) {
    int dev = 0; cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) max_smem = 48 * 1024;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  }
  static const int force_s = getenv("ROMS_B200_S3T_SLOTS") ? atoi(getenv("ROMS_B200_S3T_SLOTS")) : 0;
  static const int force_np = getenv("ROMS_B200_S3T_NP") ? atoi(getenv("ROMS_B200_S3T_NP")) : 0;
  static const int force_jch = getenv("ROMS_B200_S3T_JCH") ? atoi(getenv("ROMS_B200_S3T_JCH")) : 0;
  for (int itr0 = 1; itr0 <= b.NT; itr0 += 2) {
    const int ntr = (itr0 + 1 <= b.NT) ? 2 : 1;
    // CF/DC of the tridiagonal solve in tensor memory (4 32-bit columns per level, <= 512 columns) unless switched off
    // (mode 2, 10 columns per level: also q, Akt, dt/Hz, so that the slot is released after the forward sweep)
    static const int tm_env = getenv("ROMS_B200_S3T_TMEM") ? atoi(getenv("ROMS_B200_S3T_TMEM")) : 2;
    const int tm = (tm_env >= 2 && 10 * N <= 512) ? 2 : ((tm_env >= 1 && 4 * N <= 512) ? 1 : 0);
    // slots: as many as fit (at most 6); consumers: every slot that is neither being loaded nor being produced, at most 4
    // (with tensor memory each consumer warp needs its own lane quadrant: warps 1..4)
    int S = 0;
    for (int s = (force_s ? force_s : 6); s >= 2 && !S; --s) if (smem_v8(ntr, tm, N, s) <= (size_t)max_smem) S = s;
    if (!S || (S < 3 && !force_s)) return 2;       // fewer than 3 slots (N > ~51) cannot overlap load, advection and solve: k_step3d_t6.cu is faster there
    if (S == 3 && tm != 2 && !force_s) return 2;   // 3 slots need the early slot release of tensor-memory mode 2
    static const int force_nc = getenv("ROMS_B200_S3T_NC") ? atoi(getenv("ROMS_B200_S3T_NC")) : 0;
    // measured: 3 consumer warps are enough for 8 producer warps; with 3 slots (N = 50: 1024x512x50 in 0.88 ms against 1.26 ms
    // for k_step3d_t6.cu) two consumers work because a slot goes back to the loader after the forward sweep (one consumer: 1.50 ms)
    const int NC = force_nc ? force_nc : (S >= 5 ? 3 : (S >= 3 ? 2 : 1));
    if (NC > 4 || NC >= S) return 2;
    // producer warps x level-pair batches per warp: NP*KP >= ceil(N/2), at most 16 warps per CTA
    const int nb = (N + 1) / 2;
    int KP = 0, NP = 0;
    for (int kp = 1; kp <= 4 && !KP; kp *= 2) {
      const int np = force_np ? force_np : (nb + kp - 1) / kp;
      if (np * kp >= nb && 1 + NC + np <= (force_np ? 20 : 16)) { KP = kp; NP = np; }
    }
    if (!KP) return 2;
    // TMA: the first element of a box must be 16-byte aligned in global memory (measured on B200, tools/ubench/tma3d_test.cu: an odd
    // fp64 start column raises "illegal instruction"), so stripes start at an even array column; a column left of Istr is masked
    const int ib = b.Istr - ((b.Istr - b.LBi) & 1);
    const int rows = b.Jend - b.Jstr + 1, nstripes = (b.Iend - ib + BI) / BI;
    // j-chunks: persistent CTAs take items round-robin; cost ~ rounds * (rows per chunk + ~2 rows of pipeline start-up)
    int best_jch = rows; long best = -1;
    for (int nc = 1; nc <= rows; ++nc) {
      const int jch = (rows + nc - 1) / nc, ncr = (rows + jch - 1) / jch;
      if (jch < 4 && nc > 1) break;
      const long items = (long)nstripes * ncr, rounds = (items + nsm - 1) / nsm;
      const long cost = rounds * (jch + 2);
      if (best < 0 || cost < best) { best = cost; best_jch = jch; }
    }
    if (force_jch) best_jch = force_jch;
    A8 a;
    memset(&a, 0, sizeof(a));
    a.N = N; a.ni = D.ni; a.sk = (int)D.nij; a.LBi = b.LBi; a.LBj = b.LBj; a.Lm = b.Lm;
    a.i0 = b.Istr; a.i1 = b.Iend; a.ib = ib; a.j0 = b.Jstr; a.j1 = b.Jend; a.Jstr = b.Jstr; a.Jend = b.Jend;
    a.wallS = b.Southern_Edge && !b.NSperiodic; a.wallN = b.Northern_Edge && !b.NSperiodic; a.wrapEW = D.wrapEW;
    a.nstripes = nstripes; a.JCH = best_jch; a.nitems = nstripes * ((rows + best_jch - 1) / best_jch);
    a.S = S; a.NC = NC; a.NP = NP; a.dt = D.p.dt; a.err = D.err;
    static const int backoff = getenv("ROMS_B200_S3T_BACKOFF") ? atoi(getenv("ROMS_B200_S3T_BACKOFF")) : 0;
    a.backoff = backoff;
    static const int pfd = getenv("ROMS_B200_S3T_PFD") ? atoi(getenv("ROMS_B200_S3T_PFD")) : 0;
    a.pfd = pfd;
    a.pm = D.f[FID(pm)]; a.pn = D.f[FID(pn)];
    const size_t vol = D.nij * (size_t)N;
    int rc = 0;
    for (int qt = 0; qt < ntr; ++qt) {
      const int itrc = itr0 + qt;
      const double* t3 = D.f[FID(t)] + vol * ((3 - 1) + (size_t)3 * (itrc - 1));
      double* tw = D.f[FID(t)] + vol * ((nnew - 1) + (size_t)3 * (itrc - 1));
      const double* ak = D.f[FID(Akt)] + D.nij * (size_t)(N + 1) * (size_t)((itrc <= b.NAT ? itrc : b.NAT) - 1);
      a.out[qt] = tw;
      rc |= make_map(&a.t3[qt], t3, D.ni, D.nj, N, TW, N);
      rc |= make_map(&a.tw[qt], tw, D.ni, D.nj, N, BI, N);
      rc |= make_map(&a.ak[qt], ak, D.ni, D.nj, N + 1, BI, N + 1);
    }
    rc |= make_map(&a.hz, D.f[FID(Hz)], D.ni, D.nj, N, BI, N);
    rc |= make_map(&a.hv, D.f[FID(Hvom)], D.ni, D.nj, N, BI, N);
    rc |= make_map(&a.hu, D.f[FID(Huon)], D.ni, D.nj, N, HUW, N);
    rc |= make_map(&a.w, D.f[FID(W)], D.ni, D.nj, N + 1, BI, N + 1);
    if (rc) return 2;
    const int grid = a.nitems < nsm ? a.nitems : nsm;
#ifdef S3T_PROF
    static long long* prof = nullptr;
    if (!prof) cudaMalloc((void**)&prof, 16 * sizeof(long long));
    cudaMemsetAsync(prof, 0, 16 * sizeof(long long), c->stream);
    a.prof = prof;
#endif
    static const bool verbose = (getenv("ROMS_B200_S3T_VERBOSE") != nullptr);
    if (verbose) fprintf(stderr, "step3d_t v8: N=%d ntr=%d stripes=%d JCH=%d items=%d grid=%d slots=%d consumers=%d producers=%dx%d smem=%zu tmem=%d\n", N, ntr, nstripes,
                         a.JCH, a.nitems, grid, S, NC, NP, KP, smem_v8(ntr, tm, N, S), tm);
    const size_t smem = smem_v8(ntr, tm, N, S);
    if (ntr == 2) rc = tm == 2 ? launch_v8_kp<2, 2>(c, a, grid, smem, KP) : (tm == 1 ? launch_v8_kp<2, 1>(c, a, grid, smem, KP) : launch_v8_kp<2, 0>(c, a, grid, smem, KP));
    else rc = tm == 2 ? launch_v8_kp<1, 2>(c, a, grid, smem, KP) : (tm == 1 ? launch_v8_kp<1, 1>(c, a, grid, smem, KP) : launch_v8_kp<1, 0>(c, a, grid, smem, KP));
    if (rc) return rc;
    c->launches++;
#ifdef S3T_PROF
    {
      long long h[16]; cudaStreamSynchronize(c->stream); cudaMemcpy(h, prof, sizeof(h), cudaMemcpyDeviceToHost);
      const double L = (double)grid, P = (double)grid * NP, C = (double)grid * NC;
      fprintf(stderr, "S3T_PROF kcycles per CTA/warp: loader total %.0f wait_ring %.0f wait_slot %.0f | producer total %.0f wait_ring %.0f wait_slot %.0f work %.0f | "
                      "consumer total %.0f wait %.0f fwd %.0f bwd %.0f\n", h[0] / L / 1e3, h[1] / L / 1e3, h[2] / L / 1e3, h[4] / P / 1e3, h[5] / P / 1e3, h[6] / P / 1e3,
              h[7] / P / 1e3, h[8] / C / 1e3, h[9] / C / 1e3, h[10] / C / 1e3, h[11] / C / 1e3);
    }
#endif
  }
  return 0;
}
