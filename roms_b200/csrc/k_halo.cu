// roms_b200/csrc/k_halo.cu -- halo exchange between tiles: the replacement of
// mp_exchange2d/3d/4d (Utility/mp_exchange.F:290,2033,3373) by NCCL send/recv.
//
// Semantics kept from the reference: two phases, West/East first and then
// South/North over the FULL i-range including the freshly received W/E ghost
// columns, so corner points propagate without diagonal messages
// (mp_exchange.F:520-532,761-773); several fields are aggregated into one
// message per neighbour per phase (the reference aggregates up to 4); the E-W
// axis is periodic (tile_neighbors, mp_exchange.F:118-197).  Width = the
// mirror's halo width (>= NghostPoints).  All of it is enqueued on the
// context's stream (pack kernel -> ncclGroup{Send,Recv} -> unpack kernel), so it
// is captured in the fast-loop CUDA graph together with the step2d launches.
// NCCL is dlopen'ed: single-GPU use never needs it.
//
// DEFAULT transport (when the host connected the peers, roms_b200_p2p_*): NVLink peer MAILBOXES.  Every rank exports one
// device allocation through CUDA IPC; a neighbour maps it and ONE exchange kernel stores the boundary strips of all eight
// neighbours straight into the receivers' mailboxes over NVLink as flag-in-data messages (see below) and unpacks what
// arrives: no staging, no NCCL proxy, no fences.  The NCCL path above remains as ROMS_B200_HALO_NCCL=1 and for diag's all-reduce.
#include "common.cuh"
#include <algorithm>
#include <dlfcn.h>
#include <cstring>
#include <cstdlib>

namespace {
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void* ncclComm_p;
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId_t*) = nullptr;
  int (*CommInitRank)(ncclComm_p*, int, ncclUniqueId_t, int) = nullptr;
  int (*CommDestroy)(ncclComm_p) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_p, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
} g_nccl;
const int kNcclFloat64 = 8, kNcclSum = 0, kNcclMax = 2;
int load_nccl() {
  if (g_nccl.h) return 0;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) { g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.h) break; }
  if (!g_nccl.h) { fprintf(stderr, "roms_b200: cannot dlopen libnccl.so.2: %s\n", dlerror()); return 1; }
#define SYM(f, s) *(void**)(&g_nccl.f) = dlsym(g_nccl.h, s); if (!g_nccl.f) { fprintf(stderr, "roms_b200: NCCL symbol %s missing\n", s); return 1; }
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
  SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
  SYM(AllReduce, "ncclAllReduce") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  return 0;
}
#define NCCL_OK(x) do { int r_ = (x); if (r_ != 0) { fprintf(stderr, "roms_b200: NCCL error %s at %s:%d\n", g_nccl.GetErrorString(r_), __FILE__, __LINE__); return 2; } } while (0)

// one list entry = a contiguous stack of (i,j) planes of one field
struct HaloList { double* base[HALO_MAXF]; int nplanes[HALO_MAXF]; int nf; int total_planes; };

// phase 0: W/E strips (w columns, all rows) ; phase 1: S/N strips (w rows, all columns)
__global__ void halo_pack_kernel(const Dev D, HaloList L, int phase, int w, double* __restrict__ bufLo, double* __restrict__ bufHi) {
  const roms_b200_bounds& b = D.b;
  const int ni = D.ni, nj = D.nj;
  const int plane = blockIdx.y;           // global plane index over all fields
  int f = 0, p = plane;
  while (p >= L.nplanes[f]) { p -= L.nplanes[f]; ++f; }
  const double* src = L.base[f] + (size_t)p * D.nij;
  const int n = (phase == 0) ? w * nj : w * ni;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) {
    if (phase == 0) {
      const int c = x % w, jj = x / w;                       // column c of the strip, array row jj
      bufLo[(size_t)plane * n + x] = src[(b.Istr + c - b.LBi) + (size_t)ni * jj];            // -> west neighbour's east ghosts
      bufHi[(size_t)plane * n + x] = src[(b.Iend - w + 1 + c - b.LBi) + (size_t)ni * jj];    // -> east neighbour's west ghosts
    } else {
      const int ii = x % ni, r = x / ni;
      bufLo[(size_t)plane * n + x] = src[ii + (size_t)ni * (b.Jstr + r - b.LBj)];            // -> south neighbour
      bufHi[(size_t)plane * n + x] = src[ii + (size_t)ni * (b.Jend - w + 1 + r - b.LBj)];    // -> north neighbour
    }
  }
}
__global__ void halo_unpack_kernel(const Dev D, HaloList L, int phase, int w, const double* __restrict__ bufLo, const double* __restrict__ bufHi,
                                   int haveLo, int haveHi) {
  const roms_b200_bounds& b = D.b;
  const int ni = D.ni, nj = D.nj;
  const int plane = blockIdx.y;
  int f = 0, p = plane;
  while (p >= L.nplanes[f]) { p -= L.nplanes[f]; ++f; }
  double* dst = L.base[f] + (size_t)p * D.nij;
  const int n = (phase == 0) ? w * nj : w * ni;
  for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) {
    if (phase == 0) {
      const int c = x % w, jj = x / w;
      if (haveLo) dst[(b.Istr - w + c - b.LBi) + (size_t)ni * jj] = bufLo[(size_t)plane * n + x];   // from the west neighbour's east columns
      if (haveHi) dst[(b.Iend + 1 + c - b.LBi) + (size_t)ni * jj] = bufHi[(size_t)plane * n + x];
    } else {
      const int ii = x % ni, r = x / ni;
      if (haveLo) dst[ii + (size_t)ni * (b.Jstr - w + r - b.LBj)] = bufLo[(size_t)plane * n + x];
      if (haveHi) dst[ii + (size_t)ni * (b.Jend + 1 + r - b.LBj)] = bufHi[(size_t)plane * n + x];
    }
  }
}

// ---- NVLink peer mailboxes ----------------------------------------------------------------------------------
// exported allocation: [256-byte header][16 data regions (dir*2+slot) of `cap` elements of 16 bytes]
// dir = the side the strip ARRIVES from: 0 W, 1 E, 2 S, 3 N, 4 SW, 5 SE, 6 NW, 7 NE.
// Unlike the two-phase scheme of mp_exchange (W/E, then S/N over the full i-range so that corners propagate), all eight
// neighbours are served in ONE exchange: the corner blocks travel as their own (w x w) messages.  W/E strips span the rows
// [Jstr..Jend], extended to the array edge where the tile has no S/N neighbour (physical boundary rows); S/N strips
// likewise in i -- so after the exchange the halo frame holds exactly what the two phases would have produced.
//
// Flag-in-data messages (the idea of NCCL's LL protocol): a double travels as two 8-byte words {low half | seq << 32},
// {high half | seq << 32}, written with one 16-byte store straight into the receiver's mailbox over NVLink.  An 8-byte word
// arrives atomically, so the receiver simply polls every element until both halves carry this exchange's sequence number:
// no system fence, no flag round trip, no block that has to see all others finish, and no requirement that the blocks of the
// kernel are co-resident (round 1: data + __threadfence_system + ticket + flags + spin).  Mailboxes are double-buffered by
// sequence parity: a sender reaches exchange s+2 only after it has received its neighbour's data of s+1, which the neighbour
// sent after it had unpacked s (stream order).  A poll that lasts ~2 s (a dead peer) raises bit 1 of the device error word
// instead of hanging the node (the reference sets exit_flag=2, mp_exchange.F:544-553): roms_b200_sync reports it.
constexpr size_t P2P_HDR = 256;
constexpr int P2P_NDIR = 8;
constexpr int P2P_CH = 256;               // elements of a strip per work item (one per thread)
__host__ __device__ inline int p2p_opp(int d) { return d < 4 ? (d ^ 1) : (11 - d); }     // W<->E, S<->N, SW<->NE, SE<->NW
struct P2PView { ulonglong2* data; size_t cap; };
__host__ __device__ inline P2PView p2p_view(void* mem, size_t cap) { return P2PView{(ulonglong2*)((char*)mem + P2P_HDR), cap}; }
struct Rect { int o, w, h; };              // element offset of the first point in an (i,j) plane, width (i), height (j)
struct P2PArgs {
  void* mine; void* peer[P2P_NDIR];        // my mailbox ; the mailboxes of the 8 neighbours (null: none)
  Rect snd[P2P_NDIR], rcv[P2P_NDIR];       // what I send towards direction d / where the strip arriving from direction d goes
  size_t cap; unsigned long long* seq; unsigned int* ticket;
  int nsub, ch;                            // chunks per (plane, direction) strip, elements per chunk
};
__device__ __forceinline__ void ll_store(ulonglong2* p, double v, unsigned s) {
  const unsigned long long f = (unsigned long long)s << 32;
  const unsigned long long w0 = f | (unsigned)__double2loint(v), w1 = f | (unsigned)__double2hiint(v);
  asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ bool ll_load(const ulonglong2* p, unsigned s, double& v) {
  unsigned long long w0, w1;
  asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
  v = __hiloint2double((int)(unsigned)w1, (int)(unsigned)w0);
  return (unsigned)(w0 >> 32) == s && (unsigned)(w1 >> 32) == s;
}
// ONE kernel per exchange; work items = (plane, direction).  (1) every block pushes the strips of its items into the
// neighbours' mailboxes, (2) then polls and unpacks the strips that arrive for its items, (3) the last block to finish
// commits the sequence number (device-side, so the kernel is replayable inside the fast-loop CUDA graph).
__global__ void __launch_bounds__(256) halo_xchg_p2p_kernel(const Dev D, HaloList L, const __grid_constant__ P2PArgs a) {
  // programmatic dependent launch (see the launch site): no-ops for an ordinary launch
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int ni = D.ni;
  const unsigned long long s64 = a.seq[0] + 1;        // this exchange's sequence number
  const unsigned s = (unsigned)s64;
  const int slot = (int)(s & 1u);
  const P2PView me = p2p_view(a.mine, a.cap);
  // work items: (plane, direction, chunk of P2P_CH elements of the strip): a 512 x 6 strip handled by one block is a dozen
  // sequential remote stores and local polls per thread; chunks give every thread one or two
  const int nsub = a.nsub, nitems = L.total_planes * P2P_NDIR * nsub;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int sub = item % nsub, d = (item / nsub) % P2P_NDIR, plane = item / (nsub * P2P_NDIR);
    void* peer = a.peer[d];
    if (!peer) continue;
    const Rect r = a.snd[d];
    const int n = r.w * r.h, x0 = sub * a.ch, x1 = min(x0 + a.ch, n);
    if (x0 >= n) continue;
    int f = 0, p = plane;
    while (p >= L.nplanes[f]) { p -= L.nplanes[f]; ++f; }
    const double* fld = L.base[f] + (size_t)p * D.nij;
    ulonglong2* out = p2p_view(peer, a.cap).data + (size_t)(p2p_opp(d) * 2 + slot) * a.cap + (size_t)plane * n;
    for (int x = x0 + threadIdx.x; x < x1; x += blockDim.x) ll_store(out + x, fld[r.o + (x % r.w) + (size_t)ni * (x / r.w)], s);
  }
  bool dead = false;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int sub = item % nsub, d = (item / nsub) % P2P_NDIR, plane = item / (nsub * P2P_NDIR);
    if (!a.peer[d]) continue;
    const Rect r = a.rcv[d];
    const int n = r.w * r.h, x0 = sub * a.ch, x1 = min(x0 + a.ch, n);
    if (x0 >= n) continue;
    int f = 0, p = plane;
    while (p >= L.nplanes[f]) { p -= L.nplanes[f]; ++f; }
    double* fld = L.base[f] + (size_t)p * D.nij;
    const ulonglong2* in = me.data + (size_t)(d * 2 + slot) * a.cap + (size_t)plane * n;
    for (int x = x0 + threadIdx.x; x < x1; x += blockDim.x) {
      double v; unsigned spins = 0; long long t0 = 0;
      while (!ll_load(in + x, s, v)) {
        if (dead) break;
        if ((++spins & 0xfffu) == 0) {                                        // every 4096 polls: has this taken ~2 s?
          const long long t = clock64();
          if (!t0) t0 = t; else if (t - t0 > 4000000000ll) { dead = true; break; }
        }
      }
      fld[r.o + (x % r.w) + (size_t)ni * (x / r.w)] = v;
    }
  }
  if (dead) atomicOr(D.err, 2);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(a.ticket, 1u) == gridDim.x - 1) { a.ticket[0] = 0; __threadfence(); a.seq[0] = s64; }
  }
}
}  // namespace

extern "C" {

// ---- NVLink peer mailboxes: set-up.  Every rank calls p2p_handle, the host all-gathers the 64-byte handles
// (MPI_Allgather / torch.distributed.all_gather), every rank calls p2p_connect with the table (nranks x 64 bytes).
int roms_b200_p2p_handle(roms_b200_ctx* c, char* handle64) {
  if (!c) return 1;
  CUDA_OK(cudaSetDevice(c->device));
  const size_t strip = (size_t)c->D.halo * (size_t)((c->D.ni > c->D.nj) ? c->D.ni : c->D.nj);
  c->halo_cap = strip * HALO_MAXPLANES;
  const size_t bytes = P2P_HDR + 2 * P2P_NDIR * c->halo_cap * sizeof(ulonglong2);
  if (!c->p2p_mem) {
    CUDA_OK(cudaMalloc(&c->p2p_mem, bytes));
    CUDA_OK(cudaMemset(c->p2p_mem, 0, bytes));
    CUDA_OK(cudaMalloc((void**)&c->p2p_seq, 2 * sizeof(unsigned long long)));
    CUDA_OK(cudaMemset(c->p2p_seq, 0, 2 * sizeof(unsigned long long)));
    CUDA_OK(cudaMalloc((void**)&c->p2p_ticket, 2 * sizeof(unsigned int)));
    CUDA_OK(cudaMemset(c->p2p_ticket, 0, 2 * sizeof(unsigned int)));
  }
  cudaIpcMemHandle_t h;
  CUDA_OK(cudaIpcGetMemHandle(&h, c->p2p_mem));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return 0;
}
int roms_b200_p2p_connect(roms_b200_ctx* c, const char* handles, int nranks) {
  if (!c || !c->p2p_mem) return 1;
  CUDA_OK(cudaSetDevice(c->device));
  const roms_b200_bounds& b = c->D.b;
  if (nranks != b.NtileI * b.NtileJ) return 1;
  int nb[4]; roms_b200_tile_neighbors(&b, nb);
  const int me = b.Jtile * b.NtileI + b.Itile;
  int r8[P2P_NDIR], sr[4 * P2P_NDIR], rr[4 * P2P_NDIR];
  if (roms_b200_halo_plan(&b, c->D.halo, r8, sr, rr)) return 1;
  for (int q = 0; q < P2P_NDIR; ++q) {
    c->p2p_peer[q] = nullptr; c->p2p_rank[q] = r8[q];
    if (r8[q] < 0 || r8[q] == me) { c->p2p_rank[q] = -1; continue; }
    for (int r = 0; r < q; ++r) if (r8[r] == r8[q] && c->p2p_peer[r]) c->p2p_peer[q] = c->p2p_peer[r];   // same peer on several sides: map once
    if (c->p2p_peer[q]) continue;
    cudaIpcMemHandle_t h; memcpy(&h, handles + (size_t)64 * r8[q], 64);
    CUDA_OK(cudaIpcOpenMemHandle(&c->p2p_peer[q], h, cudaIpcMemLazyEnablePeerAccess));
  }
  c->rank = me; c->nranks = nranks; c->nbW = nb[0]; c->nbE = nb[1]; c->nbS = nb[2]; c->nbN = nb[3];
  c->p2p_on = (getenv("ROMS_B200_HALO_NCCL") == nullptr);
  return 0;
}

int roms_b200_comm_unique_id(char* id128) {
  if (load_nccl()) return 1;
  ncclUniqueId_t id; NCCL_OK(g_nccl.GetUniqueId(&id));
  memcpy(id128, id.internal, 128);
  return 0;
}

// rank = tile id (one rank = one tile = one GPU, Drivers/nl_roms.h:145-157)
int roms_b200_comm_init(roms_b200_ctx* c, int rank, int nranks, const char* id128) {
  if (!c) return 1;
  CUDA_OK(cudaSetDevice(c->device));
  const roms_b200_bounds& b = c->D.b;
  if (nranks != b.NtileI * b.NtileJ || rank != b.Jtile * b.NtileI + b.Itile) { fprintf(stderr, "roms_b200: rank/tile mismatch\n"); return 1; }
  if (load_nccl()) return 1;
  ncclUniqueId_t id; memcpy(id.internal, id128, 128);
  ncclComm_p comm = nullptr;
  NCCL_OK(g_nccl.CommInitRank(&comm, nranks, id, rank));
  c->comm = comm; c->rank = rank; c->nranks = nranks;
  c->deep = (c->D.halo >= 6 && getenv("ROMS_B200_NO_DEEP_HALO") == nullptr) ? 1 : 0;
  { int nb[4]; roms_b200_tile_neighbors(&b, nb); c->nbW = nb[0]; c->nbE = nb[1]; c->nbS = nb[2]; c->nbN = nb[3]; }
  // buffers: up to HALO_MAXPLANES planes of the larger strip
  const size_t strip = (size_t)c->D.halo * (size_t)((c->D.ni > c->D.nj) ? c->D.ni : c->D.nj);
  c->halo_cap = strip * HALO_MAXPLANES;
  for (int q = 0; q < 4; ++q) CUDA_OK(cudaMalloc((void**)&c->hbuf[q], c->halo_cap * sizeof(double)));
  return 0;
}
int roms_b200_comm_destroy(roms_b200_ctx* c) {
  if (c && c->p2p_mem) {
    cudaDeviceSynchronize();
    for (int q = 0; q < P2P_NDIR; ++q) {
      bool dup = false;
      for (int r = 0; r < q; ++r) if (c->p2p_peer[r] == c->p2p_peer[q]) dup = true;
      if (c->p2p_peer[q] && !dup) cudaIpcCloseMemHandle(c->p2p_peer[q]);
    }
    for (int q = 0; q < P2P_NDIR; ++q) c->p2p_peer[q] = nullptr;
    cudaFree(c->p2p_mem); cudaFree(c->p2p_seq); cudaFree(c->p2p_ticket); c->p2p_mem = nullptr; c->p2p_on = 0;
  }
  if (c && c->comm) { g_nccl.CommDestroy((ncclComm_p)c->comm); c->comm = nullptr; for (int q = 0; q < 4; ++q) cudaFree(c->hbuf[q]); }
  return 0;
}

}  // extern "C"

// exchange the halos of a list of fields (base pointers + plane counts)
int halo_exchange(roms_b200_ctx* c, double* const* bases, const int* nplanes, int nf) {
  if (!c->comm) return 0;                                   // single tile: periodic images are written by the kernels
  if (nf > HALO_MAXF) return 1;
  HaloList L; L.nf = nf; L.total_planes = 0;
  for (int f = 0; f < nf; ++f) { L.base[f] = bases[f]; L.nplanes[f] = nplanes[f]; L.total_planes += nplanes[f]; }
  for (int f = nf; f < HALO_MAXF; ++f) { L.base[f] = nullptr; L.nplanes[f] = 1 << 30; }
  if (L.total_planes > HALO_MAXPLANES) return 1;
  const int w = c->D.halo;
  if (c->p2p_on) {
    const roms_b200_bounds& b = c->D.b; const int ni = c->D.ni;
    int r8[P2P_NDIR], sr[4 * P2P_NDIR], rr[4 * P2P_NDIR];
    if (roms_b200_halo_plan(&b, w, r8, sr, rr)) return 1;
    bool any = false;
    for (int q = 0; q < P2P_NDIR; ++q) any = any || c->p2p_rank[q] >= 0;
    if (!any) return 0;
    auto rect = [&](const int* r) { return Rect{(r[0] - b.LBi) + ni * (r[2] - b.LBj), r[1] - r[0] + 1, r[3] - r[2] + 1}; };
    P2PArgs a{};
    a.mine = c->p2p_mem; a.cap = c->halo_cap; a.seq = c->p2p_seq; a.ticket = c->p2p_ticket;
    for (int q = 0; q < P2P_NDIR; ++q) {
      a.peer[q] = (c->p2p_rank[q] >= 0) ? c->p2p_peer[q] : nullptr;
      a.snd[q] = rect(sr + 4 * q); a.rcv[q] = rect(rr + 4 * q);
    }
    static int nsm = 0;
    if (!nsm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev); }
    // one block per (plane, direction) item, at most 8 per SM (a single block for the small 2-D swaps was measured much
    // slower: the strip copies are latency-bound and want to run side by side)
    int maxn = 1;
    for (int q = 0; q < P2P_NDIR; ++q) if (a.peer[q]) { maxn = std::max(maxn, a.snd[q].w * a.snd[q].h); maxn = std::max(maxn, a.rcv[q].w * a.rcv[q].h); }
    static const int ch_off = getenv("ROMS_B200_HALO_NOCHUNK") != nullptr;
    static const int ch_env = getenv("ROMS_B200_HALO_CHUNK") ? atoi(getenv("ROMS_B200_HALO_CHUNK")) : P2P_CH;
    a.ch = ch_off ? maxn : (ch_env > 0 ? ch_env : P2P_CH); a.nsub = (maxn + a.ch - 1) / a.ch;
    int nblk = L.total_planes * P2P_NDIR * a.nsub; if (nblk > 8 * nsm) nblk = 8 * nsm;
    // Launched as a programmatic dependent of the kernel before it (its launch latency and block scheduling overlap that
    // kernel's tail; it waits in griddepcontrol.wait before touching any field) and releasing the kernel behind it at once
    // (a step2d sub-step, which waits the same way until this exchange has completed).  ROMS_B200_HALO_PDL=0: ordinary launch.
    static const bool pdl = !(getenv("ROMS_B200_HALO_PDL") && atoi(getenv("ROMS_B200_HALO_PDL")) == 0) && !getenv("ROMS_B200_NO_PDL");
    if (pdl) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(nblk); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      CUDA_OK(cudaLaunchKernelEx(&cfg, halo_xchg_p2p_kernel, c->D, L, a));
    } else halo_xchg_p2p_kernel<<<nblk, 256, 0, c->stream>>>(c->D, L, a);
    c->launches++;
    return 0;
  }
  ncclComm_p comm = (ncclComm_p)c->comm;
  for (int phase = 0; phase < 2; ++phase) {
    const int lo = phase == 0 ? c->nbW : c->nbS, hi = phase == 0 ? c->nbE : c->nbN;
    if (lo < 0 && hi < 0) continue;
    const size_t n = (size_t)w * (phase == 0 ? c->D.nj : c->D.ni), cnt = n * L.total_planes;
    dim3 g((unsigned)((n + 255) / 256), (unsigned)L.total_planes);

    double *sLo = c->hbuf[0], *sHi = c->hbuf[1], *rLo = c->hbuf[2], *rHi = c->hbuf[3];
    halo_pack_kernel<<<g, 256, 0, c->stream>>>(c->D, L, phase, w, sLo, sHi); c->launches++;
    NCCL_OK(g_nccl.GroupStart());
    // my low strip goes to the low neighbour (it becomes its high ghosts); I receive my low ghosts from it
    if (lo >= 0) { NCCL_OK(g_nccl.Send(sLo, cnt, kNcclFloat64, lo, comm, c->stream)); NCCL_OK(g_nccl.Recv(rLo, cnt, kNcclFloat64, lo, comm, c->stream)); }
    if (hi >= 0) { NCCL_OK(g_nccl.Send(sHi, cnt, kNcclFloat64, hi, comm, c->stream)); NCCL_OK(g_nccl.Recv(rHi, cnt, kNcclFloat64, hi, comm, c->stream)); }
    NCCL_OK(g_nccl.GroupEnd());
    // what I received from the low neighbour is ITS high strip -> my low ghosts (and vice versa).
    // With only two tiles on a periodic axis lo==hi: the two messages to/from the same peer are
    // matched in issue order, so rLo holds the peer's sLo (its LOW strip = my HIGH ghosts): swap.
    const bool same = (lo >= 0 && lo == hi);
    halo_unpack_kernel<<<g, 256, 0, c->stream>>>(c->D, L, phase, w, same ? rHi : rLo, same ? rLo : rHi, lo >= 0, hi >= 0); c->launches++;
  }
  return 0;
}

// diag's mp_reduce / mp_reduce2 (Utility/distribute.F): element-wise sum of n doubles over all tiles.  k_grid.cu gives every
// tile its own slot of the vector (zeros elsewhere), so one all-reduce gathers every tile's partial results on every rank
// and the host combines them in tile order (deterministic, and MAXLOC needs no second round).
int halo_allreduce_sum(roms_b200_ctx* c, double* dev, int n) {
  if (!c->comm) return 0;
  NCCL_OK(g_nccl.AllReduce(dev, dev, (size_t)n, kNcclFloat64, kNcclSum, (ncclComm_p)c->comm, c->stream));
  return 0;
}
