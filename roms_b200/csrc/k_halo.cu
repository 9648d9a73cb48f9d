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
#include "common.cuh"
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
}  // namespace

extern "C" {

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
  { int nb[4]; roms_b200_tile_neighbors(&b, nb); c->nbW = nb[0]; c->nbE = nb[1]; c->nbS = nb[2]; c->nbN = nb[3]; }
  // buffers: up to HALO_MAXPLANES planes of the larger strip
  const size_t strip = (size_t)c->D.halo * (size_t)((c->D.ni > c->D.nj) ? c->D.ni : c->D.nj);
  c->halo_cap = strip * HALO_MAXPLANES;
  for (int q = 0; q < 4; ++q) CUDA_OK(cudaMalloc((void**)&c->hbuf[q], c->halo_cap * sizeof(double)));
  return 0;
}
int roms_b200_comm_destroy(roms_b200_ctx* c) {
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

// diag's mp_reduce (Utility/distribute.F:6880): sum of 3 doubles over all tiles
int halo_allreduce_sum(roms_b200_ctx* c, double* dev3) {
  if (!c->comm) return 0;
  NCCL_OK(g_nccl.AllReduce(dev3, dev3, 3, kNcclFloat64, kNcclSum, (ncclComm_p)c->comm, c->stream));
  return 0;
}
