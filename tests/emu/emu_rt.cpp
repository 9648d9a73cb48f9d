// tests/emu/emu_rt.cpp -- TEST INFRASTRUCTURE ONLY (see include/cuda_runtime.h): grid execution on the host and stubs for the
// parts of the library that are not built for emulation (PTX kernels, NCCL / peer-memory transport).
#include <cuda_runtime.h>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include <ucontext.h>
#include "../../roms_b200/csrc/common.cuh"

thread_local uint3 threadIdx = {0, 0, 0};
uint3 blockIdx = {0, 0, 0};
dim3 blockDim(1, 1, 1), gridDim(1, 1, 1);
double emu_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

namespace emu {
namespace {
// kernels that call __syncthreads(): their blocks run as teams of real threads; every other kernel runs its threads in a loop
const char* kTeamKernels[] = {"step2d_kernel", "diag_cols_kernel", "diag_sum_kernel", "step3d_t_v6_kernel", "step3d_t_v8_kernel", "step3d_t_v4_kernel", "t3dmix2_geo_roll_kernel", "pre_step3d_t_roll_kernel", "uv3dmix2_roll_kernel", "rhs3d_roll_kernel"};
bool needs_team(const char* k) { for (const char* t : kTeamKernels) if (strstr(k, t)) return true; return false; }
std::vector<double> g_smem(64 * 1024, 0.0);          // dynamic shared memory of the running block
thread_local bool in_team = false;

// persistent team: worker w runs thread w of the current block
struct Team {
  std::vector<std::thread> th; std::mutex mu; std::condition_variable cv_go, cv_done, cv_bar;
  const std::function<void()>* body = nullptr; dim3 bdim; long gen = 0; int nthr = 0, pending = 0; bool stop = false;
  int bar_count = 0, bar_live = 0; long bar_gen = 0;
  struct Named { int count = 0; long gen = 0; } named[16];           // PTX named barriers 0..15
  struct Vote { int count = 0; bool acc = false, result = false; long gen = 0; } vote[64];   // one per warp of the block
  std::condition_variable cv_named;
  void worker(int w) {
    long seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(mu);
      cv_go.wait(lk, [&] { return stop || gen != seen; });
      if (stop) return;
      seen = gen;
      const bool mine = w < nthr; auto f = body; const dim3 b = bdim;
      lk.unlock();
      if (mine) {
        threadIdx.x = w % b.x; threadIdx.y = (w / b.x) % b.y; threadIdx.z = w / (b.x * b.y);
        in_team = true;
        (*f)();
        in_team = false;
        lk.lock();
        --bar_live;                                   // a thread that left the kernel no longer takes part in barriers
        if (bar_live > 0 && bar_count == bar_live) { bar_count = 0; ++bar_gen; cv_bar.notify_all(); }
        if (--pending == 0) cv_done.notify_one();
      }
    }
  }
  void ensure(int n) { while ((int)th.size() < n) { const int w = (int)th.size(); th.emplace_back(&Team::worker, this, w); } }
  void run_block(dim3 b, const std::function<void()>& f) {
    const int n = (int)(b.x * b.y * b.z);
    ensure(n);
    std::unique_lock<std::mutex> lk(mu);
    body = &f; bdim = b; nthr = n; pending = n; bar_live = n; bar_count = 0; ++gen;
    for (auto& q : named) q.count = 0;
    for (auto& q : vote) { q.count = 0; q.acc = false; }
    cv_go.notify_all();
    cv_done.wait(lk, [&] { return pending == 0; });
  }
  void named_barrier(int id, int n, bool wait) {
    std::unique_lock<std::mutex> lk(mu);
    Named& b = named[id & 15];
    const long g = b.gen;
    if (++b.count == n) { b.count = 0; ++b.gen; cv_named.notify_all(); return; }
    if (wait) cv_named.wait(lk, [&] { return b.gen != g; });
  }
  double shbuf[64][32];
  double warp_shfl(int warp, int lane, double v, int src) {          // two rendezvous: all lanes have written / all lanes have read
    { std::unique_lock<std::mutex> lk(mu); shbuf[warp & 63][lane] = v; }
    (void)warp_any(warp, false);
    double r;
    { std::unique_lock<std::mutex> lk(mu); r = shbuf[warp & 63][src & 31]; }
    (void)warp_any(warp, false);
    return r;
  }
  bool warp_any(int warp, bool pred) {
    std::unique_lock<std::mutex> lk(mu);
    Vote& v = vote[warp & 63];
    const long g = v.gen;
    v.acc = v.acc || pred;
    if (++v.count == 32) { v.result = v.acc; v.acc = false; v.count = 0; ++v.gen; cv_named.notify_all(); return v.result; }
    cv_named.wait(lk, [&] { return v.gen != g; });
    return v.result;
  }
  void barrier() {
    std::unique_lock<std::mutex> lk(mu);
    const long g = bar_gen;
    if (++bar_count == bar_live) { bar_count = 0; ++bar_gen; cv_bar.notify_all(); return; }
    cv_bar.wait(lk, [&] { return bar_gen != g; });
  }
  ~Team() { { std::lock_guard<std::mutex> lk(mu); stop = true; } cv_go.notify_all(); for (auto& t : th) t.join(); }
};
Team& team() { static Team* t = new Team(); return *t; }   // leaked on purpose: workers may outlive static destruction order

// The same block execution with user-level fibers (ucontext) on the calling host thread: a barrier, a vote, a shuffle or a
// spin-wait hands control to the next thread of the block instead of going through the kernel's futexes -- an order of
// magnitude faster for 256-576 threads per block and deterministic.  Default; EMU_TEAM=threads selects the OS-thread team
// (used for the AddressSanitizer build, which does not follow swapcontext).
struct Fibers {
  struct F { ucontext_t ctx; bool done = false; };
  std::vector<F> f; std::vector<char*> stacks; ucontext_t sched; int cur = -1, live = 0, n = 0;
  const std::function<void()>* body = nullptr; dim3 bdim;
  int bar_count = 0; long bar_gen = 0;
  struct Named { int count = 0; long gen = 0; } named[16];
  struct Vote { int count = 0; bool acc = false, result = false; long gen = 0; } vote[64];
  double shbuf[64][32];
  static constexpr size_t kStack = 256 * 1024;
  static void entry(unsigned lo, unsigned hi) {
    Fibers* self = (Fibers*)(((uintptr_t)hi << 32) | (uintptr_t)lo);
    in_team = true;
    (*self->body)();
    F& me = self->f[self->cur];
    me.done = true;
    --self->live;
    if (self->live > 0 && self->bar_count == self->live) { self->bar_count = 0; ++self->bar_gen; }   // leavers drop out of barriers
    swapcontext(&me.ctx, &self->sched);
  }
  void set_tid(int w) { threadIdx.x = w % bdim.x; threadIdx.y = (w / bdim.x) % bdim.y; threadIdx.z = w / (bdim.x * bdim.y); }
  void run_block(dim3 b, const std::function<void()>& fn) {
    n = (int)(b.x * b.y * b.z); bdim = b; body = &fn; live = n; bar_count = 0;
    for (auto& q : named) q.count = 0;
    for (auto& q : vote) { q.count = 0; q.acc = false; }
    if ((int)f.size() < n) f.resize(n);
    while ((int)stacks.size() < n) stacks.push_back((char*)malloc(kStack));
    const uintptr_t me = (uintptr_t)this;
    for (int w = 0; w < n; ++w) {
      f[w].done = false;
      getcontext(&f[w].ctx);
      f[w].ctx.uc_stack.ss_sp = stacks[w]; f[w].ctx.uc_stack.ss_size = kStack; f[w].ctx.uc_link = &sched;
      makecontext(&f[w].ctx, (void (*)())entry, 2, (unsigned)(me & 0xffffffffu), (unsigned)(me >> 32));
    }
    // EMU_SCHED_SEED=s: the threads of a block are resumed in a pseudo-random order that changes at every pass (whole warps
    // stay together or not depending on the draw), to shake out orderings a round-robin schedule never produces
    static const char* seed_env = getenv("EMU_SCHED_SEED");
    static unsigned long long rng = seed_env ? (unsigned long long)atoll(seed_env) * 2654435761ull + 88172645463325252ull : 0;
    std::vector<int> order(n);
    for (int w = 0; w < n; ++w) order[w] = w;
    while (live > 0) {
      if (seed_env)
        for (int w = n - 1; w > 0; --w) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; std::swap(order[w], order[(int)(rng % (unsigned)(w + 1))]); }
      for (int q = 0; q < n; ++q) { const int w = order[q]; if (!f[w].done) { cur = w; set_tid(w); swapcontext(&sched, &f[w].ctx); } }
    }
    in_team = false;
  }
  void yield() { const int w = cur; swapcontext(&f[w].ctx, &sched); }         // the scheduler restores cur/threadIdx before resuming
  void barrier() { const long g = bar_gen; if (++bar_count == live) { bar_count = 0; ++bar_gen; return; } while (bar_gen == g) yield(); }
  void named_barrier(int id, int cnt, bool wait) {
    Named& b = named[id & 15]; const long g = b.gen;
    if (++b.count == cnt) { b.count = 0; ++b.gen; return; }
    if (wait) while (b.gen == g) yield();
  }
  bool warp_any(int warp, bool pred) {
    Vote& v = vote[warp & 63]; const long g = v.gen;
    v.acc = v.acc || pred;
    if (++v.count == 32) { v.result = v.acc; v.acc = false; v.count = 0; ++v.gen; return v.result; }
    while (v.gen == g) yield();
    return v.result;
  }
  double warp_shfl(int warp, int lane, double val, int src) {
    shbuf[warp & 63][lane] = val;
    (void)warp_any(warp, false);
    const double r = shbuf[warp & 63][src & 31];
    (void)warp_any(warp, false);
    return r;
  }
};
Fibers& fibers() { static Fibers* p = new Fibers(); return *p; }

// Cooperative grids (kernels with a grid-wide barrier: the persistent barotropic fast loop): EVERY thread of EVERY block is a
// fiber of one scheduler, resumed round-robin; a block barrier waits for the threads of that block, the kernel's own grid barrier
// (an atomic counter it spins on, yielding) for all blocks.  Each block has its own dynamic shared memory.
struct Coop {
  struct F { ucontext_t ctx; bool done = false; int blk = 0, tid = 0; };
  std::vector<F> f; std::vector<char*> stacks; ucontext_t sched; int cur = -1, live = 0;
  const std::function<void()>* body = nullptr; dim3 bdim, gdim;
  std::vector<int> bar_count, bar_live; std::vector<long> bar_gen;
  std::vector<std::vector<double>> smem;
  bool active = false;
  static constexpr size_t kStack = 96 * 1024;
  static void entry(unsigned lo, unsigned hi) {
    Coop* self = (Coop*)(((uintptr_t)hi << 32) | (uintptr_t)lo);
    in_team = true;
    (*self->body)();
    F& me = self->f[self->cur];
    me.done = true;
    --self->live;
    int& bl = self->bar_live[me.blk];
    --bl;
    if (bl > 0 && self->bar_count[me.blk] == bl) { self->bar_count[me.blk] = 0; ++self->bar_gen[me.blk]; }
    swapcontext(&me.ctx, &self->sched);
  }
  void set_ids(const F& x) {
    threadIdx.x = x.tid % bdim.x; threadIdx.y = (x.tid / bdim.x) % bdim.y; threadIdx.z = x.tid / (bdim.x * bdim.y);
    blockIdx.x = x.blk % gdim.x; blockIdx.y = (x.blk / gdim.x) % gdim.y; blockIdx.z = x.blk / (gdim.x * gdim.y);
  }
  void run(dim3 g, dim3 b, size_t smem_bytes, const std::function<void()>& fn) {
    const int nb = (int)(g.x * g.y * g.z), nt = (int)(b.x * b.y * b.z), n = nb * nt;
    gdim = g; bdim = b; body = &fn; live = n; active = true;
    bar_count.assign(nb, 0); bar_live.assign(nb, nt); bar_gen.assign(nb, 0);
    smem.assign(nb, std::vector<double>((smem_bytes + 7) / 8 + 16, 0.0));
    if ((int)f.size() < n) f.resize(n);
    while ((int)stacks.size() < n) stacks.push_back((char*)malloc(kStack));
    const uintptr_t me = (uintptr_t)this;
    for (int w = 0; w < n; ++w) {
      f[w].done = false; f[w].blk = w / nt; f[w].tid = w % nt;
      getcontext(&f[w].ctx);
      f[w].ctx.uc_stack.ss_sp = stacks[w]; f[w].ctx.uc_stack.ss_size = kStack; f[w].ctx.uc_link = &sched;
      makecontext(&f[w].ctx, (void (*)())entry, 2, (unsigned)(me & 0xffffffffu), (unsigned)(me >> 32));
    }
    // EMU_SCHED_SEED=s: in every pass a pseudo-random half of the BLOCKS is left out, so blocks drift apart by whole sub-steps
    // unless the kernel's own inter-block synchronisation holds them together (round-robin would keep them in lockstep)
    static const char* seed_env = getenv("EMU_SCHED_SEED");
    static unsigned long long rng = seed_env ? (unsigned long long)atoll(seed_env) * 2654435761ull + 88172645463325252ull : 0;
    std::vector<char> on(nb, 1);
    while (live > 0) {
      if (seed_env) for (int q = 0; q < nb; ++q) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; on[q] = (char)((rng >> 20) & 1); }
      for (int w = 0; w < n; ++w) if (!f[w].done && on[f[w].blk]) { cur = w; set_ids(f[w]); swapcontext(&sched, &f[w].ctx); }
    }
    in_team = false; active = false;
  }
  void yield() { const int w = cur; swapcontext(&f[w].ctx, &sched); set_ids(f[w]); }
  void barrier() {
    const int blk = f[cur].blk; const long g = bar_gen[blk];
    if (++bar_count[blk] == bar_live[blk]) { bar_count[blk] = 0; ++bar_gen[blk]; return; }
    while (bar_gen[blk] == g) yield();
  }
  void* dyn() { return smem[f[cur].blk].data(); }
};
Coop& coop() { static Coop* p = new Coop(); return *p; }
bool use_fibers() { static const bool v = !(getenv("EMU_TEAM") && !strcmp(getenv("EMU_TEAM"), "threads")); return v; }
}  // namespace

void* dyn_smem() { return coop().active ? coop().dyn() : (void*)g_smem.data(); }
bool coop_supported() { return use_fibers(); }
void barrier() {
  if (coop().active) { coop().barrier(); return; }
  if (!in_team) { fprintf(stderr, "emu: __syncthreads() in a kernel that is not listed in kTeamKernels (tests/emu/emu_rt.cpp)\n"); abort(); }
  if (use_fibers()) fibers().barrier(); else team().barrier();
}
void named_barrier(int id, int nthreads, bool wait) {
  if (!in_team) { fprintf(stderr, "emu: named barrier outside a team kernel\n"); abort(); }
  if (use_fibers()) fibers().named_barrier(id, nthreads, wait); else team().named_barrier(id, nthreads, wait);
}
bool warp_any(bool pred) {
  if (!in_team) { fprintf(stderr, "emu: warp vote outside a team kernel\n"); abort(); }
  const unsigned lin = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  return use_fibers() ? fibers().warp_any((int)(lin / 32), pred) : team().warp_any((int)(lin / 32), pred);
}
void sync_warp() { if (in_team) (void)warp_any(false); }
int lane_id() { return (int)((threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)) & 31); }
double warp_shfl(double v, int src) {
  if (!in_team) { fprintf(stderr, "emu: warp shuffle outside a team kernel\n"); abort(); }
  const unsigned lin = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  return use_fibers() ? fibers().warp_shfl((int)(lin / 32), (int)(lin & 31), v, src) : team().warp_shfl((int)(lin / 32), (int)(lin & 31), v, src);
}
void yield() { if (coop().active) coop().yield(); else if (in_team && use_fibers()) fibers().yield(); else std::this_thread::yield(); }
std::mutex g_kernel_lock;          // several ranks (one host thread each, see the multi-tile emulation below): one kernel at a time
void run_grid(dim3 g, dim3 b, size_t smem, const char* kernel, const std::function<void()>& body) {
  std::lock_guard<std::mutex> kl(g_kernel_lock);
  if (smem > g_smem.size() * sizeof(double)) { fprintf(stderr, "emu: %zu bytes of dynamic shared memory requested by %s\n", smem, kernel); abort(); }
  gridDim = g; blockDim = b;
  const bool tm = needs_team(kernel);
  for (unsigned bz = 0; bz < g.z; ++bz) for (unsigned by = 0; by < g.y; ++by) for (unsigned bx = 0; bx < g.x; ++bx) {
    blockIdx = uint3{bx, by, bz};
    if (tm) { if (use_fibers()) fibers().run_block(b, body); else team().run_block(b, body); continue; }
    for (unsigned tz = 0; tz < b.z; ++tz) for (unsigned ty = 0; ty < b.y; ++ty) for (unsigned tx = 0; tx < b.x; ++tx) {
      threadIdx = uint3{tx, ty, tz};
      body();
    }
  }
}
void run_grid_coop(dim3 g, dim3 b, size_t smem, const char* kernel, const std::function<void()>& body) {
  std::lock_guard<std::mutex> kl(g_kernel_lock);
  if (!use_fibers()) { fprintf(stderr, "emu: cooperative grid %s needs the fiber scheduler\n", kernel); abort(); }
  gridDim = g; blockDim = b;
  coop().run(g, b, smem, body);
}
}  // namespace emu

// ---- not built for emulation: the halo transport (k_halo.cu); see the multi-tile emulation below.
// ---- multi-tile emulation: every rank is a host thread of this process with its own context (mirror); the halo exchange of
// k_halo.cu becomes a copy between the mirrors (each rank PULLS the blocks its eight neighbours would send, rectangles from
// roms_b200_halo_plan, between two rendezvous of all ranks) and diag's all-reduce a sum over the ranks' buffers in rank order.
// This checks the distributed LOGIC of the library (tile bounds, redundant evaluation on halos, deep-halo predictor, what is
// exchanged when) -- the reference's acceptance criterion "results do not depend on the tiling" -- not the transport itself.
namespace {
struct EmuComm {
  std::mutex mu; std::condition_variable cv; int n = 0, count = 0; long gen = 0;
  std::vector<roms_b200_ctx*> ctx; std::vector<std::vector<double*>> bases; std::vector<double*> red;
  void rendezvous() {
    std::unique_lock<std::mutex> lk(mu);
    const long g = gen;
    if (++count == n) { count = 0; ++gen; cv.notify_all(); return; }
    cv.wait(lk, [&] { return gen != g; });
  }
} g_comm;
const int kOpp[8] = {1, 0, 3, 2, 7, 6, 5, 4};       // W<->E, S<->N, SW<->NE, SE<->NW (directions of roms_b200_halo_plan)
}  // namespace
int halo_exchange(roms_b200_ctx* c, double* const* bases, const int* nplanes, int nf) {
  if (!c->comm) return 0;
  EmuComm& G = g_comm;
  { std::lock_guard<std::mutex> lk(G.mu); G.bases[c->rank].assign(bases, bases + nf); }
  G.rendezvous();                                    // every rank has finished the kernels before this swap and published its list
  const roms_b200_bounds& b = c->D.b; const int w = c->D.halo;
  int r8[8], snd[32], rcv[32];
  if (roms_b200_halo_plan(&b, w, r8, snd, rcv)) return 1;
  for (int d = 0; d < 8; ++d) {
    if (r8[d] < 0) continue;
    roms_b200_ctx* p = G.ctx[r8[d]];
    int pr8[8], psnd[32], prcv[32];
    if (roms_b200_halo_plan(&p->D.b, p->D.halo, pr8, psnd, prcv)) return 1;
    const int* S = psnd + 4 * kOpp[d]; const int* R = rcv + 4 * d;          // {i0,i1,j0,j1} in global indices
    const int wi = R[1] - R[0] + 1, wj = R[3] - R[2] + 1;
    if (wi != S[1] - S[0] + 1 || wj != S[3] - S[2] + 1) { fprintf(stderr, "emu: halo plan mismatch dir %d\n", d); return 1; }
    for (int f = 0; f < nf; ++f) for (int pl = 0; pl < nplanes[f]; ++pl) {
      const double* src = G.bases[r8[d]][f] + (size_t)pl * p->D.nij;
      double* dst = bases[f] + (size_t)pl * c->D.nij;
      for (int jj = 0; jj < wj; ++jj)
        memcpy(dst + (R[0] - b.LBi) + (size_t)c->D.ni * (R[2] + jj - b.LBj),
               src + (S[0] - p->D.b.LBi) + (size_t)p->D.ni * (S[2] + jj - p->D.b.LBj), sizeof(double) * wi);
    }
  }
  G.rendezvous();                                    // nobody touches a mirror again before every rank has pulled its halos
  return 0;
}
int halo_allreduce_sum(roms_b200_ctx* c, double* dev, int n) {
  if (!c->comm) return 0;
  EmuComm& G = g_comm;
  { std::lock_guard<std::mutex> lk(G.mu); G.red[c->rank] = dev; }
  G.rendezvous();
  std::vector<double> acc(n, 0.0);
  for (int r = 0; r < G.n; ++r) for (int q = 0; q < n; ++q) acc[q] += G.red[r][q];
  G.rendezvous();
  for (int q = 0; q < n; ++q) dev[q] = acc[q];
  return 0;
}
extern "C" {
int roms_b200_comm_unique_id(char* id) { memset(id, 0, 128); return 0; }
int roms_b200_comm_init(roms_b200_ctx* c, int rank, int nranks, const char*) {
  if (!c) return 1;
  const roms_b200_bounds& b = c->D.b;
  if (nranks != b.NtileI * b.NtileJ || rank != b.Jtile * b.NtileI + b.Itile) { fprintf(stderr, "emu: rank/tile mismatch\n"); return 1; }
  EmuComm& G = g_comm;
  std::lock_guard<std::mutex> lk(G.mu);
  if (G.n != nranks) { G.n = nranks; G.ctx.assign(nranks, nullptr); G.bases.assign(nranks, {}); G.red.assign(nranks, nullptr); G.count = 0; }
  G.ctx[rank] = c;
  c->comm = &G; c->rank = rank; c->nranks = nranks;
  c->deep = (c->D.halo >= 6 && getenv("ROMS_B200_NO_DEEP_HALO") == nullptr) ? 1 : 0;
  { int nb[4]; roms_b200_tile_neighbors(&b, nb); c->nbW = nb[0]; c->nbE = nb[1]; c->nbS = nb[2]; c->nbN = nb[3]; }
  return 0;
}
int roms_b200_comm_destroy(roms_b200_ctx* c) { if (c) c->comm = nullptr; return 0; }
int roms_b200_p2p_handle(roms_b200_ctx*, char* h) { memset(h, 0, 64); return 0; }
int roms_b200_p2p_connect(roms_b200_ctx*, const char*, int) { return 0; }      // the in-process copy above is the transport
}
