// tests/emu/emu_rt.cpp -- TEST INFRASTRUCTURE ONLY (see include/cuda_runtime.h): grid execution on the host and stubs for the
// parts of the library that are not built for emulation (PTX kernels, NCCL / peer-memory transport).
#include <cuda_runtime.h>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include "../../roms_b200/csrc/common.cuh"

thread_local uint3 threadIdx = {0, 0, 0};
uint3 blockIdx = {0, 0, 0};
dim3 blockDim(1, 1, 1), gridDim(1, 1, 1);
double emu_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

namespace emu {
namespace {
// kernels that call __syncthreads(): their blocks run as teams of real threads; every other kernel runs its threads in a loop
const char* kTeamKernels[] = {"step2d_kernel", "diag_cols_kernel", "diag_sum_kernel", "step3d_t_v6_kernel", "step3d_t_v7_kernel"};
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
}  // namespace

void* dyn_smem() { return g_smem.data(); }
void barrier() {
  if (!in_team) { fprintf(stderr, "emu: __syncthreads() in a kernel that is not listed in kTeamKernels (tests/emu/emu_rt.cpp)\n"); abort(); }
  team().barrier();
}
void named_barrier(int id, int nthreads, bool wait) {
  if (!in_team) { fprintf(stderr, "emu: named barrier outside a team kernel\n"); abort(); }
  team().named_barrier(id, nthreads, wait);
}
bool warp_any(bool pred) {
  if (!in_team) { fprintf(stderr, "emu: warp vote outside a team kernel\n"); abort(); }
  const unsigned lin = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  return team().warp_any((int)(lin / 32), pred);
}
void sync_warp() { if (in_team) (void)warp_any(false); }
int lane_id() { return (int)((threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)) & 31); }
double warp_shfl(double v, int src) {
  if (!in_team) { fprintf(stderr, "emu: warp shuffle outside a team kernel\n"); abort(); }
  const unsigned lin = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
  return team().warp_shfl((int)(lin / 32), (int)(lin & 31), v, src);
}
void yield() { std::this_thread::yield(); }
void run_grid(dim3 g, dim3 b, size_t smem, const char* kernel, const std::function<void()>& body) {
  if (smem > g_smem.size() * sizeof(double)) { fprintf(stderr, "emu: %zu bytes of dynamic shared memory requested by %s\n", smem, kernel); abort(); }
  gridDim = g; blockDim = b;
  const bool tm = needs_team(kernel);
  for (unsigned bz = 0; bz < g.z; ++bz) for (unsigned by = 0; by < g.y; ++by) for (unsigned bx = 0; bx < g.x; ++bx) {
    blockIdx = uint3{bx, by, bz};
    if (tm) { team().run_block(b, body); continue; }
    for (unsigned tz = 0; tz < b.z; ++tz) for (unsigned ty = 0; ty < b.y; ++ty) for (unsigned tx = 0; tx < b.x; ++tx) {
      threadIdx = uint3{tx, ty, tz};
      body();
    }
  }
}
}  // namespace emu

// ---- not built for emulation: the shuffle-based column step3d_t (k_step3d_t4.cu, the fallback of the production kernel for
// closed W/E walls and N < 4) and the halo transport.
int k_step3d_t_v4(roms_b200_ctx*, int) { fprintf(stderr, "emu: k_step3d_t4.cu is not built for emulation (N < 4 or closed W/E walls)\n"); return 1; }
int halo_exchange(roms_b200_ctx*, double* const*, const int*, int) { return 0; }
int halo_allreduce_sum(roms_b200_ctx*, double*, int) { return 0; }
extern "C" {
int roms_b200_comm_unique_id(char*) { return 1; }
int roms_b200_comm_init(roms_b200_ctx*, int, int, const char*) { return 1; }
int roms_b200_comm_destroy(roms_b200_ctx*) { return 0; }
int roms_b200_p2p_handle(roms_b200_ctx*, char*) { return 1; }
int roms_b200_p2p_connect(roms_b200_ctx*, const char*, int) { return 1; }
}
