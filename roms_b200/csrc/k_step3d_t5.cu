// roms_b200/csrc/k_step3d_t5.cu -- step3d_t_tile with TMA-staged tiles (sm_100a).
//
// 2.5-D blocked sweep: a CTA owns a 32(i) x 8(j) patch of water columns and marches k.  For every
// level one elected thread issues seven `cp.async.bulk.tensor.3d` (TMA) loads -- the (32+4)x(8+4)
// tile of t(3) with its U3 halo, and the Huon/Hvom/W/Hz/Akt/t(nnew) rows of the patch -- into a
// 4-deep shared-memory ring; completion is tracked by one mbarrier per stage (complete_tx::bytes).
// The loads of levels k+1..k+3 are therefore in flight while level k is being computed, independent of
// occupancy or register pressure (the earlier layouts were long-scoreboard bound: profiles/).
// The stencil is evaluated from shared memory; the vertical recurrences are those of k_step3d_t4.cu
// (checkpointed Thomas, 256 B of private storage), so per-point arithmetic -- and the result bits --
// are unchanged (step3d_t.F:641-916,1150-1365,1672-1721; -fmad=false).
#include "common.cuh"
#include <cuda.h>
#include <cstdlib>
#include <map>
#include <tuple>

namespace {
constexpr int TI = 32, TJ = 8, NSTAGE = 4;
constexpr int T3_W = TI + 4, T3_H = TJ + 4;          // t(3) tile with halo 2
constexpr int HU_W = TI + 2;                          // Huon(i..i+1): 33 used, 34 keeps the row a multiple of 16 B
constexpr int HV_H = TJ + 1;
constexpr int OFF_T3 = 0, OFF_HU = OFF_T3 + T3_W * T3_H, OFF_HV = OFF_HU + HU_W * TJ, OFF_W = OFF_HV + TI * HV_H;
constexpr int OFF_HZ = OFF_W + TI * TJ, OFF_AK = OFF_HZ + TI * TJ, OFF_TW = OFF_AK + TI * TJ, STAGE_DBL = OFF_TW + TI * TJ;
constexpr uint32_t STAGE_BYTES = STAGE_DBL * 8;
static_assert((OFF_HU * 8) % 128 == 0 && (OFF_HV * 8) % 128 == 0 && (OFF_W * 8) % 128 == 0 && (STAGE_BYTES % 128) == 0, "TMA smem alignment");
constexpr int SEG = 8, MAXSEG = RB_MAXN / SEG;

struct Edges { int S, N, Jstr, Jend; };
__device__ __forceinline__ int jclamp(int j, const Edges& e) {
  if (e.S && j == e.Jstr - 1) return e.Jstr;
  if (e.N && j == e.Jend + 2) return e.Jend + 1;
  return j;
}
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(s32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
struct Maps { CUtensorMap t3, hu, hv, w, hz, ak, tw; };
}  // namespace

__global__ void __launch_bounds__(TI* TJ, 3) step3d_t_v5_kernel(const Dev D, Box bx, int nnew, int itrc, const __grid_constant__ Maps M) {
  extern __shared__ __align__(128) double sm[];
  __shared__ __align__(8) uint64_t full[NSTAGE];
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TI + tx, N = D.b.N;
  const int i0 = bx.i0 + blockIdx.x * TI, j0 = bx.j0 + blockIdx.y * TJ;
  const int i = i0 + tx, j = j0 + ty;
  const bool act = (i <= bx.i1 && j <= bx.j1);
  const int ig = min(i, bx.i1), jg = min(j, bx.j1);           // clamped indices for the few direct global accesses of idle threads
  const double dt = D.p.dt;
  const Edges e{D.b.Southern_Edge && !D.b.NSperiodic, D.b.Northern_Edge && !D.b.NSperiodic, D.b.Jstr, D.b.Jend};
  const bool south = e.S && j == e.Jstr, north = e.N && j == e.Jend;
  V3 Hzg = v3(D, FID(Hz)), t3g = v3l(D, FID(t), 3, itrc), Aktg = v3l(D, FID(Akt), min(D.b.NAT, itrc)), tw = v3l(D, FID(t), nnew, itrc);
  // tile coordinates of the tensor maps (array index space)
  const int ci = i0 - D.b.LBi, cj = j0 - D.b.LBj;
  auto issue = [&](int lvl) {          // lvl = 1..N -> stage (lvl-1)%NSTAGE ; called by thread 0 only
    const int s = (lvl - 1) % NSTAGE;
    double* st = sm + (size_t)s * STAGE_DBL;
    mbar_expect_tx(&full[s], STAGE_BYTES);
    tma_load_3d(st + OFF_T3, &M.t3, &full[s], ci - 2, cj - 2, lvl - 1);
    tma_load_3d(st + OFF_HU, &M.hu, &full[s], ci, cj, lvl - 1);
    tma_load_3d(st + OFF_HV, &M.hv, &full[s], ci, cj, lvl - 1);
    tma_load_3d(st + OFF_W, &M.w, &full[s], ci, cj, lvl);          // W(0:N): level k is plane k
    tma_load_3d(st + OFF_HZ, &M.hz, &full[s], ci, cj, lvl - 1);
    tma_load_3d(st + OFF_AK, &M.ak, &full[s], ci, cj, lvl);        // Akt(0:N)
    tma_load_3d(st + OFF_TW, &M.tw, &full[s], ci, cj, lvl - 1);
  };
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) for (int l = 1; l <= NSTAGE && l <= N; ++l) issue(l);

  const double cff = dt * v2(D, FID(pm))(ig, jg) * v2(D, FID(pn))(ig, jg);
  // shared-memory row indices of the (clamped) eta first differences
  const int rjm2 = jclamp(j - 1, e) - 1 - j0 + 2, rjm1 = jclamp(j - 1, e) - j0 + 2, rj0b = jclamp(j, e) - 1 - j0 + 2, rj0a = jclamp(j, e) - j0 + 2;
  const int rjp1b = jclamp(j + 1, e) - 1 - j0 + 2, rjp1a = jclamp(j + 1, e) - j0 + 2, rjp2b = jclamp(j + 2, e) - 1 - j0 + 2, rjp2a = jclamp(j + 2, e) - j0 + 2;
  const bool edgeX = (tx == TI - 1) || (i == bx.i1);
  double tkm1 = t3g(ig, jg, 1), tk = tkm1, tkp1 = t3g(ig, jg, 2), FCm = 0.0;
  double tkp2 = (3 <= N) ? t3g(ig, jg, 3) : 0.0;          // prefetched one level ahead of use

  // q(k) from stage of level k; also hands back Hz(k), Akt(k) and refills the stage
  auto advect = [&](int k, double& hz, double& akt) -> double {
    const int s = (k - 1) % NSTAGE;
    mbar_wait(&full[s], ((k - 1) / NSTAGE) & 1);
    const double* st = sm + (size_t)s * STAGE_DBL;
    const double* T3 = st + OFF_T3;
#define T3s(ii, jj) T3[(jj) * T3_W + (ii)]
    const int cx = tx + 2, cy = ty + 2;
    const double tkp3 = (k + 3 <= N) ? t3g(ig, jg, k + 3) : 0.0;     // issued now, consumed next level
    const double qm2 = T3s(cx - 2, cy), qm1 = T3s(cx - 1, cy), q0 = T3s(cx, cy), qp1 = T3s(cx + 1, cy);
    const double hu = st[OFF_HU + ty * HU_W + tx];
    const double d0 = qm1 - qm2, d1 = q0 - qm1, d2 = qp1 - q0;
    const double FXi = hu * 0.5 * (qm1 + q0) - (1.0 / 6.0) * ((d1 - d0) * fmax(hu, 0.0) + (d2 - d1) * fmin(hu, 0.0));
    double FXp = __shfl_down_sync(0xffffffffu, FXi, 1);
    if (edgeX) {
      const double qp2 = T3s(cx + 2, cy), hup = st[OFF_HU + ty * HU_W + tx + 1], d3 = qp2 - qp1;
      FXp = hup * 0.5 * (q0 + qp1) - (1.0 / 6.0) * ((d2 - d1) * fmax(hup, 0.0) + (d3 - d2) * fmin(hup, 0.0));
    }
    const double e_m1 = T3s(cx, rjm1) - T3s(cx, rjm2), e_0 = T3s(cx, rj0a) - T3s(cx, rj0b);
    const double e_p1 = T3s(cx, rjp1a) - T3s(cx, rjp1b), e_p2 = T3s(cx, rjp2a) - T3s(cx, rjp2b);
    const double hv = st[OFF_HV + ty * TI + tx], hvp = st[OFF_HV + (ty + 1) * TI + tx];
    const double tjm = T3s(cx, cy - 1), tjp = T3s(cx, cy + 1);
    const double FEj = hv * 0.5 * (tjm + q0) - (1.0 / 6.0) * ((e_0 - e_m1) * fmax(hv, 0.0) + (e_p1 - e_0) * fmin(hv, 0.0));
    const double FEp = hvp * 0.5 * (q0 + tjp) - (1.0 / 6.0) * ((e_p1 - e_0) * fmax(hvp, 0.0) + (e_p2 - e_p1) * fmin(hvp, 0.0));
#undef T3s
    const double c1 = cff * (FXp - FXi), c2 = cff * (FEp - FEj), c3 = c1 + c2;
    double tv = st[OFF_TW + ty * TI + tx] - c3;
    const double wk = st[OFF_W + ty * TI + tx];
    double FCk;
    if (k == N) FCk = 0.0;
    else if (k == 1) FCk = wk * (0.5 * tk + (7.0 / 12.0) * tkp1 - (1.0 / 12.0) * tkp2);
    else if (k == N - 1) FCk = wk * (0.5 * tkp1 + (7.0 / 12.0) * tk - (1.0 / 12.0) * tkm1);
    else FCk = wk * ((7.0 / 12.0) * (tk + tkp1) - (1.0 / 12.0) * (tkm1 + tkp2));
    const double cv = cff * (FCk - FCm);
    FCm = FCk;
    hz = st[OFF_HZ + ty * TI + tx];
    akt = st[OFF_AK + ty * TI + tx];
    tv = tv - cv;
    tkm1 = tk; tk = tkp1; tkp1 = tkp2; tkp2 = tkp3;
    __syncthreads();                                   // every thread is done with stage s
    if (tid == 0 && k + NSTAGE <= N) issue(k + NSTAGE);
    return tv * (1.0 / hz);
  };

  // ---- sweep 1: advection + forward elimination with checkpoints (see k_step3d_t4.cu)
  double ck_cf[MAXSEG], ck_dc[MAXSEG];
  double hz_k, hz_kp, ak_k, ak_kp;
  double q_k = advect(1, hz_k, ak_k), q_kp;
  double ohz_k = 1.0 / hz_k, ak_km = Aktg(ig, jg, 0), cf_prev = 0.0, dc_prev = 0.0;
  for (int k = 1; k <= N - 1; ++k) {
    if (((k - 1) & (SEG - 1)) == 0) { ck_cf[(k - 1) / SEG] = cf_prev; ck_dc[(k - 1) / SEG] = dc_prev; }
    q_kp = advect(k + 1, hz_kp, ak_kp);
    const double ohz_kp = 1.0 / hz_kp;
    const double FC = (1.0 / 6.0) * hz_k - dt * ak_km * ohz_k;
    const double CFk = (1.0 / 6.0) * hz_kp - dt * ak_kp * ohz_kp;
    const double BC = (1.0 / 3.0) * (hz_k + hz_kp) + dt * ak_k * (ohz_k + ohz_kp);
    const double cf = 1.0 / (BC - FC * cf_prev);
    cf_prev = cf * CFk;
    dc_prev = cf * (q_kp - q_k - FC * dc_prev);
    if (act) tw(i, j, k) = q_k;
    q_k = q_kp; hz_k = hz_kp; ohz_k = ohz_kp; ak_km = ak_k; ak_k = ak_kp;
  }
  if (!act) return;                                   // no more block-wide synchronisation below
  // ---- sweep 2: per segment (top first) recompute CF,DC, back-substitute, update
  double dc_next = 0.0, a_next = dc_next * ak_k, q_next = q_k, ohz_next = ohz_k;
  const int nseg = (N - 1 + SEG - 1) / SEG;
  for (int s = nseg - 1; s >= 0; --s) {
    const int k0 = s * SEG + 1, k1 = min(k0 + SEG - 1, N - 1);
    double scf[SEG], sdc[SEG];
    double cfp = ck_cf[s], dcp = ck_dc[s];
    double h0 = Hzg(i, j, k0), o0 = 1.0 / h0, am = Aktg(i, j, k0 - 1), a0 = Aktg(i, j, k0), q0 = tw(i, j, k0);
#pragma unroll
    for (int kk = 0; kk < SEG; ++kk) {
      const int k = k0 + kk;
      if (k <= k1) {
        const double h1 = Hzg(i, j, k + 1), o1 = 1.0 / h1, a1 = Aktg(i, j, k + 1);
        const double q1 = (k + 1 == N) ? q_next : tw(i, j, k + 1);
        const double FC = (1.0 / 6.0) * h0 - dt * am * o0;
        const double CFk = (1.0 / 6.0) * h1 - dt * a1 * o1;
        const double BC = (1.0 / 3.0) * (h0 + h1) + dt * a0 * (o0 + o1);
        const double cf = 1.0 / (BC - FC * cfp);
        cfp = cf * CFk;
        dcp = cf * (q1 - q0 - FC * dcp);
        scf[kk] = cfp; sdc[kk] = dcp;
        h0 = h1; o0 = o1; am = a0; a0 = a1; q0 = q1;
      }
    }
#pragma unroll
    for (int kk = SEG - 1; kk >= 0; --kk) {
      const int k = k0 + kk;
      if (k <= k1) {
        const double dc_k = sdc[kk] - scf[kk] * dc_next;
        const double a_k = dc_k * Aktg(i, j, k);
        const double out = q_next + dt * ohz_next * (a_next - a_k);
        const double qk = tw(i, j, k);
        st(D, tw, i, j, k + 1, out);
        if (south) st(D, tw, i, j - 1, k + 1, out);
        if (north) st(D, tw, i, j + 1, k + 1, out);
        dc_next = dc_k; a_next = a_k; q_next = qk;
        ohz_next = 1.0 / Hzg(i, j, k);
      }
    }
  }
  {
    const double out = q_next + dt * ohz_next * (a_next - 0.0);
    st(D, tw, i, j, 1, out);
    if (south) st(D, tw, i, j - 1, 1, out);
    if (north) st(D, tw, i, j + 1, 1, out);
  }
}

// ---- host: tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: no -lcuda)
namespace {
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn g_encode = nullptr;
int make_map(CUtensorMap* m, double* base, int ni, int nj, int nk, int bw, int bh) {
  if (!g_encode) {
    void* fn = nullptr; cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) return 1;
    g_encode = (EncodeFn)fn;
  }
  const cuuint64_t gdim[3] = {(cuuint64_t)ni, (cuuint64_t)nj, (cuuint64_t)nk};
  const cuuint64_t gstr[2] = {(cuuint64_t)ni * 8, (cuuint64_t)ni * nj * 8};
  const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
  return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 0 : 1;
}
std::map<std::tuple<const void*, int, int>, Maps> g_maps;     // keyed by (context, nnew, itrc)
}  // namespace

// returns 0 on success, 2 if TMA cannot be used for this mirror (row pitch not a multiple of 16 B)
int k_step3d_t_v5(roms_b200_ctx* c, int nnew) {
  const Dev& D = c->D; const roms_b200_bounds& b = D.b;
  if ((D.ni & 1) || b.N < 4) return 2;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk(TI, TJ);
  dim3 g((bx.i1 - bx.i0 + TI) / TI, (bx.j1 - bx.j0 + TJ) / TJ, 1);
  static bool attr_done = false;
  const size_t smem = (size_t)NSTAGE * STAGE_BYTES;
  if (!attr_done) { CUDA_OK(cudaFuncSetAttribute(step3d_t_v5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_done = true; }
  for (int itrc = 1; itrc <= b.NT; ++itrc) {
    auto key = std::make_tuple((const void*)c, nnew, itrc);
    auto it = g_maps.find(key);
    if (it == g_maps.end()) {
      Maps M; const int N = b.N, ni = D.ni, nj = D.nj;
      const size_t vol = D.nij * N;
      double* t3 = D.f[FID(t)] + vol * ((3 - 1) + (size_t)3 * (itrc - 1));
      double* twp = D.f[FID(t)] + vol * ((nnew - 1) + (size_t)3 * (itrc - 1));
      double* ak = D.f[FID(Akt)] + D.nij * (N + 1) * (size_t)((itrc <= b.NAT ? itrc : b.NAT) - 1);
      int rc = make_map(&M.t3, t3, ni, nj, N, T3_W, T3_H) | make_map(&M.hu, D.f[FID(Huon)], ni, nj, N, HU_W, TJ) |
               make_map(&M.hv, D.f[FID(Hvom)], ni, nj, N, TI, HV_H) | make_map(&M.w, D.f[FID(W)], ni, nj, N + 1, TI, TJ) |
               make_map(&M.hz, D.f[FID(Hz)], ni, nj, N, TI, TJ) | make_map(&M.ak, ak, ni, nj, N + 1, TI, TJ) | make_map(&M.tw, twp, ni, nj, N, TI, TJ);
      if (rc) return 2;
      it = g_maps.emplace(key, M).first;
    }
    step3d_t_v5_kernel<<<g, blk, smem, c->stream>>>(c->D, bx, nnew, itrc, it->second);
    c->launches++;
  }
  return 0;
}
void k_step3d_t_v5_forget(roms_b200_ctx* c) {
  for (auto it = g_maps.begin(); it != g_maps.end();) { if (std::get<0>(it->first) == (const void*)c) it = g_maps.erase(it); else ++it; }
}
