// roms_b200/csrc/k_step3d_t3.cu -- step3d_t_tile, third layout (see k_step3d_t.cu for the math).
//
// Phase 1 (advection, no vertical recurrence): every thread sweeps its column with the k loop
//   unrolled so the read-only loads (through the non-coherent path, __ldg) of several levels are
//   in flight together; q(k), Hz(k) and Akt(k) are parked in shared memory [k][thread].
// Phase 2 (spline tridiagonal, sequential in k): forward elimination and back substitution run
//   entirely out of shared memory; CF(k)/DC(k) overwrite the Hz/Akt slots they have consumed.
// No __syncthreads: a thread only ever touches its own shared-memory column.
#include "common.cuh"

namespace {
struct RO3 {   // read-only 3-D view (ld.global.nc)
  const double* __restrict__ p; int LBi, ni, LBj, nj, LBk;
  __device__ __forceinline__ double operator()(int i, int j, int k) const {
    return __ldg(p + ((i - LBi) + (size_t)ni * ((j - LBj) + (size_t)nj * (k - LBk))));
  }
};
__device__ __forceinline__ RO3 ro(const V3& v) { return RO3{v.p, v.LBi, v.ni, v.LBj, v.nj, v.LBk}; }
struct Edges { int S, N, Jstr, Jend; };
__device__ __forceinline__ int jclamp(int j, const Edges& e) {     // FE(i,Jstr-1)=FE(i,Jstr), FE(i,Jend+2)=FE(i,Jend+1)
  if (e.S && j == e.Jstr - 1) return e.Jstr;
  if (e.N && j == e.Jend + 2) return e.Jend + 1;
  return j;
}
}  // namespace

#define T3_BT 128
__global__ void __launch_bounds__(T3_BT) step3d_t_v3_kernel(const Dev D, Box bx, int nnew) {
  extern __shared__ double sm[];
  const int tid = threadIdx.y * 32 + threadIdx.x, lane = threadIdx.x;
  const int N = D.b.N;
  double* sQ = sm; double* sA = sm + (size_t)(N + 1) * T3_BT; double* sB = sA + (size_t)(N + 1) * T3_BT;
#define Qs(k) sQ[(k) * T3_BT + tid]
#define As(k) sA[(k) * T3_BT + tid]      /* Hz(k)  -> CF(k) */
#define Bs(k) sB[(k) * T3_BT + tid]      /* Akt(k) -> DC(k) */
  int i = bx.i0 + blockIdx.x * 32 + threadIdx.x;
  const int j = bx.j0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (j > bx.j1) return;
  const bool act = (i <= bx.i1);
  if (!act) i = bx.i1;
  const double dt = D.p.dt;
  const Edges e{D.b.Southern_Edge && !D.b.NSperiodic, D.b.Northern_Edge && !D.b.NSperiodic, D.b.Jstr, D.b.Jend};
  const bool south = e.S && j == e.Jstr, north = e.N && j == e.Jend;
  const RO3 Hz = ro(v3(D, FID(Hz))), Huon = ro(v3(D, FID(Huon))), Hvom = ro(v3(D, FID(Hvom))), W = ro(v3(D, FID(W)));
  const double cff = dt * v2(D, FID(pm))(i, j) * v2(D, FID(pn))(i, j);
  const int jm2 = jclamp(j - 1, e) - 1, jm1a = jclamp(j - 1, e), j0b = jclamp(j, e) - 1, j0a = jclamp(j, e);
  const int jp1b = jclamp(j + 1, e) - 1, jp1a = jclamp(j + 1, e), jp2b = jclamp(j + 2, e) - 1, jp2a = jclamp(j + 2, e);
  const bool edgeX = (lane == 31) || (i == bx.i1);
  for (int itrc = 1; itrc <= D.b.NT; ++itrc) {
    const RO3 t3 = ro(v3l(D, FID(t), 3, itrc)), Akt = ro(v3l(D, FID(Akt), min(D.b.NAT, itrc)));
    V3 tw = v3l(D, FID(t), nnew, itrc);
    const RO3 twr = ro(tw);
    // ---------------- phase 1: q(k) = (t(nnew) - dt*pm*pn*div(F)) / Hz
    double tkm1 = t3(i, j, 1), tk = tkm1, tkp1 = t3(i, j, 2);      // rolling t3 column values for the C4 flux
    double FCm = 0.0;
#pragma unroll 4
    for (int k = 1; k <= N; ++k) {
      // x-direction: values i-2..i+1 (and i+2 only on the warp's east edge)
      const double qm2 = t3(i - 2, j, k), qm1 = t3(i - 1, j, k), q0 = tk, qp1 = t3(i + 1, j, k);
      const double hu = Huon(i, j, k);
      const double d0 = qm1 - qm2, d1 = q0 - qm1, d2 = qp1 - q0;
      const double FXi = hu * 0.5 * (qm1 + q0) - (1.0 / 6.0) * ((d1 - d0) * fmax(hu, 0.0) + (d2 - d1) * fmin(hu, 0.0));
      double FXp = __shfl_down_sync(0xffffffffu, FXi, 1);
      if (edgeX) {
        const double qp2 = t3(i + 2, j, k), hup = Huon(i + 1, j, k), d3 = qp2 - qp1;
        FXp = hup * 0.5 * (q0 + qp1) - (1.0 / 6.0) * ((d2 - d1) * fmax(hup, 0.0) + (d3 - d2) * fmin(hup, 0.0));
      }
      // y-direction: first differences with closed-wall clamping
      const double e_m1 = t3(i, jm1a, k) - t3(i, jm2, k), e_0 = t3(i, j0a, k) - t3(i, j0b, k);
      const double e_p1 = t3(i, jp1a, k) - t3(i, jp1b, k), e_p2 = t3(i, jp2a, k) - t3(i, jp2b, k);
      const double hv = Hvom(i, j, k), hvp = Hvom(i, j + 1, k);
      const double tjm = t3(i, j - 1, k), tjp = t3(i, j + 1, k);
      const double FEj = hv * 0.5 * (tjm + q0) - (1.0 / 6.0) * ((e_0 - e_m1) * fmax(hv, 0.0) + (e_p1 - e_0) * fmin(hv, 0.0));
      const double FEp = hvp * 0.5 * (q0 + tjp) - (1.0 / 6.0) * ((e_p1 - e_0) * fmax(hvp, 0.0) + (e_p2 - e_p1) * fmin(hvp, 0.0));
      const double c1 = cff * (FXp - FXi), c2 = cff * (FEp - FEj), c3 = c1 + c2;
      double tv = twr(i, j, k) - c3;
      // vertical C4 flux at w-level k (step3d_t.F:1150-1185)
      const double tkp2 = (k + 2 <= N) ? t3(i, j, k + 2) : 0.0;
      double FCk;
      if (k == N) FCk = 0.0;
      else if (k == 1) FCk = W(i, j, 1) * (0.5 * tk + (7.0 / 12.0) * tkp1 - (1.0 / 12.0) * tkp2);
      else if (k == N - 1) FCk = W(i, j, k) * (0.5 * tkp1 + (7.0 / 12.0) * tk - (1.0 / 12.0) * tkm1);
      else FCk = W(i, j, k) * ((7.0 / 12.0) * (tk + tkp1) - (1.0 / 12.0) * (tkm1 + tkp2));
      const double cv = cff * (FCk - FCm);
      FCm = FCk;
      const double hz = Hz(i, j, k);
      tv = tv - cv;
      Qs(k) = tv * (1.0 / hz);
      As(k) = hz; Bs(k) = Akt(i, j, k);
      tkm1 = tk; tk = tkp1; tkp1 = tkp2;
    }
    // ---------------- phase 2: spline implicit vertical diffusion (step3d_t.F:1672-1721)
    double hz_k = As(1), ak_km = Akt(i, j, 0), ak_k = Bs(1), q_k = Qs(1);
    double ohz_k = 1.0 / hz_k, cf_prev = 0.0, dc_prev = 0.0;
    for (int k = 1; k <= N - 1; ++k) {
      const double hz_kp = As(k + 1), ak_kp = Bs(k + 1), q_kp = Qs(k + 1);
      const double ohz_kp = 1.0 / hz_kp;
      const double FC = (1.0 / 6.0) * hz_k - dt * ak_km * ohz_k;
      const double CFk = (1.0 / 6.0) * hz_kp - dt * ak_kp * ohz_kp;
      const double BC = (1.0 / 3.0) * (hz_k + hz_kp) + dt * ak_k * (ohz_k + ohz_kp);
      const double cf = 1.0 / (BC - FC * cf_prev);
      cf_prev = cf * CFk;
      dc_prev = cf * (q_kp - q_k - FC * dc_prev);
      As(k) = cf_prev; Bs(k) = dc_prev;            // slots k hold CF(k), DC(k); slots > k still hold Hz, Akt
      q_k = q_kp; hz_k = hz_kp; ohz_k = ohz_kp; ak_km = ak_k; ak_k = ak_kp;
    }
    double dc_next = 0.0, a_next = dc_next * ak_k, q_next = q_k, ohz_next = ohz_k;   // level N
    for (int k = N - 1; k >= 1; --k) {
      const double dc_k = Bs(k) - As(k) * dc_next;
      const double a_k = dc_k * Akt(i, j, k);
      const double out = q_next + dt * ohz_next * (a_next - a_k);
      if (act) {
        st(D, tw, i, j, k + 1, out);
        if (south) st(D, tw, i, j - 1, k + 1, out);
        if (north) st(D, tw, i, j + 1, k + 1, out);
      }
      dc_next = dc_k; a_next = a_k; q_next = Qs(k);
      ohz_next = 1.0 / Hz(i, j, k);
    }
    {
      const double out = q_next + dt * ohz_next * (a_next - 0.0);
      if (act) {
        st(D, tw, i, j, 1, out);
        if (south) st(D, tw, i, j - 1, 1, out);
        if (north) st(D, tw, i, j + 1, 1, out);
      }
    }
  }
#undef Qs
#undef As
#undef Bs
}

int k_step3d_t_v3(roms_b200_ctx* c, int nnew) {
  const roms_b200_bounds& b = c->D.b;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk(32, 4);
  dim3 g((bx.i1 - bx.i0 + 32) / 32, (bx.j1 - bx.j0 + 4) / 4, 1);
  const size_t bytes = 3 * (size_t)(b.N + 1) * T3_BT * sizeof(double);
  static bool attr_done = false;
  if (!attr_done) { cudaFuncSetAttribute(step3d_t_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024); attr_done = true; }
  step3d_t_v3_kernel<<<g, blk, bytes, c->stream>>>(c->D, bx, nnew);
  c->launches++;
  return 0;
}
