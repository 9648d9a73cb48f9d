// roms_b200/csrc/k_step3d_t.cu -- tracer corrector step3d_t_tile (step3d_t.F:393-399,
// 641-916, 1150-1365, 1672-1721, 1858-1924), the roofline-graded kernel.
//
// One thread per water column, i fastest (a warp = one 256-byte i-stripe per
// (j,k) row).  Single pass over k:
//   * U3 horizontal + C4 vertical advection of t(3), divide by Hz  -> q(k)
//   * forward elimination of the spline tridiagonal system fused in the same
//     sweep with a one-level look-ahead (needs q(k+1)-q(k))
//   * back substitution and the final update in one descending sweep.
// The per-column recurrence arrays CF, DC, q live in SHARED memory laid out
// [k][thread] (conflict-free, no __syncthreads: every thread touches only its own
// column), so nothing spills to local memory / DRAM.  The east flux FX(i+1) is
// taken from lane+1 by warp shuffle instead of being recomputed.
// HBM traffic: t(3) 8 + t(nnew) 8+8 + Akt 8 per tracer-cell, Huon,Hvom,W,Hz 32 per
// cell = 96 B per cell at NT=2 (both tracers of a column are handled by the SAME
// thread back to back so Huon/Hvom/W/Hz rows are still in L1/L2 for the second).
// Arithmetic order per point == reference (-fmad=false): bit-identical results.
#include "common.cuh"

namespace {
struct Edges { int S, N, Jstr, Jend; };
__device__ __forceinline__ double dEta(const V3& q, int i, int j, int k, const Edges& e) {
  int jj = j;
  if (e.S && j == e.Jstr - 1) jj = e.Jstr;
  if (e.N && j == e.Jend + 2) jj = e.Jend + 1;
  return q(i, jj, k) - q(i, jj - 1, k);
}
__device__ __forceinline__ double fluxX_u3(const V3& q, const V3& Huon, int i, int j, int k) {
  const double qm2 = q(i - 2, j, k), qm1 = q(i - 1, j, k), q0 = q(i, j, k), qp1 = q(i + 1, j, k);
  const double d0 = qm1 - qm2, d1 = q0 - qm1, d2 = qp1 - q0;
  const double cm = d1 - d0, cp = d2 - d1, hu = Huon(i, j, k);
  return hu * 0.5 * (qm1 + q0) - (1.0 / 6.0) * (cm * fmax(hu, 0.0) + cp * fmin(hu, 0.0));
}
__device__ __forceinline__ double fluxE_u3(const V3& q, const V3& Hvom, int i, int j, int k, const Edges& e) {
  const double d0 = dEta(q, i, j - 1, k, e), d1 = dEta(q, i, j, k, e), d2 = dEta(q, i, j + 1, k, e);
  const double cm = d1 - d0, cp = d2 - d1, hv = Hvom(i, j, k);
  return hv * 0.5 * (q(i, j - 1, k) + q(i, j, k)) - (1.0 / 6.0) * (cm * fmax(hv, 0.0) + cp * fmin(hv, 0.0));
}
__device__ __forceinline__ double fluxZ_c4(const V3& q, const V3& W, int i, int j, int k, int N) {
  const double c1 = 0.5, c2 = 7.0 / 12.0, c3 = 1.0 / 12.0;
  if (k == 0 || k == N) return 0.0;
  if (k == 1) return W(i, j, 1) * (c1 * q(i, j, 1) + c2 * q(i, j, 2) - c3 * q(i, j, 3));
  if (k == N - 1) return W(i, j, N - 1) * (c1 * q(i, j, N) + c2 * q(i, j, N - 1) - c3 * q(i, j, N - 2));
  return W(i, j, k) * (c2 * (q(i, j, k) + q(i, j, k + 1)) - c3 * (q(i, j, k - 1) + q(i, j, k + 2)));
}
}  // namespace

#define S3T_BT 128          // threads per block (32 x 4)
// QSM: q(k) kept in shared memory (3 arrays) -- else staged through t(nnew) itself (2 arrays in smem)
template <bool QSM>
__global__ void __launch_bounds__(S3T_BT) step3d_t_v2_kernel(const Dev D, Box bx, int nnew) {
  extern __shared__ double sm[];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int N = D.b.N;
  double* sCF = sm; double* sDC = sm + (size_t)(N + 1) * S3T_BT; double* sQ = sDC + (size_t)(N + 1) * S3T_BT;
#define CFs(k) sCF[(k) * S3T_BT + tid]
#define DCs(k) sDC[(k) * S3T_BT + tid]
#define Qs(k) sQ[(k) * S3T_BT + tid]
  int i = bx.i0 + blockIdx.x * 32 + threadIdx.x;
  const int j = bx.j0 + blockIdx.y * blockDim.y + threadIdx.y;
  const bool act = (i <= bx.i1 && j <= bx.j1);
  const int lane = threadIdx.x;
  if (j > bx.j1) return;                 // whole warp leaves together (warp = one j row)
  if (!act) i = bx.i1;                   // idle lanes shadow the last column (keeps shuffles defined); they never store
  const double dt = D.p.dt;
  const Edges e{D.b.Southern_Edge && !D.b.NSperiodic, D.b.Northern_Edge && !D.b.NSperiodic, D.b.Jstr, D.b.Jend};
  const bool south = e.S && j == e.Jstr, north = e.N && j == e.Jend;
  V3 Hz = v3(D, FID(Hz)), Huon = v3(D, FID(Huon)), Hvom = v3(D, FID(Hvom)), W = v3(D, FID(W));
  const double cff = dt * v2(D, FID(pm))(i, j) * v2(D, FID(pn))(i, j);
  for (int itrc = 1; itrc <= D.b.NT; ++itrc) {
    V3 t3 = v3l(D, FID(t), 3, itrc), tw = v3l(D, FID(t), nnew, itrc), Akt = v3l(D, FID(Akt), min(D.b.NAT, itrc));
    // q(k): advected tracer divided by Hz.  adv(k) needs FC(k-1): rolling.
    double FCm = 0.0;
    auto advect = [&](int k, double& ohz) -> double {
      const double FXi = fluxX_u3(t3, Huon, i, j, k);
      double FXp = __shfl_down_sync(0xffffffffu, FXi, 1);
      if (lane == 31 || i == bx.i1) FXp = fluxX_u3(t3, Huon, i + 1, j, k);
      const double FEj = fluxE_u3(t3, Hvom, i, j, k, e), FEp = fluxE_u3(t3, Hvom, i, j + 1, k, e);
      const double c1 = cff * (FXp - FXi), c2 = cff * (FEp - FEj), c3 = c1 + c2;
      double tv = tw(i, j, k) - c3;
      const double FCk = fluxZ_c4(t3, W, i, j, k, N);
      const double cv = cff * (FCk - FCm);
      FCm = FCk;
      ohz = 1.0 / Hz(i, j, k);
      tv = tv - cv;
      return tv * ohz;
    };
    double ohz_k, ohz_kp;
    double q_k = advect(1, ohz_k), q_kp;
    double hz_k = Hz(i, j, 1), ak_km = Akt(i, j, 0), ak_k = Akt(i, j, 1);
    double cf_prev = 0.0, dc_prev = 0.0;      // CF(0), DC(0)
    for (int k = 1; k <= N - 1; ++k) {
      q_kp = advect(k + 1, ohz_kp);
      const double hz_kp = Hz(i, j, k + 1), ak_kp = Akt(i, j, k + 1);
      const double FC = (1.0 / 6.0) * hz_k - dt * ak_km * ohz_k;
      const double CFk = (1.0 / 6.0) * hz_kp - dt * ak_kp * ohz_kp;
      const double BC = (1.0 / 3.0) * (hz_k + hz_kp) + dt * ak_k * (ohz_k + ohz_kp);
      const double cf = 1.0 / (BC - FC * cf_prev);
      cf_prev = cf * CFk;
      dc_prev = cf * (q_kp - q_k - FC * dc_prev);
      CFs(k) = cf_prev; DCs(k) = dc_prev;
      if (QSM) Qs(k) = q_k; else if (act) tw(i, j, k) = q_k;
      q_k = q_kp; hz_k = hz_kp; ohz_k = ohz_kp; ak_km = ak_k; ak_k = ak_kp;
    }
    // descending: back substitution (DC(N)=0) fused with t += dt*oHz*(Akt(k)*DC(k) - Akt(k-1)*DC(k-1))
    double dc_next = 0.0;                               // DC(N)
    double a_next = dc_next * ak_k;                     // DC(N)*Akt(N)   (ak_k == Akt(N) here)
    double q_next = q_k, ohz_next = ohz_k;              // level N
    for (int k = N - 1; k >= 1; --k) {
      const double dc_k = DCs(k) - CFs(k) * dc_next;
      const double a_k = dc_k * Akt(i, j, k);
      const double out = q_next + dt * ohz_next * (a_next - a_k);
      if (act) {
        st(D, tw, i, j, k + 1, out);
        if (south) st(D, tw, i, j - 1, k + 1, out);     // t3dbc_im.F:334-341,415-422 closed walls
        if (north) st(D, tw, i, j + 1, k + 1, out);
      }
      dc_next = dc_k; a_next = a_k;
      q_next = QSM ? Qs(k) : tw(i, j, k);
      ohz_next = 1.0 / Hz(i, j, k);
    }
    {
      const double out = q_next + dt * ohz_next * (a_next - 0.0);   // DC(0)=0 is not scaled by Akt
      if (act) {
        st(D, tw, i, j, 1, out);
        if (south) st(D, tw, i, j - 1, 1, out);
        if (north) st(D, tw, i, j + 1, 1, out);
      }
    }
  }
#undef CFs
#undef DCs
#undef Qs
}

int k_step3d_t_v2(roms_b200_ctx* c, int nnew) {
  const roms_b200_bounds& b = c->D.b;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk(32, 4);
  dim3 g((bx.i1 - bx.i0 + 32) / 32, (bx.j1 - bx.j0 + 4) / 4, 1);
  const size_t col = (size_t)(b.N + 1) * S3T_BT * sizeof(double);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(step3d_t_v2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(step3d_t_v2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  if (3 * col <= 110 * 1024) step3d_t_v2_kernel<true><<<g, blk, 3 * col, c->stream>>>(c->D, bx, nnew);      // 2 blocks / SM
  else step3d_t_v2_kernel<false><<<g, blk, 2 * col, c->stream>>>(c->D, bx, nnew);
  c->launches++;
  return 0;
}
