// roms_b200/csrc/k_step3d_t4.cu -- step3d_t_tile, production layout.
//
// One thread per water column (i fastest: a warp = one 256-byte i-stripe per (j,k) row),
// high occupancy, NO shared memory and only 256 B of thread-private storage:
//   sweep 1 (k ascending): U3/C4 advection -> q(k) (parked in t(nnew) itself), fused with the
//            forward elimination of the spline tridiagonal system; only every 8th (CF,DC) pair is
//            kept (a "checkpoint"), the rest is recomputed later;
//   sweep 2 (segments of 8 levels, descending): re-run the forward elimination inside the segment
//            from its checkpoint (same operations on the same inputs -> same bits), then back
//            substitution and the final update t += dt/Hz * d(Akt*DC).
// This replaces the 4 x N-double private arrays of the first version (2 KB/thread, which spilled
// through L2 to DRAM) by 4 x 8 doubles that stay in L1.  FX(i+1) comes from lane+1 by shuffle.
// Arithmetic order per point == reference (step3d_t.F:641-916,1150-1365,1672-1721), -fmad=false.
#include "common.cuh"
#include <cstdlib>

namespace {
struct RO3 {   // read-only 3-D view (ld.global.nc); only for arrays this kernel never writes
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
constexpr int SEG = 8;
constexpr int MAXSEG = RB_MAXN / SEG;
}  // namespace

template <int MINB>
__global__ void __launch_bounds__(256, MINB) step3d_t_v4_kernel(const Dev D, Box bx, int nnew) {
  const int lane = threadIdx.x, N = D.b.N;
  int i = bx.i0 + blockIdx.x * 32 + threadIdx.x;
  const int j = bx.j0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (j > bx.j1) return;                          // a warp is one j row: leaves together
  const bool act = (i <= bx.i1);
  if (!act) i = bx.i1;                            // idle lanes shadow the last column, never store
  const int itrc = 1 + blockIdx.z;
  const double dt = D.p.dt;
  const Edges e{D.b.Southern_Edge && !D.b.NSperiodic, D.b.Northern_Edge && !D.b.NSperiodic, D.b.Jstr, D.b.Jend};
  const bool south = e.S && j == e.Jstr, north = e.N && j == e.Jend;
  const RO3 Hz = ro(v3(D, FID(Hz))), Huon = ro(v3(D, FID(Huon))), Hvom = ro(v3(D, FID(Hvom))), W = ro(v3(D, FID(W)));
  const RO3 t3 = ro(v3l(D, FID(t), 3, itrc)), Akt = ro(v3l(D, FID(Akt), min(D.b.NAT, itrc)));
  V3 tw = v3l(D, FID(t), nnew, itrc);
  const double cff = dt * v2(D, FID(pm))(i, j) * v2(D, FID(pn))(i, j);
  const int jm2 = jclamp(j - 1, e) - 1, jm1a = jclamp(j - 1, e), j0b = jclamp(j, e) - 1, j0a = jclamp(j, e);
  const int jp1b = jclamp(j + 1, e) - 1, jp1a = jclamp(j + 1, e), jp2b = jclamp(j + 2, e) - 1, jp2a = jclamp(j + 2, e);
  const bool edgeX = (lane == 31) || (i == bx.i1);

  double tkm1 = t3(i, j, 1), tk = tkm1, tkp1 = t3(i, j, 2), FCm = 0.0;
  // q(k) = (t(nnew) - dt*pm*pn*(div_h F + d_k FC)) / Hz ; returns Hz(k) through hz
  auto advect = [&](int k, double& hz) -> double {
    const double qm2 = t3(i - 2, j, k), qm1 = t3(i - 1, j, k), q0 = tk, qp1 = t3(i + 1, j, k);
    const double hu = Huon(i, j, k);
    const double d0 = qm1 - qm2, d1 = q0 - qm1, d2 = qp1 - q0;
    const double FXi = hu * 0.5 * (qm1 + q0) - (1.0 / 6.0) * ((d1 - d0) * fmax(hu, 0.0) + (d2 - d1) * fmin(hu, 0.0));
    double FXp = __shfl_down_sync(0xffffffffu, FXi, 1);
    if (edgeX) {
      const double qp2 = t3(i + 2, j, k), hup = Huon(i + 1, j, k), d3 = qp2 - qp1;
      FXp = hup * 0.5 * (q0 + qp1) - (1.0 / 6.0) * ((d2 - d1) * fmax(hup, 0.0) + (d3 - d2) * fmin(hup, 0.0));
    }
    const double e_m1 = t3(i, jm1a, k) - t3(i, jm2, k), e_0 = t3(i, j0a, k) - t3(i, j0b, k);
    const double e_p1 = t3(i, jp1a, k) - t3(i, jp1b, k), e_p2 = t3(i, jp2a, k) - t3(i, jp2b, k);
    const double hv = Hvom(i, j, k), hvp = Hvom(i, j + 1, k);
    const double tjm = t3(i, j - 1, k), tjp = t3(i, j + 1, k);
    const double FEj = hv * 0.5 * (tjm + q0) - (1.0 / 6.0) * ((e_0 - e_m1) * fmax(hv, 0.0) + (e_p1 - e_0) * fmin(hv, 0.0));
    const double FEp = hvp * 0.5 * (q0 + tjp) - (1.0 / 6.0) * ((e_p1 - e_0) * fmax(hvp, 0.0) + (e_p2 - e_p1) * fmin(hvp, 0.0));
    const double c1 = cff * (FXp - FXi), c2 = cff * (FEp - FEj), c3 = c1 + c2;
    double tv = tw(i, j, k) - c3;
    const double tkp2 = (k + 2 <= N) ? t3(i, j, k + 2) : 0.0;
    double FCk;
    if (k == N) FCk = 0.0;
    else if (k == 1) FCk = W(i, j, 1) * (0.5 * tk + (7.0 / 12.0) * tkp1 - (1.0 / 12.0) * tkp2);
    else if (k == N - 1) FCk = W(i, j, k) * (0.5 * tkp1 + (7.0 / 12.0) * tk - (1.0 / 12.0) * tkm1);
    else FCk = W(i, j, k) * ((7.0 / 12.0) * (tk + tkp1) - (1.0 / 12.0) * (tkm1 + tkp2));
    const double cv = cff * (FCk - FCm);
    FCm = FCk;
    hz = Hz(i, j, k);
    tv = tv - cv;
    tkm1 = tk; tk = tkp1; tkp1 = tkp2;
    return tv * (1.0 / hz);
  };

  // ---- sweep 1: advection + forward elimination, keeping one (CF,DC) checkpoint per segment
  double ck_cf[MAXSEG], ck_dc[MAXSEG];
  double hz_k, hz_kp;
  double q_k = advect(1, hz_k), q_kp;
  double ohz_k = 1.0 / hz_k, ak_km = Akt(i, j, 0), ak_k = Akt(i, j, 1), cf_prev = 0.0, dc_prev = 0.0;
  for (int k = 1; k <= N - 1; ++k) {
    if (((k - 1) & (SEG - 1)) == 0) { ck_cf[(k - 1) / SEG] = cf_prev; ck_dc[(k - 1) / SEG] = dc_prev; }
    q_kp = advect(k + 1, hz_kp);
    const double ohz_kp = 1.0 / hz_kp, ak_kp = Akt(i, j, k + 1);
    const double FC = (1.0 / 6.0) * hz_k - dt * ak_km * ohz_k;
    const double CFk = (1.0 / 6.0) * hz_kp - dt * ak_kp * ohz_kp;
    const double BC = (1.0 / 3.0) * (hz_k + hz_kp) + dt * ak_k * (ohz_k + ohz_kp);
    const double cf = 1.0 / (BC - FC * cf_prev);
    cf_prev = cf * CFk;
    dc_prev = cf * (q_kp - q_k - FC * dc_prev);
    if (act) tw(i, j, k) = q_k;                   // park q(k); overwritten by the final value in sweep 2
    q_k = q_kp; hz_k = hz_kp; ohz_k = ohz_kp; ak_km = ak_k; ak_k = ak_kp;
  }
  // ---- sweep 2: per segment (top first) recompute CF,DC, then back-substitute and update
  double dc_next = 0.0;                           // DC(N)
  double a_next = dc_next * ak_k;                 // DC(N)*Akt(N)
  double q_next = q_k, ohz_next = ohz_k;          // level N (q(N) was never parked)
  const int nseg = (N - 1 + SEG - 1) / SEG;
  for (int s = nseg - 1; s >= 0; --s) {
    const int k0 = s * SEG + 1, k1 = min(k0 + SEG - 1, N - 1);
    double scf[SEG], sdc[SEG];
    double cfp = ck_cf[s], dcp = ck_dc[s];
    double h0 = Hz(i, j, k0), o0 = 1.0 / h0, am = Akt(i, j, k0 - 1), a0 = Akt(i, j, k0), q0 = tw(i, j, k0);
#pragma unroll
    for (int kk = 0; kk < SEG; ++kk) {
      const int k = k0 + kk;
      if (k <= k1) {
        const double h1 = Hz(i, j, k + 1), o1 = 1.0 / h1, a1 = Akt(i, j, k + 1);
        const double q1 = (k + 1 == N) ? q_next : tw(i, j, k + 1);      // q(N) lives in a register
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
        const double a_k = dc_k * Akt(i, j, k);
        const double out = q_next + dt * ohz_next * (a_next - a_k);
        const double qk = tw(i, j, k);                                  // read q(k) before level k+1.. is finalised
        if (act) {
          st(D, tw, i, j, k + 1, out);
          if (south) st(D, tw, i, j - 1, k + 1, out);                   // t3dbc_im.F:334-341,415-422
          if (north) st(D, tw, i, j + 1, k + 1, out);
        }
        dc_next = dc_k; a_next = a_k; q_next = qk;
        ohz_next = 1.0 / Hz(i, j, k);
      }
    }
  }
  {
    const double out = q_next + dt * ohz_next * (a_next - 0.0);         // DC(0)=0 is not scaled by Akt
    if (act) {
      st(D, tw, i, j, 1, out);
      if (south) st(D, tw, i, j - 1, 1, out);
      if (north) st(D, tw, i, j + 1, 1, out);
    }
  }
}

int k_step3d_t_v4(roms_b200_ctx* c, int nnew) {
  const roms_b200_bounds& b = c->D.b;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk(32, 8);
  dim3 g((bx.i1 - bx.i0 + 32) / 32, (bx.j1 - bx.j0 + 8) / 8, b.NT);
  static const int minb = getenv("ROMS_B200_S3T_MINB") ? atoi(getenv("ROMS_B200_S3T_MINB")) : 3;
  if (minb == 2) step3d_t_v4_kernel<2><<<g, blk, 0, c->stream>>>(c->D, bx, nnew);
  else if (minb == 4) step3d_t_v4_kernel<4><<<g, blk, 0, c->stream>>>(c->D, bx, nnew);
  else step3d_t_v4_kernel<3><<<g, blk, 0, c->stream>>>(c->D, bx, nnew);
  c->launches++;
  return 0;
}
