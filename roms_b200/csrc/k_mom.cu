// roms_b200/csrc/k_mom.cu -- baroclinic pressure gradient (prsgrd32), 3-D momentum
// right-hand side (rhs3d_tile), harmonic viscosity (uv3dmix2_s) and the momentum
// corrector (step3d_uv).  Column-marching threads, i fastest (coalesced rows).
#include "common.cuh"
#include <algorithm>

// ---- prsgrd32_tile, prsgrd32.h:238-433 -----------------------------------------
// pass 1: column pressure P(i,j,k) with harmonic-mean spline slopes dR,dZ.  Only the running sum over k is a recurrence:
// prsgrd_T_kernel (one thread per (i,j,k)) evaluates the increment of level k -- the harmonic means of levels k and k+1 from
// the raw differences k-1..k+1, exactly as the in-place descending sweep of prsgrd32.h:258-271 sees them -- and parks it in
// P(i,j,k); prsgrd_P_kernel (one thread per column) then adds the increments from the surface down (:279-305).
// (One column kernel doing both took 43 us on BENCHMARK1: 32 k threads, two dependent sweeps of L2 loads.)
__device__ __forceinline__ double prs_raw(const V3& a, int i, int j, int m, int N) {     // a(m+1)-a(m), with dX(N)=dX(N-1), dX(0)=dX(1)
  const int mm = (m >= N) ? N - 1 : (m <= 0 ? 1 : m);
  return a(i, j, mm + 1) - a(i, j, mm);
}
__global__ void __launch_bounds__(256) prsgrd_T_kernel(const Dev D, Box bx) {
  IJ_FROM_BOX(bx);
  const int N = D.b.N, k = 1 + blockIdx.z; const double g = D.p.g, GRho = g / D.p.rho0, HalfGRho = 0.5 * GRho;
  const double OneFifth = 0.2, OneTwelfth = 1.0 / 12.0, eps = 1.0e-10;
  V3 rho = v3(D, FID(rho)), z_r = v3(D, FID(z_r)), z_w = v3(D, FID(z_w));
  V3 P{D.P, D.b.LBi, D.ni, D.b.LBj, D.nj, 1};
  if (k == N) {
    const double zwN = z_w(i, j, N), zrN = z_r(i, j, N), rN = rho(i, j, N);
    const double cff1 = 1.0 / (zrN - z_r(i, j, N - 1));
    const double cff2 = 0.5 * (rN - rho(i, j, N - 1)) * (zwN - zrN) * cff1;
    P(i, j, N) = g * zwN + GRho * (rN + cff2) * (zwN - zrN);
    return;
  }
  // raw differences at k-1, k, k+1 and the harmonic means at k, k+1
  const double rm = prs_raw(rho, i, j, k - 1, N), r0 = prs_raw(rho, i, j, k, N), rp = prs_raw(rho, i, j, k + 1, N);
  const double zm = prs_raw(z_r, i, j, k - 1, N), z0 = prs_raw(z_r, i, j, k, N), zp = prs_raw(z_r, i, j, k + 1, N);
  const double cK = 2.0 * r0 * rm, cP = 2.0 * rp * r0;
  const double dRk = (cK > eps) ? cK / (r0 + rm) : 0.0, dRp = (cP > eps) ? cP / (rp + r0) : 0.0;
  const double dZk = 2.0 * z0 * zm / (z0 + zm), dZp = 2.0 * zp * z0 / (zp + z0);
  const double rk1 = rho(i, j, k + 1), rk = rho(i, j, k), zk1 = z_r(i, j, k + 1), zk = z_r(i, j, k);
  P(i, j, k) = HalfGRho * ((rk1 + rk) * (zk1 - zk) -
                           OneFifth * ((dRp - dRk) * (zk1 - zk - OneTwelfth * (dZp + dZk)) -
                                       (dZp - dZk) * (rk1 - rk - OneTwelfth * (dRp + dRk))));
}
__global__ void __launch_bounds__(256) prsgrd_P_kernel(const Dev D, Box bx) {
  IJ_FROM_BOX(bx);
  const int N = D.b.N;
  V3 P{D.P, D.b.LBi, D.ni, D.b.LBj, D.nj, 1};
  double Pk = P(i, j, N);
  constexpr int KB = 8;
  for (int k0 = N - 1; k0 >= 1; k0 -= KB) {
    double t[KB];
#pragma unroll
    for (int q = 0; q < KB; ++q) t[q] = P(i, j, max(k0 - q, 1));
#pragma unroll
    for (int q = 0; q < KB; ++q) { const int k = k0 - q; if (k >= 1) { Pk = Pk + t[q]; P(i, j, k) = Pk; } }
  }
}
// harmonic-mean horizontal slopes (prsgrd32.h:319-336, 383-400)
__device__ __forceinline__ void hslope(double a0, double a1, double f0, double f1, double& dZx, double& dRx) {
  const double eps = 1.0e-10;
  const double cff = 2.0 * a0 * a1;
  if (cff > eps) { const double c1 = 1.0 / (a0 + a1); dZx = cff * c1; } else dZx = 0.0;
  const double cf1 = 2.0 * f0 * f1;
  if (cf1 > eps) { const double c2 = 1.0 / (f0 + f1); dRx = cf1 * c2; } else dRx = 0.0;
}
// pass 2: ru,rv(:,:,k,nrhs) = pressure-gradient term
__global__ void __launch_bounds__(256) prsgrd_ruv_kernel(const Dev D, Box bx, int nrhs) {
  IJ_FROM_BOX(bx);
  const roms_b200_bounds& b = D.b; const int k = 1 + blockIdx.z;
  const double HalfGRho = 0.5 * (D.p.g / D.p.rho0), OneFifth = 0.2, OneTwelfth = 1.0 / 12.0;
  V3 rho = v3(D, FID(rho)), z_r = v3(D, FID(z_r)), Hz = v3(D, FID(Hz));
  V3 P{D.P, D.b.LBi, D.ni, D.b.LBj, D.nj, 1};
  if (i >= b.IstrU && i <= b.Iend && j >= b.Jstr && j <= b.Jend) {
    V3 ru = v3l(D, FID(ru), nrhs);
    const double am = z_r(i - 1, j, k) - z_r(i - 2, j, k), a0 = z_r(i, j, k) - z_r(i - 1, j, k), ap = z_r(i + 1, j, k) - z_r(i, j, k);
    const double fm = rho(i - 1, j, k) - rho(i - 2, j, k), f0 = rho(i, j, k) - rho(i - 1, j, k), fp = rho(i + 1, j, k) - rho(i, j, k);
    double dZm, dRm, dZ0, dR0;
    hslope(am, a0, fm, f0, dZm, dRm);     // dZx(i-1), dRx(i-1)
    hslope(a0, ap, f0, fp, dZ0, dR0);     // dZx(i),   dRx(i)
    ru(i, j, k) = v2(D, FID(on_u))(i, j) * 0.5 * (Hz(i, j, k) + Hz(i - 1, j, k)) *
                  (P(i - 1, j, k) - P(i, j, k) -
                   HalfGRho * ((rho(i, j, k) + rho(i - 1, j, k)) * (z_r(i, j, k) - z_r(i - 1, j, k)) -
                               OneFifth * ((dR0 - dRm) * (z_r(i, j, k) - z_r(i - 1, j, k) - OneTwelfth * (dZ0 + dZm)) -
                                           (dZ0 - dZm) * (rho(i, j, k) - rho(i - 1, j, k) - OneTwelfth * (dR0 + dRm)))));
  }
  if (i >= b.Istr && i <= b.Iend && j >= b.JstrV && j <= b.Jend) {
    V3 rv = v3l(D, FID(rv), nrhs);
    const double am = z_r(i, j - 1, k) - z_r(i, j - 2, k), a0 = z_r(i, j, k) - z_r(i, j - 1, k), ap = z_r(i, j + 1, k) - z_r(i, j, k);
    const double fm = rho(i, j - 1, k) - rho(i, j - 2, k), f0 = rho(i, j, k) - rho(i, j - 1, k), fp = rho(i, j + 1, k) - rho(i, j, k);
    double dZm, dRm, dZ0, dR0;
    hslope(am, a0, fm, f0, dZm, dRm);
    hslope(a0, ap, f0, fp, dZ0, dR0);
    rv(i, j, k) = v2(D, FID(om_v))(i, j) * 0.5 * (Hz(i, j, k) + Hz(i, j - 1, k)) *
                  (P(i, j - 1, k) - P(i, j, k) -
                   HalfGRho * ((rho(i, j, k) + rho(i, j - 1, k)) * (z_r(i, j, k) - z_r(i, j - 1, k)) -
                               OneFifth * ((dR0 - dRm) * (z_r(i, j, k) - z_r(i, j - 1, k) - OneTwelfth * (dZ0 + dZm)) -
                                           (dZ0 - dZm) * (rho(i, j, k) - rho(i, j - 1, k) - OneTwelfth * (dR0 + dRm)))));
  }
}
int k_prsgrd(roms_b200_ctx* c, int nrhs) {
  const roms_b200_bounds& b = c->D.b;
  Box bp{b.IstrU - 1, b.Iend, b.JstrV - 1, b.Jend}; dim3 blk(32, 8);
  if (b.N < 2) return 1;
  { dim3 blkT(128, 2); dim3 gT = grid2(bp, blkT); gT.z = b.N; prsgrd_T_kernel<<<gT, blkT, 0, c->stream>>>(c->D, bp); c->launches++; }
  prsgrd_P_kernel<<<grid2(bp, blk), blk, 0, c->stream>>>(c->D, bp); c->launches++;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk2(128, 2); dim3 g = grid2(bx, blk2); g.z = b.N;
  prsgrd_ruv_kernel<<<g, blk2, 0, c->stream>>>(c->D, bx, nrhs); c->launches++;
  return 0;
}

// ---- rhs3d_tile, rhs3d.F:498-1919 ----------------------------------------------
struct RQ { V3 u, v, Hz, Huon, Hvom, W; V2 fomn, dndx, dmde; int curv; int S, N, Jstr, Jend; };
// second differences with the closed-wall replacements of rhs3d.F:793-806,909-922
template <class RQT> __device__ __forceinline__ double q_uee(const RQT& Q, int i, int j, int k) {
  int jj = j; if (Q.S && j == Q.Jstr - 1) jj = Q.Jstr; if (Q.N && j == Q.Jend + 1) jj = Q.Jend;
  return Q.u(i, jj - 1, k) - 2.0 * Q.u(i, jj, k) + Q.u(i, jj + 1, k);
}
template <class RQT> __device__ __forceinline__ double q_vee(const RQT& Q, int i, int j, int k) {
  int jj = j; if (Q.S && j == Q.Jstr) jj = Q.Jstr + 1; if (Q.N && j == Q.Jend + 1) jj = Q.Jend;
  return Q.v(i, jj - 1, k) - 2.0 * Q.v(i, jj, k) + Q.v(i, jj + 1, k);
}
template <class RQT> __device__ __forceinline__ double q_Hvee(const RQT& Q, int i, int j, int k) {
  int jj = j; if (Q.S && j == Q.Jstr) jj = Q.Jstr + 1; if (Q.N && j == Q.Jend + 1) jj = Q.Jend;
  return Q.Hvom(i, jj - 1, k) - 2.0 * Q.Hvom(i, jj, k) + Q.Hvom(i, jj + 1, k);
}
template <class RQT> __device__ __forceinline__ double q_uxx(const RQT& Q, int i, int j, int k) { return Q.u(i - 1, j, k) - 2.0 * Q.u(i, j, k) + Q.u(i + 1, j, k); }
template <class RQT> __device__ __forceinline__ double q_Huxx(const RQT& Q, int i, int j, int k) { return Q.Huon(i - 1, j, k) - 2.0 * Q.Huon(i, j, k) + Q.Huon(i + 1, j, k); }
template <class RQT> __device__ __forceinline__ double q_vxx(const RQT& Q, int i, int j, int k) { return Q.v(i - 1, j, k) - 2.0 * Q.v(i, j, k) + Q.v(i + 1, j, k); }
template <class RQT> __device__ __forceinline__ double q_Hvxx(const RQT& Q, int i, int j, int k) { return Q.Hvom(i - 1, j, k) - 2.0 * Q.Hvom(i, j, k) + Q.Hvom(i + 1, j, k); }
template <class RQT> __device__ __forceinline__ double q_Huee(const RQT& Q, int i, int j, int k) { return Q.Huon(i, j - 1, k) - 2.0 * Q.Huon(i, j, k) + Q.Huon(i, j + 1, k); }
#define GADV (-0.25)
// UFx at rho-point (i,j): rhs3d.F:767-784
template <class RQT> __device__ __forceinline__ double q_UFx(const RQT& Q, int i, int j, int k) {
  const double c1 = Q.u(i, j, k) + Q.u(i + 1, j, k);
  const double c = (c1 > 0.0) ? q_uxx(Q, i, j, k) : q_uxx(Q, i + 1, j, k);
  return 0.25 * (c1 + GADV * c) * (Q.Huon(i, j, k) + Q.Huon(i + 1, j, k) + GADV * 0.5 * (q_Huxx(Q, i, j, k) + q_Huxx(Q, i + 1, j, k)));
}
// UFe at psi-point (i,j): rhs3d.F:822-840
template <class RQT> __device__ __forceinline__ double q_UFe(const RQT& Q, int i, int j, int k) {
  const double c1 = Q.u(i, j, k) + Q.u(i, j - 1, k), c2 = Q.Hvom(i, j, k) + Q.Hvom(i - 1, j, k);
  const double c = (c2 > 0.0) ? q_uee(Q, i, j - 1, k) : q_uee(Q, i, j, k);
  return 0.25 * (c1 + GADV * c) * (c2 + GADV * 0.5 * (q_Hvxx(Q, i, j, k) + q_Hvxx(Q, i - 1, j, k)));
}
// VFx at psi-point (i,j): rhs3d.F:871-889
template <class RQT> __device__ __forceinline__ double q_VFx(const RQT& Q, int i, int j, int k) {
  const double c1 = Q.v(i, j, k) + Q.v(i - 1, j, k), c2 = Q.Huon(i, j, k) + Q.Huon(i, j - 1, k);
  const double c = (c2 > 0.0) ? q_vxx(Q, i - 1, j, k) : q_vxx(Q, i, j, k);
  return 0.25 * (c1 + GADV * c) * (c2 + GADV * 0.5 * (q_Huee(Q, i, j, k) + q_Huee(Q, i, j - 1, k)));
}
// VFe at rho-point (i,j): rhs3d.F:924-942
template <class RQT> __device__ __forceinline__ double q_VFe(const RQT& Q, int i, int j, int k) {
  const double c1 = Q.v(i, j, k) + Q.v(i, j + 1, k);
  const double c = (c1 > 0.0) ? q_vee(Q, i, j, k) : q_vee(Q, i, j + 1, k);
  return 0.25 * (c1 + GADV * c) * (Q.Hvom(i, j, k) + Q.Hvom(i, j + 1, k) + GADV * 0.5 * (q_Hvee(Q, i, j, k) + q_Hvee(Q, i, j + 1, k)));
}
// Coriolis and curvilinear terms at rho-point (rhs3d.F:506-512, 570-586): returns (UFx-like, VFe-like)
__device__ __forceinline__ void q_cor(const RQ& Q, int i, int j, int k, double& cu, double& cv) {
  const double cff = 0.5 * Q.Hz(i, j, k) * Q.fomn(i, j);
  cu = cff * (Q.v(i, j, k) + Q.v(i, j + 1, k)); cv = cff * (Q.u(i, j, k) + Q.u(i + 1, j, k));
}
__device__ __forceinline__ void q_curv(const RQ& Q, int i, int j, int k, double& cu, double& cv) {
  const double c1 = 0.5 * (Q.v(i, j, k) + Q.v(i, j + 1, k)), c2 = 0.5 * (Q.u(i, j, k) + Q.u(i + 1, j, k));
  const double c3 = c1 * Q.dndx(i, j), c4 = c2 * Q.dmde(i, j);
  const double cff = Q.Hz(i, j, k) * (c3 - c4);
  cu = cff * c1; cv = cff * c2;
}
// fourth-order centred vertical momentum flux at w-level k (rhs3d.F:1133-1170, 1283-1320); (di,dj) = staggering
__device__ __forceinline__ double q_FCw(const V3& q, const V3& W, int i, int j, int k, int N, int di, int dj) {
  const double c1 = 9.0 / 16.0, c2 = 1.0 / 16.0;
  if (k == 0 || k == N) return 0.0;
  const double wt = (c1 * (W(i, j, k) + W(i - di, j - dj, k)) - c2 * (W(i + di, j + dj, k) + W(i - 2 * di, j - 2 * dj, k)));
  if (k == 1) return (c1 * (q(i, j, 1) + q(i, j, 2)) - c2 * (q(i, j, 1) + q(i, j, 3))) * wt;
  if (k == N - 1) return (c1 * (q(i, j, N - 1) + q(i, j, N)) - c2 * (q(i, j, N - 2) + q(i, j, N))) * wt;
  return (c1 * (q(i, j, k) + q(i, j, k + 1)) - c2 * (q(i, j, k - 1) + q(i, j, k + 2))) * wt;
}
// One thread per (i,j,k,component): Coriolis, curvilinear, U3 horizontal and C4 vertical advection have no vertical recurrence
// (the vertical flux at w-level k-1 is re-evaluated).  The vertical integrals rufrc/rvfrc (rhs3d.F:1707-1916) are summed in
// the reference's order by rhs3d_sum_kernel (one thread per column and component).
__global__ void __launch_bounds__(256) rhs3d_kernel(const Dev D, Box bx, int nrhs) {
  IJ_FROM_BOX(bx);
  const roms_b200_bounds& b = D.b; const int N = b.N, k = 1 + blockIdx.z % N, comp = blockIdx.z / N;
  RQ Q{v3l(D, FID(u), nrhs), v3l(D, FID(v), nrhs), v3(D, FID(Hz)), v3(D, FID(Huon)), v3(D, FID(Hvom)), v3(D, FID(W)),
       v2(D, FID(fomn)), v2(D, FID(dndx)), v2(D, FID(dmde)), D.p.app == ROMS_B200_APP_BENCHMARK,
       b.Southern_Edge && !b.NSperiodic, b.Northern_Edge && !b.NSperiodic, b.Jstr, b.Jend};
  if (comp == 0) {
    if (!(i >= b.IstrU && i <= b.Iend && j >= b.Jstr && j <= b.Jend)) return;
    V3 ru = v3l(D, FID(ru), nrhs);
    double r = ru(i, j, k), a0, a1, dmy;
    q_cor(Q, i, j, k, a0, dmy); q_cor(Q, i - 1, j, k, a1, dmy);
    r = r + 0.5 * (a0 + a1);
    if (Q.curv) { q_curv(Q, i, j, k, a0, dmy); q_curv(Q, i - 1, j, k, a1, dmy); r = r + 0.5 * (a0 + a1); }
    const double c1 = q_UFx(Q, i, j, k) - q_UFx(Q, i - 1, j, k), c2 = q_UFe(Q, i, j + 1, k) - q_UFe(Q, i, j, k);
    r = r - (c1 + c2);
    const double FCk = q_FCw(Q.u, Q.W, i, j, k, N, 1, 0), FCm = q_FCw(Q.u, Q.W, i, j, k - 1, N, 1, 0);
    r = r - (FCk - FCm);
    ru(i, j, k) = r;
  } else {
    if (!(i >= b.Istr && i <= b.Iend && j >= b.JstrV && j <= b.Jend)) return;
    V3 rv = v3l(D, FID(rv), nrhs);
    double r = rv(i, j, k), a0, a1, dmy;
    q_cor(Q, i, j, k, dmy, a0); q_cor(Q, i, j - 1, k, dmy, a1);
    r = r - 0.5 * (a0 + a1);
    if (Q.curv) { q_curv(Q, i, j, k, dmy, a0); q_curv(Q, i, j - 1, k, dmy, a1); r = r - 0.5 * (a0 + a1); }
    const double c1 = q_VFx(Q, i + 1, j, k) - q_VFx(Q, i, j, k), c2 = q_VFe(Q, i, j, k) - q_VFe(Q, i, j - 1, k);
    r = r - (c1 + c2);
    const double FCk = q_FCw(Q.v, Q.W, i, j, k, N, 0, 1), FCm = q_FCw(Q.v, Q.W, i, j, k - 1, N, 0, 1);
    r = r - (FCk - FCm);
    rv(i, j, k) = r;
  }
}
__global__ void __launch_bounds__(128) rhs3d_sum_kernel(const Dev D, Box bx, int nrhs) {
  IJ_FROM_BOX(bx);
  const roms_b200_bounds& b = D.b; const int N = b.N, comp = blockIdx.z;
  if (comp == 0) {
    if (!(i >= b.IstrU && i <= b.Iend && j >= b.Jstr && j <= b.Jend)) return;
    V3 ru = v3l(D, FID(ru), nrhs);
    double sum = ru(i, j, 1);
    for (int k = 2; k <= N; ++k) sum = sum + ru(i, j, k);
    const double cff = v2(D, FID(om_u))(i, j) * v2(D, FID(on_u))(i, j);
    const double s1 = v2(D, FID(sustr))(i, j) * cff, s2 = -v2(D, FID(bustr))(i, j) * cff;
    v2(D, FID(rufrc))(i, j) = sum + s1 + s2;
  } else {
    if (!(i >= b.Istr && i <= b.Iend && j >= b.JstrV && j <= b.Jend)) return;
    V3 rv = v3l(D, FID(rv), nrhs);
    double sum = rv(i, j, 1);
    for (int k = 2; k <= N; ++k) sum = sum + rv(i, j, k);
    const double cff = v2(D, FID(om_v))(i, j) * v2(D, FID(on_v))(i, j);
    const double s1 = v2(D, FID(svstr))(i, j) * cff, s2 = -v2(D, FID(bvstr))(i, j) * cff;
    v2(D, FID(rvfrc))(i, j) = sum + s1 + s2;
  }
}
// ---- rhs3d, production form: a block marches a 32 x 8 tile up a chunk of levels (the recipe of k_tracer.cu) ---------------------
// Per level the planes u, v, Huon, Hvom, Hz, W of the tile + two rings of halo go through shared memory (loaded into registers
// one level ahead); from them every rho point evaluates UFx, VFe, the Coriolis and the curvilinear terms ONCE, every psi point
// UFe and VFx, every u / v point the horizontal weight of its vertical flux; the u and v updates of a point read those planes.
// The vertical C4 fluxes of w-level k-1 are carried, the column values u, v (k-1..k+2) roll through registers.  With one chunk
// (whole column) rufrc/rvfrc are summed in registers in the reference's order (rhs3d.F:1707-1916) and the sum kernel disappears.
// RS: view of a raw plane in shared memory with the (i,j,k) call syntax of V3, so the flux functions above are reused as they are.
constexpr int R2_TX = 32, R2_TY = 8, R2_NT = R2_TX * R2_TY;
constexpr int R2_RW = R2_TX + 4, R2_NR = R2_RW * (R2_TY + 4);        // raw planes: i in [I0-2, I0+33], j in [J0-2, J0+9]
constexpr int R2_CW = R2_TX + 1, R2_NC = R2_CW * (R2_TY + 1);        // rho cells from (I0-1, J0-1), psi / u / v cells from (I0, J0): 33 x 9
constexpr int R2_NQ = (R2_NR + R2_NT - 1) / R2_NT, R2_CQ = (R2_NC + R2_NT - 1) / R2_NT;
constexpr int R2_NPL = 10;                                           // UFx, VFe, cor_u, cor_v, curv_u, curv_v | UFe, VFx, wt_u, wt_v
constexpr size_t R2_SMEM = (2 * R2_NPL * R2_NC + 2 * 6 * R2_NR) * sizeof(double);
struct RS {
  const double* p; int i0, j0;             // element (i,j) at (i - i0) + R2_RW * (j - j0)
  __device__ __forceinline__ double operator()(int i, int j, int) const { return p[(i - i0) + R2_RW * (j - j0)]; }
};
struct RQS { RS u, v, Hz, Huon, Hvom, W; V2 fomn, dndx, dmde; int curv; int S, N, Jstr, Jend; };
// the flux functions of the per-level kernel, on shared-memory planes (same expressions: q_* above are templates over the view)
template <bool FULL>
__global__ void __launch_bounds__(R2_NT, 2) rhs3d_roll_kernel(const Dev D, Box bx, int nrhs, int nch) {
  extern __shared__ double r2sm[];
  double* const smD = r2sm; double* const smR = smD + 2 * R2_NPL * R2_NC;
  const roms_b200_bounds& b = D.b; const int N = b.N;
  const int ch = (int)blockIdx.z, per = (N + nch - 1) / nch, k0 = 1 + per * ch, k1 = min(k0 + per - 1, N);
  if (k0 > N) return;
  const int I0 = bx.i0 + blockIdx.x * R2_TX, J0 = bx.j0 + blockIdx.y * R2_TY;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * R2_TX + tx, i = I0 + tx, j = J0 + ty;
  V3 u = v3l(D, FID(u), nrhs), v = v3l(D, FID(v), nrhs), Hz = v3(D, FID(Hz)), Huon = v3(D, FID(Huon)), Hvom = v3(D, FID(Hvom)), W = v3(D, FID(W));
  V3 ru = v3l(D, FID(ru), nrhs), rv = v3l(D, FID(rv), nrhs);
  auto DS = [&](int L) { return smD + (L & 1) * R2_NPL * R2_NC; };
  auto RAW = [&](int L) { return smR + (L & 1) * 6 * R2_NR; };           // u, v, Huon, Hvom, Hz, W
  auto view = [&](int L) {
    const double* R = RAW(L);
    return RQS{RS{R, I0 - 2, J0 - 2}, RS{R + R2_NR, I0 - 2, J0 - 2}, RS{R + 4 * R2_NR, I0 - 2, J0 - 2}, RS{R + 2 * R2_NR, I0 - 2, J0 - 2},
               RS{R + 3 * R2_NR, I0 - 2, J0 - 2}, RS{R + 5 * R2_NR, I0 - 2, J0 - 2}, v2(D, FID(fomn)), v2(D, FID(dndx)), v2(D, FID(dmde)),
               D.p.app == ROMS_B200_APP_BENCHMARK, b.Southern_Edge && !b.NSperiodic, b.Northern_Edge && !b.NSperiodic, b.Jstr, b.Jend};
  };
  int ri[R2_NQ], rj[R2_NQ]; bool ro[R2_NQ];
#pragma unroll
  for (int n = 0; n < R2_NQ; ++n) {
    const int q = tid + n * R2_NT;
    ri[n] = I0 - 2 + q % R2_RW; rj[n] = J0 - 2 + q / R2_RW;
    ro[n] = q < R2_NR && ri[n] <= bx.i1 + 2 && rj[n] <= bx.j1 + 2 && ri[n] >= b.LBi && ri[n] <= b.UBi && rj[n] >= b.LBj && rj[n] <= b.UBj;
  }
  // which derived cells this thread evaluates, and which of their quantities somebody reads (the reference's loop ranges)
  int ci_[R2_CQ], cj_[R2_CQ]; bool inq[R2_CQ];
  bool nUFx[R2_CQ], nVFe[R2_CQ], nCor[R2_CQ], nUFe[R2_CQ], nVFx[R2_CQ], nWu[R2_CQ], nWv[R2_CQ];
  double cf[R2_CQ], cdn[R2_CQ], cdm[R2_CQ];
#pragma unroll
  for (int n = 0; n < R2_CQ; ++n) {
    const int q = tid + n * R2_NT; inq[n] = q < R2_NC;
    ci_[n] = q % R2_CW; cj_[n] = q / R2_CW;
    const int ir = I0 - 1 + ci_[n], jr = J0 - 1 + cj_[n], ip = I0 + ci_[n], jp = J0 + cj_[n];
    const bool tr = inq[n] && ir <= bx.i1 && jr <= bx.j1, tp = inq[n] && ip <= bx.i1 + 1 && jp <= bx.j1 + 1;
    nUFx[n] = tr && ir >= b.IstrU - 1 && ir <= b.Iend && jr >= b.Jstr && jr <= b.Jend;
    nVFe[n] = tr && ir >= b.Istr && ir <= b.Iend && jr >= b.JstrV - 1 && jr <= b.Jend;
    nCor[n] = nUFx[n] || nVFe[n];
    nUFe[n] = tp && ip >= b.IstrU && ip <= b.Iend && jp >= b.Jstr && jp <= b.Jend + 1;
    nVFx[n] = tp && ip >= b.Istr && ip <= b.Iend + 1 && jp >= b.JstrV && jp <= b.Jend;
    nWu[n] = tp && ip >= b.IstrU && ip <= b.Iend && jp >= b.Jstr && jp <= b.Jend && ip <= bx.i1 && jp <= bx.j1;
    nWv[n] = tp && ip >= b.Istr && ip <= b.Iend && jp >= b.JstrV && jp <= b.Jend && ip <= bx.i1 && jp <= bx.j1;
    cf[n] = 0.0; cdn[n] = 0.0; cdm[n] = 0.0;
    if (nCor[n]) { cf[n] = v2(D, FID(fomn))(ir, jr); cdn[n] = v2(D, FID(dndx))(ir, jr); cdm[n] = v2(D, FID(dmde))(ir, jr); }
  }
  double r_[6][R2_NQ];
  auto loadR = [&](int L) {
#pragma unroll
    for (int n = 0; n < R2_NQ; ++n) {
#pragma unroll
      for (int f = 0; f < 6; ++f) r_[f][n] = 0.0;
      if (ro[n] && L >= 0 && L <= N) {
        r_[5][n] = W(ri[n], rj[n], L);
        if (L >= 1) { r_[0][n] = u(ri[n], rj[n], L); r_[1][n] = v(ri[n], rj[n], L); r_[2][n] = Huon(ri[n], rj[n], L); r_[3][n] = Hvom(ri[n], rj[n], L); r_[4][n] = Hz(ri[n], rj[n], L); }
      }
    }
  };
  auto commitR = [&](int L) {
    double* R = RAW(L);
#pragma unroll
    for (int n = 0; n < R2_NQ; ++n) { const int q = tid + n * R2_NT; if (q < R2_NR) {
#pragma unroll
      for (int f = 0; f < 6; ++f) R[f * R2_NR + q] = r_[f][n]; } }
  };
  auto derive = [&](int L) {
    const RQS Q = view(L); double* F = DS(L);
    const double c1 = 9.0 / 16.0, c2 = 1.0 / 16.0;
#pragma unroll
    for (int n = 0; n < R2_CQ; ++n) {
      if (!inq[n]) continue;
      const int q = tid + n * R2_NT;
      const int ir = I0 - 1 + ci_[n], jr = J0 - 1 + cj_[n], ip = I0 + ci_[n], jp = J0 + cj_[n];
      if (nUFx[n]) F[q] = q_UFx(Q, ir, jr, L);
      if (nVFe[n]) F[R2_NC + q] = q_VFe(Q, ir, jr, L);
      if (nCor[n]) {
        const double cff = 0.5 * Q.Hz(ir, jr, L) * cf[n];
        F[2 * R2_NC + q] = cff * (Q.v(ir, jr, L) + Q.v(ir, jr + 1, L)); F[3 * R2_NC + q] = cff * (Q.u(ir, jr, L) + Q.u(ir + 1, jr, L));
        if (Q.curv) {
          const double a1 = 0.5 * (Q.v(ir, jr, L) + Q.v(ir, jr + 1, L)), a2 = 0.5 * (Q.u(ir, jr, L) + Q.u(ir + 1, jr, L));
          const double a3 = a1 * cdn[n], a4 = a2 * cdm[n];
          const double cc = Q.Hz(ir, jr, L) * (a3 - a4);
          F[4 * R2_NC + q] = cc * a1; F[5 * R2_NC + q] = cc * a2;
        }
      }
      if (nUFe[n]) F[6 * R2_NC + q] = q_UFe(Q, ip, jp, L);
      if (nVFx[n]) F[7 * R2_NC + q] = q_VFx(Q, ip, jp, L);
      if (nWu[n]) F[8 * R2_NC + q] = c1 * (Q.W(ip, jp, L) + Q.W(ip - 1, jp, L)) - c2 * (Q.W(ip + 1, jp, L) + Q.W(ip - 2, jp, L));
      if (nWv[n]) F[9 * R2_NC + q] = c1 * (Q.W(ip, jp, L) + Q.W(ip, jp - 1, L)) - c2 * (Q.W(ip, jp + 1, L) + Q.W(ip, jp - 2, L));
    }
  };
  // C4 vertical flux at w-level k from the horizontal weight and the column values (q_FCw)
  auto fcw = [&](int k, double wt, double qm1, double q0, double qp1, double qp2) -> double {
    const double c1 = 9.0 / 16.0, c2 = 1.0 / 16.0;
    if (k == 0 || k == N) return 0.0;
    if (k == 1) return (c1 * (q0 + qp1) - c2 * (q0 + qp2)) * wt;
    if (k == N - 1) return (c1 * (q0 + qp1) - c2 * (qm1 + qp1)) * wt;
    return (c1 * (q0 + qp1) - c2 * (qm1 + qp2)) * wt;
  };
  const bool inb = (i <= bx.i1 && j <= bx.j1);
  const bool doU = inb && (i >= b.IstrU && i <= b.Iend && j >= b.Jstr && j <= b.Jend);
  const bool doV = inb && (i >= b.Istr && i <= b.Iend && j >= b.JstrV && j <= b.Jend);
  const bool curv = (D.p.app == ROMS_B200_APP_BENCHMARK);
  const int lr = (tx + 1) + R2_CW * (ty + 1), lp = tx + R2_CW * ty;
  auto ucol = [&](int L) -> double { return (doU && L >= 1 && L <= N) ? u(i, j, L) : 0.0; };
  auto vcol = [&](int L) -> double { return (doV && L >= 1 && L <= N) ? v(i, j, L) : 0.0; };
  // ---- prologue: raw[k0-1] (only W is used: the weights of w-level k0-1), raw[k0]; column state
  loadR(k0 - 1); commitR(k0 - 1);
  __syncthreads();
  double FCmu = 0.0, FCmv = 0.0;
  {
    // weights of w-level k0-1 straight from the raw plane (chunk start only)
    const RQS Q = view(k0 - 1); const int km = k0 - 1;
    const double c1 = 9.0 / 16.0, c2 = 1.0 / 16.0;
    if (doU) { const double wt = c1 * (Q.W(i, j, km) + Q.W(i - 1, j, km)) - c2 * (Q.W(i + 1, j, km) + Q.W(i - 2, j, km)); FCmu = fcw(km, wt, ucol(km - 1), ucol(km), ucol(km + 1), ucol(km + 2)); }
    if (doV) { const double wt = c1 * (Q.W(i, j, km) + Q.W(i, j - 1, km)) - c2 * (Q.W(i, j + 1, km) + Q.W(i, j - 2, km)); FCmv = fcw(km, wt, vcol(km - 1), vcol(km), vcol(km + 1), vcol(km + 2)); }
  }
  __syncthreads();                                  // everybody is done with raw[k0-1] before raw[k0+1] replaces it
  loadR(k0); commitR(k0);
  loadR(k0 + 1);
  double um1 = ucol(k0 - 1), u0 = ucol(k0), up1 = ucol(k0 + 1), up2 = 0.0, unx = ucol(k0 + 2);     // unx, vnx: the next level of the column, in flight for a whole level
  double vm1 = vcol(k0 - 1), v0 = vcol(k0), vp1 = vcol(k0 + 1), vp2 = 0.0, vnx = vcol(k0 + 2);
  double run = doU ? ru(i, j, k0) : 0.0, rvn = doV ? rv(i, j, k0) : 0.0;
  double sumu = 0.0, sumv = 0.0;
  __syncthreads();
  derive(k0);
  for (int k = k0; k <= k1; ++k) {
    commitR(k + 1);
    loadR(k + 2);
    const double ruk = run, rvk = rvn;
    if (k + 1 <= k1) { if (doU) run = ru(i, j, k + 1); if (doV) rvn = rv(i, j, k + 1); }
    up2 = unx; vp2 = vnx;
    unx = ucol(k + 3); vnx = vcol(k + 3);
    __syncthreads();
    if (k + 1 <= k1) derive(k + 1);
    const double* F = DS(k);
    if (doU) {
      double r = ruk;
      r = r + 0.5 * (F[2 * R2_NC + lr] + F[2 * R2_NC + lr - 1]);
      if (curv) r = r + 0.5 * (F[4 * R2_NC + lr] + F[4 * R2_NC + lr - 1]);
      const double c1 = F[lr] - F[lr - 1], c2 = F[6 * R2_NC + lp + R2_CW] - F[6 * R2_NC + lp];
      r = r - (c1 + c2);
      const double FCk = fcw(k, F[8 * R2_NC + lp], um1, u0, up1, up2);
      r = r - (FCk - FCmu);
      ru(i, j, k) = r; FCmu = FCk;
      if (FULL) sumu = (k == 1) ? r : sumu + r;
    }
    if (doV) {
      double r = rvk;
      r = r - 0.5 * (F[3 * R2_NC + lr] + F[3 * R2_NC + lr - R2_CW]);
      if (curv) r = r - 0.5 * (F[5 * R2_NC + lr] + F[5 * R2_NC + lr - R2_CW]);
      const double c1 = F[7 * R2_NC + lp + 1] - F[7 * R2_NC + lp], c2 = F[R2_NC + lr] - F[R2_NC + lr - R2_CW];
      r = r - (c1 + c2);
      const double FCk = fcw(k, F[9 * R2_NC + lp], vm1, v0, vp1, vp2);
      r = r - (FCk - FCmv);
      rv(i, j, k) = r; FCmv = FCk;
      if (FULL) sumv = (k == 1) ? r : sumv + r;
    }
    um1 = u0; u0 = up1; up1 = up2; vm1 = v0; v0 = vp1; vp1 = vp2;
  }
  if (FULL) {                                       // rhs3d.F:1707-1916
    if (doU) { const double cff = v2(D, FID(om_u))(i, j) * v2(D, FID(on_u))(i, j); const double s1 = v2(D, FID(sustr))(i, j) * cff, s2 = -v2(D, FID(bustr))(i, j) * cff; v2(D, FID(rufrc))(i, j) = sumu + s1 + s2; }
    if (doV) { const double cff = v2(D, FID(om_v))(i, j) * v2(D, FID(on_v))(i, j); const double s1 = v2(D, FID(svstr))(i, j) * cff, s2 = -v2(D, FID(bvstr))(i, j) * cff; v2(D, FID(rvfrc))(i, j) = sumv + s1 + s2; }
  }
}
int k_rhs3d_tile(roms_b200_ctx* c, int nrhs) {
  const roms_b200_bounds& b = c->D.b;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend};
  static const bool per_level = (getenv("ROMS_B200_RHS3D_PERLEVEL") != nullptr);        // the first form
  if (!per_level) {
    static AttrOnce attr;
    if (attr.need(R2_SMEM)) {
      CUDA_OK(cudaFuncSetAttribute(rhs3d_roll_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R2_SMEM));
      CUDA_OK(cudaFuncSetAttribute(rhs3d_roll_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R2_SMEM));
    }
    dim3 blk(R2_TX, R2_TY); dim3 g = grid2(bx, blk);
    const long cols = (long)g.x * g.y;
    static const int waves = getenv("ROMS_B200_RHS3D_FILL") ? atoi(getenv("ROMS_B200_RHS3D_FILL")) : 1;
    int nch = (int)std::min<long>(((long)waves * 148 + cols - 1) / cols, (b.N + 4) / 5); if (nch < 1) nch = 1;
    g.z = nch;
    if (nch == 1) { rhs3d_roll_kernel<true><<<g, blk, R2_SMEM, c->stream>>>(c->D, bx, nrhs, nch); c->launches++; return 0; }
    rhs3d_roll_kernel<false><<<g, blk, R2_SMEM, c->stream>>>(c->D, bx, nrhs, nch); c->launches++;
  } else {
    dim3 blk(32, 8); dim3 g = grid2(bx, blk); g.z = 2 * b.N;
    rhs3d_kernel<<<g, blk, 0, c->stream>>>(c->D, bx, nrhs); c->launches++;
  }
  dim3 blk2(32, 4); dim3 g2 = grid2(bx, blk2); g2.z = 2;
  rhs3d_sum_kernel<<<g2, blk2, 0, c->stream>>>(c->D, bx, nrhs); c->launches++;
  return 0;
}

// ---- uv3dmix2_s_tile, uv3dmix2_s.h:239-330 ------------------------------------------
struct MQ { V3 u, v, Hz; V2 pm, pn, pmon_r, pnom_r, pmon_p, pnom_p, om_r, on_r, om_p, on_p, visc2_r, visc2_p; };
__device__ __forceinline__ double m_cffr(const MQ& Q, int i, int j, int k) {   // rho-point strain term
  return Q.Hz(i, j, k) * 0.5 *
         (Q.pmon_r(i, j) * ((Q.pn(i, j) + Q.pn(i + 1, j)) * Q.u(i + 1, j, k) - (Q.pn(i - 1, j) + Q.pn(i, j)) * Q.u(i, j, k)) -
          Q.pnom_r(i, j) * ((Q.pm(i, j) + Q.pm(i, j + 1)) * Q.v(i, j + 1, k) - (Q.pm(i, j - 1) + Q.pm(i, j)) * Q.v(i, j, k)));
}
__device__ __forceinline__ double m_cffp(const MQ& Q, int i, int j, int k) {   // psi-point strain term
  return 0.125 * (Q.Hz(i - 1, j, k) + Q.Hz(i, j, k) + Q.Hz(i - 1, j - 1, k) + Q.Hz(i, j - 1, k)) *
         (Q.pmon_p(i, j) * ((Q.pn(i, j - 1) + Q.pn(i, j)) * Q.v(i, j, k) - (Q.pn(i - 1, j - 1) + Q.pn(i - 1, j)) * Q.v(i - 1, j, k)) +
          Q.pnom_p(i, j) * ((Q.pm(i - 1, j) + Q.pm(i, j)) * Q.u(i, j, k) - (Q.pm(i - 1, j - 1) + Q.pm(i, j - 1)) * Q.u(i, j - 1, k)));
}
__device__ __forceinline__ double m_UFx(const MQ& Q, int i, int j, int k) { return Q.on_r(i, j) * Q.on_r(i, j) * Q.visc2_r(i, j) * m_cffr(Q, i, j, k); }
__device__ __forceinline__ double m_VFe(const MQ& Q, int i, int j, int k) { return Q.om_r(i, j) * Q.om_r(i, j) * Q.visc2_r(i, j) * m_cffr(Q, i, j, k); }
__device__ __forceinline__ double m_UFe(const MQ& Q, int i, int j, int k) { return Q.om_p(i, j) * Q.om_p(i, j) * Q.visc2_p(i, j) * m_cffp(Q, i, j, k); }
__device__ __forceinline__ double m_VFx(const MQ& Q, int i, int j, int k) { return Q.on_p(i, j) * Q.on_p(i, j) * Q.visc2_p(i, j) * m_cffp(Q, i, j, k); }
// One thread per (i,j,k,component): nothing in the stress divergence is a vertical recurrence.  The two terms every level adds
// to rufrc/rvfrc (uv3dmix2_s.h:296-297,326-327) are parked in scratch volumes and summed in the reference's order
// (acc = acc + c1 + c2, k = 1..N) by uv3dmix2_sum_kernel, one thread per column and component.
__global__ void __launch_bounds__(256) uv3dmix2_kernel(const Dev D, Box bx, int nrhs, int nnew, double* scratch) {
  IJZ_FROM_BOX(bx, D.b.N);
  const roms_b200_bounds& b = D.b; const int N = b.N, k = 1 + zlev, comp = zcomp; const double dt = D.p.dt;
  MQ Q{v3l(D, FID(u), nrhs), v3l(D, FID(v), nrhs), v3(D, FID(Hz)), v2(D, FID(pm)), v2(D, FID(pn)), v2(D, FID(pmon_r)), v2(D, FID(pnom_r)),
       v2(D, FID(pmon_p)), v2(D, FID(pnom_p)), v2(D, FID(om_r)), v2(D, FID(on_r)), v2(D, FID(om_p)), v2(D, FID(on_p)), v2(D, FID(visc2_r)), v2(D, FID(visc2_p))};
  const size_t vol = D.nij * (size_t)(N + 1);
  V3 S1{scratch + (2 * comp) * vol, b.LBi, D.ni, b.LBj, D.nj, 0}, S2{scratch + (2 * comp + 1) * vol, b.LBi, D.ni, b.LBj, D.nj, 0};
  if (comp == 0) {
    if (!(i >= b.IstrU && i <= b.Iend && j >= b.Jstr && j <= b.Jend)) return;
    V3 un = v3l(D, FID(u), nnew);
    const double cff = dt * 0.25 * (Q.pm(i - 1, j) + Q.pm(i, j)) * (Q.pn(i - 1, j) + Q.pn(i, j));
    const double c1 = 0.5 * (Q.pn(i - 1, j) + Q.pn(i, j)) * (m_UFx(Q, i, j, k) - m_UFx(Q, i - 1, j, k));
    const double c2 = 0.5 * (Q.pm(i - 1, j) + Q.pm(i, j)) * (m_UFe(Q, i, j + 1, k) - m_UFe(Q, i, j, k));
    const double c3 = cff * (c1 + c2);
    S1(i, j, k) = c1; S2(i, j, k) = c2;
    un(i, j, k) = un(i, j, k) + c3;
  } else {
    if (!(i >= b.Istr && i <= b.Iend && j >= b.JstrV && j <= b.Jend)) return;
    V3 vn = v3l(D, FID(v), nnew);
    const double cff = dt * 0.25 * (Q.pm(i, j) + Q.pm(i, j - 1)) * (Q.pn(i, j) + Q.pn(i, j - 1));
    const double c1 = 0.5 * (Q.pn(i, j - 1) + Q.pn(i, j)) * (m_VFx(Q, i + 1, j, k) - m_VFx(Q, i, j, k));
    const double c2 = 0.5 * (Q.pm(i, j - 1) + Q.pm(i, j)) * (m_VFe(Q, i, j, k) - m_VFe(Q, i, j - 1, k));
    const double c3 = cff * (c1 - c2);
    S1(i, j, k) = c1; S2(i, j, k) = c2;
    vn(i, j, k) = vn(i, j, k) + c3;
  }
}
__global__ void __launch_bounds__(128) uv3dmix2_sum_kernel(const Dev D, Box bx, const double* scratch) {
  IJ_FROM_BOX(bx);
  const roms_b200_bounds& b = D.b; const int N = b.N, comp = blockIdx.z;
  if (comp == 0 ? !(i >= b.IstrU && i <= b.Iend && j >= b.Jstr && j <= b.Jend) : !(i >= b.Istr && i <= b.Iend && j >= b.JstrV && j <= b.Jend)) return;
  const size_t vol = D.nij * (size_t)(N + 1), o = (i - b.LBi) + (size_t)D.ni * (j - b.LBj);
  const double* p1 = scratch + (2 * comp) * vol + o; const double* p2 = p1 + vol;
  V2 frc = v2(D, comp == 0 ? FID(rufrc) : FID(rvfrc));
  double acc = frc(i, j);
  constexpr int KB = 6;
  for (int k0 = 1; k0 <= N; k0 += KB) {
    double a[KB], c[KB];
#pragma unroll
    for (int q = 0; q < KB; ++q) { const int k = min(k0 + q, N); a[q] = p1[D.nij * k]; c[q] = p2[D.nij * k]; }
#pragma unroll
    for (int q = 0; q < KB; ++q) if (k0 + q <= N) { if (comp == 0) acc = acc + a[q] + c[q]; else acc = acc + a[q] - c[q]; }
  }
  frc(i, j) = acc;
}
// ---- uv3dmix2, production form: a block marches a 32 x 8 tile up a chunk of levels (the recipe of k_tracer.cu) ----------------
// Per level the rho-point strain term (m_cffr) and the psi-point one (m_cffp) are evaluated ONCE per point from the u, v, Hz
// planes of the level (tile + one ring of halo, staged through shared memory, loaded into registers one level ahead), scaled
// into UFx, VFe (rho points) and UFe, VFx (psi points) in shared memory; the u and v updates of a point read them there.  The
// 2-D coefficients of a point (metric sums and products) sit in shared memory for the whole march instead of being re-read per
// level.  With one chunk (the whole column) the two terms every level adds to rufrc/rvfrc are accumulated in registers in the
// reference's order and the scratch volumes and the sum kernel disappear; with several chunks they are parked as before.
constexpr int M2_TX = 32, M2_TY = 8, M2_NT = M2_TX * M2_TY;
constexpr int M2_RW = M2_TX + 2, M2_NR = M2_RW * (M2_TY + 2);        // raw planes u, v, Hz: i in [I0-1, I0+32], j in [J0-1, J0+8]
constexpr int M2_CW = M2_TX + 1, M2_NC = M2_CW * (M2_TY + 1);        // rho cells: i in [I0-1, I0+31], j in [J0-1, J0+7]; psi cells: i in [I0, I0+32], j in [J0, J0+8]
constexpr int M2_NQ = (M2_NR + M2_NT - 1) / M2_NT, M2_CQ = (M2_NC + M2_NT - 1) / M2_NT;
constexpr int M2_DSLOT = 4 * M2_NC;                                   // UFx, VFe, UFe, VFx
constexpr int M2_NCOEF = 8;                                           // per cell: pmon, pnom, 4 metric sums, 2 scale factors
constexpr size_t M2_SMEM = (3 * M2_DSLOT + 2 * 3 * M2_NR + 2 * M2_NCOEF * M2_NC) * sizeof(double);
template <bool FULL>
__global__ void __launch_bounds__(M2_NT, 2) uv3dmix2_roll_kernel(const Dev D, Box bx, int nrhs, int nnew, int nch, double* scratch) {
  extern __shared__ double m2sm[];
  double* const smD = m2sm; double* const smR = smD + 3 * M2_DSLOT; double* const cR = smR + 2 * 3 * M2_NR; double* const cP = cR + M2_NCOEF * M2_NC;
  const roms_b200_bounds& b = D.b; const int N = b.N; const double dt = D.p.dt;
  const int ch = (int)blockIdx.z, per = (N + nch - 1) / nch, k0 = 1 + per * ch, k1 = min(k0 + per - 1, N);
  if (k0 > N) return;
  const int I0 = bx.i0 + blockIdx.x * M2_TX, J0 = bx.j0 + blockIdx.y * M2_TY;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * M2_TX + tx, i = I0 + tx, j = J0 + ty;
  V3 u = v3l(D, FID(u), nrhs), v = v3l(D, FID(v), nrhs), Hz = v3(D, FID(Hz)), un = v3l(D, FID(u), nnew), vn = v3l(D, FID(v), nnew);
  V2 pm = v2(D, FID(pm)), pn = v2(D, FID(pn));
  auto DS = [&](int L) { return smD + (L % 3) * M2_DSLOT; };
  auto RAW = [&](int L) { return smR + (L & 1) * 3 * M2_NR; };           // u; v at + M2_NR; Hz at + 2*M2_NR
  // ---- cells of this thread; coefficients of the rho and psi cells into shared memory (once)
  int ri[M2_NQ], rj[M2_NQ]; bool ro[M2_NQ];
#pragma unroll
  for (int n = 0; n < M2_NQ; ++n) {
    const int q = tid + n * M2_NT;
    ri[n] = I0 - 1 + q % M2_RW; rj[n] = J0 - 1 + q / M2_RW;
    ro[n] = q < M2_NR && ri[n] <= bx.i1 + 1 && rj[n] <= bx.j1 + 1 && ri[n] >= b.LBi && ri[n] <= b.UBi && rj[n] >= b.LBj && rj[n] <= b.UBj;
  }
  bool co_r[M2_CQ], co_p[M2_CQ]; int cq_r[M2_CQ], cq_p[M2_CQ];             // raw-plane offset of the cell's (i,j)
  {
    V2 pmon_r = v2(D, FID(pmon_r)), pnom_r = v2(D, FID(pnom_r)), pmon_p = v2(D, FID(pmon_p)), pnom_p = v2(D, FID(pnom_p));
    V2 om_r = v2(D, FID(om_r)), on_r = v2(D, FID(on_r)), om_p = v2(D, FID(om_p)), on_p = v2(D, FID(on_p)), visc2_r = v2(D, FID(visc2_r)), visc2_p = v2(D, FID(visc2_p));
#pragma unroll
    for (int n = 0; n < M2_CQ; ++n) {
      const int q = tid + n * M2_NT, ci = q % M2_CW, cj = q / M2_CW;
      // rho cell (I0-1+ci, J0-1+cj): needed for i in [Istr-1, Iend], j in [Jstr-1, Jend]
      { const int ii = I0 - 1 + ci, jj = J0 - 1 + cj;
        co_r[n] = q < M2_NC && ii <= bx.i1 && jj <= bx.j1 && ii - 1 >= b.LBi && jj - 1 >= b.LBj; cq_r[n] = ci + M2_RW * cj;   // (a row beyond a closed wall is never used)
        if (co_r[n]) {
          double* c = cR + q;
          c[0 * M2_NC] = pmon_r(ii, jj); c[1 * M2_NC] = pnom_r(ii, jj);
          c[2 * M2_NC] = pn(ii, jj) + pn(ii + 1, jj); c[3 * M2_NC] = pn(ii - 1, jj) + pn(ii, jj);
          c[4 * M2_NC] = pm(ii, jj) + pm(ii, jj + 1); c[5 * M2_NC] = pm(ii, jj - 1) + pm(ii, jj);
          c[6 * M2_NC] = on_r(ii, jj) * on_r(ii, jj) * visc2_r(ii, jj); c[7 * M2_NC] = om_r(ii, jj) * om_r(ii, jj) * visc2_r(ii, jj);
        } }
      // psi cell (I0+ci, J0+cj): needed for i in [Istr, Iend+1], j in [Jstr, Jend+1]
      { const int ii = I0 + ci, jj = J0 + cj;
        co_p[n] = q < M2_NC && ii <= bx.i1 + 1 && jj <= bx.j1 + 1; cq_p[n] = (ci + 1) + M2_RW * (cj + 1);
        if (co_p[n]) {
          double* c = cP + q;
          c[0 * M2_NC] = pmon_p(ii, jj); c[1 * M2_NC] = pnom_p(ii, jj);
          c[2 * M2_NC] = pn(ii, jj - 1) + pn(ii, jj); c[3 * M2_NC] = pn(ii - 1, jj - 1) + pn(ii - 1, jj);
          c[4 * M2_NC] = pm(ii - 1, jj) + pm(ii, jj); c[5 * M2_NC] = pm(ii - 1, jj - 1) + pm(ii, jj - 1);
          c[6 * M2_NC] = om_p(ii, jj) * om_p(ii, jj) * visc2_p(ii, jj); c[7 * M2_NC] = on_p(ii, jj) * on_p(ii, jj) * visc2_p(ii, jj);
        } }
    }
  }
  double ru_[M2_NQ], rv_[M2_NQ], rh_[M2_NQ];
  auto loadR = [&](int L) {
#pragma unroll
    for (int n = 0; n < M2_NQ; ++n) {
      ru_[n] = 0.0; rv_[n] = 0.0; rh_[n] = 0.0;
      if (ro[n] && L >= 1 && L <= N) { ru_[n] = u(ri[n], rj[n], L); rv_[n] = v(ri[n], rj[n], L); rh_[n] = Hz(ri[n], rj[n], L); }
    }
  };
  auto commitR = [&](int L) {
    double* R = RAW(L);
#pragma unroll
    for (int n = 0; n < M2_NQ; ++n) { const int q = tid + n * M2_NT; if (q < M2_NR) { R[q] = ru_[n]; R[M2_NR + q] = rv_[n]; R[2 * M2_NR + q] = rh_[n]; } }
  };
  auto derive = [&](int L) {                        // m_cffr / m_cffp of level L and the four scaled fluxes
    const double* U = RAW(L); const double* V = U + M2_NR; const double* H = V + M2_NR; double* F = DS(L);
#pragma unroll
    for (int n = 0; n < M2_CQ; ++n) {
      const int q = tid + n * M2_NT;
      if (co_r[n]) {
        const double* c = cR + q; const int r = cq_r[n];
        const double cffr = H[r] * 0.5 * (c[0] * (c[2 * M2_NC] * U[r + 1] - c[3 * M2_NC] * U[r]) - c[M2_NC] * (c[4 * M2_NC] * V[r + M2_RW] - c[5 * M2_NC] * V[r]));
        F[q] = c[6 * M2_NC] * cffr; F[M2_NC + q] = c[7 * M2_NC] * cffr;
      }
      if (co_p[n]) {
        const double* c = cP + q; const int r = cq_p[n];
        const double cffp = 0.125 * (H[r - 1] + H[r] + H[r - 1 - M2_RW] + H[r - M2_RW]) *
                            (c[0] * (c[2 * M2_NC] * V[r] - c[3 * M2_NC] * V[r - 1]) + c[M2_NC] * (c[4 * M2_NC] * U[r] - c[5 * M2_NC] * U[r - M2_RW]));
        F[2 * M2_NC + q] = c[6 * M2_NC] * cffp; F[3 * M2_NC + q] = c[7 * M2_NC] * cffp;
      }
    }
  };
  // ---- this thread's u point and v point
  const bool doU = (i <= bx.i1 && j <= bx.j1) && (i >= b.IstrU && i <= b.Iend && j >= b.Jstr && j <= b.Jend);
  const bool doV = (i <= bx.i1 && j <= bx.j1) && (i >= b.Istr && i <= b.Iend && j >= b.JstrV && j <= b.Jend);
  double cffu = 0.0, pnu = 0.0, pmu = 0.0, cffv = 0.0, pnv = 0.0, pmv = 0.0, accu = 0.0, accv = 0.0;
  if (doU) { pnu = pn(i - 1, j) + pn(i, j); pmu = pm(i - 1, j) + pm(i, j); cffu = dt * 0.25 * pmu * pnu; if (FULL) accu = v2(D, FID(rufrc))(i, j); }
  if (doV) { pnv = pn(i, j - 1) + pn(i, j); pmv = pm(i, j - 1) + pm(i, j); cffv = dt * 0.25 * (pm(i, j) + pm(i, j - 1)) * (pn(i, j) + pn(i, j - 1)); if (FULL) accv = v2(D, FID(rvfrc))(i, j); }
  const size_t vol = D.nij * (size_t)(N + 1);
  V3 S1u{scratch, b.LBi, D.ni, b.LBj, D.nj, 0}, S2u{scratch + vol, b.LBi, D.ni, b.LBj, D.nj, 0}, S1v{scratch + 2 * vol, b.LBi, D.ni, b.LBj, D.nj, 0}, S2v{scratch + 3 * vol, b.LBi, D.ni, b.LBj, D.nj, 0};
  const int lr = (tx + 1) + M2_CW * (ty + 1), lp = tx + M2_CW * ty;     // rho cell (i,j) and psi cell (i,j) of this thread
  // ---- prologue
  loadR(k0); commitR(k0);
  loadR(k0 + 1);
  double unn = doU ? un(i, j, k0) : 0.0, vnn = doV ? vn(i, j, k0) : 0.0;
  __syncthreads();                                  // coefficients and raw[k0] visible
  derive(k0);
  for (int k = k0; k <= k1; ++k) {
    commitR(k + 1);
    loadR(k + 2);
    const double unk = unn, vnk = vnn;
    if (k + 1 <= k1) { if (doU) unn = un(i, j, k + 1); if (doV) vnn = vn(i, j, k + 1); }
    __syncthreads();                                // derived[k] and raw[k+1] visible; everybody is done with derived[k-1] ... raw[k]
    if (k + 1 <= k1) derive(k + 1);
    const double* F = DS(k);
    if (doU) {
      const double c1 = 0.5 * pnu * (F[lr] - F[lr - 1]);                                        // UFx(i,j) - UFx(i-1,j)
      const double c2 = 0.5 * pmu * (F[2 * M2_NC + lp + M2_CW] - F[2 * M2_NC + lp]);           // UFe(i,j+1) - UFe(i,j)
      const double c3 = cffu * (c1 + c2);
      if (FULL) accu = accu + c1 + c2; else { S1u(i, j, k) = c1; S2u(i, j, k) = c2; }
      un(i, j, k) = unk + c3;
    }
    if (doV) {
      const double c1 = 0.5 * pnv * (F[3 * M2_NC + lp + 1] - F[3 * M2_NC + lp]);               // VFx(i+1,j) - VFx(i,j)
      const double c2 = 0.5 * pmv * (F[M2_NC + lr] - F[M2_NC + lr - M2_CW]);                   // VFe(i,j) - VFe(i,j-1)
      const double c3 = cffv * (c1 - c2);
      if (FULL) accv = accv + c1 - c2; else { S1v(i, j, k) = c1; S2v(i, j, k) = c2; }
      vn(i, j, k) = vnk + c3;
    }
  }
  if (FULL) { if (doU) v2(D, FID(rufrc))(i, j) = accu; if (doV) v2(D, FID(rvfrc))(i, j) = accv; }
}
int k_uv3dmix2(roms_b200_ctx* c, int nrhs, int nnew) {
  const roms_b200_bounds& b = c->D.b;
  if (!c->D.kpp4) return 1;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend};
  static const bool per_level = (getenv("ROMS_B200_UVMIX_PERLEVEL") != nullptr);        // the first form
  if (!per_level) {
    static AttrOnce attr;
    if (attr.need(M2_SMEM)) {
      CUDA_OK(cudaFuncSetAttribute(uv3dmix2_roll_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M2_SMEM));
      CUDA_OK(cudaFuncSetAttribute(uv3dmix2_roll_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M2_SMEM));
    }
    dim3 blk(M2_TX, M2_TY); dim3 g = grid2(bx, blk);
    const long cols = (long)g.x * g.y;
    static const int waves = getenv("ROMS_B200_UVMIX_FILL") ? atoi(getenv("ROMS_B200_UVMIX_FILL")) : 0;
    int nch = (int)std::min<long>(((long)waves * 148 + cols - 1) / cols, (b.N + 4) / 5); if (nch < 1) nch = 1;
    g.z = nch;
    if (nch == 1) { uv3dmix2_roll_kernel<true><<<g, blk, M2_SMEM, c->stream>>>(c->D, bx, nrhs, nnew, nch, c->D.kpp4); c->launches++; return 0; }
    uv3dmix2_roll_kernel<false><<<g, blk, M2_SMEM, c->stream>>>(c->D, bx, nrhs, nnew, nch, c->D.kpp4); c->launches++;
  } else {
    dim3 blk(32, 8); dim3 g = grid2(bx, blk); g.z = 2 * b.N;
    uv3dmix2_kernel<<<g, blk, 0, c->stream>>>(c->D, bx, nrhs, nnew, c->D.kpp4); c->launches++;
  }
  dim3 blk2(32, 4); dim3 g2 = grid2(bx, blk2); g2.z = 2;
  uv3dmix2_sum_kernel<<<g2, blk2, 0, c->stream>>>(c->D, bx, c->D.kpp4); c->launches++;
  return 0;
}

// ---- step3d_uv_tile, step3d_uv.F:330-1824 -------------------------------------------------
// One thread per (u or v) water column.  The column arrays live in shared memory ([level][thread], conflict-free) instead
// of thread-local memory (14 warps x 2.6 KB per SM thrashed the L1), global loads are issued in batches of UV_KB levels
// ahead of the recurrences that consume them, and the per-level reciprocals are the branch-free rcp_ieee (== 1.0/x) so
// that independent levels overlap.  Per-point operation order is the reference's.  (134 -> 77 us on BENCHMARK1.)
constexpr int UV_T = 64;      // threads per block (32 x 2)
constexpr int UV_KB = 6;      // levels per load batch
// pass 1 (interior u/v points): add ru, implicit spline vertical viscosity, replace the
// vertical mean by DU_avg1 (:357-715, :859-1182), then the closed-wall u3dbc/v3dbc rows.
__global__ void __launch_bounds__(UV_T) step3d_uv1_kernel(const Dev D, Box bx, int nrhs, int nnew, double cffab) {
  extern __shared__ double uvsm[];
  const int i = bx.i0 + blockIdx.x * blockDim.x + threadIdx.x, j = bx.j0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > bx.i1 || j > bx.j1) return;
  const roms_b200_bounds& b = D.b; const int N = b.N; const double dt = D.p.dt;
  V3 Hz = v3(D, FID(Hz)), Akv = v3(D, FID(Akv)); V2 pm = v2(D, FID(pm)), pn = v2(D, FID(pn));
  const bool S = b.Southern_Edge && !b.NSperiodic, Nn = b.Northern_Edge && !b.NSperiodic;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, L = (N + 2) * UV_T;
  double *q = uvsm + tid, *Hzk = q + L, *oHz = Hzk + L, *CF = oHz + L, *DC = CF + L, *ak = DC + L;   // element k at [k * UV_T]
  const int comp = blockIdx.z;              // u and v columns on separate threads
  int bad = 0;
  if ((comp == 0 && (i >= b.IstrU && i <= b.Iend && j >= b.Jstr && j <= b.Jend)) ||
      (comp == 1 && (i >= b.Istr && i <= b.Iend && j >= b.JstrV && j <= b.Jend))) {
    const int di = comp == 0 ? 1 : 0, dj = 1 - di;
    V3 qn = v3l(D, comp == 0 ? FID(u) : FID(v), nnew), r = v3l(D, comp == 0 ? FID(ru) : FID(rv), nrhs);
    const double DC0 = cffab * (pm(i, j) + pm(i - di, j - dj)) * (pn(i, j) + pn(i - di, j - dj));
    ak[0] = 0.5 * (Akv(i - di, j - dj, 0) + Akv(i, j, 0));
    for (int k0 = 1; k0 <= N; k0 += UV_KB) {
      double ha[UV_KB], hb[UV_KB], qv[UV_KB], rv[UV_KB], aa[UV_KB], ab[UV_KB];
#pragma unroll
      for (int t = 0; t < UV_KB; ++t) {
        const int k = min(k0 + t, N);
        ha[t] = Hz(i - di, j - dj, k); hb[t] = Hz(i, j, k); qv[t] = qn(i, j, k); rv[t] = r(i, j, k);
        aa[t] = Akv(i - di, j - dj, k); ab[t] = Akv(i, j, k);
      }
#pragma unroll
      for (int t = 0; t < UV_KB; ++t) {
        const int k = k0 + t;
        if (k <= N) {
          const double hz = 0.5 * (ha[t] + hb[t]), oh = rcp_ieee(hz, bad);
          Hzk[k * UV_T] = hz; oHz[k * UV_T] = oh;
          const double val = qv[t] + DC0 * rv[t];
          q[k * UV_T] = val * oh;
          ak[k * UV_T] = 0.5 * (aa[t] + ab[t]);
        }
      }
    }
    // spline tridiagonal (step3d_uv.F:392-438)
    {
      double cfp = 0.0, dcp = 0.0;
      double hz_k = Hzk[UV_T], oh_k = oHz[UV_T], q_k = q[UV_T], ak_km = ak[0], ak_k = ak[UV_T];
#pragma unroll 2
      for (int k = 1; k <= N - 1; ++k) {
        const double hz_p = Hzk[(k + 1) * UV_T], oh_p = oHz[(k + 1) * UV_T], q_p = q[(k + 1) * UV_T], ak_kp = ak[(k + 1) * UV_T];
        const double FC = (1.0 / 6.0) * hz_k - dt * ak_km * oh_k;
        const double CFk = (1.0 / 6.0) * hz_p - dt * ak_kp * oh_p;
        const double BC = (1.0 / 3.0) * (hz_k + hz_p) + dt * ak_k * (oh_k + oh_p);
        const double cf = rcp_ieee(BC - FC * cfp, bad);
        cfp = cf * CFk;
        dcp = cf * (q_p - q_k - FC * dcp);
        CF[k * UV_T] = cfp; DC[k * UV_T] = dcp;
        hz_k = hz_p; oh_k = oh_p; q_k = q_p; ak_km = ak_k; ak_k = ak_kp;
      }
    }
    {
      double dn = 0.0;                                   // DC(N) = 0
      DC[N * UV_T] = 0.0;
      for (int k = N - 1; k >= 1; --k) { dn = DC[k * UV_T] - CF[k * UV_T] * dn; DC[k * UV_T] = dn; }
    }
    double dcm = 0.0;
    for (int k = 1; k <= N; ++k) {
      const double dck = DC[k * UV_T] * ak[k * UV_T];
      q[k * UV_T] = q[k * UV_T] + dt * oHz[k * UV_T] * (dck - dcm);
      dcm = dck;
    }
    // vertical mean -> barotropic (step3d_uv.F:597-715)
    double CF0 = Hzk[UV_T], DCs = q[UV_T] * Hzk[UV_T];
    for (int k = 2; k <= N; ++k) { const double hz = Hzk[k * UV_T]; CF0 = CF0 + hz; DCs = DCs + q[k * UV_T] * hz; }
    const double met = v2(D, comp == 0 ? FID(on_u) : FID(om_v))(i, j), Davg = v2(D, comp == 0 ? FID(DU_avg1) : FID(DV_avg1))(i, j);
    const double c1 = 1.0 / (CF0 * met);
    const double corr = (DCs * met - Davg) * c1;
    const bool sw = comp == 0 && S && j == b.Jstr, nw = comp == 0 && Nn && j == b.Jend;
    for (int k = 1; k <= N; ++k) {
      const double val = q[k * UV_T] - corr;
      qn(i, j, k) = val;
      if (sw) qn(i, j - 1, k) = D.p.gamma2 * val;         // u3dbc_im.F:329-343,415-429 (gamma2 slip)
      if (nw) qn(i, j + 1, k) = D.p.gamma2 * val;
    }
  }
  // v3dbc_im.F:171-178,250-257: zero normal flow at the closed walls
  if (comp == 1 && i >= b.Istr && i <= b.Iend) {
    V3 vn = v3l(D, FID(v), nnew);
    if (S && j == b.Jstr) for (int k = 1; k <= N; ++k) vn(i, b.Jstr, k) = 0.0;
    if (Nn && j == b.Jend) for (int k = 1; k <= N; ++k) vn(i, b.Jend + 1, k) = 0.0;
  }
  if (bad) atomicOr(D.err, 1);
}
// pass 2: 2D/3D coupling on JstrT..JendT rows: ubar,vbar(1:2), time-centred Huon/Hvom (:1312-1756)
__global__ void __launch_bounds__(UV_T) step3d_uv2_kernel(const Dev D, Box bx, int nnew) {
  extern __shared__ double uvsm[];
  const int i = bx.i0 + blockIdx.x * blockDim.x + threadIdx.x, j = bx.j0 + blockIdx.y * blockDim.y + threadIdx.y;
  if (i > bx.i1 || j > bx.j1) return;
  const roms_b200_bounds& b = D.b; const int N = b.N;
  V3 Hz = v3(D, FID(Hz));
  const int tid = threadIdx.y * blockDim.x + threadIdx.x, L = (N + 2) * UV_T;
  double *DC = uvsm + tid, *qk = DC + L, *hq = qk + L;     // element k at [k * UV_T]
  const int comp = blockIdx.z;
  if ((comp == 0 && (i >= b.IstrP && i <= b.IendT && j >= b.JstrT && j <= b.JendT)) ||
      (comp == 1 && (i >= b.IstrT && i <= b.IendT && j >= b.Jstr && j <= b.JendT))) {
    const int di = comp == 0 ? 1 : 0, dj = 1 - di;
    V3 qn = v3l(D, comp == 0 ? FID(u) : FID(v), nnew), Hq = v3(D, comp == 0 ? FID(Huon) : FID(Hvom));
    const double met = v2(D, comp == 0 ? FID(on_u) : FID(om_v))(i, j);
    const double Davg1 = v2(D, comp == 0 ? FID(DU_avg1) : FID(DV_avg1))(i, j), Davg2 = v2(D, comp == 0 ? FID(DU_avg2) : FID(DV_avg2))(i, j);
    double DC0 = 0.0, CF0 = 0.0, FC0 = 0.0;
    const double cff = 0.5 * met;
    for (int k0 = 1; k0 <= N; k0 += UV_KB) {
      double ha[UV_KB], hb[UV_KB], qv[UV_KB], hv[UV_KB];
#pragma unroll
      for (int t = 0; t < UV_KB; ++t) {
        const int k = min(k0 + t, N);
        ha[t] = Hz(i, j, k); hb[t] = Hz(i - di, j - dj, k); qv[t] = qn(i, j, k); hv[t] = Hq(i, j, k);
      }
#pragma unroll
      for (int t = 0; t < UV_KB; ++t) {
        const int k = k0 + t;
        if (k <= N) {
          const double dc = cff * (ha[t] + hb[t]);
          DC[k * UV_T] = dc; qk[k * UV_T] = qv[t]; hq[k * UV_T] = hv[t];     // hq holds Huon/Hvom(k) until the next sweep
          DC0 = DC0 + dc;
          CF0 = CF0 + dc * qv[t];
        }
      }
    }
    DC0 = 1.0 / DC0;
    CF0 = DC0 * (CF0 - Davg1);
    const double bar = DC0 * Davg1;
    st(D, v2l(D, comp == 0 ? FID(ubar) : FID(vbar), 1), i, j, bar);
    st(D, v2l(D, comp == 0 ? FID(ubar) : FID(vbar), 2), i, j, bar);
    // boundary rows keep the barotropic mean consistent (step3d_uv.F:1383-1398, 1603-1618)
    bool fix = false;
    if (!b.NSperiodic) {
      if (comp == 0 && (j == 0 || j == b.Mm + 1) && i >= b.IstrU && i <= b.Iend) fix = true;
      if (comp == 1 && (j == 1 || j == b.Mm + 1) && i >= b.Istr && i <= b.Iend) fix = true;
    }
    if (fix) for (int k = 1; k <= N; ++k) qk[k * UV_T] = qk[k * UV_T] - CF0;
    for (int k = N; k >= 1; --k) {
      const double h = 0.5 * (hq[k * UV_T] + qk[k * UV_T] * DC[k * UV_T]);
      hq[k * UV_T] = h;
      FC0 = FC0 + h;
    }
    FC0 = DC0 * (FC0 - Davg2);
    for (int k = 1; k <= N; ++k) {
      st(D, Hq, i, j, k, hq[k * UV_T] - DC[k * UV_T] * FC0);
      st(D, qn, i, j, k, qk[k * UV_T]);
    }
  }
}
int k_step3d_uv(roms_b200_ctx* c, int nrhs, int nstp, int nnew, int iic, int ntfirst) {
  (void)nstp;
  const roms_b200_bounds& b = c->D.b; const double dt = c->D.p.dt;
  double cffab;
  if (iic == ntfirst) cffab = 0.25 * dt; else if (iic == ntfirst + 1) cffab = 0.25 * dt * 3.0 / 2.0; else cffab = 0.25 * dt * 23.0 / 12.0;
  const size_t sm1 = (size_t)6 * (b.N + 2) * UV_T * sizeof(double), sm2 = (size_t)3 * (b.N + 2) * UV_T * sizeof(double);
  static AttrOnce set1, set2;
  if (set1.need(sm1)) CUDA_OK(cudaFuncSetAttribute(step3d_uv1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
  if (set2.need(sm2)) CUDA_OK(cudaFuncSetAttribute(step3d_uv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk(32, UV_T / 32);
  dim3 g1 = grid2(bx, blk); g1.z = 2;
  step3d_uv1_kernel<<<g1, blk, sm1, c->stream>>>(c->D, bx, nrhs, nnew, cffab); c->launches++;
  Box b2{b.IstrT, b.IendT, b.JstrT, b.JendT};
  dim3 g2 = grid2(b2, blk); g2.z = 2;
  step3d_uv2_kernel<<<g2, blk, sm2, c->stream>>>(c->D, b2, nnew); c->launches++;
  return 0;
}
