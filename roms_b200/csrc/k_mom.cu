// roms_b200/csrc/k_mom.cu -- baroclinic pressure gradient (prsgrd32), 3-D momentum
// right-hand side (rhs3d_tile), harmonic viscosity (uv3dmix2_s) and the momentum
// corrector (step3d_uv).  Column-marching threads, i fastest (coalesced rows).
#include "common.cuh"

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
__device__ __forceinline__ double q_uee(const RQ& Q, int i, int j, int k) {
  int jj = j; if (Q.S && j == Q.Jstr - 1) jj = Q.Jstr; if (Q.N && j == Q.Jend + 1) jj = Q.Jend;
  return Q.u(i, jj - 1, k) - 2.0 * Q.u(i, jj, k) + Q.u(i, jj + 1, k);
}
__device__ __forceinline__ double q_vee(const RQ& Q, int i, int j, int k) {
  int jj = j; if (Q.S && j == Q.Jstr) jj = Q.Jstr + 1; if (Q.N && j == Q.Jend + 1) jj = Q.Jend;
  return Q.v(i, jj - 1, k) - 2.0 * Q.v(i, jj, k) + Q.v(i, jj + 1, k);
}
__device__ __forceinline__ double q_Hvee(const RQ& Q, int i, int j, int k) {
  int jj = j; if (Q.S && j == Q.Jstr) jj = Q.Jstr + 1; if (Q.N && j == Q.Jend + 1) jj = Q.Jend;
  return Q.Hvom(i, jj - 1, k) - 2.0 * Q.Hvom(i, jj, k) + Q.Hvom(i, jj + 1, k);
}
__device__ __forceinline__ double q_uxx(const RQ& Q, int i, int j, int k) { return Q.u(i - 1, j, k) - 2.0 * Q.u(i, j, k) + Q.u(i + 1, j, k); }
__device__ __forceinline__ double q_Huxx(const RQ& Q, int i, int j, int k) { return Q.Huon(i - 1, j, k) - 2.0 * Q.Huon(i, j, k) + Q.Huon(i + 1, j, k); }
__device__ __forceinline__ double q_vxx(const RQ& Q, int i, int j, int k) { return Q.v(i - 1, j, k) - 2.0 * Q.v(i, j, k) + Q.v(i + 1, j, k); }
__device__ __forceinline__ double q_Hvxx(const RQ& Q, int i, int j, int k) { return Q.Hvom(i - 1, j, k) - 2.0 * Q.Hvom(i, j, k) + Q.Hvom(i + 1, j, k); }
__device__ __forceinline__ double q_Huee(const RQ& Q, int i, int j, int k) { return Q.Huon(i, j - 1, k) - 2.0 * Q.Huon(i, j, k) + Q.Huon(i, j + 1, k); }
#define GADV (-0.25)
// UFx at rho-point (i,j): rhs3d.F:767-784
__device__ __forceinline__ double q_UFx(const RQ& Q, int i, int j, int k) {
  const double c1 = Q.u(i, j, k) + Q.u(i + 1, j, k);
  const double c = (c1 > 0.0) ? q_uxx(Q, i, j, k) : q_uxx(Q, i + 1, j, k);
  return 0.25 * (c1 + GADV * c) * (Q.Huon(i, j, k) + Q.Huon(i + 1, j, k) + GADV * 0.5 * (q_Huxx(Q, i, j, k) + q_Huxx(Q, i + 1, j, k)));
}
// UFe at psi-point (i,j): rhs3d.F:822-840
__device__ __forceinline__ double q_UFe(const RQ& Q, int i, int j, int k) {
  const double c1 = Q.u(i, j, k) + Q.u(i, j - 1, k), c2 = Q.Hvom(i, j, k) + Q.Hvom(i - 1, j, k);
  const double c = (c2 > 0.0) ? q_uee(Q, i, j - 1, k) : q_uee(Q, i, j, k);
  return 0.25 * (c1 + GADV * c) * (c2 + GADV * 0.5 * (q_Hvxx(Q, i, j, k) + q_Hvxx(Q, i - 1, j, k)));
}
// VFx at psi-point (i,j): rhs3d.F:871-889
__device__ __forceinline__ double q_VFx(const RQ& Q, int i, int j, int k) {
  const double c1 = Q.v(i, j, k) + Q.v(i - 1, j, k), c2 = Q.Huon(i, j, k) + Q.Huon(i, j - 1, k);
  const double c = (c2 > 0.0) ? q_vxx(Q, i - 1, j, k) : q_vxx(Q, i, j, k);
  return 0.25 * (c1 + GADV * c) * (c2 + GADV * 0.5 * (q_Huee(Q, i, j, k) + q_Huee(Q, i, j - 1, k)));
}
// VFe at rho-point (i,j): rhs3d.F:924-942
__device__ __forceinline__ double q_VFe(const RQ& Q, int i, int j, int k) {
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
int k_rhs3d_tile(roms_b200_ctx* c, int nrhs) {
  const roms_b200_bounds& b = c->D.b;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk(32, 8); dim3 g = grid2(bx, blk); g.z = 2 * b.N;
  rhs3d_kernel<<<g, blk, 0, c->stream>>>(c->D, bx, nrhs); c->launches++;
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
int k_uv3dmix2(roms_b200_ctx* c, int nrhs, int nnew) {
  const roms_b200_bounds& b = c->D.b;
  if (!c->D.kpp4) return 1;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk(32, 8); dim3 g = grid2(bx, blk); g.z = 2 * b.N;
  uv3dmix2_kernel<<<g, blk, 0, c->stream>>>(c->D, bx, nrhs, nnew, c->D.kpp4); c->launches++;
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
  static size_t set1 = 0, set2 = 0;
  if (sm1 > set1) { CUDA_OK(cudaFuncSetAttribute(step3d_uv1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1)); set1 = sm1; }
  if (sm2 > set2) { CUDA_OK(cudaFuncSetAttribute(step3d_uv2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2)); set2 = sm2; }
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk(32, UV_T / 32);
  dim3 g1 = grid2(bx, blk); g1.z = 2;
  step3d_uv1_kernel<<<g1, blk, sm1, c->stream>>>(c->D, bx, nrhs, nnew, cffab); c->launches++;
  Box b2{b.IstrT, b.IendT, b.JstrT, b.JendT};
  dim3 g2 = grid2(b2, blk); g2.z = 2;
  step3d_uv2_kernel<<<g2, blk, sm2, c->stream>>>(c->D, b2, nnew); c->launches++;
  return 0;
}
