// roms_b200/csrc/k_kpp.cu -- in-loop physics of the BENCHMARK option set that is
// column-local: KPP vertical mixing (lmd_vmix + lmd_skpp + lmd_finish), COARE
// bulk fluxes (bulk_flux) and the analytical forcing of set_data.
// These use exp/log/pow/atan: results agree with the CPU restatement to a few
// ulp per call (device libm vs glibc), not bit-for-bit.
#include "common.cuh"
#include <cmath>

#define VONKAR 0.41              /* mod_scalars.F:469 */
// mod_scalars.F:1635-1712
#define LMD_RI0 0.7
#define LMD_BVFCON (-2.0e-5)
#define LMD_NU0C 0.01
#define LMD_NU0M 10.0e-4
#define LMD_NU0S 10.0e-4
#define LMD_CSTAR 10.0
#define LMD_CV 1.25
#define LMD_RIC 0.3
#define LMD_AM 1.257
#define LMD_AS (-28.86)
#define LMD_BETAT (-0.2)
#define LMD_CEKMAN 0.7
#define LMD_CMONOB 1.0
#define LMD_CM 8.36
#define LMD_CS 98.96
#define LMD_EPSILON 0.1
#define LMD_ZETAM (-0.2)
#define LMD_ZETAS (-1.0)

__constant__ double k_mu1[9] = {0.35, 0.6, 1.0, 1.5, 1.4, 0.42, 0.37, 0.33, 0.00468592};
__constant__ double k_mu2[9] = {23.0, 20.0, 17.0, 14.0, 7.9, 5.13, 3.54, 2.34, 1.51};
__constant__ double k_r1[9] = {0.58, 0.62, 0.67, 0.77, 0.78, 0.57, 0.57, 0.57, 0.55};
__device__ __forceinline__ double k_swfrac(int Jindex, double Z) {     // lmd_swfrac.F, Zscale=-1
  const double fac1 = -1.0 / k_mu1[Jindex - 1], fac2 = -1.0 / k_mu2[Jindex - 1], fac3 = k_r1[Jindex - 1];
  return exp(Z * fac1) * fac3 + exp(Z * fac2) * (1.0 - fac3);
}
// turbulent velocity scales (lmd_skpp.F, three inlined copies)
__device__ __forceinline__ void wscale(double Ustar, double Ustar3, double zetahat, double zetapar, double& wm, double& ws) {
  const double r3 = 1.0 / 3.0;
  if (zetahat >= 0.0) { wm = VONKAR * Ustar / (1.0 + 5.0 * zetapar); ws = wm; }
  else {
    if (zetapar > LMD_ZETAM) wm = VONKAR * Ustar * pow(1.0 - 16.0 * zetapar, 0.25);
    else wm = VONKAR * pow(LMD_AM * Ustar3 - LMD_CM * zetahat, r3);
    if (zetapar > LMD_ZETAS) ws = VONKAR * Ustar * pow(1.0 - 16.0 * zetapar, 0.5);
    else ws = VONKAR * pow(LMD_AS * Ustar3 - LMD_CS * zetahat, r3);
  }
}
// ---- KPP: lmd_vmix_tile (lmd_vmix.F:99-434) + lmd_skpp_tile (lmd_skpp.F) + lmd_finish_tile (lmd_vmix.F:437-660) ----
// Four kernels; only what is a vertical recurrence or a search runs one thread per water column, the rest runs one thread
// per (i,j,k).  Every point is evaluated with the operations (and their order) of the reference.
//   1 kpp_spline  (column)  spline derivatives dR (of pden), dU, dV: forward/backward recurrence (lmd_skpp.F, RI_SPLINES)
//   2 kpp_levels  (i,j,k)   Bflux, initial ghats, interior Richardson-number mixing, bulk Richardson function FC
//   3 kpp_sbl     (column)  boundary-layer depth hsbl (first zero crossing of FC from the surface, Ekman / Monin-Obukhov
//                           limits), ksbl, shape-function constants G1, dG1dS at the base of the layer
//   4 kpp_finish  (i,j,k)   boundary-layer profiles above ksbl, nonlocal flux ghats, convective adjustment, lateral
//                           conditions (bc_w3d gradient rows + periodic images)
// Scratch: D.kpp4 = {dR, dU, dV, FC} (ni,nj,0:N) each, D.swdk = Bflux (ni,nj,0:N), D.scratch2 planes 0..5 = G constants.
// (The reference also splines rho for lmd_vmix's Rig but then uses bvf, lmd_vmix.F:253: that solve has no effect.)
namespace {
struct KppC { double lmd_Cg, Vtc; };      // mod_scalars.F:4592 ; lmd_skpp.F Vtc -- evaluated once on the host
__device__ __forceinline__ V3 scr3(const Dev& D, double* p) { return V3{p, D.b.LBi, D.ni, D.b.LBj, D.nj, 0}; }
__device__ __forceinline__ V2 scr2(const Dev& D, int plane) { return V2{D.scratch2 + D.nij * plane, D.b.LBi, D.ni, D.b.LBj}; }
__device__ __forceinline__ double kpp_ustar(const Dev& D, int i, int j) {
  V2 sustr = v2(D, FID(sustr)), svstr = v2(D, FID(svstr));
  const double ta = 0.5 * (sustr(i, j) + sustr(i + 1, j)), tb = 0.5 * (svstr(i, j) + svstr(i, j + 1));
  return sqrt(sqrt(ta * ta + tb * tb));
}
struct KppSurf { double Bo, Bosol, stT, stS, srf; int Jw; };
__device__ __forceinline__ KppSurf kpp_surf(const Dev& D, int i, int j) {
  KppSurf q;
  const double al = v2(D, FID(alpha))(i, j), be = v2(D, FID(beta))(i, j);
  q.srf = v2(D, FID(srflx))(i, j); q.stT = v2l(D, FID(stflx), 1)(i, j); q.stS = v2l(D, FID(stflx), 2)(i, j);
  q.Bo = D.p.g * (al * (q.stT - q.srf) - be * q.stS); q.Bosol = D.p.g * al * q.srf;
  q.Jw = (int)v2(D, FID(Jwtype))(i, j);
  return q;
}
}  // namespace

__global__ void __launch_bounds__(128) kpp_spline_kernel(const Dev D, Box bx, int nstp) {
  IJ_FROM_BOX(bx);
  const int N = D.b.N; const size_t vol = D.nij * (size_t)(N + 1);
  V3 Hz = v3(D, FID(Hz)), pden = v3(D, FID(pden)), u = v3l(D, FID(u), nstp), v = v3l(D, FID(v), nstp);
  V3 sR = scr3(D, D.kpp4), sU = scr3(D, D.kpp4 + vol), sV = scr3(D, D.kpp4 + 2 * vol);
  double FC[RB_MAXN + 1], dR[RB_MAXN + 1], dU[RB_MAXN + 1], dV[RB_MAXN + 1];
  FC[0] = 0.0; dR[0] = 0.0; dU[0] = 0.0; dV[0] = 0.0;
  double hzk = Hz(i, j, 1), dk = pden(i, j, 1), uk = u(i, j, 1), upk = u(i + 1, j, 1), vk = v(i, j, 1), vpk = v(i, j + 1, 1);
  double fcp = 0.0, drp = 0.0, dup = 0.0, dvp = 0.0;
  constexpr int KB = 6;                         // levels per load batch
  for (int k0 = 1; k0 <= N - 1; k0 += KB) {
    double hzn[KB], dn[KB], un[KB], upn[KB], vn[KB], vpn[KB];
#pragma unroll
    for (int q = 0; q < KB; ++q) {
      const int kk = min(k0 + q + 1, N);
      hzn[q] = Hz(i, j, kk); dn[q] = pden(i, j, kk); un[q] = u(i, j, kk); upn[q] = u(i + 1, j, kk); vn[q] = v(i, j, kk); vpn[q] = v(i, j + 1, kk);
    }
#pragma unroll
    for (int q = 0; q < KB; ++q) {
      const int k = k0 + q;
      if (k <= N - 1) {
        const double cff = 1.0 / (2.0 * hzn[q] + hzk * (2.0 - fcp));
        fcp = cff * hzn[q];
        drp = cff * (6.0 * (dn[q] - dk) - hzk * drp);
        dup = cff * (3.0 * (un[q] - uk + upn[q] - upk) - hzk * dup);
        dvp = cff * (3.0 * (vn[q] - vk + vpn[q] - vpk) - hzk * dvp);
        FC[k] = fcp; dR[k] = drp; dU[k] = dup; dV[k] = dvp;
        hzk = hzn[q]; dk = dn[q]; uk = un[q]; upk = upn[q]; vk = vn[q]; vpk = vpn[q];
      }
    }
  }
  double rn = 0.0, un_ = 0.0, vn_ = 0.0;         // dR, dU, dV at k+1
  sR(i, j, N) = 0.0; sU(i, j, N) = 0.0; sV(i, j, N) = 0.0;
  for (int k = N - 1; k >= 1; --k) {
    const double fc = FC[k];
    rn = dR[k] - fc * rn; un_ = dU[k] - fc * un_; vn_ = dV[k] - fc * vn_;
    sR(i, j, k) = rn; sU(i, j, k) = un_; sV(i, j, k) = vn_;
  }
  sR(i, j, 0) = 0.0; sU(i, j, 0) = 0.0; sV(i, j, 0) = 0.0;
}

__global__ void __launch_bounds__(256) kpp_levels_kernel(const Dev D, Box bx, int nstp, KppC kc) {
  IJZ_FROM_BOX(bx, D.b.N + 1);
  const int N = D.b.N, k = zlev; const size_t vol = D.nij * (size_t)(N + 1);
  const double gorho0 = D.p.g / D.p.rho0;
  V3 Hz = v3(D, FID(Hz)), z_w = v3(D, FID(z_w)), pden = v3(D, FID(pden)), bvf = v3(D, FID(bvf));
  V3 u = v3l(D, FID(u), nstp), v = v3l(D, FID(v), nstp);
  V3 Akv = v3(D, FID(Akv)), Akt1 = v3l(D, FID(Akt), 1), gh1 = v3l(D, FID(ghats), 1), gh2 = v3l(D, FID(ghats), 2);
  V3 sR = scr3(D, D.kpp4), sU = scr3(D, D.kpp4 + vol), sV = scr3(D, D.kpp4 + 2 * vol), sF = scr3(D, D.kpp4 + 3 * vol), sB = scr3(D, D.swdk);
  const double zwN = z_w(i, j, N), zwk = z_w(i, j, k);
  const KppSurf q = kpp_surf(D, i, j);
  // ---- surface buoyancy flux profile and the initial nonlocal-flux factors (lmd_skpp.F)
  const double swdk = k_swfrac(q.Jw, zwN - zwk);
  const double Bf = (q.Bo + q.Bosol * (1.0 - swdk));
  sB(i, j, k) = Bf;
  {
    const double cff = 1.0 - (0.5 + copysign(0.5, Bf));
    gh1(i, j, k) = -cff * (q.stT - q.srf + q.srf * (1.0 - swdk));
    gh2(i, j, k) = cff * q.stS;
  }
  // ---- interior shear / internal-wave mixing (lmd_vmix.F:232-330); Akt(:,:,:,isalt) = Akt(:,:,:,itemp) is set by kpp_finish
  if (k >= 1 && k <= N - 1) {
    const double eps = 1.0e-14, bv = bvf(i, j, k), dUk = sU(i, j, k), dVk = sV(i, j, k);
    double shear2 = dUk * dUk + dVk * dVk;
    const double Rig = bv / (shear2 + eps);
    double cff = fmin(1.0, fmax(0.0, Rig) / LMD_RI0);
    double nu_sx = 1.0 - cff * cff;
    nu_sx = nu_sx * nu_sx * nu_sx;
    shear2 = bv / (Rig + eps);
    cff = shear2 * shear2 / (shear2 * shear2 + 16.0e-10);
    nu_sx = cff * nu_sx;
    cff = 1.0 / sqrt(fmax(bv, 1.0e-7));
    Akv(i, j, k) = 1.0e-6 * cff + LMD_NU0M * nu_sx;
    Akt1(i, j, k) = 1.0e-7 * cff + LMD_NU0S * nu_sx;
  }
  // ---- bulk Richardson function FC(k) = Ritop - Ric*Ribot between the surface reference and level k+1 (lmd_skpp.F)
  if (k == N) { sF(i, j, N) = 0.0; return; }
  {
    const double small = 1.0e-20, c13 = 1.0 / 3.0, c16 = 1.0 / 6.0;
    const int kk = k + 1;
    const double sl_dpth = LMD_EPSILON * (zwN - v2(D, FID(hsbl))(i, j));          // hsbl of the previous step
    const double Ustar = kpp_ustar(D, i, j), Ustar3 = Ustar * Ustar * Ustar;
    const double hzN = Hz(i, j, N);
    const double Rref = pden(i, j, N) + hzN * (c13 * sR(i, j, N) + c16 * sR(i, j, N - 1));
    const double Uref = 0.5 * (u(i, j, N) + u(i + 1, j, N)) + hzN * (c13 * sU(i, j, N) + c16 * sU(i, j, N - 1));
    const double Vref = 0.5 * (v(i, j, N) + v(i, j + 1, N)) + hzN * (c13 * sV(i, j, N) + c16 * sV(i, j, N - 1));
    const double depth = zwN - zwk;
    const double sigma = (Bf < 0.0) ? fmin(sl_dpth, depth) : depth;
    const double zetahat = VONKAR * sigma * Bf, zetapar = zetahat / (Ustar3 + small);
    double wm, ws;
    wscale(Ustar, Ustar3, zetahat, zetapar, wm, ws);
    const double hzk = Hz(i, j, kk);
    const double Rk = pden(i, j, kk) - hzk * (c13 * sR(i, j, kk - 1) + c16 * sR(i, j, kk));
    const double Uk = 0.5 * (u(i, j, kk) + u(i + 1, j, kk)) - hzk * (c13 * sU(i, j, kk - 1) + c16 * sU(i, j, kk));
    const double Vk = 0.5 * (v(i, j, kk) + v(i, j + 1, kk)) - hzk * (c13 * sV(i, j, kk - 1) + c16 * sV(i, j, kk));
    const double Ritop = -gorho0 * (Rref - Rk) * depth;
    const double dUr = Uref - Uk, dVr = Vref - Vk;
    const double Ribot = dUr * dUr + dVr * dVr + kc.Vtc * depth * ws * sqrt(fabs(bvf(i, j, kk - 1)));
    sF(i, j, k) = Ritop - LMD_RIC * Ribot;
  }
}

__global__ void __launch_bounds__(128) kpp_sbl_kernel(const Dev D, Box bx) {
  IJ_FROM_BOX(bx);
  const roms_b200_bounds& b = D.b; const int N = b.N; const size_t vol = D.nij * (size_t)(N + 1);
  const bool south = b.Southern_Edge && !b.NSperiodic && j == b.Jstr, north = b.Northern_Edge && !b.NSperiodic && j == b.Jend;
  V3 z_w = v3(D, FID(z_w)), Akv = v3(D, FID(Akv)), Akt1 = v3l(D, FID(Akt), 1), sF = scr3(D, D.kpp4 + 3 * vol);
  V2 hsbl = v2(D, FID(hsbl));
  const double eps = 1.0e-10, small = 1.0e-20;
  const double zwN = z_w(i, j, N);
  const double Ustar = kpp_ustar(D, i, j), Ustar3 = Ustar * Ustar * Ustar;
  const KppSurf q = kpp_surf(D, i, j);
  // first zero crossing of FC from the surface (k = N..2), linear interpolation between the two w-levels
  int ks = 1; double hs = z_w(i, j, 1);
  {
    constexpr int KB = 6;
    double fup = sF(i, j, N), zup = zwN;
    for (int k0 = N; k0 >= 2 && ks == 1; k0 -= KB) {
      double f[KB], z[KB];
#pragma unroll
      for (int r = 0; r < KB; ++r) { const int km = max(k0 - r - 1, 1); f[r] = sF(i, j, km); z[r] = z_w(i, j, km); }
#pragma unroll
      for (int r = 0; r < KB; ++r) {
        const int k = k0 - r;
        if (k >= 2 && ks == 1) {
          if (f[r] > 0.0) { hs = (zup * f[r] - z[r] * fup) / (f[r] - fup); ks = k; }
          fup = f[r]; zup = z[r];
        }
      }
    }
  }
  double Bfsfc = (q.Bo + q.Bosol * (1.0 - k_swfrac(q.Jw, zwN - hs)));
  if (Ustar > 0.0 && Bfsfc > 0.0) {
    const double hekman = LMD_CEKMAN * Ustar / fmax(fabs(v2(D, FID(f))(i, j)), eps);
    const double hmonob = LMD_CMONOB * Ustar * Ustar * Ustar / fmax(VONKAR * Bfsfc, eps);
    hs = (zwN - fmin(fmin(hekman, hmonob), zwN - hs));
  }
  hs = fmin(hs, zwN);
  hs = fmax(hs, z_w(i, j, 0));
  st(D, hsbl, i, j, hs);                               // bc_r2d_tile: gradient + periodic images
  if (south) st(D, hsbl, i, j - 1, hs);
  if (north) st(D, hsbl, i, j + 1, hs);
  ks = 1;
  {
    constexpr int KB = 6;
    for (int k0 = N; k0 >= 2 && ks == 1; k0 -= KB) {
      double z[KB];
#pragma unroll
      for (int r = 0; r < KB; ++r) z[r] = z_w(i, j, max(k0 - r - 1, 1));
#pragma unroll
      for (int r = 0; r < KB; ++r) { const int k = k0 - r; if (k >= 2 && ks == 1 && z[r] < hs) ks = k; }
    }
  }
  Bfsfc = (q.Bo + q.Bosol * (1.0 - k_swfrac(q.Jw, zwN - hs)));
  double wm, ws;
  {
    const double cff = (Bfsfc > 0.0) ? 1.0 : LMD_EPSILON;
    const double sigma = cff * (zwN - hs);
    const double zetahat = VONKAR * sigma * Bfsfc, zetapar = zetahat / (Ustar3 + small);
    wscale(Ustar, Ustar3, zetahat, zetapar, wm, ws);
  }
  const double f1 = 5.0 * fmax(0.0, Bfsfc) * VONKAR / (Ustar * Ustar * Ustar * Ustar + eps);
  double Gm1, Gt1, Gs1, dGm1dS, dGt1dS, dGs1dS;
  const double zbl = zwN - hs;
  if (hs > z_w(i, j, 1)) {
    const int k = ks;
    const double zk = z_w(i, j, k), zkm = z_w(i, j, k - 1);
    const double akvk = Akv(i, j, k), akvm = Akv(i, j, k - 1), aktk = Akt1(i, j, k), aktm = Akt1(i, j, k - 1);
    // salinity: interior values equal the temperature ones; at k = N the (untouched) surface value of Akt(:,:,N,isalt)
    const double aksk = (k == N) ? v3l(D, FID(Akt), 2)(i, j, N) : aktk, aksm = aktm;
    const double cff = 1.0 / (zk - zkm);
    const double cff_dn = cff * (hs - zkm), cff_up = cff * (zk - hs);
    double K_bl = cff_dn * akvk + cff_up * akvm, dK_bl = cff * (akvk - akvm);
    Gm1 = K_bl / (zbl * wm + eps); dGm1dS = fmin(0.0, -dK_bl / (wm + eps) - K_bl * f1);
    K_bl = cff_dn * aktk + cff_up * aktm; dK_bl = cff * (aktk - aktm);
    Gt1 = K_bl / (zbl * ws + eps); dGt1dS = fmin(0.0, -dK_bl / (ws + eps) - K_bl * f1);
    K_bl = cff_dn * aksk + cff_up * aksm; dK_bl = cff * (aksk - aksm);
    Gs1 = K_bl / (zbl * ws + eps); dGs1dS = fmin(0.0, -dK_bl / (ws + eps) - K_bl * f1);
  } else {
    ks = 0;
    V2 bustr = v2(D, FID(bustr)), bvstr = v2(D, FID(bvstr));
    const double a = 0.5 * (bustr(i, j) + bustr(i + 1, j)), bb = 0.5 * (bvstr(i, j) + bvstr(i, j + 1));
    const double Ustarb = sqrt(sqrt(a * a + bb * bb));
    const double dK_bl = VONKAR * Ustarb, K_bl = dK_bl * (hs - z_w(i, j, 0));
    Gm1 = K_bl / (zbl * wm + eps); dGm1dS = fmin(0.0, -dK_bl / (wm + eps) - K_bl * f1);
    Gt1 = K_bl / (zbl * ws + eps); dGt1dS = fmin(0.0, -dK_bl / (ws + eps) - K_bl * f1);
    Gs1 = Gt1; dGs1dS = dGt1dS;
  }
  D.ksbl[(i - b.LBi) + D.ni * (j - b.LBj)] = ks;
  scr2(D, 0)(i, j) = Gm1; scr2(D, 1)(i, j) = Gt1; scr2(D, 2)(i, j) = Gs1;
  scr2(D, 3)(i, j) = dGm1dS; scr2(D, 4)(i, j) = dGt1dS; scr2(D, 5)(i, j) = dGs1dS;
}

__global__ void __launch_bounds__(256) kpp_finish_kernel(const Dev D, Box bx, KppC kc) {
  IJ_FROM_BOX(bx);
  const roms_b200_bounds& b = D.b; const int N = b.N, k = blockIdx.z;
  const bool south = b.Southern_Edge && !b.NSperiodic && j == b.Jstr, north = b.Northern_Edge && !b.NSperiodic && j == b.Jend;
  V3 z_w = v3(D, FID(z_w)), bvf = v3(D, FID(bvf)), sB = scr3(D, D.swdk);
  V3 Akv = v3(D, FID(Akv)), Akt1 = v3l(D, FID(Akt), 1), Akt2 = v3l(D, FID(Akt), 2), gh1 = v3l(D, FID(ghats), 1), gh2 = v3l(D, FID(ghats), 2);
  double akv, akt1, akt2;
  if (k >= 1 && k <= N - 1) {
    const int ks = D.ksbl[(i - b.LBi) + D.ni * (j - b.LBj)];
    double g1, g2;
    if (k > ks) {
      const double eps = 1.0e-10, small = 1.0e-20;
      const double zwN = z_w(i, j, N), hs = v2(D, FID(hsbl))(i, j), zbl = zwN - hs, sl_dpth = LMD_EPSILON * (zwN - hs);
      const double Ustar = kpp_ustar(D, i, j), Ustar3 = Ustar * Ustar * Ustar;
      const double Gm1 = scr2(D, 0)(i, j), Gt1 = scr2(D, 1)(i, j), Gs1 = scr2(D, 2)(i, j);
      const double dGm1dS = scr2(D, 3)(i, j), dGt1dS = scr2(D, 4)(i, j), dGs1dS = scr2(D, 5)(i, j);
      const double Bf = sB(i, j, k);
      const double depth = zwN - z_w(i, j, k);
      double sigma = (Bf < 0.0) ? fmin(sl_dpth, depth) : depth;
      const double zetahat = VONKAR * sigma * Bf, zetapar = zetahat / (Ustar3 + small);
      double wm, ws;
      wscale(Ustar, Ustar3, zetahat, zetapar, wm, ws);
      sigma = depth / (zbl + eps);
      const double a1 = sigma - 2.0, a2 = 3.0 - 2.0 * sigma, a3 = sigma - 1.0;
      const double Gm = a1 + a2 * Gm1 + a3 * dGm1dS, Gt = a1 + a2 * Gt1 + a3 * dGt1dS, Gs = a1 + a2 * Gs1 + a3 * dGs1dS;
      akv = depth * wm * (1.0 + sigma * Gm);
      akt1 = depth * ws * (1.0 + sigma * Gt);
      akt2 = depth * ws * (1.0 + sigma * Gs);
      const double cff = kc.lmd_Cg * (1.0 - (0.5 + copysign(0.5, Bf))) / (zbl * ws + eps);
      g1 = cff * gh1(i, j, k); g2 = cff * gh2(i, j, k);
    } else { akv = Akv(i, j, k); akt1 = Akt1(i, j, k); akt2 = akt1; g1 = 0.0; g2 = 0.0; }
    gh1(i, j, k) = g1; gh2(i, j, k) = g2;
    // lmd_finish: convective adjustment
    double cff = fmax(bvf(i, j, k), LMD_BVFCON);
    cff = fmin(1.0, (LMD_BVFCON - cff) / LMD_BVFCON);
    double nu_sxc = 1.0 - cff * cff;
    nu_sxc = nu_sxc * nu_sxc * nu_sxc;
    akv = akv + LMD_NU0C * nu_sxc; akt1 = akt1 + LMD_NU0C * nu_sxc; akt2 = akt2 + LMD_NU0C * nu_sxc;
  } else { akv = Akv(i, j, k); akt1 = Akt1(i, j, k); akt2 = Akt2(i, j, k); }
  // lateral conditions (bc_w3d: gradient + periodic images)
  st(D, Akv, i, j, k, akv); st(D, Akt1, i, j, k, akt1); st(D, Akt2, i, j, k, akt2);
  if (south) { st(D, Akv, i, j - 1, k, akv); st(D, Akt1, i, j - 1, k, akt1); st(D, Akt2, i, j - 1, k, akt2); }
  if (north) { st(D, Akv, i, j + 1, k, akv); st(D, Akt1, i, j + 1, k, akt1); st(D, Akt2, i, j + 1, k, akt2); }
}
// part: 1 = the spline derivatives of pden, u, v only (they need rho_eos, not the surface fluxes: main3d runs them beside
// bulk_flux / set_vbc on the second stream), 2 = everything else, 3 = both
int k_lmd_vmix_part(roms_b200_ctx* c, int nstp, int part) {
  const roms_b200_bounds& b = c->D.b;
  if (!c->D.kpp4 || !c->D.swdk) { fprintf(stderr, "roms_b200: lmd_vmix needs the BENCHMARK option set (KPP scratch not allocated)\n"); return 1; }
  if (b.N < 3) return 1;
  static const KppC kc = {LMD_CSTAR * VONKAR * pow(LMD_CS * VONKAR * LMD_EPSILON, 1.0 / 3.0),
                          LMD_CV * sqrt(-LMD_BETAT) / (sqrt(LMD_CS * LMD_EPSILON) * LMD_RIC * VONKAR * VONKAR)};
  // With neighbour tiles the column physics is also evaluated on the first ring of halo points (same inputs, same bits as the
  // neighbour's interior): step3d_uv and pre_step3d read Akv at (i-1,j), (i,j-1), so the mp_exchange3d of Akv
  // (lmd_vmix.F:640-650) needs no message.  Inputs reach one point further (u(i+1), sustr(i+1), ...): see k_bulk_flux / k_set_vbc.
  const Dev De = widened(c, 1); const roms_b200_bounds& e = De.b;
  Box bx{e.Istr, e.Iend, e.Jstr, e.Jend};
  dim3 blkc(32, 4), blkl(64, 4);
  dim3 gc = grid2(bx, blkc), gl = grid2(bx, blkl); gl.z = b.N + 1;
  if (part & 1) { kpp_spline_kernel<<<gc, blkc, 0, c->stream>>>(De, bx, nstp); c->launches++; }
  if (part & 2) {
    kpp_levels_kernel<<<gl, blkl, 0, c->stream>>>(De, bx, nstp, kc); c->launches++;
    kpp_sbl_kernel<<<gc, blkc, 0, c->stream>>>(De, bx); c->launches++;
    kpp_finish_kernel<<<gl, blkl, 0, c->stream>>>(De, bx, kc); c->launches++;
  }
  return 0;
}
int k_lmd_vmix(roms_b200_ctx* c, int nstp) { return k_lmd_vmix_part(c, nstp, 3); }

// ---- bulk_flux_tile (COARE 3.0), bulk_flux.F -----------------------------------------------------------
#define BLK_CPA 1004.67
#define BLK_CPW 4000.0
#define BLK_RGAS 287.1
#define BLK_ZABL 600.0
#define BLK_BETA 1.2
#define PI_D 3.14159265358979323846
__device__ double bulk_psiu(double ZoL) {
  const double r3 = 1.0 / 3.0;
  if (ZoL < 0.0) {
    const double x = pow(1.0 - 15.0 * ZoL, 0.25);
    const double psik = 2.0 * log(0.5 * (1.0 + x)) + log(0.5 * (1.0 + x * x)) - 2.0 * atan(x) + 0.5 * PI_D;
    double cff = sqrt(3.0);
    const double y = pow(1.0 - 10.15 * ZoL, r3);
    const double psic = 1.5 * log(r3 * (1.0 + y + y * y)) - cff * atan((1.0 + 2.0 * y) / cff) + PI_D / cff;
    cff = ZoL * ZoL;
    const double Fw = cff / (1.0 + cff);
    return (1.0 - Fw) * psik + Fw * psic;
  }
  const double cff = fmin(50.0, 0.35 * ZoL);
  return -((1.0 + ZoL) + 0.6667 * (ZoL - 14.28) / exp(cff) + 8.525);
}
__device__ double bulk_psit(double ZoL) {
  const double r3 = 1.0 / 3.0;
  if (ZoL < 0.0) {
    const double x = pow(1.0 - 15.0 * ZoL, 0.5);
    const double psik = 2.0 * log(0.5 * (1.0 + x));
    double cff = sqrt(3.0);
    const double y = pow(1.0 - 34.15 * ZoL, r3);
    const double psic = 1.5 * log(r3 * (1.0 + y + y * y)) - cff * atan((1.0 + 2.0 * y) / cff) + PI_D / cff;
    cff = ZoL * ZoL;
    const double Fw = cff / (1.0 + cff);
    return (1.0 - Fw) * psik + Fw * psic;
  }
  const double cff = fmin(50.0, 0.35 * ZoL);
  return -(pow(1.0 + 2.0 * ZoL, 1.5) + 0.6667 * (ZoL - 14.28) / exp(cff) + 8.525);
}
__global__ void __launch_bounds__(128) bulk_flux1_kernel(const Dev D, Box bx, int nrhs) {
  IJ_FROM_BOX(bx);
  const roms_b200_bounds& b = D.b; const int N = b.N; const double g = D.p.g, rho0 = D.p.rho0;
  const double eps = 1.0e-20, r3 = 1.0 / 3.0, Cp = 3985.0, StefBo = 5.67e-8, emmiss = 0.97;
  const double ZW = D.p.blk_ZW, ZT = D.p.blk_ZT, ZQ = D.p.blk_ZQ;
  V2 Taux{D.scratch2, b.LBi, D.ni, b.LBj}, Tauy{D.scratch2 + D.nij, b.LBi, D.ni, b.LBj};
  const double Uair = v2(D, FID(Uwind))(i, j), Vair = v2(D, FID(Vwind))(i, j);
  double Hscale = rho0 * Cp;
  const double Wmag = sqrt(Uair * Uair + Vair * Vair);
  const double PairM = v2(D, FID(Pair))(i, j), TairC = v2(D, FID(Tair))(i, j), TairK = TairC + 273.16;
  const double TseaC = v3l(D, FID(t), nrhs, 1)(i, j, N), TseaK = TseaC + 273.16;
  const double RH = v2(D, FID(Hair))(i, j), cloud = v2(D, FID(cloud))(i, j), rain = v2(D, FID(rain))(i, j);
  const double delTc = 0.0, delQc = 0.0;
  double cff = (0.7859 + 0.03477 * TairC) / (1.0 + 0.00412 * TairC);
  const double e_sat = pow(10.0, cff), vap_p = e_sat * RH;
  const double cff2 = TairK * TairK * TairK, cff1 = cff2 * TairK;
  const double LRad = -emmiss * StefBo * (cff1 * (0.39 - 0.05 * sqrt(vap_p)) * (1.0 - 0.6823 * cloud * cloud) + cff2 * 4.0 * (TseaK - TairK));
  cff = (1.0007 + 3.46e-6 * PairM) * 6.1121 * exp(17.502 * TairC / (240.97 + TairC));
  const double Qair = 0.62197 * (cff / (PairM - 0.378 * cff + eps));
  double Q;
  if (RH < 2.0) { cff = cff * RH; Q = 0.62197 * (cff / (PairM - 0.378 * cff + eps)); } else Q = RH / 1000.0;
  cff = (1.0007 + 3.46e-6 * PairM) * 6.1121 * exp(17.502 * TseaC / (240.97 + TseaC));
  cff = cff * 0.98;
  const double Qsea = 0.62197 * (cff / (PairM - 0.378 * cff));
  const double rhoAir = PairM * 100.0 / (BLK_RGAS * TairK * (1.0 + 0.61 * Q));
  const double VisAir = 1.326e-5 * (1.0 + TairC * (6.542e-3 + TairC * (8.301e-6 - 4.84e-9 * TairC)));
  const double Hlv = (2.501 - 0.00237 * TseaC) * 1.0e+6;
  double Wgus = 0.5;
  double delW = sqrt(Wmag * Wmag + Wgus * Wgus);
  const double delQ = Qsea - Q, delT = TseaC - TairC;
  double ZoW = 0.0001;
  const double u10 = delW * log(10.0 / ZoW) / log(ZW / ZoW);
  double Wstar = 0.035 * u10;
  const double Zo10 = 0.011 * Wstar * Wstar / g + 0.11 * VisAir / Wstar;
  const double t0 = VONKAR / log(10.0 / Zo10), Cd10 = t0 * t0, Ch10 = 0.00115, Ct10 = Ch10 / sqrt(Cd10);
  const double ZoT10 = 10.0 / exp(VONKAR / Ct10);
  const double t1 = VONKAR / log(ZW / Zo10), Cd = t1 * t1;
  const double Ct = VONKAR / log(ZT / ZoT10), CC = VONKAR * Ct / Cd;
  const double Ribcu = -ZW / (BLK_ZABL * 0.004 * (BLK_BETA * BLK_BETA * BLK_BETA));
  const double Ri = -g * ZW * ((delT - delTc) + 0.61 * TairK * delQ) / (TairK * delW * delW + eps);
  const double Zetu = (Ri < 0.0) ? CC * Ri / (1.0 + Ri / Ribcu) : CC * Ri / (1.0 + 3.0 * Ri / CC);
  const double L10 = ZW / Zetu;
  Wstar = delW * VONKAR / (log(ZW / Zo10) - bulk_psiu(ZW / L10));
  double Tstar = -(delT - delTc) * VONKAR / (log(ZT / ZoT10) - bulk_psit(ZT / L10));
  double Qstar = -(delQ - delQc) * VONKAR / (log(ZQ / ZoT10) - bulk_psit(ZQ / L10));
  const double charn = fmin(0.028, -0.005 + 0.0017 * delW);
  for (int Iter = 1; Iter <= 3; ++Iter) {
    ZoW = charn * Wstar * Wstar / g + 0.11 * VisAir / (Wstar + eps);
    const double Rr = ZoW * Wstar / VisAir;
    const double ZoQ = fmin(1.6e-4, 5.8e-5 / pow(Rr, 0.72)), ZoT = ZoQ;
    const double ZoL = VONKAR * g * ZW * (Tstar * (1.0 + 0.61 * Q) + 0.61 * TairK * Qstar) / (TairK * Wstar * Wstar * (1.0 + 0.61 * Q) + eps);
    const double L = ZW / (ZoL + eps);
    const double Wpsi = bulk_psiu(ZoL), Tpsi = bulk_psit(ZT / L), Qpsi = bulk_psit(ZQ / L);
    Wstar = fmax(eps, delW * VONKAR / (log(ZW / ZoW) - Wpsi));
    Tstar = -(delT - delTc) * VONKAR / (log(ZT / ZoT) - Tpsi);
    Qstar = -(delQ - delQc) * VONKAR / (log(ZQ / ZoQ) - Qpsi);
    const double Bf = -g / TairK * Wstar * (Tstar + 0.61 * TairK * Qstar);
    Wgus = (Bf > 0.0) ? BLK_BETA * pow(Bf * BLK_ZABL, r3) : 0.2;
    delW = sqrt(Wmag * Wmag + Wgus * Wgus);
  }
  const double Hs = -BLK_CPA * rhoAir * Wstar * Tstar;
  const double diffw = 2.11e-5 * pow(TairK / 273.16, 1.94);
  const double diffh = 0.02411 * (1.0 + TairC * (3.309e-3 - 1.44e-6 * TairC)) / (rhoAir * BLK_CPA + eps);
  cff = Qair * Hlv / (BLK_RGAS * TairK * TairK);
  const double wet_bulb = 1.0 / (1.0 + 0.622 * (cff * Hlv * diffw) / (BLK_CPA * diffh));
  const double Hsr = fabs(rain) * wet_bulb * BLK_CPW * ((TseaC - TairC) + (Qsea - Q) * Hlv / BLK_CPA);
  const double SHeat = (Hs + Hsr);
  const double Hl = -Hlv * rhoAir * Wstar * Qstar;
  const double upvel = -1.61 * Wstar * Qstar - (1.0 + 1.61 * Q) * Wstar * Tstar / TairK;
  const double Hlw = rhoAir * Hlv * upvel * Q;
  const double LHeat = (Hl + Hlw);
  const double Taur = 0.85 * fabs(rain) * Wmag;
  cff = rhoAir * (Wstar * Wstar + Taur / rhoAir) / (Wmag + eps);
  Taux(i, j) = cff * Uair; Tauy(i, j) = cff * Vair;
  if (i >= b.IstrR && i <= b.IendR && j >= b.JstrR && j <= b.JendR) {
    Hscale = 1.0 / (rho0 * Cp);
    const double lr = LRad * Hscale, lh = -LHeat * Hscale, sh = -SHeat * Hscale;
    st(D, v2(D, FID(lrflx)), i, j, lr); st(D, v2(D, FID(lhflx)), i, j, lh); st(D, v2(D, FID(shflx)), i, j, sh);
    st(D, v2l(D, FID(stflux), 1), i, j, (v2(D, FID(srflx))(i, j) + lr + lh + sh));
  }
}
__global__ void bulk_flux2_kernel(const Dev D, Box bx) {
  IJ_FROM_BOX(bx);
  const roms_b200_bounds& b = D.b; const double cff = 0.5 / D.p.rho0;
  V2 Taux{D.scratch2, b.LBi, D.ni, b.LBj}, Tauy{D.scratch2 + D.nij, b.LBi, D.ni, b.LBj};
  if (i >= b.Istr && i <= b.IendR && j >= b.JstrR && j <= b.JendR) st(D, v2(D, FID(sustr)), i, j, cff * (Taux(i - 1, j) + Taux(i, j)));
  if (i >= b.IstrR && i <= b.IendR && j >= b.Jstr && j <= b.JendR) st(D, v2(D, FID(svstr)), i, j, cff * (Tauy(i, j - 1) + Tauy(i, j)));
}
int k_bulk_flux(roms_b200_ctx* c, int nrhs) {
  // With neighbour tiles the fluxes are evaluated two points into the halo (point-wise in the surface state and the forcing,
  // which are valid on the whole halo), so sustr, svstr, stflux need no halo swap: KPP on the first halo ring reads them one
  // point further out.
  const Dev De = widened(c, 2); const roms_b200_bounds& b = De.b;
  Box b1{b.Istr - 1, b.IendR, b.Jstr - 1, b.JendR}; dim3 blk(32, 4);
  bulk_flux1_kernel<<<grid2(b1, blk), blk, 0, c->stream>>>(De, b1, nrhs); c->launches++;
  Box b2{b.IstrR, b.IendR, b.JstrR, b.JendR};
  bulk_flux2_kernel<<<grid2(b2, blk), blk, 0, c->stream>>>(De, b2); c->launches++;
  return 0;
}

// ---- set_data (analytical forcing), set_data.F + Functionals/ana_*.h -----------------------------
struct ForcingArgs { double Dangle, Hangle, Rsolar, windamp; };
__global__ void set_data_kernel(const Dev D, Box bx, ForcingArgs fa) {
  IJ_FROM_BOX(bx);
  const roms_b200_bounds& b = D.b;
  if (D.p.app == ROMS_B200_APP_BENCHMARK) {
    const double deg2rad = PI_D / 180.0;
    st(D, v2(D, FID(cloud)), i, j, 0.6); st(D, v2(D, FID(Tair)), i, j, 4.0); st(D, v2(D, FID(Hair)), i, j, 0.8);
    const double latr = v2(D, FID(latr))(i, j), lonr = v2(D, FID(lonr))(i, j);
    const double LatRad = latr * deg2rad;
    const double cff1 = sin(LatRad) * sin(fa.Dangle), cff2 = cos(LatRad) * cos(fa.Dangle);
    double sr = 0.0;
    const double zenith = cff1 + cff2 * cos(fa.Hangle - lonr * deg2rad);
    if (zenith > 0.0) {
      const double Ta = 4.0, cl = 0.6;
      const double cff = (0.7859 + 0.03477 * Ta) / (1.0 + 0.00412 * Ta);
      const double e_sat = pow(10.0, cff), vap_p = e_sat * 0.8;
      sr = fa.Rsolar * zenith * zenith * (1.0 - 0.6 * (cl * cl * cl)) / ((zenith + 2.7) * vap_p * 1.0e-3 + 1.085 * zenith + 0.1);
    }
    sr = (1.0 - 0.06) * sr;
    st(D, v2(D, FID(srflx)), i, j, sr);
    const double c = 0.2 * (60.0 + latr);
    st(D, v2(D, FID(Uwind)), i, j, 15.0 * exp(-c * c)); st(D, v2(D, FID(Vwind)), i, j, 0.0);
    st(D, v2(D, FID(rain)), i, j, 0.0); st(D, v2(D, FID(Pair)), i, j, 1025.0);
    v2l(D, FID(btflux), 1)(i, j) = 0.0; v2l(D, FID(btflux), 2)(i, j) = 0.0;
    st(D, v2l(D, FID(stflux), 2), i, j, 0.0);
  } else {
    {
      st(D, v2l(D, FID(stflux), 1), i, j, 0.0); st(D, v2l(D, FID(stflux), 2), i, j, 0.0);
      v2l(D, FID(btflux), 1)(i, j) = 0.0; v2l(D, FID(btflux), 2)(i, j) = 0.0;
    }
    if (i >= D.uI0) st(D, v2(D, FID(sustr)), i, j, fa.windamp);
    if (j >= D.vJ0) st(D, v2(D, FID(svstr)), i, j, 0.0);
  }
}
// host part of caldate/datevec (Utility/dateclock.F) for time_ref=0
static double h_ufloor(double X) { return X - fmod(X, 1.0) - fmod(2.0 + copysign(1.0, X), 3.0); }
static double h_tfloor(double X, double CT) {
  double Q = 1.0; if (X < 0.0) Q = 1.0 - CT;
  const double RMAX = Q / (2.0 - CT), EPS5 = CT / Q;
  const double Y = h_ufloor(X + fmax(CT, fmin(RMAX, EPS5 * fabs(1.0 + h_ufloor(X)))));
  if (X <= 0.0 || (Y - X) < RMAX) return Y;
  return Y - 1.0;
}
int k_set_data(roms_b200_ctx* c, double tdays) {
  const roms_b200_bounds& b = c->D.b; const roms_b200_params& p = c->D.p;
  ForcingArgs fa{0, 0, 0, 0};
  if (p.app == ROMS_B200_APP_BENCHMARK) {
    const double DateNumber = 367.0 + tdays, DayFraction = fabs(DateNumber - trunc(DateNumber));
    double seconds = h_tfloor(DayFraction * 86400.0 + 0.5, 3.0 * 2.220446049250313e-16);
    const double hour = seconds / 3600.0, yday = (double)(1 + (int)floor(tdays)) + DayFraction;
    fa.Dangle = 23.44 * cos((172.0 - yday) * 2.0 * PI_D / 365.2425);
    fa.Dangle = fa.Dangle * (PI_D / 180.0);
    fa.Hangle = (12.0 - hour) * PI_D / 12.0;
    fa.Rsolar = 1353.0 / (p.rho0 * 3985.0);
  } else {
    if ((tdays - p.dstart) <= 2.0) fa.windamp = -0.1 * sin(PI_D * (tdays - p.dstart) / 4.0) / p.rho0;
    else fa.windamp = -0.1 / p.rho0;
  }
  Box bx{c->D.rI0, c->D.rI1, c->D.rJ0, c->D.rJ1}; dim3 blk(128, 2);
  set_data_kernel<<<grid2(bx, blk), blk, 0, c->stream>>>(c->D, bx, fa); c->launches++;
  return 0;
}
