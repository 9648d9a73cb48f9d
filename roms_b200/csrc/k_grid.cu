// roms_b200/csrc/k_grid.cu -- depth / mass-flux / omega / EOS / vertical BC kernels.
// One thread per water column (i fastest -> coalesced rows), marching k.
// Reference routines (file:line) are cited at each kernel; arithmetic per point
// follows the reference operation order so results are bit-identical to a
// non-fast-math build (compiled with -fmad=false).
#include "common.cuh"
#include <cmath>

// ---- set_depth_tile, Nonlinear/set_depth.F:192-245 (Vtransform=2) -----------
__global__ void set_depth_kernel(const Dev D, Box bx) {
  IJ_FROM_BOX(bx);
  const int N = D.b.N; const double hc = D.p.hc;
  V2 h = v2(D, FID(h)), Zt = v2(D, FID(Zt_avg1));
  V3 z_w = v3(D, FID(z_w)), z_r = v3(D, FID(z_r)), Hz = v3(D, FID(Hz));
  const double hwater = h(i, j), zt = Zt(i, j);
  const double hinv = 1.0 / (hc + hwater);
  double zw_prev = -hwater;
  st(D, z_w, i, j, 0, zw_prev);
  for (int k = 1; k <= N; ++k) {
    double cff_r = hc * D.sc_r[k], cff_w = hc * D.sc_w[k];
    double cff2_r = (cff_r + D.Cs_r[k] * hwater) * hinv;
    double cff2_w = (cff_w + D.Cs_w[k] * hwater) * hinv;
    double zw = zt + (zt + hwater) * cff2_w;
    double zr = zt + (zt + hwater) * cff2_r;
    st(D, z_w, i, j, k, zw);
    st(D, z_r, i, j, k, zr);
    st(D, Hz, i, j, k, zw - zw_prev);
    zw_prev = zw;
  }
}
int k_set_depth(roms_b200_ctx* c) {
  const roms_b200_bounds& b = c->D.b;
  Box bx{c->D.rI0, c->D.rI1, c->D.rJ0, c->D.rJ1}; dim3 blk(64, 4);
  set_depth_kernel<<<grid2(bx, blk), blk, 0, c->stream>>>(c->D, bx); c->launches++;
  return 0;
}

// ---- set_massflux_tile, Nonlinear/set_massflux.F:140-163 -------------------
__global__ void set_massflux_kernel(const Dev D, Box bx, int nrhs) {
  IJ_FROM_BOX(bx);
  const int k = 1 + blockIdx.z;
  const roms_b200_bounds& b = D.b;
  V3 Hz = v3(D, FID(Hz)), Huon = v3(D, FID(Huon)), Hvom = v3(D, FID(Hvom));
  V3 u = v3l(D, FID(u), nrhs), v = v3l(D, FID(v), nrhs);
  V2 on_u = v2(D, FID(on_u)), om_v = v2(D, FID(om_v));
  if (i >= D.uI0)
    st(D, Huon, i, j, k, 0.5 * (Hz(i, j, k) + Hz(i - 1, j, k)) * u(i, j, k) * on_u(i, j));
  if (j >= D.vJ0)
    st(D, Hvom, i, j, k, 0.5 * (Hz(i, j, k) + Hz(i, j - 1, k)) * v(i, j, k) * om_v(i, j));
}
int k_set_massflux(roms_b200_ctx* c, int nrhs) {
  const roms_b200_bounds& b = c->D.b;
  Box bx{c->D.rI0, c->D.rI1, c->D.rJ0, c->D.rJ1}; dim3 blk(128, 2); dim3 g = grid2(bx, blk); g.z = b.N;
  set_massflux_kernel<<<g, blk, 0, c->stream>>>(c->D, bx, nrhs); c->launches++;
  return 0;
}

// ---- omega_tile, Nonlinear/omega.F:215-355 + bc_w3d (bc_3d.F:588-723) --------
__global__ void omega_kernel(const Dev D, Box bx) {
  IJ_FROM_BOX(bx);
  const int N = D.b.N; const roms_b200_bounds& b = D.b;
  V3 W = v3(D, FID(W)), Huon = v3(D, FID(Huon)), Hvom = v3(D, FID(Hvom)), z_w = v3(D, FID(z_w));
  const bool south = b.Southern_Edge && j == b.Jstr, north = b.Northern_Edge && j == b.Jend;
  double w[RB_MAXN + 1], zw[RB_MAXN + 1];
  w[0] = 0.0;
  const double zw0 = z_w(i, j, 0);
  constexpr int KB = 6;                          // levels per load batch (the column sum itself stays sequential)
  double wk = 0.0;
  for (int k0 = 1; k0 <= N; k0 += KB) {
    double ue[KB], uw[KB], vn[KB], vs[KB], z[KB];
#pragma unroll
    for (int q = 0; q < KB; ++q) {
      const int k = min(k0 + q, N);
      ue[q] = Huon(i + 1, j, k); uw[q] = Huon(i, j, k); vn[q] = Hvom(i, j + 1, k); vs[q] = Hvom(i, j, k); z[q] = z_w(i, j, k);
    }
#pragma unroll
    for (int q = 0; q < KB; ++q) {
      const int k = k0 + q;
      if (k <= N) { wk = wk - (ue[q] - uw[q] + vn[q] - vs[q]); w[k] = wk; zw[k] = z[q]; }
    }
  }
  const double wrk = w[N] / (zw[N] - zw0);
  for (int k = N - 1; k >= 1; --k) w[k] = w[k] - wrk * (zw[k] - zw0);
  w[N] = 0.0;
  for (int k = 0; k <= N; ++k) {
    st(D, W, i, j, k, w[k]);
    if (south) st(D, W, i, j - 1, k, w[k]);
    if (north) st(D, W, i, j + 1, k, w[k]);
  }
}
int k_omega(roms_b200_ctx* c) {
  const roms_b200_bounds& b = c->D.b;
  Box bx{c->D.oI0, c->D.oI1, c->D.oJ0, c->D.oJ1}; dim3 blk(64, 2);
  omega_kernel<<<grid2(bx, blk), blk, 0, c->stream>>>(c->D, bx); c->launches++;
  return 0;
}

// ---- wvelocity_tile, Nonlinear/wvelocity.F:171-274 (+ bc_w3d): "true" vertical velocity at W-points --------
// One thread per water column: the s-surface contribution vert(k) (4 staggered products per level) into a private
// array, then the cubic shift to W-points.  Same operation order as the reference, point by point.
__global__ void wvelocity_kernel(const Dev D, Box bx, int ninp) {
  IJ_FROM_BOX(bx);
  const int N = D.b.N; const roms_b200_bounds& b = D.b;
  V3 u = v3l(D, FID(u), ninp), v = v3l(D, FID(v), ninp), z_r = v3(D, FID(z_r)), z_w = v3(D, FID(z_w)), W = v3(D, FID(W)), wvel = v3(D, FID(wvel));
  V2 pm = v2(D, FID(pm)), pn = v2(D, FID(pn)), DU = v2(D, FID(DU_avg1)), DV = v2(D, FID(DV_avg1));
  const bool south = b.Southern_Edge && j == b.Jstr, north = b.Northern_Edge && j == b.Jend;
  const double pmW = pm(i - 1, j) + pm(i, j), pmE = pm(i, j) + pm(i + 1, j), pnS = pn(i, j - 1) + pn(i, j), pnN = pn(i, j) + pn(i, j + 1);
  double vert[RB_MAXN + 1], zwv[RB_MAXN + 1], Wv[RB_MAXN + 1];
  constexpr int KB = 4;                          // levels per load batch
  for (int k0 = 1; k0 <= N; k0 += KB) {
    double zc[KB], zW[KB], zE[KB], zS[KB], zN[KB], uW[KB], uE[KB], vS[KB], vN[KB], zz[KB], ww[KB];
#pragma unroll
    for (int q = 0; q < KB; ++q) {
      const int k = min(k0 + q, N);
      zc[q] = z_r(i, j, k); zW[q] = z_r(i - 1, j, k); zE[q] = z_r(i + 1, j, k); zS[q] = z_r(i, j - 1, k); zN[q] = z_r(i, j + 1, k);
      uW[q] = u(i, j, k); uE[q] = u(i + 1, j, k); vS[q] = v(i, j, k); vN[q] = v(i, j + 1, k);
      zz[q] = z_w(i, j, k); ww[q] = W(i, j, k);
    }
#pragma unroll
    for (int q = 0; q < KB; ++q) {
      const int k = k0 + q;
      if (k <= N) {
        const double wW = uW[q] * (zc[q] - zW[q]) * pmW, wE = uE[q] * (zE[q] - zc[q]) * pmE;
        const double wS = vS[q] * (zc[q] - zS[q]) * pnS, wN = vN[q] * (zN[q] - zc[q]) * pnN;
        double vt = 0.25 * (wW + wE);
        vt = vt + 0.25 * (wS + wN);
        vert[k] = vt; zwv[k] = zz[q]; Wv[k] = ww[q];
      }
    }
  }
  const double cff1 = 3.0 / 8.0, cff2 = 3.0 / 4.0, cff3 = 1.0 / 8.0, cff4 = 9.0 / 16.0, cff5 = 1.0 / 16.0;
  const double zw0 = z_w(i, j, 0), zwN = zwv[N];
  const double wrk = (DU(i, j) - DU(i + 1, j) + DV(i, j) - DV(i, j + 1)) / (zwN - zw0);
  const double pmn = pm(i, j) * pn(i, j);
  const double zr1 = z_r(i, j, 1), zr2 = z_r(i, j, 2), zrN = z_r(i, j, N), zrNm = z_r(i, j, N - 1);
  auto put = [&](int k, double val) {
    st(D, wvel, i, j, k, val);
    if (south) st(D, wvel, i, j - 1, k, val);
    if (north) st(D, wvel, i, j + 1, k, val);
  };
  {
    const double slope = (zr1 - zw0) / (zr2 - zr1);
    put(0, cff1 * (vert[1] - slope * (vert[2] - vert[1])) + cff2 * vert[1] - cff3 * vert[2]);
    put(1, pmn * (Wv[1] + wrk * (zwv[1] - zw0)) + cff1 * vert[1] + cff2 * vert[2] - cff3 * vert[3]);
  }
  for (int k = 2; k <= N - 2; ++k)
    put(k, pmn * (Wv[k] + wrk * (zwv[k] - zw0)) + cff4 * (vert[k] + vert[k + 1]) - cff5 * (vert[k - 1] + vert[k + 2]));
  {
    const double slope = (zwN - zrN) / (zrN - zrNm);
    put(N, pmn * wrk * (zwN - zw0) + cff1 * (vert[N] + slope * (vert[N] - vert[N - 1])) + cff2 * vert[N] - cff3 * vert[N - 1]);
    put(N - 1, pmn * (Wv[N - 1] + wrk * (zwv[N - 1] - zw0)) + cff1 * vert[N] + cff2 * vert[N - 1] - cff3 * vert[N - 2]);
  }
}
// Interior columns of the tile only (plus E-W periodic images and closed-wall rows): wvel is read by diag on the interior and
// by the output path; its tile halos are not exchanged (the reference's mp_exchange3d of wvel, wvelocity.F:275-281, serves output).
int k_wvelocity(roms_b200_ctx* c, int ninp) {
  const roms_b200_bounds& b = c->D.b;
  if (b.N < 3) return 1;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk(64, 2);
  wvelocity_kernel<<<grid2(bx, blk), blk, 0, c->stream>>>(c->D, bx, ninp); c->launches++;
  return 0;
}

// ---- set_zeta_tile, Nonlinear/set_zeta.F:101-118 -----------------------------
__global__ void set_zeta_kernel(const Dev D, Box bx) {
  IJ_FROM_BOX(bx);
  V2 Zt = v2(D, FID(Zt_avg1)), z1 = v2l(D, FID(zeta), 1), z2 = v2l(D, FID(zeta), 2);
  const double zt = Zt(i, j);
  st(D, z1, i, j, zt); st(D, z2, i, j, zt);
}
int k_set_zeta(roms_b200_ctx* c) {
  const roms_b200_bounds& b = c->D.b;
  Box bx{c->D.rI0, c->D.rI1, c->D.rJ0, c->D.rJ1}; dim3 blk(128, 2);   // == IstrR..IendR x JstrR..JendR on a single tile
  set_zeta_kernel<<<grid2(bx, blk), blk, 0, c->stream>>>(c->D, bx); c->launches++;
  return 0;
}

// ---- rho_eos_tile -----------------------------------------------------------
// linear: Nonlinear/rho_eos.F:688-740 ; nonlinear (UNESCO / Jackett-McDougall):
// rho_eos.F:293-440 with coefficients of Modules/mod_eoscoef.F:24-64.
#define A00 (+1.909256e+04)
#define A01 (+2.098925e+02)
#define A02 (-3.041638e+00)
#define A03 (-1.852732e-03)
#define A04 (-1.361629e-05)
#define B00 (+1.044077e+02)
#define B01 (-6.500517e+00)
#define B02 (+1.553190e-01)
#define B03 (+2.326469e-04)
#define D00 (-5.587545e+00)
#define D01 (+7.390729e-01)
#define D02 (-1.909078e-02)
#define E00 (+4.721788e-01)
#define E01 (+1.028859e-02)
#define E02 (-2.512549e-04)
#define E03 (-5.939910e-07)
#define F00 (-1.571896e-02)
#define F01 (-2.598241e-04)
#define F02 (+7.267926e-06)
#define G00 (+2.042967e-03)
#define G01 (+1.045941e-05)
#define G02 (-5.782165e-10)
#define G03 (+1.296821e-07)
#define H00 (-2.595994e-07)
#define H01 (-1.248266e-09)
#define H02 (-3.508914e-09)
#define Q00 (+9.99842594e+02)
#define Q01 (+6.793952e-02)
#define Q02 (-9.095290e-03)
#define Q03 (+1.001685e-04)
#define Q04 (-1.120083e-06)
#define Q05 (+6.536332e-09)
#define U00 (+8.24493e-01)
#define U01 (-4.08990e-03)
#define U02 (+7.64380e-05)
#define U03 (-8.24670e-07)
#define U04 (+5.38750e-09)
#define V00 (-5.72466e-03)
#define V01 (+1.02270e-04)
#define V02 (-1.65460e-06)
#define W00 (+4.8314e-04)

__global__ void rho_eos_kernel(const Dev D, Box bx, int nrhs) {
  IJ_FROM_BOX(bx);
  const int N = D.b.N; const double g = D.p.g, rho0 = D.p.rho0;
  V3 Hz = v3(D, FID(Hz)), z_r = v3(D, FID(z_r)), z_w = v3(D, FID(z_w)), rho = v3(D, FID(rho)), pden = v3(D, FID(pden));
  V3 T = v3l(D, FID(t), nrhs, 1), S = v3l(D, FID(t), nrhs, 2);
  V2 rhoA = v2(D, FID(rhoA)), rhoS = v2(D, FID(rhoS));
  double den[RB_MAXN + 1], hzv[RB_MAXN + 1];
  if (D.p.app == ROMS_B200_APP_UPWELLING) {
    const double R0 = D.p.R0;
    for (int k = 1; k <= N; ++k) {
      hzv[k] = Hz(i, j, k);
      double r = R0 - R0 * D.p.Tcoef * (T(i, j, k) - D.p.T0);
      r = r + R0 * D.p.Scoef * (S(i, j, k) - D.p.S0);
      r = r - 1000.0;
      den[k] = r;
      st(D, rho, i, j, k, r); st(D, pden, i, j, k, r);
    }
  } else {
    V3 bvf = v3(D, FID(bvf)); V2 alpha = v2(D, FID(alpha)), beta = v2(D, FID(beta));
    double p_den1 = 0, p_b0 = 0, p_b1 = 0, p_b2 = 0, p_zr = 0;   // level k-1
    constexpr int KB = 5;                      // levels per load batch; the polynomial evaluations of a batch are independent
    for (int k0 = 1; k0 <= N; k0 += KB) {
    double bT[KB], bS[KB], bZr[KB], bZw[KB];
#pragma unroll
    for (int q = 0; q < KB; ++q) { const int k = min(k0 + q, N); bT[q] = T(i, j, k); bS[q] = S(i, j, k); bZr[q] = z_r(i, j, k); bZw[q] = z_w(i, j, k - 1); hzv[k] = Hz(i, j, k); }
#pragma unroll
    for (int q = 0; q < KB; ++q) {
      const int k = k0 + q;
      if (k > N) break;
      const double Tt = fmax(-2.0, bT[q]), Ts = fmax(0.0, bS[q]);
      const double sqrtTs = sqrt(Ts), Tp = bZr[q], Tpr10 = 0.1 * Tp;
      const double C0 = Q00 + Tt * (Q01 + Tt * (Q02 + Tt * (Q03 + Tt * (Q04 + Tt * Q05))));
      const double C1 = U00 + Tt * (U01 + Tt * (U02 + Tt * (U03 + Tt * U04)));
      const double C2 = V00 + Tt * (V01 + Tt * V02);
      const double den1 = C0 + Ts * (C1 + sqrtTs * C2 + Ts * W00);
      const double C3 = A00 + Tt * (A01 + Tt * (A02 + Tt * (A03 + Tt * A04)));
      const double C4 = B00 + Tt * (B01 + Tt * (B02 + Tt * B03));
      const double C5 = D00 + Tt * (D01 + Tt * D02);
      const double C6 = E00 + Tt * (E01 + Tt * (E02 + Tt * E03));
      const double C7 = F00 + Tt * (F01 + Tt * F02);
      const double C8 = G01 + Tt * (G02 + Tt * G03);
      const double C9 = H00 + Tt * (H01 + Tt * H02);
      const double bulk0 = C3 + Ts * (C4 + sqrtTs * C5);
      const double bulk1 = C6 + Ts * (C7 + sqrtTs * G00);
      const double bulk2 = C8 + Ts * C9;
      const double bulk = bulk0 - Tp * (bulk1 - Tp * bulk2);
      const double cff = 1.0 / (bulk + Tpr10);
      double dn = den1 * bulk * cff;
      dn = dn - 1000.0;
      den[k] = dn;
      st(D, rho, i, j, k, dn); st(D, pden, i, j, k, (den1 - 1000.0));
      if (k > 1) {                     // bvf at interface k-1 (rho_eos.F:402-424)
        const double zw = bZw[q];
        const double bulk_up = bulk0 - zw * (bulk1 - bulk2 * zw);
        const double bulk_dn = p_b0 - zw * (p_b1 - p_b2 * zw);
        const double cff1 = 1.0 / (bulk_up + 0.1 * zw), cff2 = 1.0 / (bulk_dn + 0.1 * zw);
        const double den_up = cff1 * (den1 * bulk_up), den_dn = cff2 * (p_den1 * bulk_dn);
        st(D, bvf, i, j, k - 1, -g * (den_up - den_dn) / (0.5 * (den_up + den_dn) * (Tp - p_zr)));
      }
      p_den1 = den1; p_b0 = bulk0; p_b1 = bulk1; p_b2 = bulk2; p_zr = Tp;
      if (k == N) {                    // thermal expansion / saline contraction at the surface (:426-470)
        const double dC0 = Q01 + Tt * (2.0 * Q02 + Tt * (3.0 * Q03 + Tt * (4.0 * Q04 + Tt * 5.0 * Q05)));
        const double dC1 = U01 + Tt * (2.0 * U02 + Tt * (3.0 * U03 + Tt * 4.0 * U04));
        const double dC2 = V01 + Tt * 2.0 * V02;
        const double Dden1DS = C1 + 1.5 * C2 * sqrtTs + 2.0 * W00 * Ts;
        const double Dden1DT = dC0 + Ts * (dC1 + sqrtTs * dC2);
        const double dC3 = A01 + Tt * (2.0 * A02 + Tt * (3.0 * A03 + Tt * 4.0 * A04));
        const double dC4 = B01 + Tt * (2.0 * B02 + Tt * 3.0 * B03);
        const double dC5 = D01 + Tt * 2.0 * D02;
        const double dC6 = E01 + Tt * (2.0 * E02 + Tt * 3.0 * E03);
        const double dC7 = F01 + Tt * 2.0 * F02;
        const double dC8 = G02 + Tt * 2.0 * G03;
        const double dC9 = H01 + Tt * 2.0 * H02;
        const double DbulkDS = C4 + sqrtTs * 1.5 * C5 - Tp * (C7 + sqrtTs * 1.5 * G00 - Tp * C9);
        const double DbulkDT = dC3 + Ts * (dC4 + sqrtTs * dC5) - Tp * (dC6 + Ts * dC7 - Tp * (dC8 + Ts * dC9));
        const double c0 = bulk + Tpr10, c1 = Tpr10 * den1, c2 = bulk * c0;
        const double wrk = (dn + 1000.0) * c0 * c0;
        const double Tcof = -(DbulkDT * c1 + Dden1DT * c2);
        const double Scof = (DbulkDS * c1 + Dden1DS * c2);
        const double ci = 1.0 / wrk;
        st(D, alpha, i, j, ci * Tcof); st(D, beta, i, j, ci * Scof);
      }
    }
    }
    st(D, bvf, i, j, 0, 0.0); st(D, bvf, i, j, N, 0.0);
  }
  // vertical averages for the barotropic pressure gradient (rho_eos.F:370-387 / :702-718)
  double cff1 = den[N] * hzv[N];
  double rS = 0.5 * cff1 * hzv[N], rA = cff1;
  for (int k = N - 1; k >= 1; --k) {
    const double hz = hzv[k];
    cff1 = den[k] * hz;
    rS = rS + hz * (rA + 0.5 * cff1);
    rA = rA + cff1;
  }
  const double cff2 = 1.0 / rho0;
  cff1 = 1.0 / (z_w(i, j, N) - z_w(i, j, 0));
  st(D, rhoA, i, j, cff2 * cff1 * rA);
  st(D, rhoS, i, j, 2.0 * cff1 * cff1 * cff2 * rS);
}
int k_rho_eos(roms_b200_ctx* c, int nrhs) {
  const roms_b200_bounds& b = c->D.b;
  Box bx{c->D.rI0, c->D.rI1, c->D.rJ0, c->D.rJ1}; dim3 blk(64, 2);
  rho_eos_kernel<<<grid2(bx, blk), blk, 0, c->stream>>>(c->D, bx, nrhs); c->launches++;
  return 0;
}

// ---- set_vbc_tile, Nonlinear/set_vbc.F:620-720 + bc_u2d/bc_v2d (bc_2d.F) ------
__global__ void set_vbc_kernel(const Dev D, Box bx, int nrhs) {
  IJ_FROM_BOX(bx);
  const roms_b200_bounds& b = D.b; const int N = b.N;
  V3 u = v3l(D, FID(u), nrhs), v = v3l(D, FID(v), nrhs), S = v3l(D, FID(t), nrhs, 2);
  if (i >= b.IstrR && i <= b.IendR && j >= b.JstrR && j <= b.JendR) {
    V2 st1 = v2l(D, FID(stflx), 1), bt1 = v2l(D, FID(btflx), 1), st2 = v2l(D, FID(stflx), 2), bt2 = v2l(D, FID(btflx), 2);
    V2 sx1 = v2l(D, FID(stflux), 1), bx1 = v2l(D, FID(btflux), 1), sx2 = v2l(D, FID(stflux), 2);
    st1(i, j) = sx1(i, j); bt1(i, j) = bx1(i, j);
    const double EmP = sx2(i, j);
    st2(i, j) = EmP * S(i, j, N);
    bt2(i, j) = bt2(i, j) * S(i, j, 1);
  }
  V2 bustr = v2(D, FID(bustr)), bvstr = v2(D, FID(bvstr));
  const bool quad = (D.p.app == ROMS_B200_APP_BENCHMARK);
  if (i >= b.IstrU && i <= b.Iend && j >= b.Jstr && j <= b.Jend) {
    double val;
    if (quad) {
      V2 rd = v2(D, FID(rdrag2));
      const double cff1 = 0.25 * (v(i, j, 1) + v(i, j + 1, 1) + v(i - 1, j, 1) + v(i - 1, j + 1, 1));
      const double cff2 = sqrt(u(i, j, 1) * u(i, j, 1) + cff1 * cff1);
      val = 0.5 * (rd(i - 1, j) + rd(i, j)) * u(i, j, 1) * cff2;
    } else {
      V2 rd = v2(D, FID(rdrag));
      val = 0.5 * (rd(i - 1, j) + rd(i, j)) * u(i, j, 1);
    }
    st(D, bustr, i, j, val);
    if (b.Southern_Edge && j == b.Jstr) st(D, bustr, i, j - 1, D.p.gamma2 * val);
    if (b.Northern_Edge && j == b.Jend) st(D, bustr, i, j + 1, D.p.gamma2 * val);
  }
  if (i >= b.Istr && i <= b.Iend && j >= b.JstrV && j <= b.Jend) {
    double val;
    if (quad) {
      V2 rd = v2(D, FID(rdrag2));
      const double cff1 = 0.25 * (u(i, j, 1) + u(i + 1, j, 1) + u(i, j - 1, 1) + u(i + 1, j - 1, 1));
      const double cff2 = sqrt(cff1 * cff1 + v(i, j, 1) * v(i, j, 1));
      val = 0.5 * (rd(i, j - 1) + rd(i, j)) * v(i, j, 1) * cff2;
    } else {
      V2 rd = v2(D, FID(rdrag));
      val = 0.5 * (rd(i, j - 1) + rd(i, j)) * v(i, j, 1);
    }
    st(D, bvstr, i, j, val);
  }
  if (i >= b.Istr && i <= b.Iend) {                // bc_v2d: closed walls
    if (b.Southern_Edge && j == b.Jstr) st(D, bvstr, i, b.Jstr, 0.0);
    if (b.Northern_Edge && j == b.Jend) st(D, bvstr, i, b.Jend + 1, 0.0);
  }
}
int k_set_vbc(roms_b200_ctx* c, int nrhs) {
  // with neighbour tiles: also on two halo points (see k_bulk_flux), so stflx, btflx, bustr, bvstr need no halo swap
  const Dev De = widened(c, 2); const roms_b200_bounds& b = De.b;
  Box bx{b.IstrR, b.IendR, b.JstrR, b.JendR}; dim3 blk(128, 2);
  set_vbc_kernel<<<grid2(bx, blk), blk, 0, c->stream>>>(De, bx, nrhs); c->launches++;
  return 0;
}

// ---- ana_vmix_tile, Functionals/ana_vmix.h:200-206,327-337 -------------------
__global__ void ana_vmix_kernel(const Dev D, Box bx) {
  IJ_FROM_BOX(bx);
  const int N = D.b.N;
  V3 Akv = v3(D, FID(Akv)), z_w = v3(D, FID(z_w)), Akt1 = v3l(D, FID(Akt), 1), Akt2 = v3l(D, FID(Akt), 2);
  for (int k = 1; k <= N - 1; ++k) {
    st(D, Akv, i, j, k, 2.0e-03 + 8.0e-03 * exp(z_w(i, j, k) / 150.0));
    st(D, Akt1, i, j, k, D.p.Akt_bak[0]); st(D, Akt2, i, j, k, D.p.Akt_bak[1]);
  }
}
int k_ana_vmix(roms_b200_ctx* c) {
  const roms_b200_bounds& b = c->D.b;
  Box bx{c->D.rI0, c->D.rI1, c->D.rJ0, c->D.rJ1}; dim3 blk(64, 4);
  ana_vmix_kernel<<<grid2(bx, blk), blk, 0, c->stream>>>(c->D, bx); c->launches++;
  return 0;
}

// ---- diag_tile, Nonlinear/diag.F:209-411,512-542 -------------------------------------------------------------
// Stage 1 (one thread per water column, k = N..1 as the reference): ke2d, pe2d of the column into two scratch planes; the
// column's largest Courant number (strict '>' in descending k keeps the first = the reference's scan order inside a
// column), speed and density maxima; a block (128 columns of one row) reduces the maxima with a total order, so the result
// does not depend on the reduction tree.
// Stage 2: the reference's two-stage horizontal sum (diag.F:296-322): one thread per i collapses j sequentially; the host
// then adds the per-i sums in i order -- bit-identical to the serial reference.  One extra block reduces stage 1's maxima.
namespace {
struct CMax { double C, Cu, Cv, Cw; int i, j, k; };
// true if a precedes b: larger C, or the same non-zero C met earlier in the reference's scan (j ascending, k descending, i ascending)
__device__ __forceinline__ bool cmax_before(double aC, int ai, int aj, int ak, double bC, int bi, int bj, int bk) {
  if (aC > bC) return true;
  if (!(aC == bC) || !(aC > 0.0)) return false;
  if (aj != bj) return aj < bj;
  if (ak != bk) return ak > bk;
  return ai < bi;
}
constexpr int DG_NT = 128;    // threads per block of both stages
constexpr int DG_NP = 9;      // doubles per partial: C, Cu, Cv, Cw, i, j, k, maxspeed, maxrho
}  // namespace
__global__ void __launch_bounds__(DG_NT) diag_cols_kernel(const Dev D, int nstp, V2 ke2d, V2 pe2d, V2 vo2d, double* __restrict__ partial) {
  const roms_b200_bounds& b = D.b; const int N = b.N; const double g = D.p.g, dt = D.p.dt;
  V3 Hz = v3(D, FID(Hz)), z_w = v3(D, FID(z_w)), z_r = v3(D, FID(z_r)), rho = v3(D, FID(rho)), wvel = v3(D, FID(wvel));
  V3 u = v3l(D, FID(u), nstp), v = v3l(D, FID(v), nstp); V2 pm = v2(D, FID(pm)), pn = v2(D, FID(pn)), omn = v2(D, FID(omn));
  const int j = b.Jstr + blockIdx.y, i = b.Istr + blockIdx.x * DG_NT + threadIdx.x;
  double bC = 0.0, bCu = 0.0, bCv = 0.0, bCw = 0.0, spd2 = 0.0, mrho = -1.0e37; int bk = 0, bi = 0, bj = 0;
  if (i <= b.Iend) {
    const double zwN = z_w(i, j, N), zw0 = z_w(i, j, 0), cff = g / D.p.rho0, pmi = pm(i, j), pni = pn(i, j);
    double ke2 = 0.0, pe2 = 0.5 * g * zwN * zwN;
    double wup = wvel(i, j, N);
    constexpr int KB = 5;                       // levels per load batch (the loads of a batch are issued together)
    for (int k0 = N; k0 >= 1; k0 -= KB) {
      double ua[KB], ub[KB], va[KB], vb[KB], hz[KB], r[KB], zr[KB], wl[KB];
#pragma unroll
      for (int q = 0; q < KB; ++q) {
        const int k = max(k0 - q, 1);
        ua[q] = u(i, j, k); ub[q] = u(i + 1, j, k); va[q] = v(i, j, k); vb[q] = v(i, j + 1, k);
        hz[q] = Hz(i, j, k); r[q] = rho(i, j, k); zr[q] = z_r(i, j, k); wl[q] = wvel(i, j, k - 1);
      }
#pragma unroll
      for (int q = 0; q < KB; ++q) {
        const int k = k0 - q;
        if (k >= 1) {
          const double u2v2 = ua[q] * ua[q] + ub[q] * ub[q] + va[q] * va[q] + vb[q] * vb[q];
          ke2 = ke2 + hz[q] * 0.25 * u2v2;
          pe2 = pe2 + cff * hz[q] * (r[q] + 1000.0) * (zr[q] - zw0);
          const double Cu = 0.5 * fabs(ua[q] + ub[q]) * dt * pmi, Cv = 0.5 * fabs(va[q] + vb[q]) * dt * pni, Cw = 0.5 * fabs(wl[q] + wup) * dt / hz[q];
          wup = wl[q];
          const double Cc = Cu + Cv + Cw;
          if (Cc > bC) { bC = Cc; bCu = Cu; bCv = Cv; bCw = Cw; bk = k; bi = i; bj = j; }
          spd2 = fmax(spd2, 0.5 * u2v2);        // sqrt is monotonic and correctly rounded: max sqrt(x) == sqrt(max x)
          mrho = fmax(mrho, r[q]);
        }
      }
    }
    const double o = omn(i, j);
    ke2d(i, j) = o * ke2; pe2d(i, j) = o * pe2; vo2d(i, j) = o * (zwN - zw0);     // the products of diag.F:305-313
  }
  const double spd = sqrt(spd2);
  __shared__ double sC[DG_NT], sS[DG_NT], sR[DG_NT]; __shared__ int sK[DG_NT], sI[DG_NT], sJ[DG_NT], sO[DG_NT];
  const int t = threadIdx.x;
  sC[t] = bC; sK[t] = bk; sI[t] = bi; sJ[t] = bj; sO[t] = t; sS[t] = spd; sR[t] = mrho;
  __syncthreads();
  for (int o = DG_NT / 2; o > 0; o >>= 1) {
    if (t < o) {
      if (cmax_before(sC[t + o], sI[t + o], sJ[t + o], sK[t + o], sC[t], sI[t], sJ[t], sK[t])) { sC[t] = sC[t + o]; sI[t] = sI[t + o]; sJ[t] = sJ[t + o]; sK[t] = sK[t + o]; sO[t] = sO[t + o]; }
      sS[t] = fmax(sS[t], sS[t + o]); sR[t] = fmax(sR[t], sR[t + o]);
    }
    __syncthreads();
  }
  double* out = partial + (size_t)DG_NP * (blockIdx.y * gridDim.x + blockIdx.x);
  if (t == sO[0]) { out[0] = bC; out[1] = bCu; out[2] = bCv; out[3] = bCw; out[4] = bi; out[5] = bj; out[6] = bk; }
  if (t == 0) { out[7] = sS[0]; out[8] = sR[0]; }
}
// red layout: [0,wi) ke sums per i ; [wi,2wi) pe ; [2wi,3wi) volume ; [3wi,3wi+DG_NP) maxima ; then stage 1's partials
__global__ void __launch_bounds__(DG_NT) diag_sum_kernel(const Dev D, V2 ke2d, V2 pe2d, V2 vo2d, double* __restrict__ red, int nparts) {
  const roms_b200_bounds& b = D.b; const int wi = b.Iend - b.Istr + 1, t = threadIdx.x;
  if (blockIdx.x + 1 < gridDim.x) {
    const int i = b.Istr + blockIdx.x * DG_NT + t;
    if (i > b.Iend) return;
    double vol = 0.0, pe = 0.0, ke = 0.0;
    constexpr int JB = 8;                       // rows per load batch; the sums stay sequential in j (diag.F:303-316)
    for (int j0 = b.Jstr; j0 <= b.Jend; j0 += JB) {
      double a[JB], p[JB], k[JB];
#pragma unroll
      for (int q = 0; q < JB; ++q) { const int j = min(j0 + q, b.Jend); a[q] = vo2d(i, j); p[q] = pe2d(i, j); k[q] = ke2d(i, j); }
#pragma unroll
      for (int q = 0; q < JB; ++q) if (j0 + q <= b.Jend) { vol = vol + a[q]; pe = pe + p[q]; ke = ke + k[q]; }
    }
    red[i - b.Istr] = ke; red[wi + i - b.Istr] = pe; red[2 * wi + i - b.Istr] = vol;
    return;
  }
  // last block: maxima over stage 1's partials
  const double* partial = red + 3 * wi + DG_NP;
  double bC = 0.0, spd = 0.0, mrho = -1.0e37; int bi = 0, bj = 0, bk = 0, bq = -1;
  for (int q = t; q < nparts; q += DG_NT) {
    const double* p = partial + (size_t)DG_NP * q;
    if (cmax_before(p[0], (int)p[4], (int)p[5], (int)p[6], bC, bi, bj, bk)) { bC = p[0]; bi = (int)p[4]; bj = (int)p[5]; bk = (int)p[6]; bq = q; }
    spd = fmax(spd, p[7]); mrho = fmax(mrho, p[8]);
  }
  __shared__ double sC[DG_NT], sS[DG_NT], sR[DG_NT]; __shared__ int sK[DG_NT], sI[DG_NT], sJ[DG_NT], sQ[DG_NT];
  sC[t] = bC; sK[t] = bk; sI[t] = bi; sJ[t] = bj; sQ[t] = bq; sS[t] = spd; sR[t] = mrho;
  __syncthreads();
  for (int o = DG_NT / 2; o > 0; o >>= 1) {
    if (t < o) {
      if (cmax_before(sC[t + o], sI[t + o], sJ[t + o], sK[t + o], sC[t], sI[t], sJ[t], sK[t])) { sC[t] = sC[t + o]; sI[t] = sI[t + o]; sJ[t] = sJ[t + o]; sK[t] = sK[t + o]; sQ[t] = sQ[t + o]; }
      sS[t] = fmax(sS[t], sS[t + o]); sR[t] = fmax(sR[t], sR[t + o]);
    }
    __syncthreads();
  }
  if (t == 0) {
    double* out = red + 3 * wi;
    if (sQ[0] >= 0) { const double* p = partial + (size_t)DG_NP * sQ[0]; for (int q = 0; q < 7; ++q) out[q] = p[q]; }
    else for (int q = 0; q < 7; ++q) out[q] = 0.0;
    out[7] = sS[0]; out[8] = sR[0];
  }
}
// diag in two halves so that a host driver can overlap its own work with the step: begin = the two kernels + asynchronous
// D2H of the per-i sums and the maxima into pinned memory; end = wait, final sums in i order, mp_reduce across tiles.
int k_diag_begin(roms_b200_ctx* c, int nstp) {
  const Dev& D = c->D; const roms_b200_bounds& b = D.b; const int wi = b.Iend - b.Istr + 1, wj = b.Jend - b.Jstr + 1;
  const int nbx = (wi + DG_NT - 1) / DG_NT, nparts = nbx * wj;
  V2 ke2d{D.scratch2 + 8 * D.nij, b.LBi, D.ni, b.LBj}, pe2d{D.scratch2 + 9 * D.nij, b.LBi, D.ni, b.LBj}, vo2d{D.scratch2 + 10 * D.nij, b.LBi, D.ni, b.LBj};
  diag_cols_kernel<<<dim3(nbx, wj), DG_NT, 0, c->stream>>>(D, nstp, ke2d, pe2d, vo2d, D.red + 3 * wi + DG_NP); c->launches++;
  diag_sum_kernel<<<nbx + 1, DG_NT, 0, c->stream>>>(D, ke2d, pe2d, vo2d, D.red, nparts); c->launches++;
  CUDA_OK(cudaMemcpyAsync(c->h_red, D.red, sizeof(double) * (3 * wi + DG_NP), cudaMemcpyDeviceToHost, c->stream));
  return 0;
}
int k_diag(roms_b200_ctx* c, int nstp, double* out3) {
  if (k_diag_begin(c, nstp)) return 1;
  return k_diag_end(c, out3);
}
int k_diag_end(roms_b200_ctx* c, double* out3) {
  const roms_b200_bounds& b = c->D.b; const int wi = b.Iend - b.Istr + 1;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  double ke = 0, pe = 0, vol = 0;
  for (int q = 0; q < wi; ++q) { vol = vol + c->h_red[2 * wi + q]; pe = pe + c->h_red[wi + q]; ke = ke + c->h_red[q]; }     // diag.F:318-322
  double mx[DG_NP];
  for (int q = 0; q < DG_NP; ++q) mx[q] = c->h_red[3 * wi + q];
  if (c->comm) {
    // mp_reduce (SUM,SUM,SUM,MAX,MAX) and mp_reduce2 (MAXLOC) of diag.F:385-411 as ONE all-reduce: every tile fills its own
    // slot of 12 doubles, the sum gathers all slots everywhere, the host combines them in tile order as diag.F:363-384 does
    double* h = c->h_red; const int nr = c->nranks, n = 12 * nr;
    if (nr > 64) return 1;
    for (int q = 0; q < n; ++q) h[q] = 0.0;
    double* me = h + 12 * c->rank;
    me[0] = ke; me[1] = pe; me[2] = vol;
    for (int q = 0; q < DG_NP; ++q) me[3 + q] = mx[q];
    CUDA_OK(cudaMemcpyAsync(c->D.red, h, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (halo_allreduce_sum(c, c->D.red, n)) return 1;
    CUDA_OK(cudaMemcpyAsync(h, c->D.red, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
    ke = 0.0; pe = 0.0; vol = 0.0;
    double C = 0.0, Cu = 0.0, Cv = 0.0, Cw = 0.0, Ci = 0.0, Cj = 0.0, Ck = 0.0, spd = -1.0e35, mr = -1.0e35;
    for (int r = 0; r < nr; ++r) {
      const double* t = h + 12 * r;
      vol = vol + t[2]; ke = ke + t[0]; pe = pe + t[1];
      spd = std::fmax(spd, t[10]); mr = std::fmax(mr, t[11]);
      if (t[3] == C) { Ci = std::fmin(Ci, t[7]); Cj = std::fmin(Cj, t[8]); Ck = std::fmin(Ck, t[9]); }
      else if (t[3] > C) { C = t[3]; Cu = t[4]; Cv = t[5]; Cw = t[6]; Ci = t[7]; Cj = t[8]; Ck = t[9]; }
    }
    mx[0] = C; mx[1] = Cu; mx[2] = Cv; mx[3] = Cw; mx[4] = Ci; mx[5] = Cj; mx[6] = Ck; mx[7] = spd; mx[8] = mr;
  }
  double* d = c->last_diag;
  d[0] = ke / vol; d[1] = pe / vol; d[2] = vol;
  for (int q = 0; q < 7; ++q) d[3 + q] = mx[q];
  d[10] = mx[7]; d[11] = mx[8];
  // diag.F:512-542: blow-up test (the reference inspects the printed energies for NaN/Inf/overflow characters)
  d[12] = (!std::isfinite(d[0]) || !std::isfinite(d[1]) || d[10] > 20.0 || d[11] > 200.0) ? 1.0 : 0.0;
  out3[0] = d[0]; out3[1] = d[1]; out3[2] = d[2];
  return 0;
}

// ---- ana_initial (Functionals/ana_initial.h: BENCHMARK :545-558, UPWELLING :828-848) -------
// Analytical initial conditions evaluated on the device from the device z_r:
// fluid at rest, zeta=0, T(z), S.  Writes time level 1 (Nonlinear/initial.F:358).
__global__ void ana_initial_kernel(const Dev D, Box bx) {
  IJ_FROM_BOX(bx);
  const roms_b200_bounds& b = D.b; const int N = b.N;
  V3 z_r = v3(D, FID(z_r)), T = v3l(D, FID(t), 1, 1), S = v3l(D, FID(t), 1, 2);
  // u, v, ubar, vbar, zeta start from the zero-initialised mirror (fluid at rest)
  const double val1 = (44.69 / 39.382) * (44.69 / 39.382);
  const double val2 = val1 * (D.p.rho0 * 800.0 / D.p.g) * (5.0e-05 / ((42.689 / 44.69) * (42.689 / 44.69)));
  for (int k = 1; k <= N; ++k) {
    const double z = z_r(i, j, k);
    if (D.p.app == ROMS_B200_APP_BENCHMARK) { st(D, T, i, j, k, val2 * exp(z / 800.0) * (0.6 - 0.4 * tanh(z / 800.0))); st(D, S, i, j, k, 35.0); }
    else { st(D, T, i, j, k, D.p.T0 + 8.0 * exp(z / 50.0)); st(D, S, i, j, k, D.p.S0); }
  }
}
// ---- post_initial: ini_zeta_tile + ini_fields_tile (Nonlinear/ini_fields.F) -----------------
// lateral BCs + periodic images on time level (nstp,kstp), Zt_avg1=zeta(kstp), ubar/vbar = vertical mean of u,v.
__global__ void ini_fields_kernel(const Dev D, Box bx, int nstp, int kstp) {
  IJ_FROM_BOX(bx);
  const roms_b200_bounds& b = D.b; const int N = b.N;
  const bool S = b.Southern_Edge && !b.NSperiodic, Nn = b.Northern_Edge && !b.NSperiodic;
  const bool south = S && j == b.Jstr, north = Nn && j == b.Jend;
  V3 Hz = v3(D, FID(Hz)), u = v3l(D, FID(u), nstp), v = v3l(D, FID(v), nstp);
  V2 zeta = v2l(D, FID(zeta), kstp), Zt = v2(D, FID(Zt_avg1)), ubar = v2l(D, FID(ubar), kstp), vbar = v2l(D, FID(vbar), kstp);
  {
    const double z = zeta(i, j);
    st(D, zeta, i, j, z); st(D, Zt, i, j, z);
    if (south) { st(D, zeta, i, j - 1, z); st(D, Zt, i, j - 1, z); }
    if (north) { st(D, zeta, i, j + 1, z); st(D, Zt, i, j + 1, z); }
  }
  for (int itrc = 1; itrc <= b.NT; ++itrc) {
    V3 T = v3l(D, FID(t), nstp, itrc);
    for (int k = 1; k <= N; ++k) {
      const double val = T(i, j, k);
      st(D, T, i, j, k, val);
      if (south) st(D, T, i, j - 1, k, val);
      if (north) st(D, T, i, j + 1, k, val);
    }
  }
  double DC0 = 0.0, CF0 = 0.0;
  for (int k = 1; k <= N; ++k) {
    const double uu = u(i, j, k), dc = 0.5 * (Hz(i, j, k) + Hz(i - 1, j, k));
    DC0 = DC0 + dc; CF0 = CF0 + dc * uu;
    st(D, u, i, j, k, uu);
    if (south) st(D, u, i, j - 1, k, D.p.gamma2 * uu);
    if (north) st(D, u, i, j + 1, k, D.p.gamma2 * uu);
  }
  { const double c1 = 1.0 / DC0; const double ub = CF0 * c1; st(D, ubar, i, j, ub);
    if (south) st(D, ubar, i, j - 1, D.p.gamma2 * ub); if (north) st(D, ubar, i, j + 1, D.p.gamma2 * ub); }
  if (j >= b.JstrM) {
    DC0 = 0.0; CF0 = 0.0;
    for (int k = 1; k <= N; ++k) {
      const double vv = v(i, j, k), dc = 0.5 * (Hz(i, j, k) + Hz(i, j - 1, k));
      DC0 = DC0 + dc; CF0 = CF0 + dc * vv;
      st(D, v, i, j, k, vv);
    }
    const double c1 = 1.0 / DC0; st(D, vbar, i, j, CF0 * c1);
  }
  if (south) { for (int k = 1; k <= N; ++k) st(D, v, i, b.Jstr, k, 0.0); st(D, vbar, i, b.Jstr, 0.0); }
  if (north) { for (int k = 1; k <= N; ++k) st(D, v, i, b.Jend + 1, k, 0.0); st(D, vbar, i, b.Jend + 1, 0.0); }
}
int k_ana_initial(roms_b200_ctx* c) {
  const roms_b200_bounds& b = c->D.b;
  Box bx{c->D.rI0, c->D.rI1, c->D.rJ0, c->D.rJ1}; dim3 blk(64, 4);
  ana_initial_kernel<<<grid2(bx, blk), blk, 0, c->stream>>>(c->D, bx); c->launches++;
  return 0;
}
int k_ini_fields(roms_b200_ctx* c, int nstp, int kstp) {
  const roms_b200_bounds& b = c->D.b;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk(64, 4);
  ini_fields_kernel<<<grid2(bx, blk), blk, 0, c->stream>>>(c->D, bx, nstp, kstp); c->launches++;
  return 0;
}
