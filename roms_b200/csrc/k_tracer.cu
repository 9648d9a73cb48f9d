// roms_b200/csrc/k_tracer.cu -- tracer predictor (pre_step3d), corrector (step3d_t)
// and harmonic tracer mixing (t3dmix2).  Column-marching kernels: one thread per
// water column, i fastest so every row access of a warp is one coalesced stripe;
// the +-2 j-neighbour rows of the U3 stencil are re-read through L1/L2 by the
// other warps of the (32 x 8) block.  Per-column tridiagonal work lives in
// thread-private arrays.  Arithmetic order per point == reference (-fmad=false).
#include "common.cuh"
#include <algorithm>
#include <cstdlib>

struct Edges { int S, N, Jstr, Jend; };
__device__ __forceinline__ Edges edges(const Dev& D) { return Edges{D.b.Southern_Edge && !D.b.NSperiodic, D.b.Northern_Edge && !D.b.NSperiodic, D.b.Jstr, D.b.Jend}; }

// first difference in eta with the closed-wall replacement FE(i,Jstr-1)=FE(i,Jstr),
// FE(i,Jend+2)=FE(i,Jend+1)  (pre_step3d.F:480-493, step3d_t.F:705-723)
__device__ __forceinline__ double dEta(const V3& q, int i, int j, int k, const Edges& e) {
  int jj = j;
  if (e.S && j == e.Jstr - 1) jj = e.Jstr;
  if (e.N && j == e.Jend + 2) jj = e.Jend + 1;
  return q(i, jj, k) - q(i, jj - 1, k);
}
// third-order upstream (U3) horizontal tracer fluxes, pre_step3d.F:406-533 == step3d_t.F:641-767
__device__ __forceinline__ double fluxX_u3(const V3& q, const V3& Huon, int i, int j, int k) {
  const double d0 = q(i - 1, j, k) - q(i - 2, j, k), d1 = q(i, j, k) - q(i - 1, j, k), d2 = q(i + 1, j, k) - q(i, j, k);
  const double cm = d1 - d0, cp = d2 - d1, hu = Huon(i, j, k);
  return hu * 0.5 * (q(i - 1, j, k) + q(i, j, k)) - (1.0 / 6.0) * (cm * fmax(hu, 0.0) + cp * fmin(hu, 0.0));
}
__device__ __forceinline__ double fluxE_u3(const V3& q, const V3& Hvom, int i, int j, int k, const Edges& e) {
  const double d0 = dEta(q, i, j - 1, k, e), d1 = dEta(q, i, j, k, e), d2 = dEta(q, i, j + 1, k, e);
  const double cm = d1 - d0, cp = d2 - d1, hv = Hvom(i, j, k);
  return hv * 0.5 * (q(i, j - 1, k) + q(i, j, k)) - (1.0 / 6.0) * (cm * fmax(hv, 0.0) + cp * fmin(hv, 0.0));
}
// fourth-order centred vertical flux at w-level k (pre_step3d.F:773-808, step3d_t.F:1150-1185)
__device__ __forceinline__ double fluxZ_c4(const V3& q, const V3& W, int i, int j, int k, int N) {
  const double c1 = 0.5, c2 = 7.0 / 12.0, c3 = 1.0 / 12.0;
  if (k == 0 || k == N) return 0.0;
  if (k == 1) return W(i, j, 1) * (c1 * q(i, j, 1) + c2 * q(i, j, 2) - c3 * q(i, j, 3));
  if (k == N - 1) return W(i, j, N - 1) * (c1 * q(i, j, N) + c2 * q(i, j, N - 1) - c3 * q(i, j, N - 2));
  return W(i, j, k) * (c2 * (q(i, j, k) + q(i, j, k + 1)) - c3 * (q(i, j, k - 1) + q(i, j, k + 2)));
}
// store with the closed-wall gradient condition of t3dbc_im.F:334-341,415-422
__device__ __forceinline__ void st_tbc(const Dev& D, const V3& A, int i, int j, int k, double val, const Edges& e) {
  st(D, A, i, j, k, val);
  if (e.S && j == e.Jstr) st(D, A, i, j - 1, k, val);
  if (e.N && j == e.Jend) st(D, A, i, j + 1, k, val);
}

__constant__ double c_mu1[9] = {0.35, 0.6, 1.0, 1.5, 1.4, 0.42, 0.37, 0.33, 0.00468592};   // mod_scalars.F:1585
__constant__ double c_mu2[9] = {23.0, 20.0, 17.0, 14.0, 7.9, 5.13, 3.54, 2.34, 1.51};      // mod_scalars.F:1589
__constant__ double c_r1[9] = {0.58, 0.62, 0.67, 0.77, 0.78, 0.57, 0.57, 0.57, 0.55};      // mod_scalars.F:1593
// lmd_swfrac_tile with Zscale=-1 (lmd_swfrac.F)
__device__ __forceinline__ double swfrac(int Jindex, double Z) {
  const double fac1 = -1.0 / c_mu1[Jindex - 1], fac2 = -1.0 / c_mu2[Jindex - 1], fac3 = c_r1[Jindex - 1];
  return exp(Z * fac1) * fac3 + exp(Z * fac2) * (1.0 - fac3);
}

// ---- pre_step3d_tile, tracer part: pre_step3d.F:329-344,406-932,1152-1168 -----
// One thread per (i,j,k,tracer).  Nothing here is a vertical recurrence: the vertical flux differences only need the
// flux at w-level k-1, which the thread re-evaluates (same operations on the same operands -> same bits as the column
// march of the reference), so all levels run in parallel (30x more warps on a BENCHMARK-size tile).
__global__ void __launch_bounds__(256) pre_step3d_t_kernel(const Dev D, Box bx, int nstp, int nnew, int first) {
  IJZ_FROM_BOX(bx, D.b.N);
  const int N = D.b.N, k = 1 + zlev, itrc = 1 + zcomp; const double dt = D.p.dt;
  const Edges e = edges(D);
  V3 Hz = v3(D, FID(Hz)), Huon = v3(D, FID(Huon)), Hvom = v3(D, FID(Hvom)), W = v3(D, FID(W)), z_r = v3(D, FID(z_r));
  V3 tn = v3l(D, FID(t), nstp, itrc), tw = v3l(D, FID(t), nnew, itrc), t3 = v3l(D, FID(t), 3, itrc);
  V3 Akt = v3l(D, FID(Akt), min(D.b.NAT, itrc));
  const double pmn_pm = v2(D, FID(pm))(i, j), pmn_pn = v2(D, FID(pn))(i, j);
  const double Gamma = 1.0 / 6.0;
  double cff, cff1, cff2;
  if (first) { cff = 0.5 * dt; cff1 = 1.0; cff2 = 0.0; } else { cff = (1.0 - Gamma) * dt; cff1 = 0.5 + Gamma; cff2 = 0.5 - Gamma; }
  const double cpm = cff * pmn_pm * pmn_pn;
  const double hz = Hz(i, j, k), tnk = tn(i, j, k);
  double t3out;                           // both results are stored at the end: a store in the middle would keep the loads of the second half behind it
  {
    // horizontal predictor, then the vertical part with artificial continuity
    const double FXi = fluxX_u3(tn, Huon, i, j, k), FXp = fluxX_u3(tn, Huon, i + 1, j, k);
    const double FEj = fluxE_u3(tn, Hvom, i, j, k, e), FEp = fluxE_u3(tn, Hvom, i, j + 1, k, e);
    double t3h = hz * (cff1 * tnk + cff2 * tw(i, j, k)) - cpm * (FXp - FXi + FEp - FEj);
    const double FCk = fluxZ_c4(tn, W, i, j, k, N), FCm = fluxZ_c4(tn, W, i, j, k - 1, N);
    const double DC = 1.0 / (hz - cpm * (Huon(i + 1, j, k) - Huon(i, j, k) + Hvom(i, j + 1, k) - Hvom(i, j, k) + (W(i, j, k) - W(i, j, k - 1))));
    t3h = DC * (t3h - cpm * (FCk - FCm));
    t3out = t3h;
  }
  // t(nnew) = Hz*t(nstp) + explicit vertical terms (pre_step3d.F:863-932)
  const double cff3 = dt * (1.0 - 1.0 /*lambda*/);
  const bool bench = (D.p.app == ROMS_B200_APP_BENCHMARK);
  double srf = 0.0, zwN = 0.0; int Jw = 1; V3 gh = v3l(D, FID(ghats), min(D.b.NAT, itrc)); V3 z_w = v3(D, FID(z_w));
  if (bench && itrc == 1) { srf = v2(D, FID(srflx))(i, j); zwN = z_w(i, j, N); Jw = (int)v2(D, FID(Jwtype))(i, j); }
  auto vflx = [&](int kk) -> double {                 // FC at w-level kk (0: bottom flux, N: surface flux)
    if (kk == 0) return dt * v2l(D, FID(btflx), itrc)(i, j);
    if (kk == N) return dt * v2l(D, FID(stflx), itrc)(i, j);
    const double c = 1.0 / (z_r(i, j, kk + 1) - z_r(i, j, kk));
    double F = cff3 * c * Akt(i, j, kk) * (tn(i, j, kk + 1) - tn(i, j, kk));
    if (bench) {
      if (itrc <= D.b.NAT) F = F - dt * Akt(i, j, kk) * gh(i, j, kk);
      if (itrc == 1) F = F + dt * srf * swfrac(Jw, zwN - z_w(i, j, kk));
    }
    return F;
  };
  const double Fk = vflx(k), Fm = vflx(k - 1);
  const double a = hz * tnk, bdiff = Fk - Fm;
  st_tbc(D, t3, i, j, k, t3out, e);
  tw(i, j, k) = a + bdiff;
}

// ---- pre_step3d_tile, momentum part: pre_step3d.F:943-1144 ------------------------
// One thread per (i,j,k,component); the flux at w-level k-1 is re-evaluated instead of carried.
__global__ void __launch_bounds__(256) pre_step3d_uv_kernel(const Dev D, Box bx, int nrhs, int nstp, int nnew, int mode) {
  IJZ_FROM_BOX(bx, D.b.N);
  const roms_b200_bounds& b = D.b; const int N = b.N, k = 1 + zlev, comp = zcomp; const double dt = D.p.dt;
  V3 Hz = v3(D, FID(Hz)), z_r = v3(D, FID(z_r)), Akv = v3(D, FID(Akv));
  V2 pm = v2(D, FID(pm)), pn = v2(D, FID(pn));
  const int indx = 3 - nrhs;
  const double cff3 = dt * (1.0 - 1.0 /*lambda*/);
  const int di = comp == 0 ? 1 : 0, dj = 1 - di;   // neighbour offset of the staggered point
  if (comp == 0 && !(i >= b.IstrU && i <= b.Iend && j >= b.Jstr && j <= b.Jend)) return;
  if (comp == 1 && !(i >= b.Istr && i <= b.Iend && j >= b.JstrV && j <= b.Jend)) return;
  V3 q = v3l(D, comp == 0 ? FID(u) : FID(v), nstp), qn = v3l(D, comp == 0 ? FID(u) : FID(v), nnew);
  V3 r1 = v3l(D, comp == 0 ? FID(ru) : FID(rv), nrhs), r2 = v3l(D, comp == 0 ? FID(ru) : FID(rv), indx);
  const double DC0 = (dt * 0.25) * (pm(i, j) + pm(i - di, j - dj)) * (pn(i, j) + pn(i - di, j - dj));
  auto vflx = [&](int kk) -> double {
    if (kk == 0) return dt * v2(D, comp == 0 ? FID(bustr) : FID(bvstr))(i, j);
    if (kk == N) return dt * v2(D, comp == 0 ? FID(sustr) : FID(svstr))(i, j);
    const double c = 1.0 / (z_r(i, j, kk + 1) + z_r(i - di, j - dj, kk + 1) - z_r(i, j, kk) - z_r(i - di, j - dj, kk));
    return cff3 * c * (q(i, j, kk + 1) - q(i, j, kk)) * (Akv(i, j, kk) + Akv(i - di, j - dj, kk));
  };
  const double Fk = vflx(k), Fm = vflx(k - 1);
  const double a = q(i, j, k) * 0.5 * (Hz(i, j, k) + Hz(i - di, j - dj, k));
  const double d = Fk - Fm;
  double val;
  if (mode == 0) val = a + d;
  else if (mode == 1) { const double c3 = 0.5 * DC0; val = a - c3 * r2(i, j, k) + d; }
  else val = a + DC0 * ((5.0 / 12.0) * r1(i, j, k) - (16.0 / 12.0) * r2(i, j, k)) + d;
  qn(i, j, k) = val;
}

// ---- pre_step3d tracers, production form: a block marches a 32 x 8 tile up a chunk of levels ---------------------------------
// (the recipe of t3dmix2_geo_roll_kernel below).  Per level the U3 fluxes FX (u-points) and FE (v-points) are evaluated once
// per point into shared memory from the level's t(nstp) plane (tile + two rings of halo, staged through shared memory and
// loaded into registers one level ahead of its use); the vertical C4 flux and the vertical diffusive flux of w-level k-1 are
// carried from the level below; the column values t(k-1..k+2) roll through registers.  The per-level kernel above evaluates
// every horizontal flux twice, every vertical flux twice (four exp() per cell of tracer 1) and re-reads 5 levels of t.
// Same operations on the same operands in the same order -> same bits.
constexpr int P2_TX = 32, P2_TY = 8, P2_NT = P2_TX * P2_TY;
constexpr int P2_RW = P2_TX + 4, P2_NR = P2_RW * (P2_TY + 4);           // raw t plane: i in [I0-2, I0+33], j in [J0-2, J0+9]
constexpr int P2_XW = P2_TX + 1, P2_NX = P2_XW * P2_TY;                 // u-point plane: i in [I0, I0+32], j in [J0, J0+7]
constexpr int P2_NE = P2_TX * (P2_TY + 1);                               // v-point plane: i in [I0, I0+31], j in [J0, J0+8]
constexpr int P2_FSLOT = 2 * P2_NX + 2 * P2_NE;                          // FX, Huon, FE, Hvom
constexpr int P2_NQ = (P2_NR + P2_NT - 1) / P2_NT;
constexpr size_t P2_SMEM = (3 * P2_FSLOT + 4 * P2_NR) * sizeof(double);
__global__ void __launch_bounds__(P2_NT, 2) pre_step3d_t_roll_kernel(const Dev D, Box bx, int nstp, int nnew, int first, int nch) {
  extern __shared__ double p2sm[];
  double* const smF = p2sm; double* const smR = p2sm + 3 * P2_FSLOT;
  const int N = D.b.N, itrc = 1 + (int)blockIdx.z / nch, ch = (int)blockIdx.z % nch; const double dt = D.p.dt;
  const int per = (N + nch - 1) / nch, k0 = 1 + per * ch, k1 = min(k0 + per - 1, N);
  if (k0 > N) return;
  const Edges e = edges(D);
  const int I0 = bx.i0 + blockIdx.x * P2_TX, J0 = bx.j0 + blockIdx.y * P2_TY;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * P2_TX + tx, i = I0 + tx, j = J0 + ty;
  const bool mine = (i <= bx.i1 && j <= bx.j1);
  V3 Hz = v3(D, FID(Hz)), Huon = v3(D, FID(Huon)), Hvom = v3(D, FID(Hvom)), W = v3(D, FID(W)), z_r = v3(D, FID(z_r)), z_w = v3(D, FID(z_w));
  V3 tn = v3l(D, FID(t), nstp, itrc), tw = v3l(D, FID(t), nnew, itrc), t3 = v3l(D, FID(t), 3, itrc);
  V3 Akt = v3l(D, FID(Akt), min(D.b.NAT, itrc)), gh = v3l(D, FID(ghats), min(D.b.NAT, itrc));
  auto FL = [&](int L) { return smF + (L % 3) * P2_FSLOT; };             // FX; Huon at + P2_NX; FE at + 2*P2_NX; Hvom at + 2*P2_NX + P2_NE
  auto RAW = [&](int L) { return smR + (L & 3) * P2_NR; };
  // the cells this thread handles: raw (up to P2_NQ), u-points and v-points (up to two each)
  int ri[P2_NQ], rj[P2_NQ]; bool ro[P2_NQ];
#pragma unroll
  for (int n = 0; n < P2_NQ; ++n) {
    const int q = tid + n * P2_NT;
    ri[n] = I0 - 2 + q % P2_RW; rj[n] = J0 - 2 + q / P2_RW; ro[n] = q < P2_NR && ri[n] <= bx.i1 + 2 && rj[n] <= bx.j1 + 2 &&
            ri[n] >= D.b.LBi && ri[n] <= D.b.UBi && rj[n] >= D.b.LBj && rj[n] <= D.b.UBj;   // rows beyond a closed wall do not exist (and dEta never uses them)
  }
  int xg_i[2], xg_j[2], eg_i[2], eg_j[2], xr[2], er[2], ed[2][3]; bool xo[2], eo[2];
#pragma unroll
  for (int n = 0; n < 2; ++n) {
    const int q = tid + n * P2_NT;
    const int xi = q % P2_XW, xj = q / P2_XW; xg_i[n] = I0 + xi; xg_j[n] = J0 + xj; xo[n] = q < P2_NX && xg_i[n] <= bx.i1 + 1 && xg_j[n] <= bx.j1;
    xr[n] = (xi + 2) + P2_RW * (xj + 2);
    const int ei = q % P2_TX, ej = q / P2_TX; eg_i[n] = I0 + ei; eg_j[n] = J0 + ej; eo[n] = q < P2_NE && eg_i[n] <= bx.i1 && eg_j[n] <= bx.j1 + 1;
    er[n] = (ei + 2) + P2_RW * (ej + 2);
    // dEta at rows j-1, j, j+1 with the closed-wall replacement: raw offset of the upper row of each difference
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      int jj = eg_j[n] - 1 + m;
      if (e.S && jj == e.Jstr - 1) jj = e.Jstr;
      if (e.N && jj == e.Jend + 2) jj = e.Jend + 1;
      ed[n][m] = (ei + 2) + P2_RW * (jj - J0 + 2);
    }
  }
  double rq[P2_NQ], hu[2], hv[2];
  auto loadR = [&](int L) {
#pragma unroll
    for (int n = 0; n < P2_NQ; ++n) { rq[n] = 0.0; if (ro[n] && L >= 1 && L <= N) rq[n] = tn(ri[n], rj[n], L); }
  };
  auto commitR = [&](int L) {
    double* R = RAW(L);
#pragma unroll
    for (int n = 0; n < P2_NQ; ++n) { const int q = tid + n * P2_NT; if (q < P2_NR) R[q] = rq[n]; }
  };
  auto loadH = [&](int L) {
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      hu[n] = 0.0; hv[n] = 0.0;
      if (L >= 1 && L <= N) { if (xo[n]) hu[n] = Huon(xg_i[n], xg_j[n], L); if (eo[n]) hv[n] = Hvom(eg_i[n], eg_j[n], L); }
    }
  };
  auto deriveF = [&](int L) {                       // fluxX_u3 / fluxE_u3 of level L from its raw plane and hu/hv (of level L)
    const double* R = RAW(L); double* F = FL(L);
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      const int q = tid + n * P2_NT;
      if (xo[n]) {
        const int r = xr[n];
        const double d0 = R[r - 1] - R[r - 2], d1 = R[r] - R[r - 1], d2 = R[r + 1] - R[r];
        const double cm = d1 - d0, cp = d2 - d1, h = hu[n];
        F[q] = h * 0.5 * (R[r - 1] + R[r]) - (1.0 / 6.0) * (cm * fmax(h, 0.0) + cp * fmin(h, 0.0));
        F[P2_NX + q] = h;
      }
      if (eo[n]) {
        const int r = er[n];
        const double d0 = R[ed[n][0]] - R[ed[n][0] - P2_RW], d1 = R[ed[n][1]] - R[ed[n][1] - P2_RW], d2 = R[ed[n][2]] - R[ed[n][2] - P2_RW];
        const double cm = d1 - d0, cp = d2 - d1, h = hv[n];
        F[2 * P2_NX + q] = h * 0.5 * (R[r - P2_RW] + R[r]) - (1.0 / 6.0) * (cm * fmax(h, 0.0) + cp * fmin(h, 0.0));
        F[2 * P2_NX + P2_NE + q] = h;
      }
    }
  };
  const int lx = tx + P2_XW * ty, le = tx + P2_TX * ty, lr = (tx + 2) + P2_RW * (ty + 2);
  // ---- per-column constants
  const double Gamma = 1.0 / 6.0;
  double cff, cff1, cff2;
  if (first) { cff = 0.5 * dt; cff1 = 1.0; cff2 = 0.0; } else { cff = (1.0 - Gamma) * dt; cff1 = 0.5 + Gamma; cff2 = 0.5 - Gamma; }
  const double cff3 = dt * (1.0 - 1.0 /*lambda*/);
  const bool bench = (D.p.app == ROMS_B200_APP_BENCHMARK);
  double cpm = 0.0, srf = 0.0, zwN = 0.0, btf = 0.0, stf = 0.0; int Jw = 1;
  if (mine) {
    cpm = cff * v2(D, FID(pm))(i, j) * v2(D, FID(pn))(i, j);
    if (bench && itrc == 1) { srf = v2(D, FID(srflx))(i, j); zwN = z_w(i, j, N); Jw = (int)v2(D, FID(Jwtype))(i, j); }
    btf = v2l(D, FID(btflx), itrc)(i, j); stf = v2l(D, FID(stflx), itrc)(i, j);
  }
  // vertical diffusive flux at w-level kk from the column values (vflx of the per-level kernel)
  auto vflx = [&](int kk, double zr_hi, double zr_lo, double akt, double ghk, double zwk, double t_hi, double t_lo) -> double {
    if (kk == 0) return dt * btf;
    if (kk == N) return dt * stf;
    const double c = 1.0 / (zr_hi - zr_lo);
    double F = cff3 * c * akt * (t_hi - t_lo);
    if (bench) {
      if (itrc <= D.b.NAT) F = F - dt * akt * ghk;
      if (itrc == 1) F = F + dt * srf * swfrac(Jw, zwN - zwk);
    }
    return F;
  };
  auto fcz = [&](int k, double w, double qm1, double q0, double qp1, double qp2) -> double {    // fluxZ_c4 at w-level k
    const double c1 = 0.5, c2 = 7.0 / 12.0, c3 = 1.0 / 12.0;
    if (k == 0 || k == N) return 0.0;
    if (k == 1) return w * (c1 * q0 + c2 * qp1 - c3 * qp2);
    if (k == N - 1) return w * (c1 * qp1 + c2 * q0 - c3 * qm1);
    return w * (c2 * (q0 + qp1) - c3 * (qm1 + qp2));
  };
  // ---- prologue
  loadR(k0); commitR(k0);
  loadR(k0 + 1); commitR(k0 + 1);
  loadH(k0);
  loadR(k0 + 2);
  // column state at the chunk start: t(k0-2 .. k0+1), W(k0-1), z_r(k0-1), z_r(k0), the two carried fluxes of w-level k0-1
  auto tcol = [&](int L) -> double { return (mine && L >= 1 && L <= N) ? tn(i, j, L) : 0.0; };
  double tm1 = tcol(k0 - 1), t0 = tcol(k0), tp1 = tcol(k0 + 1), tp2 = 0.0;
  double Wm = 0.0, FCm = 0.0, Fm = 0.0, zr0 = 0.0;
  if (mine) {
    const int km = k0 - 1;
    Wm = W(i, j, km);
    FCm = fcz(km, Wm, tcol(km - 1), tm1, t0, tp1);
    zr0 = z_r(i, j, k0);
    Fm = vflx(km, zr0, km >= 1 ? z_r(i, j, km) : 0.0, (km >= 1 && km < N) ? Akt(i, j, km) : 0.0, (km >= 1 && km < N) ? gh(i, j, km) : 0.0,
              (km >= 1 && km < N) ? z_w(i, j, km) : 0.0, t0, tm1);
  }
  // per-level scalars of the column, loaded one level ahead
  double nW = 0.0, nHz = 0.0, ntw = 0.0, nzr = 0.0, nAkt = 0.0, ngh = 0.0, nzw = 0.0;
  auto loadC = [&](int L) {
    if (mine && L <= k1) {
      nW = W(i, j, L); nHz = Hz(i, j, L); ntw = tw(i, j, L);
      if (L < N) { nzr = z_r(i, j, L + 1); nAkt = Akt(i, j, L); ngh = gh(i, j, L); nzw = z_w(i, j, L); }
    }
  };
  loadC(k0);
  __syncthreads();
  deriveF(k0);
  loadH(k0 + 1);
  for (int k = k0; k <= k1; ++k) {
    if (k + 1 <= N) deriveF(k + 1);
    loadH(k + 2);
    commitR(k + 2);
    loadR(k + 3);
    const double Wk = nW, hz = nHz, twk = ntw, zr1 = nzr, akt = nAkt, ghk = ngh, zwk = nzw;
    loadC(k + 1);
    __syncthreads();
    if (mine) {
      tp2 = (k + 2 <= N) ? RAW(k + 2)[lr] : 0.0;
      const double* F = FL(k);
      const double FXi = F[lx], FXp = F[lx + 1], hui = F[P2_NX + lx], hup = F[P2_NX + lx + 1];
      const double FEj = F[2 * P2_NX + le], FEp = F[2 * P2_NX + le + P2_TX], hvj = F[2 * P2_NX + P2_NE + le], hvp = F[2 * P2_NX + P2_NE + le + P2_TX];
      // horizontal predictor, then the vertical part with artificial continuity (pre_step3d.F:406-861)
      double t3h = hz * (cff1 * t0 + cff2 * twk) - cpm * (FXp - FXi + FEp - FEj);
      const double FCk = fcz(k, Wk, tm1, t0, tp1, tp2);
      const double DC = 1.0 / (hz - cpm * (hup - hui + hvp - hvj + (Wk - Wm)));
      t3h = DC * (t3h - cpm * (FCk - FCm));
      // t(nnew) = Hz*t(nstp) + explicit vertical terms (pre_step3d.F:863-932)
      const double Fk = vflx(k, zr1, zr0, akt, ghk, zwk, tp1, t0);
      const double a = hz * t0, bdiff = Fk - Fm;
      st_tbc(D, t3, i, j, k, t3h, e);
      tw(i, j, k) = a + bdiff;
      FCm = FCk; Fm = Fk; Wm = Wk; zr0 = zr1; tm1 = t0; t0 = tp1; tp1 = tp2;
    }
  }
}

// ---- pre_step3d momentum, production form: one thread per (u or v) column and chunk of levels, marching upward --------------
// The vertical viscous flux of w-level k-1 and the z_r pair of level k are carried, everything a level needs is loaded one
// level ahead; the per-level kernel evaluates every flux (a division) twice and reads 19 values per cell instead of 9.
__global__ void __launch_bounds__(256) pre_step3d_uv_march_kernel(const Dev D, Box bx, int nrhs, int nstp, int nnew, int mode, int nch) {
  IJ_FROM_BOX(bx);
  const roms_b200_bounds& b = D.b; const int N = b.N, comp = (int)blockIdx.z / nch, ch = (int)blockIdx.z % nch; const double dt = D.p.dt;
  const int per = (N + nch - 1) / nch, k0 = 1 + per * ch, k1 = min(k0 + per - 1, N);
  if (k0 > N) return;
  if (comp == 0 && !(i >= b.IstrU && i <= b.Iend && j >= b.Jstr && j <= b.Jend)) return;
  if (comp == 1 && !(i >= b.Istr && i <= b.Iend && j >= b.JstrV && j <= b.Jend)) return;
  V3 Hz = v3(D, FID(Hz)), z_r = v3(D, FID(z_r)), Akv = v3(D, FID(Akv));
  V2 pm = v2(D, FID(pm)), pn = v2(D, FID(pn));
  const int indx = 3 - nrhs, di = comp == 0 ? 1 : 0, dj = 1 - di, in = i - di, jn = j - dj;
  const double cff3 = dt * (1.0 - 1.0 /*lambda*/);
  V3 q = v3l(D, comp == 0 ? FID(u) : FID(v), nstp), qn = v3l(D, comp == 0 ? FID(u) : FID(v), nnew);
  V3 r1 = v3l(D, comp == 0 ? FID(ru) : FID(rv), nrhs), r2 = v3l(D, comp == 0 ? FID(ru) : FID(rv), indx);
  const double DC0 = (dt * 0.25) * (pm(i, j) + pm(in, jn)) * (pn(i, j) + pn(in, jn));
  const double Fbot = dt * v2(D, comp == 0 ? FID(bustr) : FID(bvstr))(i, j), Ftop = dt * v2(D, comp == 0 ? FID(sustr) : FID(svstr))(i, j);
  // flux at w-level kk from the z_r pairs of levels kk, kk+1, q(kk), q(kk+1) and the Akv pair of level kk
  auto vflx = [&](int kk, double zo1, double zn1, double zo0, double zn0, double q1, double q0, double ako, double akn) -> double {
    if (kk == 0) return Fbot;
    if (kk == N) return Ftop;
    const double c = 1.0 / (zo1 + zn1 - zo0 - zn0);
    return cff3 * c * (q1 - q0) * (ako + akn);
  };
  // column state at the chunk start: level k0 pair of z_r, q(k0), the flux of w-level k0-1
  double zo0 = z_r(i, j, k0), zn0 = z_r(in, jn, k0), q0 = q(i, j, k0), Fm;
  { const int km = k0 - 1;
    if (km >= 1) Fm = vflx(km, zo0, zn0, z_r(i, j, km), z_r(in, jn, km), q0, q(i, j, km), Akv(i, j, km), Akv(in, jn, km));
    else Fm = Fbot; }
  // level k operands, loaded one level ahead
  double nzo, nzn, nq, nako, nakn, nhz, nhzn, nr1, nr2;
  auto loadL = [&](int k) {
    nzo = 0.0; nzn = 0.0; nq = 0.0; nako = 0.0; nakn = 0.0; nr1 = 0.0; nr2 = 0.0;
    if (k < N) { nzo = z_r(i, j, k + 1); nzn = z_r(in, jn, k + 1); nq = q(i, j, k + 1); nako = Akv(i, j, k); nakn = Akv(in, jn, k); }
    nhz = Hz(i, j, k); nhzn = Hz(in, jn, k);
    if (mode >= 1) nr2 = r2(i, j, k);
    if (mode == 2) nr1 = r1(i, j, k);
  };
  loadL(k0);
  for (int k = k0; k <= k1; ++k) {
    const double zo1 = nzo, zn1 = nzn, q1 = nq, ako = nako, akn = nakn, hz = nhz, hzn = nhzn, r1k = nr1, r2k = nr2;
    if (k + 1 <= k1) loadL(k + 1);
    const double Fk = vflx(k, zo1, zn1, zo0, zn0, q1, q0, ako, akn);
    const double a = q0 * 0.5 * (hz + hzn);
    const double d = Fk - Fm;
    double val;
    if (mode == 0) val = a + d;
    else if (mode == 1) { const double c3 = 0.5 * DC0; val = a - c3 * r2k + d; }
    else val = a + DC0 * ((5.0 / 12.0) * r1k - (16.0 / 12.0) * r2k) + d;
    qn(i, j, k) = val;
    Fm = Fk; zo0 = zo1; zn0 = zn1; q0 = q1;
  }
}

// pre_step3d_tile in its two independent halves: tracers (pre_step3d.F:329-957) and momentum (:960-1168)
int k_pre_step3d_t(roms_b200_ctx* c, int nrhs, int nstp, int nnew, int iic, int ntfirst) {
  (void)nrhs;
  const roms_b200_bounds& b = c->D.b;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk(32, 8); dim3 g = grid2(bx, blk); g.z = b.NT * b.N;
  static const bool per_level = (getenv("ROMS_B200_PRE3D_PERLEVEL") != nullptr);      // the first form
  if (!per_level) {
    static AttrOnce attr;
    if (attr.need(P2_SMEM)) CUDA_OK(cudaFuncSetAttribute(pre_step3d_t_roll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P2_SMEM));
    dim3 blk2(P2_TX, P2_TY); dim3 g2 = grid2(bx, blk2);
    const long cols = (long)g2.x * g2.y * b.NT;
    static const int waves = getenv("ROMS_B200_PRE3D_FILL") ? atoi(getenv("ROMS_B200_PRE3D_FILL")) : 1;
    int nch = (int)std::min<long>(((long)waves * 148 + cols - 1) / cols, (b.N + 4) / 5); if (nch < 1) nch = 1;
    g2.z = b.NT * nch;
    pre_step3d_t_roll_kernel<<<g2, blk2, P2_SMEM, c->stream>>>(c->D, bx, nstp, nnew, iic == ntfirst ? 1 : 0, nch); c->launches++;
    return 0;
  }
  pre_step3d_t_kernel<<<g, blk, 0, c->stream>>>(c->D, bx, nstp, nnew, iic == ntfirst ? 1 : 0); c->launches++;
  return 0;
}
int k_pre_step3d_uv(roms_b200_ctx* c, int nrhs, int nstp, int nnew, int iic, int ntfirst) {
  const roms_b200_bounds& b = c->D.b;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk(32, 8); dim3 g = grid2(bx, blk); g.z = 2 * b.N;
  const int mode = (iic == ntfirst) ? 0 : (iic == ntfirst + 1 ? 1 : 2);
  static const bool per_level = (getenv("ROMS_B200_PRE3D_PERLEVEL") != nullptr);      // the first form
  if (!per_level) {
    dim3 g2 = grid2(bx, blk);
    const long cols = (long)g2.x * g2.y * 2;
    static const int waves = getenv("ROMS_B200_PRE3DUV_FILL") ? atoi(getenv("ROMS_B200_PRE3DUV_FILL")) : 0;
    int nch = (int)std::min<long>(((long)waves * 148 + cols - 1) / cols, (b.N + 4) / 5); if (nch < 1) nch = 1;
    g2.z = 2 * nch;
    pre_step3d_uv_march_kernel<<<g2, blk, 0, c->stream>>>(c->D, bx, nrhs, nstp, nnew, mode, nch); c->launches++;
    return 0;
  }
  pre_step3d_uv_kernel<<<g, blk, 0, c->stream>>>(c->D, bx, nrhs, nstp, nnew, mode); c->launches++;
  return 0;
}
int k_pre_step3d(roms_b200_ctx* c, int nrhs, int nstp, int nnew, int iic, int ntfirst) {
  return k_pre_step3d_t(c, nrhs, nstp, nnew, iic, ntfirst) | k_pre_step3d_uv(c, nrhs, nstp, nnew, iic, ntfirst);
}

// ---- step3d_t_tile: step3d_t.F:393-1924 -> k_step3d_t8.cu (production), k_step3d_t6.cu, k_step3d_t4.cu
int k_step3d_t(roms_b200_ctx* c, int nrhs, int nstp, int nnew) {
  (void)nrhs; (void)nstp;
  // production: the TMA/mbarrier layout of k_step3d_t8.cu; it declines (rc 2) closed W/E walls, N < 4 and level counts whose
  // slots do not fit shared memory (N > ~60), which fall back to the warp-specialised j-march of k_step3d_t6.cu (round 1) and
  // from there to the column march of k_step3d_t4.cu.  ROMS_B200_S3T_V8=0 / ROMS_B200_STEP3D_T_V4=1 select them for A/B timing.
  static const bool use_v4 = (getenv("ROMS_B200_STEP3D_T_V4") != nullptr);   // one-thread-per-column checkpointed Thomas
  static const bool no_v8 = (getenv("ROMS_B200_S3T_V8") != nullptr && atoi(getenv("ROMS_B200_S3T_V8")) == 0);
  if (!use_v4 && !no_v8) { const int rc = k_step3d_t_v8(c, nnew); if (rc != 2) return rc; }
  if (!use_v4) { const int rc = k_step3d_t_v6(c, nnew); if (rc != 2) return rc; }
  return k_step3d_t_v4(c, nnew);
}

// ---- t3dmix2_s_tile, t3dmix2_s.h:198-301 (MIX_S_TS, UPWELLING) ---------------
__global__ void t3dmix2_s_kernel(const Dev D, Box bx, int nrhs, int nnew) {
  IJ_FROM_BOX(bx);
  const int N = D.b.N, itrc = 1 + blockIdx.z; const double dt = D.p.dt;
  V3 Hz = v3(D, FID(Hz)), tr = v3l(D, FID(t), nrhs, itrc), tw = v3l(D, FID(t), nnew, itrc);
  V2 d2 = v2l(D, FID(diff2), itrc), pmon_u = v2(D, FID(pmon_u)), pnom_v = v2(D, FID(pnom_v));
  const double cff = dt * v2(D, FID(pm))(i, j) * v2(D, FID(pn))(i, j);
  const double cx0 = 0.25 * (d2(i, j) + d2(i - 1, j)) * pmon_u(i, j), cx1 = 0.25 * (d2(i + 1, j) + d2(i, j)) * pmon_u(i + 1, j);
  const double ce0 = 0.25 * (d2(i, j) + d2(i, j - 1)) * pnom_v(i, j), ce1 = 0.25 * (d2(i, j + 1) + d2(i, j)) * pnom_v(i, j + 1);
  for (int k = 1; k <= N; ++k) {
    const double FX0 = cx0 * (Hz(i, j, k) + Hz(i - 1, j, k)) * (tr(i, j, k) - tr(i - 1, j, k));
    const double FX1 = cx1 * (Hz(i + 1, j, k) + Hz(i, j, k)) * (tr(i + 1, j, k) - tr(i, j, k));
    const double FE0 = ce0 * (Hz(i, j, k) + Hz(i, j - 1, k)) * (tr(i, j, k) - tr(i, j - 1, k));
    const double FE1 = ce1 * (Hz(i, j + 1, k) + Hz(i, j, k)) * (tr(i, j + 1, k) - tr(i, j, k));
    const double c1 = cff * (FX1 - FX0), c2 = cff * (FE1 - FE0), c3 = c1 + c2;
    tw(i, j, k) = tw(i, j, k) + c3;
  }
}

// ---- t3dmix2_geo_tile, t3dmix2_geo.h:219-419 (MIX_GEO_TS, BENCHMARK) ----------
// Geopotential rotation of the mixing tensor.  The reference rolls two k-levels
// (k1,k2) of dZdx,dTdx,dZde,dTde,dTdz,FS through scratch planes; here each thread
// re-evaluates the slopes it needs for level pair (k,k+1) straight from z_r and t.
struct GeoQ { V3 z_r, tr; V2 pm, pn; int N; V3 dtz; };
// dTdz at w-level kw (0 or N -> 0).  Every cell needs it at 17 (point, level) pairs around it, each a division; the reference
// keeps it in a scratch plane pair (t3dmix2_geo.h:262-285).  Here geo_dTdz_kernel evaluates it once per point into a scratch
// volume (G.dtz) and the flux kernel loads it; without a scratch volume (G.dtz.p == nullptr) it is re-evaluated in place.
__device__ __forceinline__ double g_dTdz_eval(const V3& z_r, const V3& tr, int N, int i, int j, int kw) {
  if (kw == 0 || kw == N) return 0.0;
  const double c = 1.0 / (z_r(i, j, kw + 1) - z_r(i, j, kw));
  return c * (tr(i, j, kw + 1) - tr(i, j, kw));
}
__device__ __forceinline__ double g_dTdz(const GeoQ& G, int i, int j, int kw) {
  if (G.dtz.p) return G.dtz(i, j, kw);
  return g_dTdz_eval(G.z_r, G.tr, G.N, i, j, kw);
}
__global__ void __launch_bounds__(256) geo_dTdz_kernel(const Dev D, Box bx, int nrhs, double* scratch) {
  IJZ_FROM_BOX(bx, D.b.N + 1);
  const int N = D.b.N, kw = zlev, itrc = 1 + zcomp;
  V3 S{scratch + D.nij * (size_t)(N + 1) * (itrc - 1), D.b.LBi, D.ni, D.b.LBj, D.nj, 0};
  S(i, j, kw) = g_dTdz_eval(v3(D, FID(z_r)), v3l(D, FID(t), nrhs, itrc), N, i, j, kw);
}
__device__ __forceinline__ void g_dx(const GeoQ& G, int i, int j, int k, double& dZ, double& dT) {   // at u-point, rho-level k
  const double c = 0.5 * (G.pm(i, j) + G.pm(i - 1, j));
  dZ = c * (G.z_r(i, j, k) - G.z_r(i - 1, j, k)); dT = c * (G.tr(i, j, k) - G.tr(i - 1, j, k));
}
__device__ __forceinline__ void g_de(const GeoQ& G, int i, int j, int k, double& dZ, double& dT) {   // at v-point
  const double c = 0.5 * (G.pn(i, j) + G.pn(i, j - 1));
  dZ = c * (G.z_r(i, j, k) - G.z_r(i, j - 1, k)); dT = c * (G.tr(i, j, k) - G.tr(i, j - 1, k));
}
// FX at u-point (i,j), level k: uses dTdz at w-levels k-1 (index k1 side) and k (k2 side)
__device__ __forceinline__ double g_FX(const GeoQ& G, const V3& Hz, const V2& d2, const V2& on_u, int i, int j, int k) {
  double dZ, dT; g_dx(G, i, j, k, dZ, dT);
  const double c = 0.25 * (d2(i, j) + d2(i - 1, j)) * on_u(i, j);
  return c * (Hz(i, j, k) + Hz(i - 1, j, k)) *
         (dT - 0.5 * (fmin(dZ, 0.0) * (g_dTdz(G, i - 1, j, k - 1) + g_dTdz(G, i, j, k)) +
                      fmax(dZ, 0.0) * (g_dTdz(G, i - 1, j, k) + g_dTdz(G, i, j, k - 1))));
}
__device__ __forceinline__ double g_FE(const GeoQ& G, const V3& Hz, const V2& d2, const V2& om_v, int i, int j, int k) {
  double dZ, dT; g_de(G, i, j, k, dZ, dT);
  const double c = 0.25 * (d2(i, j) + d2(i, j - 1)) * om_v(i, j);
  return c * (Hz(i, j, k) + Hz(i, j - 1, k)) *
         (dT - 0.5 * (fmin(dZ, 0.0) * (g_dTdz(G, i, j - 1, k - 1) + g_dTdz(G, i, j, k)) +
                      fmax(dZ, 0.0) * (g_dTdz(G, i, j - 1, k) + g_dTdz(G, i, j, k - 1))));
}
// FS at w-level k (between rho-levels k and k+1); zero at k=0 and k=N
__device__ __forceinline__ double g_FS(const GeoQ& G, const V2& d2, int i, int j, int k) {
  if (k == 0 || k == G.N) return 0.0;
  const double cff = 0.5 * d2(i, j), tz = g_dTdz(G, i, j, k);
  double zx_a, tx_a, zx_b, tx_b, zx_c, tx_c, zx_d, tx_d;
  g_dx(G, i, j, k, zx_a, tx_a);         // dZdx(i  ,j,k1)
  g_dx(G, i + 1, j, k + 1, zx_b, tx_b); // dZdx(i+1,j,k2)
  g_dx(G, i, j, k + 1, zx_c, tx_c);     // dZdx(i  ,j,k2)
  g_dx(G, i + 1, j, k, zx_d, tx_d);     // dZdx(i+1,j,k1)
  double c1 = fmin(zx_a, 0.0), c2 = fmin(zx_b, 0.0), c3 = fmax(zx_c, 0.0), c4 = fmax(zx_d, 0.0);
  double FS = cff * (c1 * (c1 * tz - tx_a) + c2 * (c2 * tz - tx_b) + c3 * (c3 * tz - tx_c) + c4 * (c4 * tz - tx_d));
  g_de(G, i, j, k, zx_a, tx_a); g_de(G, i, j + 1, k + 1, zx_b, tx_b); g_de(G, i, j, k + 1, zx_c, tx_c); g_de(G, i, j + 1, k, zx_d, tx_d);
  c1 = fmin(zx_a, 0.0); c2 = fmin(zx_b, 0.0); c3 = fmax(zx_c, 0.0); c4 = fmax(zx_d, 0.0);
  FS = FS + cff * (c1 * (c1 * tz - tx_a) + c2 * (c2 * tz - tx_b) + c3 * (c3 * tz - tx_c) + c4 * (c4 * tz - tx_d));
  return FS;
}
// One thread per (i,j,chunk of GEO_KCH levels,tracer): nothing in this operator is a vertical recurrence, so chunks of levels
// run in parallel; FS(k-1) is carried inside a chunk and re-evaluated at its first level (same operations -> same bits).
// (One level per thread doubles the expensive FS work and was measured slower: 168 vs 129 us on 512x64x30.)
constexpr int GEO_KCH = 5;
__global__ void __launch_bounds__(256) t3dmix2_geo_kernel(const Dev D, Box bx, int nrhs, int nnew, double* scratch) {
  IJZ_FROM_BOX(bx, (D.b.N + GEO_KCH - 1) / GEO_KCH);
  const int N = D.b.N, nch = (N + GEO_KCH - 1) / GEO_KCH, k0 = 1 + GEO_KCH * zlev, itrc = 1 + zcomp; const double dt = D.p.dt;
  const int k1 = min(k0 + GEO_KCH - 1, N);
  V3 Hz = v3(D, FID(Hz)), tw = v3l(D, FID(t), nnew, itrc);
  GeoQ G{v3(D, FID(z_r)), v3l(D, FID(t), nrhs, itrc), v2(D, FID(pm)), v2(D, FID(pn)), N,
         V3{scratch ? scratch + D.nij * (size_t)(N + 1) * (itrc - 1) : nullptr, D.b.LBi, D.ni, D.b.LBj, D.nj, 0}};
  V2 d2 = v2l(D, FID(diff2), itrc), on_u = v2(D, FID(on_u)), om_v = v2(D, FID(om_v));
  const double cff = dt * G.pm(i, j) * G.pn(i, j);
  double FSm = g_FS(G, d2, i, j, k0 - 1);        // FS at w-level k-1
  for (int k = k0; k <= k1; ++k) {
    const double FX0 = g_FX(G, Hz, d2, on_u, i, j, k), FX1 = g_FX(G, Hz, d2, on_u, i + 1, j, k);
    const double FE0 = g_FE(G, Hz, d2, om_v, i, j, k), FE1 = g_FE(G, Hz, d2, om_v, i, j + 1, k);
    const double FSk = g_FS(G, d2, i, j, k);
    const double c1 = cff * (FX1 - FX0), c2 = cff * (FE1 - FE0), c3 = dt * (FSk - FSm), c4 = c1 + c2 + c3;
    tw(i, j, k) = tw(i, j, k) + c4;
    FSm = FSk;
  }
}
// ---- t3dmix2_geo, production form: the reference's rolling window, in shared memory -------------------------------------
// A block owns a 32 x 8 tile for a chunk of levels and marches upward.  Per level the slopes dZdx, dTdx (u-points), dZde, dTde
// (v-points) and dTdz (w-points, one ring of halo) are evaluated ONCE per point into shared memory -- the scratch plane pairs
// (k1,k2) of t3dmix2_geo.h:219-419 -- and FX, FE, FS read them there; the per-level kernel above re-evaluated every slope at each
// of its ~6 uses from global memory (ncu: ~100 loads per cell, L1-bound at 0.09 of the HBM roofline) and needed dTdz parked
// in a scratch volume by a kernel of its own.  The operands z_r, t, Hz of a level go through shared memory too (tile + one ring
// of halo: 4 global loads per thread and level instead of ~20) and are loaded into registers one whole level AHEAD of their
// use, so the global-memory latency overlaps a level of work.  Per level k:
//   derive  slopes of level k+1 and dTdz(k) from the raw planes of levels k, k+1      (shared -> shared)
//   commit  the registers (raw level k+2) to their slot; issue the loads of level k+3  (global -> registers, in flight)
//   barrier
//   fluxes  FX, FE (level k), FS (w-level k) from the derived planes, update t(nnew)
// One barrier per level: derived planes rotate through 3 slots, raw planes through 4 (a slot is rewritten two / three levels
// after its last reader).  Same operations on the same operands in the same order -> same bits.
constexpr int G2_TX = 32, G2_TY = 8, G2_NT = G2_TX * G2_TY;
constexpr int G2_XW = G2_TX + 1, G2_NDX = G2_XW * G2_TY;              // u-point plane: i in [I0, I0+32], j in [J0, J0+7]
constexpr int G2_NDE = G2_TX * (G2_TY + 1);                             // v-point plane: i in [I0, I0+31], j in [J0, J0+8]
constexpr int G2_ZW = G2_TX + 2, G2_NTZ = G2_ZW * (G2_TY + 2);          // halo'd plane: i in [I0-1, I0+32], j in [J0-1, J0+8]
constexpr int G2_DSLOT = 2 * G2_NDX + 2 * G2_NDE + G2_NTZ;              // derived: dZdx, dTdx, dZde, dTde, dTdz
constexpr int G2_RSLOT = 3 * G2_NTZ;                                    // raw: z_r, t, Hz
constexpr int G2_NQ = (G2_NTZ + G2_NT - 1) / G2_NT;                     // halo'd cells per thread (2)
constexpr size_t G2_SMEM = (3 * G2_DSLOT + 4 * G2_RSLOT) * sizeof(double);
template <int MINB>
__global__ void __launch_bounds__(G2_NT, MINB) t3dmix2_geo_roll_kernel(const Dev D, Box bx, int nrhs, int nnew, int nch) {
  extern __shared__ double g2sm[];
  double* const smD = g2sm; double* const smR = g2sm + 3 * G2_DSLOT;
  const int N = D.b.N, itrc = 1 + (int)blockIdx.z / nch, ch = (int)blockIdx.z % nch; const double dt = D.p.dt;
  const int per = (N + nch - 1) / nch, k0 = 1 + per * ch, k1 = min(k0 + per - 1, N);
  if (k0 > N) return;
  const int I0 = bx.i0 + blockIdx.x * G2_TX, J0 = bx.j0 + blockIdx.y * G2_TY;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * G2_TX + tx, i = I0 + tx, j = J0 + ty;
  const bool mine = (i <= bx.i1 && j <= bx.j1);
  V3 Hz = v3(D, FID(Hz)), tw = v3l(D, FID(t), nnew, itrc), z_r = v3(D, FID(z_r)), tr = v3l(D, FID(t), nrhs, itrc);
  V2 pm = v2(D, FID(pm)), pn = v2(D, FID(pn)), d2 = v2l(D, FID(diff2), itrc), on_u = v2(D, FID(on_u)), om_v = v2(D, FID(om_v));
  auto DXz = [&](int L) { return smD + (L % 3) * G2_DSLOT; };                 // dZdx of level L; dTdx follows at + G2_NDX
  auto DEz = [&](int L) { return smD + (L % 3) * G2_DSLOT + 2 * G2_NDX; };    // dZde; dTde at + G2_NDE
  auto TZ = [&](int kw) { return smD + (kw % 3) * G2_DSLOT + 2 * G2_NDX + 2 * G2_NDE; };
  auto RAW = [&](int L) { return smR + (L & 3) * G2_RSLOT; };                 // z_r of level L; t at + G2_NTZ; Hz at + 2 * G2_NTZ
  // cells this thread handles in each plane (up to two), as offsets into the halo'd raw planes, and their metric factors
  int xq[2], eq[2]; bool xo[2], eo[2], zo[G2_NQ]; int zi[G2_NQ], zj[G2_NQ]; double xc[2], ec[2];
#pragma unroll
  for (int n = 0; n < 2; ++n) {
    const int q = tid + n * G2_NT;
    const int xi = q % G2_XW, xj = q / G2_XW; xo[n] = q < G2_NDX && I0 + xi <= bx.i1 + 1 && J0 + xj <= bx.j1; xq[n] = (xi + 1) + G2_ZW * (xj + 1);
    const int ei = q % G2_TX, ej = q / G2_TX; eo[n] = q < G2_NDE && I0 + ei <= bx.i1 && J0 + ej <= bx.j1 + 1; eq[n] = (ei + 1) + G2_ZW * (ej + 1);
    xc[n] = xo[n] ? 0.5 * (pm(I0 + xi, J0 + xj) + pm(I0 + xi - 1, J0 + xj)) : 0.0;
    ec[n] = eo[n] ? 0.5 * (pn(I0 + ei, J0 + ej) + pn(I0 + ei, J0 + ej - 1)) : 0.0;
  }
#pragma unroll
  for (int n = 0; n < G2_NQ; ++n) {
    const int q = tid + n * G2_NT;
    zi[n] = I0 - 1 + q % G2_ZW; zj[n] = J0 - 1 + q / G2_ZW; zo[n] = q < G2_NTZ && zi[n] <= bx.i1 + 1 && zj[n] <= bx.j1 + 1;
    zo[n] = zo[n] && zi[n] >= D.b.LBi && zi[n] <= D.b.UBi && zj[n] >= D.b.LBj && zj[n] <= D.b.UBj;
  }
  double rz[G2_NQ], rt[G2_NQ], rh[G2_NQ];          // a raw level in flight
  auto loadR = [&](int L) {                         // global -> registers (levels outside 1..N: nothing)
#pragma unroll
    for (int n = 0; n < G2_NQ; ++n) {
      rz[n] = 0.0; rt[n] = 0.0; rh[n] = 0.0;
      if (zo[n] && L >= 1 && L <= N) { rz[n] = z_r(zi[n], zj[n], L); rt[n] = tr(zi[n], zj[n], L); rh[n] = Hz(zi[n], zj[n], L); }
    }
  };
  auto commitR = [&](int L) {                       // registers -> raw slot of level L
    double* R = RAW(L);
#pragma unroll
    for (int n = 0; n < G2_NQ; ++n) { const int q = tid + n * G2_NT; if (q < G2_NTZ) { R[q] = rz[n]; R[G2_NTZ + q] = rt[n]; R[2 * G2_NTZ + q] = rh[n]; } }
  };
  auto deriveD = [&](int L) {                       // slopes of rho-level L (g_dx, g_de) from its raw planes
    const double* Rz = RAW(L); const double* Rt = Rz + G2_NTZ; double* X = DXz(L); double* E = DEz(L);
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      const int q = tid + n * G2_NT;
      if (xo[n]) { X[q] = xc[n] * (Rz[xq[n]] - Rz[xq[n] - 1]); X[G2_NDX + q] = xc[n] * (Rt[xq[n]] - Rt[xq[n] - 1]); }
      if (eo[n]) { E[q] = ec[n] * (Rz[eq[n]] - Rz[eq[n] - G2_ZW]); E[G2_NDE + q] = ec[n] * (Rt[eq[n]] - Rt[eq[n] - G2_ZW]); }
    }
  };
  auto deriveTZ = [&](int kw) {                     // dTdz at w-level kw (g_dTdz_eval) from the raw planes of levels kw, kw+1
    double* T = TZ(kw); const double* Rl = RAW(kw); const double* Rh = RAW(kw + 1);
#pragma unroll
    for (int n = 0; n < G2_NQ; ++n) {
      const int q = tid + n * G2_NT;
      if (!zo[n]) continue;
      if (kw == 0 || kw == N) { T[q] = 0.0; continue; }
      const double c = 1.0 / (Rh[q] - Rl[q]);
      T[q] = c * (Rh[G2_NTZ + q] - Rl[G2_NTZ + q]);
    }
  };
  const int lx = tx + G2_XW * ty, le = tx + G2_TX * ty, lz = (tx + 1) + G2_ZW * (ty + 1);   // this thread's point in each plane
  auto FSat = [&](int k, double d2c) -> double {                         // g_FS
    if (k == 0 || k == N) return 0.0;
    const double cff = 0.5 * d2c, tz = TZ(k)[lz];
    const double *Xa = DXz(k), *Xb = DXz(k + 1), *Ea = DEz(k), *Eb = DEz(k + 1);
    double c1 = fmin(Xa[lx], 0.0), c2 = fmin(Xb[lx + 1], 0.0), c3 = fmax(Xb[lx], 0.0), c4 = fmax(Xa[lx + 1], 0.0);
    double FS = cff * (c1 * (c1 * tz - Xa[G2_NDX + lx]) + c2 * (c2 * tz - Xb[G2_NDX + lx + 1]) + c3 * (c3 * tz - Xb[G2_NDX + lx]) + c4 * (c4 * tz - Xa[G2_NDX + lx + 1]));
    c1 = fmin(Ea[le], 0.0); c2 = fmin(Eb[le + G2_TX], 0.0); c3 = fmax(Eb[le], 0.0); c4 = fmax(Ea[le + G2_TX], 0.0);
    FS = FS + cff * (c1 * (c1 * tz - Ea[G2_NDE + le]) + c2 * (c2 * tz - Eb[G2_NDE + le + G2_TX]) + c3 * (c3 * tz - Eb[G2_NDE + le]) + c4 * (c4 * tz - Ea[G2_NDE + le + G2_TX]));
    return FS;
  };
  // per-column constants of this thread's point (g_FX / g_FE coefficients, t3dmix2_geo.h:300-330)
  double cff = 0.0, d2c = 0.0, cx0 = 0.0, cx1 = 0.0, ce0 = 0.0, ce1 = 0.0;
  if (mine) {
    cff = dt * pm(i, j) * pn(i, j); d2c = d2(i, j);
    cx0 = 0.25 * (d2c + d2(i - 1, j)) * on_u(i, j); cx1 = 0.25 * (d2(i + 1, j) + d2c) * on_u(i + 1, j);
    ce0 = 0.25 * (d2c + d2(i, j - 1)) * om_v(i, j); ce1 = 0.25 * (d2(i, j + 1) + d2c) * om_v(i, j + 1);
  }
  // ---- prologue: raw levels k0-1 .. k0+1 into their slots, level k0+2 into the registers; derived planes of k0-1, k0
  loadR(k0 - 1); commitR(k0 - 1);
  loadR(k0); commitR(k0);
  loadR(k0 + 1); commitR(k0 + 1);
  loadR(k0 + 2);
  double twn = mine ? tw(i, j, k0) : 0.0;
  __syncthreads();
  if (k0 > 1) deriveD(k0 - 1);
  deriveD(k0);
  deriveTZ(k0 - 1);
  __syncthreads();
  double FSm = mine ? FSat(k0 - 1, d2c) : 0.0;
  for (int k = k0; k <= k1; ++k) {
    if (k + 1 <= N) deriveD(k + 1);
    deriveTZ(k);
    commitR(k + 2);
    loadR(k + 3);
    const double twk = twn;
    if (mine && k + 1 <= k1) twn = tw(i, j, k + 1);
    __syncthreads();
    if (mine) {
      const double *X = DXz(k), *E = DEz(k), *Tm = TZ(k - 1), *Tk = TZ(k), *Rh = RAW(k) + 2 * G2_NTZ;
      const double hz = Rh[lz], hzw = Rh[lz - 1], hze = Rh[lz + 1], hzs = Rh[lz - G2_ZW], hzn = Rh[lz + G2_ZW];
      // g_FX at (i,j) and (i+1,j); g_FE at (i,j) and (i,j+1)
      const double dZ0 = X[lx], dT0 = X[G2_NDX + lx], dZ1 = X[lx + 1], dT1 = X[G2_NDX + lx + 1];
      const double FX0 = cx0 * (hz + hzw) * (dT0 - 0.5 * (fmin(dZ0, 0.0) * (Tm[lz - 1] + Tk[lz]) + fmax(dZ0, 0.0) * (Tk[lz - 1] + Tm[lz])));
      const double FX1 = cx1 * (hze + hz) * (dT1 - 0.5 * (fmin(dZ1, 0.0) * (Tm[lz] + Tk[lz + 1]) + fmax(dZ1, 0.0) * (Tk[lz] + Tm[lz + 1])));
      const double eZ0 = E[le], eT0 = E[G2_NDE + le], eZ1 = E[le + G2_TX], eT1 = E[G2_NDE + le + G2_TX];
      const double FE0 = ce0 * (hz + hzs) * (eT0 - 0.5 * (fmin(eZ0, 0.0) * (Tm[lz - G2_ZW] + Tk[lz]) + fmax(eZ0, 0.0) * (Tk[lz - G2_ZW] + Tm[lz])));
      const double FE1 = ce1 * (hzn + hz) * (eT1 - 0.5 * (fmin(eZ1, 0.0) * (Tm[lz] + Tk[lz + G2_ZW]) + fmax(eZ1, 0.0) * (Tk[lz] + Tm[lz + G2_ZW])));
      const double FSk = FSat(k, d2c);
      const double c1 = cff * (FX1 - FX0), c2 = cff * (FE1 - FE0), c3 = dt * (FSk - FSm), c4 = c1 + c2 + c3;
      tw(i, j, k) = twk + c4;
      FSm = FSk;
    }
  }
}
int k_t3dmix2(roms_b200_ctx* c, int nrhs, int nstp, int nnew) {
  (void)nstp;
  const roms_b200_bounds& b = c->D.b;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk(32, 8); dim3 g = grid2(bx, blk); g.z = b.NT;
  if (c->D.p.app == ROMS_B200_APP_UPWELLING) t3dmix2_s_kernel<<<g, blk, 0, c->stream>>>(c->D, bx, nrhs, nnew);
  else {
    // dTdz once per point into its own scratch volumes (NT volumes of (ni,nj,0:N); not the KPP scratch: the tracer branch of
    // main3d runs beside uv3dmix2, which parks its column terms there), on the points the fluxes of the interior reach:
    // i-1..i+1, j-1..j+1
    static const bool per_level = (getenv("ROMS_B200_T3DMIX_PERLEVEL") != nullptr);     // the first form (two kernels + scratch volume)
    if (!per_level) {
      // chunks of levels only as far as needed to fill the machine (each chunk re-evaluates one level of planes to start)
      dim3 blk2(G2_TX, G2_TY); dim3 g2 = grid2(bx, blk2);
      const long cols = (long)g2.x * g2.y * b.NT;
      static const int waves = getenv("ROMS_B200_T3DMIX_FILL") ? atoi(getenv("ROMS_B200_T3DMIX_FILL")) : 2;
      int nch = (int)std::min<long>(((long)waves * 148 + cols - 1) / cols, (b.N + 4) / 5); if (nch < 1) nch = 1;
      g2.z = b.NT * nch;
      static const int minb = getenv("ROMS_B200_T3DMIX_MINB") ? atoi(getenv("ROMS_B200_T3DMIX_MINB")) : 2;
      static AttrOnce attr;
      if (attr.need(G2_SMEM)) {
        CUDA_OK(cudaFuncSetAttribute(t3dmix2_geo_roll_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G2_SMEM));
        CUDA_OK(cudaFuncSetAttribute(t3dmix2_geo_roll_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G2_SMEM));
        CUDA_OK(cudaFuncSetAttribute(t3dmix2_geo_roll_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G2_SMEM));
      }
      if (minb == 4) t3dmix2_geo_roll_kernel<4><<<g2, blk2, G2_SMEM, c->stream>>>(c->D, bx, nrhs, nnew, nch);
      else if (minb == 3) t3dmix2_geo_roll_kernel<3><<<g2, blk2, G2_SMEM, c->stream>>>(c->D, bx, nrhs, nnew, nch);
      else t3dmix2_geo_roll_kernel<2><<<g2, blk2, G2_SMEM, c->stream>>>(c->D, bx, nrhs, nnew, nch);
      c->launches++;
      return 0;
    }
    double* scratch = c->D.dtdz;
    if (scratch) {
      Box bd{b.Istr - 1, b.Iend + 1, b.Jstr - 1, b.Jend + 1}; dim3 gd = grid2(bd, blk); gd.z = b.NT * (b.N + 1);
      geo_dTdz_kernel<<<gd, blk, 0, c->stream>>>(c->D, bd, nrhs, scratch); c->launches++;
    }
    g.z = b.NT * ((b.N + GEO_KCH - 1) / GEO_KCH); t3dmix2_geo_kernel<<<g, blk, 0, c->stream>>>(c->D, bx, nrhs, nnew, scratch);
  }
  c->launches++;
  return 0;
}
