// roms_b200/csrc/k_tracer.cu -- tracer predictor (pre_step3d), corrector (step3d_t)
// and harmonic tracer mixing (t3dmix2).  Column-marching kernels: one thread per
// water column, i fastest so every row access of a warp is one coalesced stripe;
// the +-2 j-neighbour rows of the U3 stencil are re-read through L1/L2 by the
// other warps of the (32 x 8) block.  Per-column tridiagonal work lives in
// thread-private arrays.  Arithmetic order per point == reference (-fmad=false).
#include "common.cuh"
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
  IJ_FROM_BOX(bx);
  const int N = D.b.N, k = 1 + blockIdx.z % N, itrc = 1 + blockIdx.z / N; const double dt = D.p.dt;
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
  {
    // horizontal predictor, then the vertical part with artificial continuity
    const double FXi = fluxX_u3(tn, Huon, i, j, k), FXp = fluxX_u3(tn, Huon, i + 1, j, k);
    const double FEj = fluxE_u3(tn, Hvom, i, j, k, e), FEp = fluxE_u3(tn, Hvom, i, j + 1, k, e);
    double t3h = hz * (cff1 * tnk + cff2 * tw(i, j, k)) - cpm * (FXp - FXi + FEp - FEj);
    const double FCk = fluxZ_c4(tn, W, i, j, k, N), FCm = fluxZ_c4(tn, W, i, j, k - 1, N);
    const double DC = 1.0 / (hz - cpm * (Huon(i + 1, j, k) - Huon(i, j, k) + Hvom(i, j + 1, k) - Hvom(i, j, k) + (W(i, j, k) - W(i, j, k - 1))));
    t3h = DC * (t3h - cpm * (FCk - FCm));
    st_tbc(D, t3, i, j, k, t3h, e);
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
  tw(i, j, k) = a + bdiff;
}

// ---- pre_step3d_tile, momentum part: pre_step3d.F:943-1144 ------------------------
// One thread per (i,j,k,component); the flux at w-level k-1 is re-evaluated instead of carried.
__global__ void __launch_bounds__(256) pre_step3d_uv_kernel(const Dev D, Box bx, int nrhs, int nstp, int nnew, int mode) {
  IJ_FROM_BOX(bx);
  const roms_b200_bounds& b = D.b; const int N = b.N, k = 1 + blockIdx.z % N, comp = blockIdx.z / N; const double dt = D.p.dt;
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

// pre_step3d_tile in its two independent halves: tracers (pre_step3d.F:329-957) and momentum (:960-1168)
int k_pre_step3d_t(roms_b200_ctx* c, int nrhs, int nstp, int nnew, int iic, int ntfirst) {
  (void)nrhs;
  const roms_b200_bounds& b = c->D.b;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk(32, 8); dim3 g = grid2(bx, blk); g.z = b.NT * b.N;
  pre_step3d_t_kernel<<<g, blk, 0, c->stream>>>(c->D, bx, nstp, nnew, iic == ntfirst ? 1 : 0); c->launches++;
  return 0;
}
int k_pre_step3d_uv(roms_b200_ctx* c, int nrhs, int nstp, int nnew, int iic, int ntfirst) {
  const roms_b200_bounds& b = c->D.b;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk(32, 8); dim3 g = grid2(bx, blk); g.z = 2 * b.N;
  const int mode = (iic == ntfirst) ? 0 : (iic == ntfirst + 1 ? 1 : 2);
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
  IJ_FROM_BOX(bx);
  const int N = D.b.N, kw = blockIdx.z % (N + 1), itrc = 1 + blockIdx.z / (N + 1);
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
  IJ_FROM_BOX(bx);
  const int N = D.b.N, nch = (N + GEO_KCH - 1) / GEO_KCH, k0 = 1 + GEO_KCH * (blockIdx.z % nch), itrc = 1 + blockIdx.z / nch; const double dt = D.p.dt;
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
int k_t3dmix2(roms_b200_ctx* c, int nrhs, int nstp, int nnew) {
  (void)nstp;
  const roms_b200_bounds& b = c->D.b;
  Box bx{b.Istr, b.Iend, b.Jstr, b.Jend}; dim3 blk(32, 8); dim3 g = grid2(bx, blk); g.z = b.NT;
  if (c->D.p.app == ROMS_B200_APP_UPWELLING) t3dmix2_s_kernel<<<g, blk, 0, c->stream>>>(c->D, bx, nrhs, nnew);
  else {
    // dTdz once per point into its own scratch volumes (NT volumes of (ni,nj,0:N); not the KPP scratch: the tracer branch of
    // main3d runs beside uv3dmix2, which parks its column terms there), on the points the fluxes of the interior reach:
    // i-1..i+1, j-1..j+1
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
