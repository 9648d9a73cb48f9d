// roms_b200/csrc/k_step2d.cu -- one barotropic sub-step (LF predictor or AM3
// corrector) of step2d_tile, Nonlinear/step2d_LF_AM3.h:606-3056, as ONE kernel.
//
// The reference stages ~20 tile-sized scratch planes (Drhs, DUon, DVom, zeta_new,
// gzeta..., UFx, UFe, ...) between loop nests.  Here every thread owns one (i,j)
// and re-evaluates the staggered transports / provisional free surface it needs
// from the time levels it only READS (krhs,kstp), so the whole sub-step is a
// single launch with no inter-block dependency: the 2*nfast+1 launches of a
// baroclinic step are captured in a CUDA graph (roms_b200.cu).  All 2-D state of
// a benchmark-size tile (~50 planes x 277 KB) stays resident in the 126 MB L2.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <vector>

// Block tile: 32 x 4 points.  Drhs, DUon, DVom (step2d_LF_AM3.h:664-702) are evaluated ONCE per point of the tile plus the
// halo the 4th-order fluxes reach (3 west/south, 2 east/north) into shared memory; everything downstream reads them there
// (the first layout re-evaluated them per use: ~100 DUon/DVom per thread).  Same operations on the same operands -> same bits.
constexpr int S2_TX = 32, S2_TY = 4, S2_TW = S2_TX + 5, S2_TH = S2_TY + 5;
// ubar/vbar(:,:,krhs) and pm, pn of the tile + halo live in shared memory too: every flux reads them at up to 5 x 5 neighbours
// (about half of the ~90 global loads of the main phase), and filling them together with Drhs makes it ONE round of L2 loads
// before the main phase instead of two (ncu: the sub-step is bound by dependent load rounds, long-scoreboard 7.8 warps per issue).
struct T2 {
  const double* t; int ti0, tj0;
  __device__ __forceinline__ double operator()(int i, int j) const { return t[(i - ti0) + S2_TW * (j - tj0)]; }
};
// M: how the time-invariant planes (h, metrics, rhoA/rhoS) are reached: V2 = global memory (per-sub-step kernel),
// T2 = shared-memory tiles filled once per fast loop (persistent kernel)
template <class M>
struct S2 {
  T2 zk; T2 ub, vb;                       // zeta, ubar, vbar (:,:,krhs) tiles
  M h; T2 pm, pn; M on_u, om_v, rhoA, rhoS;
  double fac, dtfast; int mode;           // mode: 0 iif==1, 1 predictor, 2 corrector
  int S, N, Jstr, Jend;
  const double *tDr, *tDU, *tDV;          // shared-memory tiles, element (i,j) at (i-ti0) + S2_TW*(j-tj0)
  int ti0, tj0;
};
template <class S> __device__ __forceinline__ double Drhs(const S& s, int i, int j) { return s.tDr[(i - s.ti0) + S2_TW * (j - s.tj0)]; }
template <class S> __device__ __forceinline__ double DUon(const S& s, int i, int j) { return s.tDU[(i - s.ti0) + S2_TW * (j - s.tj0)]; }
template <class S> __device__ __forceinline__ double DVom(const S& s, int i, int j) { return s.tDV[(i - s.ti0) + S2_TW * (j - s.tj0)]; }
struct Zst { double rhs_zeta, zn, Dnew, zwrk, gzeta, gzeta2, gzetaSA; };
// step2d_LF_AM3.h:899-980; zs = zeta(i,j,kstp), rzs = rzeta(i,j,kstp), rzp = rzeta(i,j,ptsk) (the last two are read by the corrector only)
template <class S> __device__ __forceinline__ Zst zstate(const S& s, int i, int j, double zs, double rzs, double rzp) {
  Zst z;
  const double div = (DUon(s, i, j) - DUon(s, i + 1, j)) + (DVom(s, i, j) - DVom(s, i, j + 1));
  const double pmn = s.pm(i, j) * s.pn(i, j);
  z.rhs_zeta = div;
  if (s.mode == 0) {
    z.zn = zs + pmn * s.dtfast * div;
    z.zwrk = 0.5 * (zs + z.zn);
  } else if (s.mode == 1) {
    const double cff1 = 2.0 * s.dtfast, cff4 = 4.0 / 25.0, cff5 = 1.0 - 2.0 * cff4;
    z.zn = zs + pmn * cff1 * div;
    z.zwrk = cff5 * s.zk(i, j) + cff4 * (zs + z.zn);
  } else {
    const double cff1 = s.dtfast * 5.0 / 12.0, cff2 = s.dtfast * 8.0 / 12.0, cff3 = s.dtfast * 1.0 / 12.0, cff4 = 2.0 / 5.0, cff5 = 1.0 - cff4;
    const double cff = cff1 * div;
    z.zn = zs + pmn * (cff + cff2 * rzs - cff3 * rzp);
    z.zwrk = cff5 * z.zn + cff4 * s.zk(i, j);
  }
  z.Dnew = z.zn + s.h(i, j);
  z.gzeta = (s.fac + s.rhoS(i, j)) * z.zwrk;
  z.gzeta2 = z.gzeta * z.zwrk;
  z.gzetaSA = z.zwrk * (s.rhoS(i, j) - s.rhoA(i, j));
  return z;
}
// second differences with the closed-wall replacements (:1283-1296, 1361-1376)
template <class S> __device__ __forceinline__ double g_ux(const S& s, int i, int j) { return s.ub(i - 1, j) - 2.0 * s.ub(i, j) + s.ub(i + 1, j); }
template <class S> __device__ __forceinline__ double g_Dux(const S& s, int i, int j) { return DUon(s, i - 1, j) - 2.0 * DUon(s, i, j) + DUon(s, i + 1, j); }
template <class S> __device__ __forceinline__ double g_ue(const S& s, int i, int j) {
  int jj = j; if (s.S && j == s.Jstr - 1) jj = s.Jstr; if (s.N && j == s.Jend + 1) jj = s.Jend;
  return s.ub(i, jj - 1) - 2.0 * s.ub(i, jj) + s.ub(i, jj + 1);
}
template <class S> __device__ __forceinline__ double g_Dvx(const S& s, int i, int j) { return DVom(s, i - 1, j) - 2.0 * DVom(s, i, j) + DVom(s, i + 1, j); }
template <class S> __device__ __forceinline__ double g_vx(const S& s, int i, int j) { return s.vb(i - 1, j) - 2.0 * s.vb(i, j) + s.vb(i + 1, j); }
template <class S> __device__ __forceinline__ double g_Due(const S& s, int i, int j) { return DUon(s, i, j - 1) - 2.0 * DUon(s, i, j) + DUon(s, i, j + 1); }
template <class S> __device__ __forceinline__ double g_ve(const S& s, int i, int j) {
  int jj = j; if (s.S && j == s.Jstr) jj = s.Jstr + 1; if (s.N && j == s.Jend + 1) jj = s.Jend;
  return s.vb(i, jj - 1) - 2.0 * s.vb(i, jj) + s.vb(i, jj + 1);
}
template <class S> __device__ __forceinline__ double g_Dve(const S& s, int i, int j) {
  int jj = j; if (s.S && j == s.Jstr) jj = s.Jstr + 1; if (s.N && j == s.Jend + 1) jj = s.Jend;
  return DVom(s, i, jj - 1) - 2.0 * DVom(s, i, jj) + DVom(s, i, jj + 1);
}
#define C6 (1.0 / 6.0)
// fourth-order centred advective fluxes (:1298-1393)
template <class S> __device__ __forceinline__ double a_UFx(const S& s, int i, int j) {
  return 0.25 * (s.ub(i, j) + s.ub(i + 1, j) - C6 * (g_ux(s, i, j) + g_ux(s, i + 1, j))) *
         (DUon(s, i, j) + DUon(s, i + 1, j) - C6 * (g_Dux(s, i, j) + g_Dux(s, i + 1, j)));
}
template <class S> __device__ __forceinline__ double a_UFe(const S& s, int i, int j) {
  return 0.25 * (s.ub(i, j) + s.ub(i, j - 1) - C6 * (g_ue(s, i, j) + g_ue(s, i, j - 1))) *
         (DVom(s, i, j) + DVom(s, i - 1, j) - C6 * (g_Dvx(s, i, j) + g_Dvx(s, i - 1, j)));
}
template <class S> __device__ __forceinline__ double a_VFx(const S& s, int i, int j) {
  return 0.25 * (s.vb(i, j) + s.vb(i - 1, j) - C6 * (g_vx(s, i, j) + g_vx(s, i - 1, j))) *
         (DUon(s, i, j) + DUon(s, i, j - 1) - C6 * (g_Due(s, i, j) + g_Due(s, i, j - 1)));
}
template <class S> __device__ __forceinline__ double a_VFe(const S& s, int i, int j) {
  return 0.25 * (s.vb(i, j) + s.vb(i, j + 1) - C6 * (g_ve(s, i, j) + g_ve(s, i, j + 1))) *
         (DVom(s, i, j) + DVom(s, i, j + 1) - C6 * (g_Dve(s, i, j) + g_Dve(s, i, j + 1)));
}
template <class M> struct V2D { M fomn, dndx, dmde, visc2_r, visc2_p, pmon_r, pnom_r, pmon_p, pnom_p, om_r, on_r, om_p, on_p; };
// harmonic viscosity fluxes (:1591-1625): rho-point and psi-point strain terms
template <class S, class MM> __device__ __forceinline__ double v_r(const S& s, const MM& m, int i, int j) {
  return m.visc2_r(i, j) * Drhs(s, i, j) * 0.5 *
         (m.pmon_r(i, j) * ((s.pn(i, j) + s.pn(i + 1, j)) * s.ub(i + 1, j) - (s.pn(i - 1, j) + s.pn(i, j)) * s.ub(i, j)) -
          m.pnom_r(i, j) * ((s.pm(i, j) + s.pm(i, j + 1)) * s.vb(i, j + 1) - (s.pm(i, j - 1) + s.pm(i, j)) * s.vb(i, j)));
}
template <class S, class MM> __device__ __forceinline__ double v_p(const S& s, const MM& m, int i, int j) {
  const double Dp = 0.25 * (Drhs(s, i, j) + Drhs(s, i - 1, j) + Drhs(s, i, j - 1) + Drhs(s, i - 1, j - 1));
  return m.visc2_p(i, j) * Dp * 0.5 *
         (m.pmon_p(i, j) * ((s.pn(i, j - 1) + s.pn(i, j)) * s.vb(i, j) - (s.pn(i - 1, j - 1) + s.pn(i - 1, j)) * s.vb(i - 1, j)) +
          m.pnom_p(i, j) * ((s.pm(i - 1, j) + s.pm(i, j)) * s.ub(i, j) - (s.pm(i - 1, j - 1) + s.pm(i, j - 1)) * s.ub(i, j - 1)));
}
template <class S, class MM> __device__ __forceinline__ void c_cor(const S& s, const MM& m, int i, int j, double& cu, double& cv) {
  const double cff = 0.5 * Drhs(s, i, j) * m.fomn(i, j);
  cu = cff * (s.vb(i, j) + s.vb(i, j + 1)); cv = cff * (s.ub(i, j) + s.ub(i + 1, j));
}
template <class S, class MM> __device__ __forceinline__ void c_curv(const S& s, const MM& m, int i, int j, double& cu, double& cv) {
  const double c1 = 0.5 * (s.vb(i, j) + s.vb(i, j + 1)), c2 = 0.5 * (s.ub(i, j) + s.ub(i + 1, j));
  const double c3 = c1 * m.dndx(i, j), c4 = c2 * m.dmde(i, j);
  const double cff = Drhs(s, i, j) * (c3 - c4);
  cu = cff * c1; cv = cff * c2;
}

struct Step2dArgs { int krhs, kstp, knew, nstp, nnew, iif, pred, stepmode; };  // stepmode: 0 iic==ntfirst, 1 ntfirst+1, 2 later

// Shared-memory tiles of one block, each S2_TW x S2_TH: six that change with the sub-step (Drhs, DUon, DVom and zeta, ubar, vbar of
// level krhs), and the time-invariant planes: pm, pn always; the other eighteen only in the persistent kernel.
constexpr int S2_TS = S2_TW * S2_TH;
enum { TL_Dr, TL_DU, TL_DV, TL_U, TL_V, TL_Z, TL_pm, TL_pn, TL_NDYN,
       TL_h = TL_NDYN, TL_rhoA, TL_rhoS, TL_on_u, TL_om_v, TL_fomn, TL_dndx, TL_dmde, TL_visc2_r, TL_visc2_p, TL_pmon_r, TL_pnom_r, TL_pmon_p, TL_pnom_p,
       TL_om_r, TL_on_r, TL_om_p, TL_on_p, TL_NALL };
template <bool PERS> struct Acc;
template <> struct Acc<false> {
  typedef V2 M;
  static __device__ __forceinline__ V2 get(const Dev& D, const double*, int, int, int fid, int) { return v2(D, fid); }
};
template <> struct Acc<true> {
  typedef T2 M;
  static __device__ __forceinline__ T2 get(const Dev&, const double* sm, int ti0, int tj0, int, int tl) { return T2{sm + tl * S2_TS, ti0, tj0}; }
};
#define S2M(name) Acc<PERS>::get(D, sm, ti0, tj0, FID(name), TL_##name)

// Threads (x,y) = point of the tile, z = momentum component: z=0 advances zeta, the fast-time averages and ubar,
// z=1 advances vbar (both evaluate the free-surface state they need from the shared tiles).
// One sub-step on block tile (bX, bY) of box bx.  Every value a thread needs from global memory beyond the tiles (its own point
// and the point behind it of the other time levels, the forcing, the running averages) is loaded BEFORE the tile barrier, in the
// same round of loads as the tiles: the sub-step is latency-bound and each dependent round of L2 loads costs as much as all of
// its arithmetic.  PERS: the time-invariant planes come from shared memory (filled once per fast loop by the persistent kernel).
struct Boxes { Box b[4]; };               // blockIdx.z selects the box (the frame of a tile is up to four strips)
template <bool PERS>
__device__ __forceinline__ void step2d_body(const Dev& D, const Box& bx, const Step2dArgs& a, int bX, int bY, double* sm) {
  typedef typename Acc<PERS>::M M;
  if (bx.i0 + bX * S2_TX > bx.i1 || bx.j0 + bY * S2_TY > bx.j1) return;   // whole block outside this box
  const int i = bx.i0 + bX * S2_TX + threadIdx.x, j = bx.j0 + bY * S2_TY + threadIdx.y;
  const roms_b200_bounds& b = D.b;
  const int krhs = a.krhs, kstp = a.kstp, knew = a.knew, iif = a.iif, ptsk = 3 - kstp;
  const bool PRED = a.pred != 0;
  const int ti0 = bx.i0 + bX * S2_TX - 3, tj0 = bx.j0 + bY * S2_TY - 3;
  double *tDr = sm + TL_Dr * S2_TS, *tDU = sm + TL_DU * S2_TS, *tDV = sm + TL_DV * S2_TS, *tU = sm + TL_U * S2_TS, *tV = sm + TL_V * S2_TS,
         *tZ = sm + TL_Z * S2_TS, *tPm = sm + TL_pm * S2_TS, *tPn = sm + TL_pn * S2_TS;
  S2<M> s{T2{tZ, ti0, tj0}, T2{tU, ti0, tj0}, T2{tV, ti0, tj0},
          S2M(h), T2{tPm, ti0, tj0}, T2{tPn, ti0, tj0}, S2M(on_u), S2M(om_v), S2M(rhoA), S2M(rhoS),
          1000.0 / D.p.rho0, D.p.dtfast, (iif == 1) ? 0 : (PRED ? 1 : 2),
          b.Southern_Edge && !b.NSperiodic, b.Northern_Edge && !b.NSperiodic, b.Jstr, b.Jend,
          tDr, tDU, tDV, ti0, tj0};
  const int mycomp = threadIdx.z, di = mycomp == 0 ? 1 : 0, dj = 1 - di;
  const bool act = (i <= bx.i1 && j <= bx.j1);
  const bool inner = act && iif <= D.p.nfast && (i >= b.Istr && i <= b.Iend && j >= b.Jstr && j <= b.Jend);
  const bool doC = inner && (mycomp == 0 ? (i >= b.IstrU) : (j >= b.JstrV));
  const bool avg = act && mycomp == 0;
  const bool last = (iif == D.p.nfast + 1) && PRED;    // auxiliary pass: periodic images of the averages (:821-855)
  const bool inR = avg && (i >= b.IstrR && i <= b.IendR && j >= b.JstrR && j <= b.JendR);
  const bool inU = avg && (i >= b.Istr && i <= b.IendR && j >= b.JstrR && j <= b.JendR);
  const bool inV = avg && (i >= b.IstrR && i <= b.IendR && j >= b.Jstr && j <= b.JendR);
  V2 Zt = v2(D, FID(Zt_avg1)), DU1 = v2(D, FID(DU_avg1)), DU2 = v2(D, FID(DU_avg2)), DV1 = v2(D, FID(DV_avg1)), DV2 = v2(D, FID(DV_avg2));
  V2 frc = v2(D, mycomp == 0 ? FID(rufrc) : FID(rvfrc));
  V2 qs = v2l(D, mycomp == 0 ? FID(ubar) : FID(vbar), kstp), qn = v2l(D, mycomp == 0 ? FID(ubar) : FID(vbar), knew);
  // ---- the early loads (nothing below the barriers reads global memory again except the start-up forms of the first predictor)
  double e_zs0 = 0.0, e_zsm = 0.0, e_rzs0 = 0.0, e_rzsm = 0.0, e_rzp0 = 0.0, e_rzpm = 0.0, e_frc = 0.0, e_qs = 0.0, e_rs = 0.0, e_rp = 0.0;
  double e_Zt = 0.0, e_DU1 = 0.0, e_DU2 = 0.0, e_DV1 = 0.0, e_DV2 = 0.0, e_w1 = 0.0, e_w2a = 0.0, e_w2b = 0.0;
  if (inner) {
    V2 zs = v2l(D, FID(zeta), kstp);
    e_zs0 = zs(i, j);
    if (doC) e_zsm = zs(i - di, j - dj);
    if (s.mode == 2) {
      V2 rzs = v2l(D, FID(rzeta), kstp > 2 ? 1 : kstp), rzp = v2l(D, FID(rzeta), ptsk < 1 ? 1 : ptsk);
      e_rzs0 = rzs(i, j); e_rzp0 = rzp(i, j);
      if (doC) { e_rzsm = rzs(i - di, j - dj); e_rzpm = rzp(i - di, j - dj); }
    }
    if (doC) {
      e_frc = frc(i, j); e_qs = qs(i, j);
      if (!(iif == 1 || PRED)) {
        V2 rs = v2l(D, mycomp == 0 ? FID(rubar) : FID(rvbar), kstp), rp = v2l(D, mycomp == 0 ? FID(rubar) : FID(rvbar), ptsk);
        e_rs = rs(i, j); e_rp = rp(i, j);
      }
    }
  }
  if (avg) {
    e_w2a = D.w2[iif]; e_w2b = D.w2[iif + 1];
    if (PRED && iif > 1) {
      e_w1 = D.w1[iif - 1];
      if (inR) e_Zt = Zt(i, j);
      if (inU) e_DU1 = DU1(i, j);
      if (inV) e_DV1 = DV1(i, j);
    }
    if (!(PRED && iif == 1)) {
      if (inU) e_DU2 = DU2(i, j);
      if (inV) e_DV2 = DV2(i, j);
    }
  }
  {
    // ---- shared tiles on [I0-3,I1+2]x[J0-3,J1+2]: zeta, ubar, vbar of level krhs and Drhs (one round of loads; on_u, om_v of the
    // cell ride along in registers), then DUon/DVom where Drhs(i-1)/(j-1) exist (:664-702).  A thread owns at most S2_NQ cells.
    constexpr int S2_NT = S2_TX * S2_TY * 2, S2_NQ = (S2_TS + S2_NT - 1) / S2_NT;
    const int tid = (threadIdx.z * S2_TY + threadIdx.y) * S2_TX + threadIdx.x;
    V2 zkg = v2l(D, FID(zeta), krhs), ubg = v2l(D, FID(ubar), krhs), vbg = v2l(D, FID(vbar), krhs);
    double r_ou[S2_NQ], r_ov[S2_NQ], r_ub[S2_NQ], r_vb[S2_NQ];
#pragma unroll
    for (int n = 0; n < S2_NQ; ++n) {
      const int q = tid + n * S2_NT;
      r_ou[n] = 0.0; r_ov[n] = 0.0; r_ub[n] = 0.0; r_vb[n] = 0.0;
      if (q < S2_TS) {
        const int ii = s.ti0 + q % S2_TW, jj = s.tj0 + q / S2_TW;
        const bool in = (ii >= b.LBi && ii <= b.UBi && jj >= b.LBj && jj <= b.UBj);
        double dr = 0.0, zv = 0.0;
        if constexpr (PERS) {
          if (in) { zv = zkg(ii, jj); dr = zv + s.h(ii, jj); r_ub[n] = ubg(ii, jj); r_vb[n] = vbg(ii, jj); r_ou[n] = s.on_u(ii, jj); r_ov[n] = s.om_v(ii, jj); }
        } else {
          double pmv = 0.0, pnv = 0.0;
          if (in) {
            zv = zkg(ii, jj); dr = zv + s.h(ii, jj); r_ub[n] = ubg(ii, jj); r_vb[n] = vbg(ii, jj); r_ou[n] = s.on_u(ii, jj); r_ov[n] = s.om_v(ii, jj);
            pmv = v2(D, FID(pm))(ii, jj); pnv = v2(D, FID(pn))(ii, jj);
          }
          tPm[q] = pmv; tPn[q] = pnv;
        }
        tDr[q] = dr; tZ[q] = zv; tU[q] = r_ub[n]; tV[q] = r_vb[n];
      }
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < S2_NQ; ++n) {
      const int q = tid + n * S2_NT;
      if (q < S2_TS) {
        const int qi = q % S2_TW, qj = q / S2_TW, ii = s.ti0 + qi, jj = s.tj0 + qj;
        const bool in = (ii >= b.LBi && ii <= b.UBi && jj >= b.LBj && jj <= b.UBj);
        double du = 0.0, dv = 0.0;
        if (in && qi >= 1 && ii - 1 >= b.LBi) {
          const double cff = 0.5 * r_ou[n]; const double cff1 = cff * (tDr[q] + tDr[q - 1]);
          du = r_ub[n] * cff1;
        }
        if (in && qj >= 1 && jj - 1 >= b.LBj) {
          const double cff = 0.5 * r_ov[n]; const double cff1 = cff * (tDr[q] + tDr[q - S2_TW]);
          dv = r_vb[n] * cff1;
        }
        tDU[q] = du; tDV[q] = dv;
      }
    }
    __syncthreads();
  }
  if (!act) return;
  // ---- fast-time averaging (:742-810)
  if (mycomp == 0) {
    if (PRED) {
      if (iif == 1) {
        const double cff2 = (-1.0 / 12.0) * e_w2b;
        if (inR) Zt(i, j) = 0.0;
        if (inU) { DU1(i, j) = 0.0; DU2(i, j) = cff2 * DUon(s, i, j); }
        if (inV) { DV1(i, j) = 0.0; DV2(i, j) = cff2 * DVom(s, i, j); }
      } else {
        const double cff1 = e_w1;
        const double cff2 = (8.0 / 12.0) * e_w2a - (1.0 / 12.0) * e_w2b;
        if (inR) { const double val = e_Zt + cff1 * s.zk(i, j); if (last) st(D, Zt, i, j, val); else Zt(i, j) = val; }
        if (inU) { const double du = DUon(s, i, j); const double val = e_DU1 + cff1 * du; if (last) st(D, DU1, i, j, val); else DU1(i, j) = val; DU2(i, j) = e_DU2 + cff2 * du; }
        if (inV) { const double dv = DVom(s, i, j); const double val = e_DV1 + cff1 * dv; if (last) st(D, DV1, i, j, val); else DV1(i, j) = val; DV2(i, j) = e_DV2 + cff2 * dv; }
      }
    } else {
      const double cff2 = (iif == 1) ? e_w2a : (5.0 / 12.0) * e_w2a;
      if (inU) DU2(i, j) = e_DU2 + cff2 * DUon(s, i, j);
      if (inV) DV2(i, j) = e_DV2 + cff2 * DVom(s, i, j);
    }
  }
  if (!inner) return;
  const bool doU = (i >= b.IstrU), doV = (j >= b.JstrV);
  const bool south = s.S && j == b.Jstr, north = s.N && j == b.Jend;
  // ---- free surface (:899-1072)
  const Zst z0 = zstate(s, i, j, e_zs0, e_rzs0, e_rzp0);
  if (mycomp == 0) {
    V2 zn = v2l(D, FID(zeta), knew);
    st(D, zn, i, j, z0.zn);
    if (south) st(D, zn, i, j - 1, z0.zn);           // zetabc_tile closed: zero gradient (zetabc.F:353-358,437-442)
    if (north) st(D, zn, i, j + 1, z0.zn);
    if (PRED) st(D, v2l(D, FID(rzeta), krhs), i, j, z0.rhs_zeta);
  }
  V2D<M> m{S2M(fomn), S2M(dndx), S2M(dmde), S2M(visc2_r), S2M(visc2_p), S2M(pmon_r), S2M(pnom_r),
           S2M(pmon_p), S2M(pnom_p), S2M(om_r), S2M(on_r), S2M(om_p), S2M(on_p)};
  const bool curv = (D.p.app == ROMS_B200_APP_BENCHMARK);
  const double cg = 0.5 * D.p.g, c3 = 1.0 / 3.0;
  const int comp = mycomp;
  if (comp == 0 ? doU : doV) {
    const Zst zm = zstate(s, i - di, j - dj, e_zsm, e_rzsm, e_rzpm);
    const double hm = s.h(i - di, j - dj), h0 = s.h(i, j);
    // pressure gradient with variable-density terms (:1088-1205)
    double rhs = cg * (comp == 0 ? s.on_u(i, j) : s.om_v(i, j)) *
                 ((hm + h0) * (zm.gzeta - z0.gzeta) +
                  (hm - h0) * (zm.gzetaSA + z0.gzetaSA + c3 * (s.rhoA(i - di, j - dj) - s.rhoA(i, j)) * (zm.zwrk - z0.zwrk)) +
                  (zm.gzeta2 - z0.gzeta2));
    double cu0, cv0, cu1, cv1;
    if (comp == 0) {
      { const double c1 = a_UFx(s, i, j) - a_UFx(s, i - 1, j), c2 = a_UFe(s, i, j + 1) - a_UFe(s, i, j); rhs = rhs - (c1 + c2); }
      c_cor(s, m, i, j, cu0, cv0); c_cor(s, m, i - 1, j, cu1, cv1); rhs = rhs + 0.5 * (cu0 + cu1);
      if (curv) { c_curv(s, m, i, j, cu0, cv0); c_curv(s, m, i - 1, j, cu1, cv1); rhs = rhs + 0.5 * (cu0 + cu1); }
      const double UFx0 = m.on_r(i, j) * m.on_r(i, j) * v_r(s, m, i, j), UFx1 = m.on_r(i - 1, j) * m.on_r(i - 1, j) * v_r(s, m, i - 1, j);
      const double UFe1 = m.om_p(i, j + 1) * m.om_p(i, j + 1) * v_p(s, m, i, j + 1), UFe0 = m.om_p(i, j) * m.om_p(i, j) * v_p(s, m, i, j);
      const double c1 = 0.5 * (s.pn(i - 1, j) + s.pn(i, j)) * (UFx0 - UFx1), c2 = 0.5 * (s.pm(i - 1, j) + s.pm(i, j)) * (UFe1 - UFe0);
      rhs = rhs + (c1 + c2);
    } else {
      { const double c1 = a_VFx(s, i + 1, j) - a_VFx(s, i, j), c2 = a_VFe(s, i, j) - a_VFe(s, i, j - 1); rhs = rhs - (c1 + c2); }
      c_cor(s, m, i, j, cu0, cv0); c_cor(s, m, i, j - 1, cu1, cv1); rhs = rhs - 0.5 * (cv0 + cv1);
      if (curv) { c_curv(s, m, i, j, cu0, cv0); c_curv(s, m, i, j - 1, cu1, cv1); rhs = rhs - 0.5 * (cv0 + cv1); }
      const double VFx1 = m.on_p(i + 1, j) * m.on_p(i + 1, j) * v_p(s, m, i + 1, j), VFx0 = m.on_p(i, j) * m.on_p(i, j) * v_p(s, m, i, j);
      const double VFe0 = m.om_r(i, j) * m.om_r(i, j) * v_r(s, m, i, j), VFe1 = m.om_r(i, j - 1) * m.om_r(i, j - 1) * v_r(s, m, i, j - 1);
      const double c1 = 0.5 * (s.pn(i, j - 1) + s.pn(i, j)) * (VFx1 - VFx0), c2 = 0.5 * (s.pm(i, j - 1) + s.pm(i, j)) * (VFe0 - VFe1);
      rhs = rhs + (c1 - c2);
    }
    // coupling with the 3-D forcing (:2241-2459)
    if (iif == 1 && PRED) {
      V3 r0n = v3l(D, comp == 0 ? FID(ru) : FID(rv), a.nstp), r0w = v3l(D, comp == 0 ? FID(ru) : FID(rv), a.nnew);
      const double fr = e_frc - rhs;
      frc(i, j) = fr;
      if (a.stepmode == 0) rhs = rhs + fr;
      else if (a.stepmode == 1) rhs = rhs + 1.5 * fr - 0.5 * r0w(i, j, 0);
      else rhs = rhs + (23.0 / 12.0) * fr - (16.0 / 12.0) * r0w(i, j, 0) + (5.0 / 12.0) * r0n(i, j, 0);
      r0n(i, j, 0) = fr;
    } else rhs = rhs + e_frc;
    // time stepping (:2493-2674)
    const double Dstp = (e_zs0 + h0) + (e_zsm + hm);
    const double cff = (s.pm(i, j) + s.pm(i - di, j - dj)) * (s.pn(i, j) + s.pn(i - di, j - dj));
    const double fc = 1.0 / (z0.Dnew + zm.Dnew);
    double val;
    if (iif == 1 || PRED) {
      const double cff1 = (iif == 1) ? 0.5 * D.p.dtfast : D.p.dtfast;
      val = (e_qs * Dstp + cff * cff1 * rhs) * fc;
    } else {
      const double cff1 = 0.5 * D.p.dtfast * 5.0 / 12.0, cff2 = 0.5 * D.p.dtfast * 8.0 / 12.0, cff3 = 0.5 * D.p.dtfast * 1.0 / 12.0;
      val = (e_qs * Dstp + cff * (cff1 * rhs + cff2 * e_rs - cff3 * e_rp)) * fc;
    }
    if (PRED) v2l(D, comp == 0 ? FID(rubar) : FID(rvbar), krhs)(i, j) = rhs;
    st(D, qn, i, j, val);
    if (comp == 0) {                                  // u2dbc_im.F:483-496 (gamma2 slip on closed walls)
      if (south) st(D, qn, i, j - 1, D.p.gamma2 * val);
      if (north) st(D, qn, i, j + 1, D.p.gamma2 * val);
    }
  }
  if (mycomp == 1) {                                  // v2dbc_im.F:253-258,395-400 (no normal flow)
    V2 vn = v2l(D, FID(vbar), knew);
    if (south) st(D, vn, i, b.Jstr, 0.0);
    if (north) st(D, vn, i, b.Jend + 1, 0.0);
  }
}
template <int MINB>
__global__ void __launch_bounds__(S2_TX * S2_TY * 2, MINB) step2d_kernel(const Dev D, const Boxes bxs, Step2dArgs a) {
  __shared__ double sm[TL_NDYN * S2_TS];
  // Programmatic dependent launch (launch_boxes): let the next sub-step's grid be scheduled now, and do not touch any
  // field before the previous sub-step has completed and flushed.  Both are no-ops for an ordinary launch.
#ifndef ROMS_B200_EMU
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
  step2d_body<false>(D, bxs.b[blockIdx.z], a, (int)blockIdx.x, (int)blockIdx.y, sm);
}

// ---- the whole fast loop of a baroclinic step as ONE persistent kernel (single-tile runs) -----------------------------------
// Every block owns one 32 x 4 tile for all 2*nfast+1 sub-steps.  A sub-step reads, besides its own points, only points within
// three cells of its tile (plus, on a periodic axis, the images of the opposite edge), so instead of a kernel boundary -- or a
// grid-wide barrier -- between two sub-steps a block waits for the blocks of the ADJACENT tiles: each block publishes the
// number of sub-steps it has completed (its stores, then __syncthreads, then a release store of the counter by one thread);
// before sub-step p a block polls the counters of its neighbours with acquire loads until all have completed p-1.  That is
// sufficient for every hazard of the three rotating time levels: a block starts p only after its neighbours finished p-1
// (reads of their results; their reads of the level this block overwrites in p were in p-1 or earlier), and no neighbour
// starts p+1 before this block finished p.  Points are advanced by step2d_body, the code of the per-sub-step kernel: same bits.
// The counters run on from launch to launch (`base`), so they are never reset; `abort` (the word behind the counters) is set
// by a block whose wait times out (device error bit 16) and makes every other block leave its wait at once.
constexpr int S2_MAXPH = 250;
struct PersistArgs {
  Box bx; int ntx, nty, nphase, xr, wrap; unsigned base; unsigned* done;   // xr: reach of the neighbourhood in x (tiles)
  int nstp, nnew, stepmode;
  long long* prof;                        // ROMS_B200_S2_PROF=1: clock64 stamps of one block, 4 per sub-step (before wait, after wait, after body, after release)
  unsigned ph[S2_MAXPH];                  // per sub-step: krhs | kstp << 2 | knew << 4 | pred << 6 | iif << 8
};
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
#ifdef ROMS_B200_EMU
  return __atomic_load_n(p, __ATOMIC_ACQUIRE);
#else
  unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
#endif
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
#ifdef ROMS_B200_EMU
  __atomic_store_n(p, v, __ATOMIC_RELEASE);
#else
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}
__global__ void __launch_bounds__(S2_TX * S2_TY * 2, 2) step2d_persist_kernel(const __grid_constant__ Dev D, const __grid_constant__ PersistArgs P) {
  extern __shared__ double s2sm[];
  const int blk = blockIdx.x, bX = blk % P.ntx, bY = blk / P.ntx;
  const int tid = (threadIdx.z * S2_TY + threadIdx.y) * S2_TX + threadIdx.x;
  {
    // the time-invariant planes of this block's tile (+ halo), once: pm, pn, h, rhoA, rhoS (constant during the fast loop) and the metrics
    const int ti0 = P.bx.i0 + bX * S2_TX - 3, tj0 = P.bx.j0 + bY * S2_TY - 3;
    const roms_b200_bounds& b = D.b;
    for (int q = tid; q < S2_TS; q += S2_TX * S2_TY * 2) {
      const int ii = ti0 + q % S2_TW, jj = tj0 + q / S2_TW;
      const bool in = (ii >= b.LBi && ii <= b.UBi && jj >= b.LBj && jj <= b.UBj);
#define S2FILL(name) s2sm[TL_##name * S2_TS + q] = in ? v2(D, FID(name))(ii, jj) : 0.0
      S2FILL(pm); S2FILL(pn); S2FILL(h); S2FILL(rhoA); S2FILL(rhoS); S2FILL(on_u); S2FILL(om_v); S2FILL(fomn); S2FILL(dndx); S2FILL(dmde);
      S2FILL(visc2_r); S2FILL(visc2_p); S2FILL(pmon_r); S2FILL(pnom_r); S2FILL(pmon_p); S2FILL(pnom_p); S2FILL(om_r); S2FILL(on_r); S2FILL(om_p); S2FILL(on_p);
#undef S2FILL
    }
    __syncthreads();
  }
  // the neighbour this thread polls (threads 0 .. 3*(2*xr+1)-1), -1: none
  int nb = -1;
  {
    const int span = 2 * P.xr + 1;
    if (tid < 3 * span) {
      int dx = tid % span - P.xr, dy = tid / span - 1, x = bX + dx, y = bY + dy;
      if (P.wrap) { if (x < 0) x += P.ntx; else if (x >= P.ntx) x -= P.ntx; }
      // on a wrapped axis with few tiles the same neighbour appears more than once: harmless
      if (x >= 0 && x < P.ntx && y >= 0 && y < P.nty && !(x == bX && y == bY)) nb = y * P.ntx + x;
    }
  }
  unsigned* abort_w = P.done + P.ntx * P.nty;
  const bool prof = P.prof && blk == (P.nty / 2) * P.ntx + P.ntx / 2 && tid == 0;
  for (int p = 0; p < P.nphase; ++p) {
    if (prof) P.prof[4 * p] = clock64();
    if (p > 0) {
      if (nb >= 0) {
        const unsigned want = P.base + (unsigned)p;
        long long t0 = 0;
        while ((int)(ld_acquire(P.done + nb) - want) < 0) {
#ifdef ROMS_B200_EMU
          static const bool nowait = getenv("EMU_S2P_NOWAIT") != nullptr;    // negative control of the emulation tests
          if (nowait) break;
          emu::yield();
#else
          if (*(volatile unsigned*)abort_w) break;
          const long long t = clock64();
          if (!t0) t0 = t; else if (t - t0 > 4000000000ll) { atomicOr(D.err, 16); *(volatile unsigned*)abort_w = 1u; __threadfence(); break; }
#endif
        }
      }
      __syncthreads();
    }
    if (prof) P.prof[4 * p + 1] = clock64();
    const unsigned w = P.ph[p];
    const Step2dArgs a{(int)(w & 3u), (int)((w >> 2) & 3u), (int)((w >> 4) & 3u), P.nstp, P.nnew, (int)(w >> 8), (int)((w >> 6) & 1u), P.stepmode};
    step2d_body<true>(D, P.bx, a, bX, bY, s2sm);
    __syncthreads();                                  // every store of this block for sub-step p has been issued
    if (prof) P.prof[4 * p + 2] = clock64();
    if (tid == 0) st_release(P.done + blk, P.base + (unsigned)p + 1u);
    if (prof) P.prof[4 * p + 3] = clock64();
  }
}

static inline void launch_boxes(roms_b200_ctx* c, cudaStream_t st, const Box* bx, int n, const Step2dArgs& a, const Dev* Duse = nullptr) {
  Boxes B{}; int gx = 1, gy = 1;
  for (int q = 0; q < n; ++q) {
    B.b[q] = bx[q];
    gx = std::max(gx, (bx[q].i1 - bx[q].i0 + S2_TX) / S2_TX); gy = std::max(gy, (bx[q].j1 - bx[q].j0 + S2_TY) / S2_TY);
  }
  // Consecutive sub-steps are chained by programmatic dependent launch, which overlaps the launch latency and block scheduling
  // of sub-step n+1 with the tail of sub-step n (the kernel waits in griddepcontrol.wait before its first read): 9.36 -> 9.07 us
  // per sub-step on BENCHMARK1, bit-identical.  With neighbours the halo kernel between two sub-steps is part of the chain
  // (k_halo.cu): 2 GPUs 1.361 -> 1.345 ms per step with the sub-steps alone.  ROMS_B200_NO_PDL=1 switches it off.
  static const bool pdl_off = (getenv("ROMS_B200_NO_PDL") != nullptr);
  const bool pdl = !pdl_off;
  // resident blocks per SM the kernel is compiled for: 2 (128 registers, no spills) is fastest while the grid is a wave or two
  // (BENCHMARK1: 6.97 us per sub-step against 8.84 with 3); on grids of many waves the third block's occupancy wins although
  // it costs 40 bytes of spills (2048x256: 97.8 against 107.2 us).  ROMS_B200_S2_MINB=2|3 forces one.
  static const int minb_env = getenv("ROMS_B200_S2_MINB") ? atoi(getenv("ROMS_B200_S2_MINB")) : 0;
  const int minb = minb_env ? minb_env : ((long)gx * gy * n >= 6 * 148 ? 3 : 2);
  auto kern = (minb == 3) ? step2d_kernel<3> : step2d_kernel<2>;
  if (pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(gx, gy, n); cfg.blockDim = dim3(S2_TX, S2_TY, 2); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, kern, Duse ? *Duse : c->D, B, a) != cudaSuccess) fprintf(stderr, "roms_b200: programmatic launch of step2d failed\n");
  } else if (minb == 3) step2d_kernel<3><<<dim3(gx, gy, n), dim3(S2_TX, S2_TY, 2), 0, st>>>(Duse ? *Duse : c->D, B, a);
  else step2d_kernel<2><<<dim3(gx, gy, n), dim3(S2_TX, S2_TY, 2), 0, st>>>(Duse ? *Duse : c->D, B, a);
  c->launches++;
}
// With neighbours, the sub-step is split so that the halo exchange of the frame overlaps the interior stencil: the frame
// (the strips of width `halo` the neighbours need) runs on the launch stream and is followed by the exchange; the interior
// runs on a second stream.  Every point is advanced by one thread from levels it only reads, so the split changes no bits.
int k_step2d(roms_b200_ctx* c, int krhs, int kstp, int knew, int nstp, int nnew, int iif, int pred, int iic, int ntfirst) {
  const roms_b200_bounds& b = c->D.b;
  Step2dArgs a{krhs, kstp, knew, nstp, nnew, iif, pred, (iic == ntfirst) ? 0 : (iic == ntfirst + 1 ? 1 : 2)};
  const Box full{b.IstrR, b.IendR, b.JstrR, b.JendR};
  static const bool overlap = (getenv("ROMS_B200_OVERLAP") != nullptr);   // measured slower (2-stream graph nodes cost more than the exchange they hide): opt-in
  const bool hW = c->comm && c->nbW >= 0, hE = c->comm && c->nbE >= 0, hS = c->comm && c->nbS >= 0, hN = c->comm && c->nbN >= 0;
  const int w = c->D.halo;
  c->forked = 0;
  if (c->deep && pred && (hW || hE || hS || hN)) {
    // Deep-halo predictor: every point is advanced from levels it only reads, so the predictor can also be evaluated on the
    // 3 halo points next to each neighbour (the halo is >= 6 wide and was refreshed after the last corrector): the corrector
    // then finds zeta/ubar/vbar(3), rzeta, rubar of its whole stencil locally and the swap after the predictor disappears
    // (59 -> 30 swaps per fast loop; same operations on the same operands on both sides of a tile edge -> same bits).
    const int e = 3;
    Dev De = c->D; roms_b200_bounds& q = De.b;
    if (hW) { q.Istr -= e; q.IstrU -= e; q.IstrR -= e; }
    if (hE) { q.Iend += e; q.IendR += e; }
    if (hS) { q.Jstr -= e; q.JstrV -= e; q.JstrR -= e; }
    if (hN) { q.Jend += e; q.JendR += e; }
    const Box ext{q.IstrR, q.IendR, q.JstrR, q.JendR};
    launch_boxes(c, c->stream, &ext, 1, a, &De);
    return 0;
  }
  if (!overlap || !(hW || hE || hS || hN) || (b.Iend - b.Istr + 1) <= 2 * w + 1 || (b.Jend - b.Jstr + 1) <= 2 * w + 1) {
    launch_boxes(c, c->stream, &full, 1, a);
    return 0;
  }
  const int xa = hW ? b.Istr + w - 1 : full.i0 - 1, xb = hE ? b.Iend - w + 1 : full.i1 + 1;
  const int ya = hS ? b.Jstr + w - 1 : full.j0 - 1, yb = hN ? b.Jend - w + 1 : full.j1 + 1;
  Box fr[4]; int nf = 0;
  if (hW) fr[nf++] = Box{full.i0, xa, full.j0, full.j1};
  if (hE) fr[nf++] = Box{xb, full.i1, full.j0, full.j1};
  if (hS) fr[nf++] = Box{xa + 1, xb - 1, full.j0, ya};
  if (hN) fr[nf++] = Box{xa + 1, xb - 1, yb, full.j1};
  const Box inner{xa + 1, xb - 1, ya + 1, yb - 1};
  CUDA_OK(cudaEventRecord(c->ev_fork, c->stream));            // everything before this sub-step is complete
  CUDA_OK(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
  launch_boxes(c, c->stream2, &inner, 1, a);
  CUDA_OK(cudaEventRecord(c->ev_join, c->stream2));
  launch_boxes(c, c->stream, fr, nf, a);
  c->forked = 1;
  return 0;
}
int k_step2d_join(roms_b200_ctx* c) {
  if (c->forked) { CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev_join, 0)); c->forked = 0; }
  return 0;
}

// Host side of the persistent fast loop.  rc 0: launched; 2: declined (the caller runs the per-sub-step path): not asked for,
// tiles with neighbours (the halo exchange sits between two sub-steps), more block tiles than can be resident at once.
// MEASURED on BENCHMARK1 (512x64, 289 blocks on 148 SMs, B200): 7.5-9.1 us per sub-step against 6.97 us for the graph of
// per-sub-step launches chained by programmatic dependent launch.  Clock stamps of the middle block (ROMS_B200_S2_PROF=1):
// 3.4 us in step2d_body (about 350 fp64 instructions per thread with -fmad=false: 2 resident blocks keep the SM's fp64 pipe
// busy for ~1.4 us of it), 0.7 us in the release (MEMBAR.ALL.GPU), 3.3 us waiting for the slowest neighbour -- the hand-shake
// costs what the kernel boundary costs.  Hence OPT-IN (ROMS_B200_S2_PERSIST=1); the emulation tests run it (bit-identical).
int k_step2d_persist(roms_b200_ctx* c, const int* ph, int nphase, int nstp, int nnew, int iic, int ntfirst) {
  static const char* env = getenv("ROMS_B200_S2_PERSIST");
  if (c->comm || !env || atoi(env) == 0 || nphase > S2_MAXPH) return 2;
  const roms_b200_bounds& b = c->D.b;
  PersistArgs P{};
  P.bx = Box{b.IstrR, b.IendR, b.JstrR, b.JendR};
  P.ntx = (P.bx.i1 - P.bx.i0 + S2_TX) / S2_TX; P.nty = (P.bx.j1 - P.bx.j0 + S2_TY) / S2_TY;
  const int ntiles = P.ntx * P.nty;
  const size_t smem = TL_NALL * S2_TS * sizeof(double);
  if (c->s2p_cap < 0) return 2;
  if (c->s2p_cap == 0) {
#ifdef ROMS_B200_EMU
    if (!emu::coop_supported()) { c->s2p_cap = -1; return 2; }
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    c->s2p_cap = 2 * sms;
#else
    int per_sm = 0, sms = 0, dev = 0, coop = 0;
    CUDA_OK(cudaGetDevice(&dev));
    CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    CUDA_OK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    CUDA_OK(cudaFuncSetAttribute(step2d_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, step2d_persist_kernel, S2_TX * S2_TY * 2, smem));
    c->s2p_cap = coop ? per_sm * sms : -1;
    if (c->s2p_cap <= 0) { c->s2p_cap = -1; return 2; }
#endif
  }
  if (ntiles > c->s2p_cap) return 2;
  if (!c->s2p_done) {
    CUDA_OK(cudaMalloc((void**)&c->s2p_done, (ntiles + 1) * sizeof(unsigned)));
    CUDA_OK(cudaMemset(c->s2p_done, 0, (ntiles + 1) * sizeof(unsigned)));
  }
  // reach of the neighbourhood in x: a sub-step reads 3 cells beyond its tile; the last tile of the row may be narrower than that
  // only when the axis wraps (its far side then belongs to the tile before it)
  const int lastw = (P.bx.i1 - P.bx.i0 + 1) - (P.ntx - 1) * S2_TX;
  P.wrap = c->D.wrapEW; P.xr = (P.wrap && lastw < 3) ? 2 : 1;
  P.nphase = nphase; P.base = c->s2p_base; P.done = c->s2p_done;
  P.nstp = nstp; P.nnew = nnew; P.stepmode = (iic == ntfirst) ? 0 : (iic == ntfirst + 1 ? 1 : 2);
  for (int q = 0; q < nphase; ++q) P.ph[q] = (unsigned)ph[q];
  c->s2p_base += (unsigned)nphase;
#ifndef ROMS_B200_EMU
  static const bool prof = getenv("ROMS_B200_S2_PROF") != nullptr;
  static long long* profbuf = nullptr;
  if (prof) { if (!profbuf) CUDA_OK(cudaMalloc((void**)&profbuf, 4 * S2_MAXPH * sizeof(long long))); P.prof = profbuf; }
#endif
#ifdef ROMS_B200_EMU
  emu::run_grid_coop(dim3(ntiles), dim3(S2_TX, S2_TY, 2), smem, "step2d_persist_kernel", [&]() { step2d_persist_kernel(c->D, P); });
#else
  void* args[2] = {(void*)&c->D, (void*)&P};
  CUDA_OK(cudaLaunchCooperativeKernel((const void*)step2d_persist_kernel, dim3(ntiles), dim3(S2_TX, S2_TY, 2), args, smem, c->stream));
#endif
  c->launches++;
#ifndef ROMS_B200_EMU
  if (prof) {
    static int calls = 0;
    if (++calls == 20) {
      std::vector<long long> h(4 * nphase);
      CUDA_OK(cudaStreamSynchronize(c->stream));
      CUDA_OK(cudaMemcpy(h.data(), profbuf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      double w = 0, bd = 0, rl = 0;
      for (int q = 1; q < nphase; ++q) { w += h[4 * q + 1] - h[4 * q]; bd += h[4 * q + 2] - h[4 * q + 1]; rl += h[4 * q + 3] - h[4 * q + 2]; }
      fprintf(stderr, "step2d persistent, cycles per sub-step of the middle block: wait %.0f body %.0f release %.0f total %.0f\n", w / (nphase - 1), bd / (nphase - 1),
              rl / (nphase - 1), (double)(h[4 * (nphase - 1) + 3] - h[0]) / nphase);
    }
  }
#endif
  return 0;
}
