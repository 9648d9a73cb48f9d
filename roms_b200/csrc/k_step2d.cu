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
struct S2 {
  V2 zk, zs; T2 ub, vb;                   // zeta(:,:,krhs), zeta(:,:,kstp) ; ubar/vbar(:,:,krhs) tiles
  V2 h; T2 pm, pn; V2 on_u, om_v, rhoA, rhoS, rzs, rzp;   // pm, pn tiles ; rzeta(:,:,kstp), rzeta(:,:,ptsk)
  double fac, dtfast; int mode;           // mode: 0 iif==1, 1 predictor, 2 corrector
  int S, N, Jstr, Jend;
  const double *tDr, *tDU, *tDV;          // shared-memory tiles, element (i,j) at (i-ti0) + S2_TW*(j-tj0)
  int ti0, tj0;
};
__device__ __forceinline__ double Drhs(const S2& s, int i, int j) { return s.tDr[(i - s.ti0) + S2_TW * (j - s.tj0)]; }
__device__ __forceinline__ double DUon(const S2& s, int i, int j) { return s.tDU[(i - s.ti0) + S2_TW * (j - s.tj0)]; }
__device__ __forceinline__ double DVom(const S2& s, int i, int j) { return s.tDV[(i - s.ti0) + S2_TW * (j - s.tj0)]; }
struct Zst { double rhs_zeta, zn, Dnew, zwrk, gzeta, gzeta2, gzetaSA; };
// step2d_LF_AM3.h:899-980
__device__ __forceinline__ Zst zstate(const S2& s, int i, int j) {
  Zst z;
  const double div = (DUon(s, i, j) - DUon(s, i + 1, j)) + (DVom(s, i, j) - DVom(s, i, j + 1));
  const double pmn = s.pm(i, j) * s.pn(i, j);
  z.rhs_zeta = div;
  if (s.mode == 0) {
    z.zn = s.zs(i, j) + pmn * s.dtfast * div;
    z.zwrk = 0.5 * (s.zs(i, j) + z.zn);
  } else if (s.mode == 1) {
    const double cff1 = 2.0 * s.dtfast, cff4 = 4.0 / 25.0, cff5 = 1.0 - 2.0 * cff4;
    z.zn = s.zs(i, j) + pmn * cff1 * div;
    z.zwrk = cff5 * s.zk(i, j) + cff4 * (s.zs(i, j) + z.zn);
  } else {
    const double cff1 = s.dtfast * 5.0 / 12.0, cff2 = s.dtfast * 8.0 / 12.0, cff3 = s.dtfast * 1.0 / 12.0, cff4 = 2.0 / 5.0, cff5 = 1.0 - cff4;
    const double cff = cff1 * div;
    z.zn = s.zs(i, j) + pmn * (cff + cff2 * s.rzs(i, j) - cff3 * s.rzp(i, j));
    z.zwrk = cff5 * z.zn + cff4 * s.zk(i, j);
  }
  z.Dnew = z.zn + s.h(i, j);
  z.gzeta = (s.fac + s.rhoS(i, j)) * z.zwrk;
  z.gzeta2 = z.gzeta * z.zwrk;
  z.gzetaSA = z.zwrk * (s.rhoS(i, j) - s.rhoA(i, j));
  return z;
}
// second differences with the closed-wall replacements (:1283-1296, 1361-1376)
__device__ __forceinline__ double g_ux(const S2& s, int i, int j) { return s.ub(i - 1, j) - 2.0 * s.ub(i, j) + s.ub(i + 1, j); }
__device__ __forceinline__ double g_Dux(const S2& s, int i, int j) { return DUon(s, i - 1, j) - 2.0 * DUon(s, i, j) + DUon(s, i + 1, j); }
__device__ __forceinline__ double g_ue(const S2& s, int i, int j) {
  int jj = j; if (s.S && j == s.Jstr - 1) jj = s.Jstr; if (s.N && j == s.Jend + 1) jj = s.Jend;
  return s.ub(i, jj - 1) - 2.0 * s.ub(i, jj) + s.ub(i, jj + 1);
}
__device__ __forceinline__ double g_Dvx(const S2& s, int i, int j) { return DVom(s, i - 1, j) - 2.0 * DVom(s, i, j) + DVom(s, i + 1, j); }
__device__ __forceinline__ double g_vx(const S2& s, int i, int j) { return s.vb(i - 1, j) - 2.0 * s.vb(i, j) + s.vb(i + 1, j); }
__device__ __forceinline__ double g_Due(const S2& s, int i, int j) { return DUon(s, i, j - 1) - 2.0 * DUon(s, i, j) + DUon(s, i, j + 1); }
__device__ __forceinline__ double g_ve(const S2& s, int i, int j) {
  int jj = j; if (s.S && j == s.Jstr) jj = s.Jstr + 1; if (s.N && j == s.Jend + 1) jj = s.Jend;
  return s.vb(i, jj - 1) - 2.0 * s.vb(i, jj) + s.vb(i, jj + 1);
}
__device__ __forceinline__ double g_Dve(const S2& s, int i, int j) {
  int jj = j; if (s.S && j == s.Jstr) jj = s.Jstr + 1; if (s.N && j == s.Jend + 1) jj = s.Jend;
  return DVom(s, i, jj - 1) - 2.0 * DVom(s, i, jj) + DVom(s, i, jj + 1);
}
#define C6 (1.0 / 6.0)
// fourth-order centred advective fluxes (:1298-1393)
__device__ __forceinline__ double a_UFx(const S2& s, int i, int j) {
  return 0.25 * (s.ub(i, j) + s.ub(i + 1, j) - C6 * (g_ux(s, i, j) + g_ux(s, i + 1, j))) *
         (DUon(s, i, j) + DUon(s, i + 1, j) - C6 * (g_Dux(s, i, j) + g_Dux(s, i + 1, j)));
}
__device__ __forceinline__ double a_UFe(const S2& s, int i, int j) {
  return 0.25 * (s.ub(i, j) + s.ub(i, j - 1) - C6 * (g_ue(s, i, j) + g_ue(s, i, j - 1))) *
         (DVom(s, i, j) + DVom(s, i - 1, j) - C6 * (g_Dvx(s, i, j) + g_Dvx(s, i - 1, j)));
}
__device__ __forceinline__ double a_VFx(const S2& s, int i, int j) {
  return 0.25 * (s.vb(i, j) + s.vb(i - 1, j) - C6 * (g_vx(s, i, j) + g_vx(s, i - 1, j))) *
         (DUon(s, i, j) + DUon(s, i, j - 1) - C6 * (g_Due(s, i, j) + g_Due(s, i, j - 1)));
}
__device__ __forceinline__ double a_VFe(const S2& s, int i, int j) {
  return 0.25 * (s.vb(i, j) + s.vb(i, j + 1) - C6 * (g_ve(s, i, j) + g_ve(s, i, j + 1))) *
         (DVom(s, i, j) + DVom(s, i, j + 1) - C6 * (g_Dve(s, i, j) + g_Dve(s, i, j + 1)));
}
struct V2D { V2 fomn, dndx, dmde, visc2_r, visc2_p, pmon_r, pnom_r, pmon_p, pnom_p, om_r, on_r, om_p, on_p; };
// harmonic viscosity fluxes (:1591-1625): rho-point and psi-point strain terms
__device__ __forceinline__ double v_r(const S2& s, const V2D& m, int i, int j) {
  return m.visc2_r(i, j) * Drhs(s, i, j) * 0.5 *
         (m.pmon_r(i, j) * ((s.pn(i, j) + s.pn(i + 1, j)) * s.ub(i + 1, j) - (s.pn(i - 1, j) + s.pn(i, j)) * s.ub(i, j)) -
          m.pnom_r(i, j) * ((s.pm(i, j) + s.pm(i, j + 1)) * s.vb(i, j + 1) - (s.pm(i, j - 1) + s.pm(i, j)) * s.vb(i, j)));
}
__device__ __forceinline__ double v_p(const S2& s, const V2D& m, int i, int j) {
  const double Dp = 0.25 * (Drhs(s, i, j) + Drhs(s, i - 1, j) + Drhs(s, i, j - 1) + Drhs(s, i - 1, j - 1));
  return m.visc2_p(i, j) * Dp * 0.5 *
         (m.pmon_p(i, j) * ((s.pn(i, j - 1) + s.pn(i, j)) * s.vb(i, j) - (s.pn(i - 1, j - 1) + s.pn(i - 1, j)) * s.vb(i - 1, j)) +
          m.pnom_p(i, j) * ((s.pm(i - 1, j) + s.pm(i, j)) * s.ub(i, j) - (s.pm(i - 1, j - 1) + s.pm(i, j - 1)) * s.ub(i, j - 1)));
}
__device__ __forceinline__ void c_cor(const S2& s, const V2D& m, int i, int j, double& cu, double& cv) {
  const double cff = 0.5 * Drhs(s, i, j) * m.fomn(i, j);
  cu = cff * (s.vb(i, j) + s.vb(i, j + 1)); cv = cff * (s.ub(i, j) + s.ub(i + 1, j));
}
__device__ __forceinline__ void c_curv(const S2& s, const V2D& m, int i, int j, double& cu, double& cv) {
  const double c1 = 0.5 * (s.vb(i, j) + s.vb(i, j + 1)), c2 = 0.5 * (s.ub(i, j) + s.ub(i + 1, j));
  const double c3 = c1 * m.dndx(i, j), c4 = c2 * m.dmde(i, j);
  const double cff = Drhs(s, i, j) * (c3 - c4);
  cu = cff * c1; cv = cff * c2;
}

struct Step2dArgs { int krhs, kstp, knew, nstp, nnew, iif, pred, stepmode; };  // stepmode: 0 iic==ntfirst, 1 ntfirst+1, 2 later

// Threads (x,y) = point of the tile, z = momentum component: z=0 advances zeta, the fast-time averages and ubar,
// z=1 advances vbar (both evaluate the free-surface state they need from the shared tiles).
struct Boxes { Box b[4]; };               // blockIdx.z selects the box (the frame of a tile is up to four strips)
__global__ void __launch_bounds__(S2_TX * S2_TY * 2) step2d_kernel(const Dev D, const Boxes bxs, Step2dArgs a) {
  __shared__ double tDr[S2_TW * S2_TH], tDU[S2_TW * S2_TH], tDV[S2_TW * S2_TH], tU[S2_TW * S2_TH], tV[S2_TW * S2_TH], tPm[S2_TW * S2_TH], tPn[S2_TW * S2_TH];
  // Programmatic dependent launch (launch_boxes): let the next sub-step's grid be scheduled now, and do not touch any
  // field before the previous sub-step has completed and flushed.  Both are no-ops for an ordinary launch.
#ifndef ROMS_B200_EMU
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
  const Box bx = bxs.b[blockIdx.z];
  if (bx.i0 + (int)blockIdx.x * S2_TX > bx.i1 || bx.j0 + (int)blockIdx.y * S2_TY > bx.j1) return;   // whole block outside this box
  const int i = bx.i0 + blockIdx.x * S2_TX + threadIdx.x, j = bx.j0 + blockIdx.y * S2_TY + threadIdx.y;
  const roms_b200_bounds& b = D.b;
  const int krhs = a.krhs, kstp = a.kstp, knew = a.knew, iif = a.iif, ptsk = 3 - kstp;
  const bool PRED = a.pred != 0;
  const int ti0 = bx.i0 + (int)blockIdx.x * S2_TX - 3, tj0 = bx.j0 + (int)blockIdx.y * S2_TY - 3;
  S2 s{v2l(D, FID(zeta), krhs), v2l(D, FID(zeta), kstp), T2{tU, ti0, tj0}, T2{tV, ti0, tj0},
       v2(D, FID(h)), T2{tPm, ti0, tj0}, T2{tPn, ti0, tj0}, v2(D, FID(on_u)), v2(D, FID(om_v)), v2(D, FID(rhoA)), v2(D, FID(rhoS)),
       v2l(D, FID(rzeta), kstp > 2 ? 1 : kstp), v2l(D, FID(rzeta), ptsk < 1 ? 1 : ptsk),
       1000.0 / D.p.rho0, D.p.dtfast, (iif == 1) ? 0 : (PRED ? 1 : 2),
       b.Southern_Edge && !b.NSperiodic, b.Northern_Edge && !b.NSperiodic, b.Jstr, b.Jend,
       tDr, tDU, tDV, ti0, tj0};
  {
    // ---- shared tiles on [I0-3,I1+2]x[J0-3,J1+2]: Drhs, ubar, vbar (one round of loads; on_u, om_v of the cell ride along in
    // registers), then DUon/DVom where Drhs(i-1)/(j-1) exist (:664-702).  A thread owns at most S2_NQ cells of the tile.
    constexpr int S2_NT = S2_TX * S2_TY * 2, S2_NQ = (S2_TW * S2_TH + S2_NT - 1) / S2_NT;
    const int tid = (threadIdx.z * S2_TY + threadIdx.y) * S2_TX + threadIdx.x;
    V2 ubg = v2l(D, FID(ubar), krhs), vbg = v2l(D, FID(vbar), krhs), pmg = v2(D, FID(pm)), png = v2(D, FID(pn));
    double r_ou[S2_NQ], r_ov[S2_NQ], r_ub[S2_NQ], r_vb[S2_NQ];
#pragma unroll
    for (int n = 0; n < S2_NQ; ++n) {
      const int q = tid + n * S2_NT;
      r_ou[n] = 0.0; r_ov[n] = 0.0; r_ub[n] = 0.0; r_vb[n] = 0.0;
      if (q < S2_TW * S2_TH) {
        const int ii = s.ti0 + q % S2_TW, jj = s.tj0 + q / S2_TW;
        const bool in = (ii >= b.LBi && ii <= b.UBi && jj >= b.LBj && jj <= b.UBj);
        double dr = 0.0, pmv = 0.0, pnv = 0.0;
        if (in) {
          dr = s.zk(ii, jj) + s.h(ii, jj); r_ub[n] = ubg(ii, jj); r_vb[n] = vbg(ii, jj); r_ou[n] = s.on_u(ii, jj); r_ov[n] = s.om_v(ii, jj);
          pmv = pmg(ii, jj); pnv = png(ii, jj);
        }
        tDr[q] = dr; tU[q] = r_ub[n]; tV[q] = r_vb[n]; tPm[q] = pmv; tPn[q] = pnv;
      }
    }
    __syncthreads();
#pragma unroll
    for (int n = 0; n < S2_NQ; ++n) {
      const int q = tid + n * S2_NT;
      if (q < S2_TW * S2_TH) {
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
  if (i > bx.i1 || j > bx.j1) return;
  const int mycomp = threadIdx.z;
  // ---- fast-time averaging (:742-810)
  if (mycomp == 0) {
    V2 Zt = v2(D, FID(Zt_avg1)), DU1 = v2(D, FID(DU_avg1)), DU2 = v2(D, FID(DU_avg2)), DV1 = v2(D, FID(DV_avg1)), DV2 = v2(D, FID(DV_avg2));
    const bool last = (iif == D.p.nfast + 1) && PRED;    // auxiliary pass: periodic images of the averages (:821-855)
    const bool inR = (i >= b.IstrR && i <= b.IendR && j >= b.JstrR && j <= b.JendR);
    const bool inU = (i >= b.Istr && i <= b.IendR && j >= b.JstrR && j <= b.JendR);
    const bool inV = (i >= b.IstrR && i <= b.IendR && j >= b.Jstr && j <= b.JendR);
    if (PRED) {
      if (iif == 1) {
        const double cff2 = (-1.0 / 12.0) * D.w2[iif + 1];
        if (inR) Zt(i, j) = 0.0;
        if (inU) { DU1(i, j) = 0.0; DU2(i, j) = cff2 * DUon(s, i, j); }
        if (inV) { DV1(i, j) = 0.0; DV2(i, j) = cff2 * DVom(s, i, j); }
      } else {
        const double cff1 = D.w1[iif - 1];
        const double cff2 = (8.0 / 12.0) * D.w2[iif] - (1.0 / 12.0) * D.w2[iif + 1];
        if (inR) { const double val = Zt(i, j) + cff1 * s.zk(i, j); if (last) st(D, Zt, i, j, val); else Zt(i, j) = val; }
        if (inU) { const double du = DUon(s, i, j); const double val = DU1(i, j) + cff1 * du; if (last) st(D, DU1, i, j, val); else DU1(i, j) = val; DU2(i, j) = DU2(i, j) + cff2 * du; }
        if (inV) { const double dv = DVom(s, i, j); const double val = DV1(i, j) + cff1 * dv; if (last) st(D, DV1, i, j, val); else DV1(i, j) = val; DV2(i, j) = DV2(i, j) + cff2 * dv; }
      }
    } else {
      const double cff2 = (iif == 1) ? D.w2[iif] : (5.0 / 12.0) * D.w2[iif];
      if (inU) DU2(i, j) = DU2(i, j) + cff2 * DUon(s, i, j);
      if (inV) DV2(i, j) = DV2(i, j) + cff2 * DVom(s, i, j);
    }
  }
  if (iif > D.p.nfast) return;
  if (!(i >= b.Istr && i <= b.Iend && j >= b.Jstr && j <= b.Jend)) return;
  const bool doU = (i >= b.IstrU), doV = (j >= b.JstrV);
  const bool south = s.S && j == b.Jstr, north = s.N && j == b.Jend;
  // ---- free surface (:899-1072)
  const Zst z0 = zstate(s, i, j);
  if (mycomp == 0) {
    V2 zn = v2l(D, FID(zeta), knew);
    st(D, zn, i, j, z0.zn);
    if (south) st(D, zn, i, j - 1, z0.zn);           // zetabc_tile closed: zero gradient (zetabc.F:353-358,437-442)
    if (north) st(D, zn, i, j + 1, z0.zn);
    if (PRED) st(D, v2l(D, FID(rzeta), krhs), i, j, z0.rhs_zeta);
  }
  V2D m{v2(D, FID(fomn)), v2(D, FID(dndx)), v2(D, FID(dmde)), v2(D, FID(visc2_r)), v2(D, FID(visc2_p)), v2(D, FID(pmon_r)), v2(D, FID(pnom_r)),
        v2(D, FID(pmon_p)), v2(D, FID(pnom_p)), v2(D, FID(om_r)), v2(D, FID(on_r)), v2(D, FID(om_p)), v2(D, FID(on_p))};
  const bool curv = (D.p.app == ROMS_B200_APP_BENCHMARK);
  const double cg = 0.5 * D.p.g, c3 = 1.0 / 3.0;
  const int comp = mycomp;
  if (comp == 0 ? doU : doV) {
    const int di = comp == 0 ? 1 : 0, dj = 1 - di;
    const Zst zm = zstate(s, i - di, j - dj);
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
    V2 frc = v2(D, comp == 0 ? FID(rufrc) : FID(rvfrc));
    if (iif == 1 && PRED) {
      V3 r0n = v3l(D, comp == 0 ? FID(ru) : FID(rv), a.nstp), r0w = v3l(D, comp == 0 ? FID(ru) : FID(rv), a.nnew);
      const double fr = frc(i, j) - rhs;
      frc(i, j) = fr;
      if (a.stepmode == 0) rhs = rhs + fr;
      else if (a.stepmode == 1) rhs = rhs + 1.5 * fr - 0.5 * r0w(i, j, 0);
      else rhs = rhs + (23.0 / 12.0) * fr - (16.0 / 12.0) * r0w(i, j, 0) + (5.0 / 12.0) * r0n(i, j, 0);
      r0n(i, j, 0) = fr;
    } else rhs = rhs + frc(i, j);
    // time stepping (:2493-2674)
    const double Dstp = (s.zs(i, j) + h0) + (s.zs(i - di, j - dj) + hm);
    const double cff = (s.pm(i, j) + s.pm(i - di, j - dj)) * (s.pn(i, j) + s.pn(i - di, j - dj));
    const double fc = 1.0 / (z0.Dnew + zm.Dnew);
    V2 qs = v2l(D, comp == 0 ? FID(ubar) : FID(vbar), kstp), qn = v2l(D, comp == 0 ? FID(ubar) : FID(vbar), knew);
    double val;
    if (iif == 1 || PRED) {
      const double cff1 = (iif == 1) ? 0.5 * D.p.dtfast : D.p.dtfast;
      val = (qs(i, j) * Dstp + cff * cff1 * rhs) * fc;
    } else {
      const double cff1 = 0.5 * D.p.dtfast * 5.0 / 12.0, cff2 = 0.5 * D.p.dtfast * 8.0 / 12.0, cff3 = 0.5 * D.p.dtfast * 1.0 / 12.0;
      V2 rs = v2l(D, comp == 0 ? FID(rubar) : FID(rvbar), kstp), rp = v2l(D, comp == 0 ? FID(rubar) : FID(rvbar), ptsk);
      val = (qs(i, j) * Dstp + cff * (cff1 * rhs + cff2 * rs(i, j) - cff3 * rp(i, j))) * fc;
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

static inline void launch_boxes(roms_b200_ctx* c, cudaStream_t st, const Box* bx, int n, const Step2dArgs& a, const Dev* Duse = nullptr) {
  Boxes B{}; int gx = 1, gy = 1;
  for (int q = 0; q < n; ++q) {
    B.b[q] = bx[q];
    gx = std::max(gx, (bx[q].i1 - bx[q].i0 + S2_TX) / S2_TX); gy = std::max(gy, (bx[q].j1 - bx[q].j0 + S2_TY) / S2_TY);
  }
  // Consecutive sub-steps are chained by programmatic dependent launch, which overlaps the launch latency and block scheduling
  // of sub-step n+1 with the tail of sub-step n (the kernel waits in griddepcontrol.wait before its first read): 9.36 -> 9.07 us
  // per sub-step on BENCHMARK1, bit-identical.  Single tile only (with neighbours a halo kernel sits between two sub-steps and
  // the combination has not been measured on hardware: ROMS_B200_PDL=1 forces it on, ROMS_B200_NO_PDL=1 off).
  static const bool pdl_off = (getenv("ROMS_B200_NO_PDL") != nullptr), pdl_on = (getenv("ROMS_B200_PDL") != nullptr);
  const bool pdl = !pdl_off && (pdl_on || !c->comm);
  if (pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(gx, gy, n); cfg.blockDim = dim3(S2_TX, S2_TY, 2); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, step2d_kernel, Duse ? *Duse : c->D, B, a) != cudaSuccess) fprintf(stderr, "roms_b200: programmatic launch of step2d failed\n");
  } else step2d_kernel<<<dim3(gx, gy, n), dim3(S2_TX, S2_TY, 2), 0, st>>>(Duse ? *Duse : c->D, B, a);
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
