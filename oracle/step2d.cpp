// oracle/step2d.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h header).
// Nonlinear/step2d_LF_AM3.h:606-3056 (serial / shared-memory branches), the
// legacy leap-frog predictor / Adams-Moulton corrector barotropic kernel that
// step2d.F:20 actually includes.
#include "oracle.h"

namespace orc {

void step2d(Model& M, const Tile& T) {
  const Config& c = M.c;
  const int krhs = M.krhs, kstp = M.kstp, knew = M.knew, nstp = M.nstp, nnew = M.nnew, iif = M.iif, nfast = M.nfast;
  const bool PRED = M.PREDICTOR_2D_STEP, CORR = !PRED;
  const bool curv = (c.app == BENCHMARK);
  const double dtfast = M.dtfast, g = c.g, rho0 = c.rho0;
  F3 &zeta = M.zeta, &ubar = M.ubar, &vbar = M.vbar, &rzeta = M.rzeta, &rubar = M.rubar, &rvbar = M.rvbar;
  F2 &h = M.h, &pm = M.pm, &pn = M.pn, &on_u = M.on_u, &om_v = M.om_v, &rhoA = M.rhoA, &rhoS = M.rhoS;
  F2 &Zt_avg1 = M.Zt_avg1, &DU_avg1 = M.DU_avg1, &DU_avg2 = M.DU_avg2, &DV_avg1 = M.DV_avg1, &DV_avg2 = M.DV_avg2;
  F2 &rufrc = M.rufrc, &rvfrc = M.rvfrc; F4 &ru = M.ru, &rv = M.rv;
  const std::vector<double>&w1 = M.weight1, &w2 = M.weight2;
#define SS(x) S2 x(T.IminS, T.ImaxS, T.JminS, T.JmaxS)
  SS(Dgrad); SS(Dnew); SS(Drhs); SS(Drhs_p); SS(Dstp); SS(DUon); SS(DVom); SS(UFe); SS(UFx); SS(VFe); SS(VFx);
  SS(grad); SS(gzeta); SS(gzeta2); SS(gzetaSA); SS(rhs_ubar); SS(rhs_vbar); SS(rhs_zeta); SS(zeta_new); SS(zwrk);
#undef SS
  const int ptsk = 3 - kstp;
  // :664-702 total depth and transports at krhs
  for (int j = T.JstrVm2 - 1; j <= T.Jendp2; ++j) for (int i = T.IstrUm2 - 1; i <= T.Iendp2; ++i) Drhs(i, j) = zeta(i, j, krhs) + h(i, j);
  for (int j = T.JstrVm2 - 1; j <= T.Jendp2; ++j) for (int i = T.IstrUm2; i <= T.Iendp2; ++i) {
    double cff = 0.5 * on_u(i, j); double cff1 = cff * (Drhs(i, j) + Drhs(i - 1, j));
    DUon(i, j) = ubar(i, j, krhs) * cff1;
  }
  for (int j = T.JstrVm2; j <= T.Jendp2; ++j) for (int i = T.IstrUm2 - 1; i <= T.Iendp2; ++i) {
    double cff = 0.5 * om_v(i, j); double cff1 = cff * (Drhs(i, j) + Drhs(i, j - 1));
    DVom(i, j) = vbar(i, j, krhs) * cff1;
  }
  // :742-810 fast-time averaging
  if (PRED) {
    if (iif == 1) {
      double cff2 = (-1.0 / 12.0) * w2[iif + 1];
      for (int j = T.JstrR; j <= T.JendR; ++j) {
        for (int i = T.IstrR; i <= T.IendR; ++i) Zt_avg1(i, j) = 0.0;
        for (int i = T.Istr; i <= T.IendR; ++i) { DU_avg1(i, j) = 0.0; DU_avg2(i, j) = cff2 * DUon(i, j); }
      }
      for (int j = T.Jstr; j <= T.JendR; ++j) for (int i = T.IstrR; i <= T.IendR; ++i) { DV_avg1(i, j) = 0.0; DV_avg2(i, j) = cff2 * DVom(i, j); }
    } else {
      double cff1 = w1[iif - 1];
      double cff2 = (8.0 / 12.0) * w2[iif] - (1.0 / 12.0) * w2[iif + 1];
      for (int j = T.JstrR; j <= T.JendR; ++j) {
        for (int i = T.IstrR; i <= T.IendR; ++i) Zt_avg1(i, j) = Zt_avg1(i, j) + cff1 * zeta(i, j, krhs);
        for (int i = T.Istr; i <= T.IendR; ++i) { DU_avg1(i, j) = DU_avg1(i, j) + cff1 * DUon(i, j); DU_avg2(i, j) = DU_avg2(i, j) + cff2 * DUon(i, j); }
      }
      for (int j = T.Jstr; j <= T.JendR; ++j) for (int i = T.IstrR; i <= T.IendR; ++i) {
        DV_avg1(i, j) = DV_avg1(i, j) + cff1 * DVom(i, j); DV_avg2(i, j) = DV_avg2(i, j) + cff2 * DVom(i, j);
      }
    }
  } else {
    double cff2 = (iif == 1) ? w2[iif] : (5.0 / 12.0) * w2[iif];
    for (int j = T.JstrR; j <= T.JendR; ++j) for (int i = T.Istr; i <= T.IendR; ++i) DU_avg2(i, j) = DU_avg2(i, j) + cff2 * DUon(i, j);
    for (int j = T.Jstr; j <= T.JendR; ++j) for (int i = T.IstrR; i <= T.IendR; ++i) DV_avg2(i, j) = DV_avg2(i, j) + cff2 * DVom(i, j);
  }
  // :821-883 auxiliary last pass
  if (iif == (nfast + 1) && PRED) { exchange_r2d(M, T, Zt_avg1); exchange_u2d(M, T, DU_avg1); exchange_v2d(M, T, DV_avg1); }
  if (iif > nfast) return;
  // :899-980 free surface
  double fac = 1000.0 / rho0;
  if (iif == 1) {
    double cff1 = dtfast;
    for (int j = T.JstrV - 1; j <= T.Jend; ++j) for (int i = T.IstrU - 1; i <= T.Iend; ++i) {
      rhs_zeta(i, j) = (DUon(i, j) - DUon(i + 1, j)) + (DVom(i, j) - DVom(i, j + 1));
      zeta_new(i, j) = zeta(i, j, kstp) + pm(i, j) * pn(i, j) * cff1 * rhs_zeta(i, j);
      Dnew(i, j) = zeta_new(i, j) + h(i, j);
      zwrk(i, j) = 0.5 * (zeta(i, j, kstp) + zeta_new(i, j));
      gzeta(i, j) = (fac + rhoS(i, j)) * zwrk(i, j);
      gzeta2(i, j) = gzeta(i, j) * zwrk(i, j);
      gzetaSA(i, j) = zwrk(i, j) * (rhoS(i, j) - rhoA(i, j));
    }
  } else if (PRED) {
    double cff1 = 2.0 * dtfast, cff4 = 4.0 / 25.0, cff5 = 1.0 - 2.0 * cff4;
    for (int j = T.JstrV - 1; j <= T.Jend; ++j) for (int i = T.IstrU - 1; i <= T.Iend; ++i) {
      rhs_zeta(i, j) = (DUon(i, j) - DUon(i + 1, j)) + (DVom(i, j) - DVom(i, j + 1));
      zeta_new(i, j) = zeta(i, j, kstp) + pm(i, j) * pn(i, j) * cff1 * rhs_zeta(i, j);
      Dnew(i, j) = zeta_new(i, j) + h(i, j);
      zwrk(i, j) = cff5 * zeta(i, j, krhs) + cff4 * (zeta(i, j, kstp) + zeta_new(i, j));
      gzeta(i, j) = (fac + rhoS(i, j)) * zwrk(i, j);
      gzeta2(i, j) = gzeta(i, j) * zwrk(i, j);
      gzetaSA(i, j) = zwrk(i, j) * (rhoS(i, j) - rhoA(i, j));
    }
  } else if (CORR) {
    double cff1 = dtfast * 5.0 / 12.0, cff2 = dtfast * 8.0 / 12.0, cff3 = dtfast * 1.0 / 12.0, cff4 = 2.0 / 5.0, cff5 = 1.0 - cff4;
    for (int j = T.JstrV - 1; j <= T.Jend; ++j) for (int i = T.IstrU - 1; i <= T.Iend; ++i) {
      double cff = cff1 * ((DUon(i, j) - DUon(i + 1, j)) + (DVom(i, j) - DVom(i, j + 1)));
      zeta_new(i, j) = zeta(i, j, kstp) + pm(i, j) * pn(i, j) * (cff + cff2 * rzeta(i, j, kstp) - cff3 * rzeta(i, j, ptsk));
      Dnew(i, j) = zeta_new(i, j) + h(i, j);
      zwrk(i, j) = cff5 * zeta_new(i, j) + cff4 * zeta(i, j, krhs);
      gzeta(i, j) = (fac + rhoS(i, j)) * zwrk(i, j);
      gzeta2(i, j) = gzeta(i, j) * zwrk(i, j);
      gzetaSA(i, j) = zwrk(i, j) * (rhoS(i, j) - rhoA(i, j));
    }
  }
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) zeta(i, j, knew) = zeta_new(i, j);
  if (PRED) {
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) rzeta(i, j, krhs) = rhs_zeta(i, j);
    exchange_r2d(M, T, rzeta.slab(krhs));
  }
  zetabc(M, T, knew);
  exchange_r2d(M, T, zeta.slab(knew));
  // :1088-1205 pressure gradient (VAR_RHO_2D)
  {
    double cff1 = 0.5 * g, cff2 = 1.0 / 3.0;
    for (int j = T.Jstr; j <= T.Jend; ++j) {
      for (int i = T.IstrU; i <= T.Iend; ++i)
        rhs_ubar(i, j) = cff1 * on_u(i, j) *
                         ((h(i - 1, j) + h(i, j)) * (gzeta(i - 1, j) - gzeta(i, j)) +
                          (h(i - 1, j) - h(i, j)) * (gzetaSA(i - 1, j) + gzetaSA(i, j) + cff2 * (rhoA(i - 1, j) - rhoA(i, j)) * (zwrk(i - 1, j) - zwrk(i, j))) +
                          (gzeta2(i - 1, j) - gzeta2(i, j)));
      if (j >= T.JstrV)
        for (int i = T.Istr; i <= T.Iend; ++i)
          rhs_vbar(i, j) = cff1 * om_v(i, j) *
                           ((h(i, j - 1) + h(i, j)) * (gzeta(i, j - 1) - gzeta(i, j)) +
                            (h(i, j - 1) - h(i, j)) * (gzetaSA(i, j - 1) + gzetaSA(i, j) + cff2 * (rhoA(i, j - 1) - rhoA(i, j)) * (zwrk(i, j - 1) - zwrk(i, j))) +
                            (gzeta2(i, j - 1) - gzeta2(i, j)));
    }
  }
  // :1251-1423 fourth-order centred advection
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrUm1; i <= T.Iendp1; ++i) {
    grad(i, j) = ubar(i - 1, j, krhs) - 2.0 * ubar(i, j, krhs) + ubar(i + 1, j, krhs);
    Dgrad(i, j) = DUon(i - 1, j) - 2.0 * DUon(i, j) + DUon(i + 1, j);
  }
  if (!M.EWperiodic) {
    if (T.W) for (int j = T.Jstr; j <= T.Jend; ++j) { grad(T.Istr, j) = grad(T.Istr + 1, j); Dgrad(T.Istr, j) = Dgrad(T.Istr + 1, j); }
    if (T.E) for (int j = T.Jstr; j <= T.Jend; ++j) { grad(T.Iend + 1, j) = grad(T.Iend, j); Dgrad(T.Iend + 1, j) = Dgrad(T.Iend, j); }
  }
  double cff = 1.0 / 6.0;
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU - 1; i <= T.Iend; ++i)
    UFx(i, j) = 0.25 * (ubar(i, j, krhs) + ubar(i + 1, j, krhs) - cff * (grad(i, j) + grad(i + 1, j))) *
                (DUon(i, j) + DUon(i + 1, j) - cff * (Dgrad(i, j) + Dgrad(i + 1, j)));
  for (int j = T.Jstrm1; j <= T.Jendp1; ++j) for (int i = T.IstrU; i <= T.Iend; ++i)
    grad(i, j) = ubar(i, j - 1, krhs) - 2.0 * ubar(i, j, krhs) + ubar(i, j + 1, krhs);
  if (!M.NSperiodic) {
    if (T.S) for (int i = T.IstrU; i <= T.Iend; ++i) grad(i, T.Jstr - 1) = grad(i, T.Jstr);
    if (T.N) for (int i = T.IstrU; i <= T.Iend; ++i) grad(i, T.Jend + 1) = grad(i, T.Jend);
  }
  for (int j = T.Jstr; j <= T.Jend + 1; ++j) for (int i = T.IstrU - 1; i <= T.Iend; ++i)
    Dgrad(i, j) = DVom(i - 1, j) - 2.0 * DVom(i, j) + DVom(i + 1, j);
  for (int j = T.Jstr; j <= T.Jend + 1; ++j) for (int i = T.IstrU; i <= T.Iend; ++i)
    UFe(i, j) = 0.25 * (ubar(i, j, krhs) + ubar(i, j - 1, krhs) - cff * (grad(i, j) + grad(i, j - 1))) *
                (DVom(i, j) + DVom(i - 1, j) - cff * (Dgrad(i, j) + Dgrad(i - 1, j)));
  for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istrm1; i <= T.Iendp1; ++i)
    grad(i, j) = vbar(i - 1, j, krhs) - 2.0 * vbar(i, j, krhs) + vbar(i + 1, j, krhs);
  if (!M.EWperiodic) {
    if (T.W) for (int j = T.JstrV; j <= T.Jend; ++j) grad(T.Istr - 1, j) = grad(T.Istr, j);
    if (T.E) for (int j = T.JstrV; j <= T.Jend; ++j) grad(T.Iend + 1, j) = grad(T.Iend, j);
  }
  for (int j = T.JstrV - 1; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend + 1; ++i)
    Dgrad(i, j) = DUon(i, j - 1) - 2.0 * DUon(i, j) + DUon(i, j + 1);
  for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend + 1; ++i)
    VFx(i, j) = 0.25 * (vbar(i, j, krhs) + vbar(i - 1, j, krhs) - cff * (grad(i, j) + grad(i - 1, j))) *
                (DUon(i, j) + DUon(i, j - 1) - cff * (Dgrad(i, j) + Dgrad(i, j - 1)));
  for (int j = T.JstrVm1; j <= T.Jendp1; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
    grad(i, j) = vbar(i, j - 1, krhs) - 2.0 * vbar(i, j, krhs) + vbar(i, j + 1, krhs);
    Dgrad(i, j) = DVom(i, j - 1) - 2.0 * DVom(i, j) + DVom(i, j + 1);
  }
  if (!M.NSperiodic) {
    if (T.S) for (int i = T.Istr; i <= T.Iend; ++i) { grad(i, T.Jstr) = grad(i, T.Jstr + 1); Dgrad(i, T.Jstr) = Dgrad(i, T.Jstr + 1); }
    if (T.N) for (int i = T.Istr; i <= T.Iend; ++i) { grad(i, T.Jend + 1) = grad(i, T.Jend); Dgrad(i, T.Jend + 1) = Dgrad(i, T.Jend); }
  }
  for (int j = T.JstrV - 1; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i)
    VFe(i, j) = 0.25 * (vbar(i, j, krhs) + vbar(i, j + 1, krhs) - cff * (grad(i, j) + grad(i, j + 1))) *
                (DVom(i, j) + DVom(i, j + 1) - cff * (Dgrad(i, j) + Dgrad(i, j + 1)));
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) {
    double cff1 = UFx(i, j) - UFx(i - 1, j), cff2 = UFe(i, j + 1) - UFe(i, j); fac = cff1 + cff2;
    rhs_ubar(i, j) = rhs_ubar(i, j) - fac;
  }
  for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
    double cff1 = VFx(i + 1, j) - VFx(i, j), cff2 = VFe(i, j) - VFe(i, j - 1); fac = cff1 + cff2;
    rhs_vbar(i, j) = rhs_vbar(i, j) - fac;
  }
  // :1432-1458 Coriolis
  for (int j = T.JstrV - 1; j <= T.Jend; ++j) for (int i = T.IstrU - 1; i <= T.Iend; ++i) {
    double cf = 0.5 * Drhs(i, j) * M.fomn(i, j);
    UFx(i, j) = cf * (vbar(i, j, krhs) + vbar(i, j + 1, krhs));
    VFe(i, j) = cf * (ubar(i, j, krhs) + ubar(i + 1, j, krhs));
  }
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) { double fac1 = 0.5 * (UFx(i, j) + UFx(i - 1, j)); rhs_ubar(i, j) = rhs_ubar(i, j) + fac1; }
  for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) { double fac1 = 0.5 * (VFe(i, j) + VFe(i, j - 1)); rhs_vbar(i, j) = rhs_vbar(i, j) - fac1; }
  // :1497-1562 curvilinear terms
  if (curv) {
    for (int j = T.JstrV - 1; j <= T.Jend; ++j) for (int i = T.IstrU - 1; i <= T.Iend; ++i) {
      double cff1 = 0.5 * (vbar(i, j, krhs) + vbar(i, j + 1, krhs));
      double cff2 = 0.5 * (ubar(i, j, krhs) + ubar(i + 1, j, krhs));
      double cff3 = cff1 * M.dndx(i, j), cff4 = cff2 * M.dmde(i, j);
      double cf = Drhs(i, j) * (cff3 - cff4);
      UFx(i, j) = cf * cff1; VFe(i, j) = cf * cff2;
    }
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) { double fac1 = 0.5 * (UFx(i, j) + UFx(i - 1, j)); rhs_ubar(i, j) = rhs_ubar(i, j) + fac1; }
    for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) { double fac1 = 0.5 * (VFe(i, j) + VFe(i, j - 1)); rhs_vbar(i, j) = rhs_vbar(i, j) - fac1; }
  }
  // :1574-1651 harmonic viscosity
  for (int j = T.Jstr; j <= T.Jend + 1; ++j) for (int i = T.Istr; i <= T.Iend + 1; ++i)
    Drhs_p(i, j) = 0.25 * (Drhs(i, j) + Drhs(i - 1, j) + Drhs(i, j - 1) + Drhs(i - 1, j - 1));
  for (int j = T.JstrV - 1; j <= T.Jend; ++j) for (int i = T.IstrU - 1; i <= T.Iend; ++i) {
    double cf = M.visc2_r(i, j) * Drhs(i, j) * 0.5 *
                (M.pmon_r(i, j) * ((pn(i, j) + pn(i + 1, j)) * ubar(i + 1, j, krhs) - (pn(i - 1, j) + pn(i, j)) * ubar(i, j, krhs)) -
                 M.pnom_r(i, j) * ((pm(i, j) + pm(i, j + 1)) * vbar(i, j + 1, krhs) - (pm(i, j - 1) + pm(i, j)) * vbar(i, j, krhs)));
    UFx(i, j) = M.on_r(i, j) * M.on_r(i, j) * cf;
    VFe(i, j) = M.om_r(i, j) * M.om_r(i, j) * cf;
  }
  for (int j = T.Jstr; j <= T.Jend + 1; ++j) for (int i = T.Istr; i <= T.Iend + 1; ++i) {
    double cf = M.visc2_p(i, j) * Drhs_p(i, j) * 0.5 *
                (M.pmon_p(i, j) * ((pn(i, j - 1) + pn(i, j)) * vbar(i, j, krhs) - (pn(i - 1, j - 1) + pn(i - 1, j)) * vbar(i - 1, j, krhs)) +
                 M.pnom_p(i, j) * ((pm(i - 1, j) + pm(i, j)) * ubar(i, j, krhs) - (pm(i - 1, j - 1) + pm(i, j - 1)) * ubar(i, j - 1, krhs)));
    UFe(i, j) = M.om_p(i, j) * M.om_p(i, j) * cf;
    VFx(i, j) = M.on_p(i, j) * M.on_p(i, j) * cf;
  }
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) {
    double cff1 = 0.5 * (pn(i - 1, j) + pn(i, j)) * (UFx(i, j) - UFx(i - 1, j));
    double cff2 = 0.5 * (pm(i - 1, j) + pm(i, j)) * (UFe(i, j + 1) - UFe(i, j));
    fac = cff1 + cff2; rhs_ubar(i, j) = rhs_ubar(i, j) + fac;
  }
  for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
    double cff1 = 0.5 * (pn(i, j - 1) + pn(i, j)) * (VFx(i + 1, j) - VFx(i, j));
    double cff2 = 0.5 * (pm(i, j - 1) + pm(i, j)) * (VFe(i, j) - VFe(i, j - 1));
    fac = cff1 - cff2; rhs_vbar(i, j) = rhs_vbar(i, j) + fac;
  }
  // :2241-2459 coupling with the 3-D forcing
  if (iif == 1 && PRED) {
    if (M.iic == M.ntfirst) {
      for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) {
        rufrc(i, j) = rufrc(i, j) - rhs_ubar(i, j); rhs_ubar(i, j) = rhs_ubar(i, j) + rufrc(i, j); ru(i, j, 0, nstp) = rufrc(i, j);
      }
      for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
        rvfrc(i, j) = rvfrc(i, j) - rhs_vbar(i, j); rhs_vbar(i, j) = rhs_vbar(i, j) + rvfrc(i, j); rv(i, j, 0, nstp) = rvfrc(i, j);
      }
    } else if (M.iic == M.ntfirst + 1) {
      for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) {
        rufrc(i, j) = rufrc(i, j) - rhs_ubar(i, j);
        rhs_ubar(i, j) = rhs_ubar(i, j) + 1.5 * rufrc(i, j) - 0.5 * ru(i, j, 0, nnew);
        ru(i, j, 0, nstp) = rufrc(i, j);
      }
      for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
        rvfrc(i, j) = rvfrc(i, j) - rhs_vbar(i, j);
        rhs_vbar(i, j) = rhs_vbar(i, j) + 1.5 * rvfrc(i, j) - 0.5 * rv(i, j, 0, nnew);
        rv(i, j, 0, nstp) = rvfrc(i, j);
      }
    } else {
      double cff1 = 23.0 / 12.0, cff2 = 16.0 / 12.0, cff3 = 5.0 / 12.0;
      for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) {
        rufrc(i, j) = rufrc(i, j) - rhs_ubar(i, j);
        rhs_ubar(i, j) = rhs_ubar(i, j) + cff1 * rufrc(i, j) - cff2 * ru(i, j, 0, nnew) + cff3 * ru(i, j, 0, nstp);
        ru(i, j, 0, nstp) = rufrc(i, j);
      }
      for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
        rvfrc(i, j) = rvfrc(i, j) - rhs_vbar(i, j);
        rhs_vbar(i, j) = rhs_vbar(i, j) + cff1 * rvfrc(i, j) - cff2 * rv(i, j, 0, nnew) + cff3 * rv(i, j, 0, nstp);
        rv(i, j, 0, nstp) = rvfrc(i, j);
      }
    }
  } else {
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) rhs_ubar(i, j) = rhs_ubar(i, j) + rufrc(i, j);
    for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) rhs_vbar(i, j) = rhs_vbar(i, j) + rvfrc(i, j);
  }
  // :2493-2674 time-step the momentum
  for (int j = T.JstrV - 1; j <= T.Jend; ++j) for (int i = T.IstrU - 1; i <= T.Iend; ++i) Dstp(i, j) = zeta(i, j, kstp) + h(i, j);
  if (iif == 1 || PRED) {
    double cff1 = (iif == 1) ? 0.5 * dtfast : dtfast;
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) {
      double cf = (pm(i, j) + pm(i - 1, j)) * (pn(i, j) + pn(i - 1, j));
      double fc = 1.0 / (Dnew(i, j) + Dnew(i - 1, j));
      ubar(i, j, knew) = (ubar(i, j, kstp) * (Dstp(i, j) + Dstp(i - 1, j)) + cf * cff1 * rhs_ubar(i, j)) * fc;
    }
    for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
      double cf = (pm(i, j) + pm(i, j - 1)) * (pn(i, j) + pn(i, j - 1));
      double fc = 1.0 / (Dnew(i, j) + Dnew(i, j - 1));
      vbar(i, j, knew) = (vbar(i, j, kstp) * (Dstp(i, j) + Dstp(i, j - 1)) + cf * cff1 * rhs_vbar(i, j)) * fc;
    }
  } else {
    double cff1 = 0.5 * dtfast * 5.0 / 12.0, cff2 = 0.5 * dtfast * 8.0 / 12.0, cff3 = 0.5 * dtfast * 1.0 / 12.0;
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) {
      double cf = (pm(i, j) + pm(i - 1, j)) * (pn(i, j) + pn(i - 1, j));
      double fc = 1.0 / (Dnew(i, j) + Dnew(i - 1, j));
      ubar(i, j, knew) = (ubar(i, j, kstp) * (Dstp(i, j) + Dstp(i - 1, j)) +
                          cf * (cff1 * rhs_ubar(i, j) + cff2 * rubar(i, j, kstp) - cff3 * rubar(i, j, ptsk))) * fc;
    }
    for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
      double cf = (pm(i, j) + pm(i, j - 1)) * (pn(i, j) + pn(i, j - 1));
      double fc = 1.0 / (Dnew(i, j) + Dnew(i, j - 1));
      vbar(i, j, knew) = (vbar(i, j, kstp) * (Dstp(i, j) + Dstp(i, j - 1)) +
                          cf * (cff1 * rhs_vbar(i, j) + cff2 * rvbar(i, j, kstp) - cff3 * rvbar(i, j, ptsk))) * fc;
    }
  }
  if (PRED) {
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) rubar(i, j, krhs) = rhs_ubar(i, j);
    for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) rvbar(i, j, krhs) = rhs_vbar(i, j);
  }
  u2dbc(M, T, knew); v2dbc(M, T, knew);
  exchange_u2d(M, T, ubar.slab(knew)); exchange_v2d(M, T, vbar.slab(knew));
}

}  // namespace orc
