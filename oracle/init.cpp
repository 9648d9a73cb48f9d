// oracle/init.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h header).
// Analytical grid / initial state / forcing and the main3d sequencing.
#include "oracle.h"
#include <algorithm>
#include <stdexcept>
#include <thread>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>

namespace orc {

static const double pi = 3.14159265358979323846;      // mod_scalars.F:813
static const double deg2rad = pi / 180.0;             // mod_scalars.F:814
static const double Eradius = 6371315.0;              // mod_scalars.F:459
static const double Cp = 3985.0;                      // mod_scalars.F:456
static const double Csolar = 1353.0;                  // mod_scalars.F:457

// Utility/set_scoord.F:165-178 (hc) and :393-433 (Vstretching=4)
void set_scoord(Model& M) {
  const int N = M.N; const double theta_s = M.c.theta_s, theta_b = M.c.theta_b;
  M.hc = M.c.Tcline;                                   // Vtransform=2
  M.sc_r.assign(N + 1, 0.0); M.sc_w.assign(N + 1, 0.0); M.Cs_r.assign(N + 1, 0.0); M.Cs_w.assign(N + 1, 0.0);
  double ds = 1.0 / (double)N;
  M.sc_w[N] = 0.0; M.Cs_w[N] = 0.0;
  for (int k = N - 1; k >= 1; --k) {
    double sc_w = ds * (double)(k - N), Csur;
    M.sc_w[k] = sc_w;
    if (theta_s > 0.0) Csur = (1.0 - std::cosh(theta_s * sc_w)) / (std::cosh(theta_s) - 1.0);
    else Csur = -(sc_w * sc_w);
    if (theta_b > 0.0) M.Cs_w[k] = (std::exp(theta_b * Csur) - 1.0) / (1.0 - std::exp(-theta_b));
    else M.Cs_w[k] = Csur;
  }
  M.sc_w[0] = -1.0; M.Cs_w[0] = -1.0;
  for (int k = 1; k <= N; ++k) {
    double sc_r = ds * ((double)(k - N) - 0.5), Csur;
    M.sc_r[k] = sc_r;
    if (theta_s > 0.0) Csur = (1.0 - std::cosh(theta_s * sc_r)) / (std::cosh(theta_s) - 1.0);
    else Csur = -(sc_r * sc_r);
    if (theta_b > 0.0) M.Cs_r[k] = (std::exp(theta_b * Csur) - 1.0) / (1.0 - std::exp(-theta_b));
    else M.Cs_r[k] = Csur;
  }
}

// Utility/set_weights.F:47-195 (POWER_LAW; real(r16) accumulators -> __float128)
void set_weights(Model& M) {
  typedef __float128 r16;
  const int ndtfast = M.c.ndtfast;
  const double Falpha = 2.0, Fbeta = 4.0, Fgamma = 0.284;    // mod_scalars.F:327-329
  std::vector<double> w1(2 * ndtfast + 2, 0.0), w2(2 * ndtfast + 2, 0.0);
  int nfast = 0;
  double scale = (Falpha + 1.0) * (Falpha + Fbeta + 1.0) / ((Falpha + 2.0) * (Falpha + Fbeta + 2.0) * (double)ndtfast);
  double gamma = Fgamma * std::max(0.0, 1.0 - 10.0 / (double)ndtfast);
  for (int iter = 1; iter <= 16; ++iter) {
    nfast = 0;
    for (int i = 1; i <= 2 * ndtfast; ++i) {
      r16 cff = (r16)scale * (r16)(double)i;
      r16 c2 = cff * cff;                 // cff**Falpha,  Falpha=2
      r16 c6 = c2 * c2 * c2;              // cff**(Falpha+Fbeta) = cff**6
      w1[i] = (double)(c2 - c6 - (r16)gamma * cff);
      if (w1[i] > 0.0) nfast = i;
      if (nfast > 0 && w1[i] < 0.0) w1[i] = 0.0;
    }
    r16 wsum = 0, shift = 0;
    for (int i = 1; i <= nfast; ++i) { wsum = wsum + (r16)w1[i]; shift = shift + (r16)w1[i] * (r16)(double)i; }
    scale = (double)((r16)scale * shift / (wsum * (r16)(double)ndtfast));
  }
  for (int iter = 1; iter <= ndtfast; ++iter) {
    r16 wsum = 0, shift = 0;
    for (int i = 1; i <= nfast; ++i) { wsum = wsum + (r16)w1[i]; shift = shift + (r16)(double)i * (r16)w1[i]; }
    shift = shift / wsum;
    r16 cff = (r16)(double)ndtfast - shift;
    if (cff > (r16)1.0) {
      nfast = nfast + 1;
      for (int i = nfast; i >= 2; --i) w1[i] = w1[i - 1];
      w1[1] = 0.0;
    } else if (cff > (r16)0.0) {
      wsum = (r16)1.0 - cff;
      for (int i = nfast; i >= 2; --i) w1[i] = (double)(wsum * (r16)w1[i] + cff * (r16)w1[i - 1]);
      w1[1] = (double)(wsum * (r16)w1[1]);
    } else if (cff < (r16)(-1.0)) {
      nfast = nfast - 1;
      for (int i = 1; i <= nfast; ++i) w1[i] = w1[i + 1];
      w1[nfast + 1] = 0.0;
    } else if (cff < (r16)0.0) {
      wsum = (r16)1.0 + cff;
      for (int i = 1; i <= nfast - 1; ++i) w1[i] = (double)(wsum * (r16)w1[i] - cff * (r16)w1[i + 1]);
      w1[nfast] = (double)(wsum * (r16)w1[nfast]);
    }
  }
  for (int j = 1; j <= nfast; ++j) {
    r16 cff = (r16)w1[j];
    for (int i = 1; i <= j; ++i) w2[i] = (double)((r16)w2[i] + cff);
  }
  r16 wsum = 0, cff = 0;
  for (int i = 1; i <= nfast; ++i) { wsum = wsum + (r16)w1[i]; cff = cff + (r16)w2[i]; }
  wsum = (r16)1.0 / wsum; cff = (r16)1.0 / cff;
  for (int i = 1; i <= nfast; ++i) { w1[i] = (double)(wsum * (r16)w1[i]); w2[i] = (double)(cff * (r16)w2[i]); }
  M.nfast = nfast; M.weight1 = w1; M.weight2 = w2;
  M.dtfast = M.c.dt / (double)ndtfast;              // Utility/inp_par.F (dtfast=dt/ndtfast)
}

// Functionals/ana_grid.h: BENCHMARK :243-248,462-482,677-691,870-876,931-937 ;
//                         UPWELLING :386-391,560-575,1058-1070
void ana_grid(Model& M, const Tile& T) {
  const int Lm = M.Lm, Mm = M.Mm;
  int Imin = T.W ? T.Istr - 1 : T.Istr, Imax = T.E ? T.Iend + 1 : T.Iend;
  int Jmin = T.S ? T.Jstr - 1 : T.Jstr, Jmax = T.N ? T.Jend + 1 : T.Jend;
  S2 wrkX(T.IminS, T.ImaxS, T.JminS, T.JmaxS), wrkY(T.IminS, T.ImaxS, T.JminS, T.JmaxS);
  const int j0 = std::min(T.JstrT, T.Jstr - 1), j1 = std::max(T.Jend + 1, T.JendT);
  const int i0 = std::min(T.IstrT, T.Istr - 1), i1 = std::max(T.Iend + 1, T.IendT);
  if (M.c.app == BENCHMARK) {
    double Xsize = 360.0, Esize = 20.0;
    double dx = Xsize / (double)Lm, dy = Esize / (double)Mm;
    for (int j = Jmin; j <= Jmax; ++j) {
      double val1 = -70.0 + dy * ((double)j - 0.5);
      for (int i = Imin; i <= Imax; ++i) { M.lonr(i, j) = dx * ((double)i - 0.5); M.latr(i, j) = val1; }
    }
    double val1 = (double)Lm / (2.0 * pi * Eradius);
    double val2 = (double)Mm * 360.0 / (2.0 * pi * Eradius * Esize);
    for (int j = j0; j <= j1; ++j) {
      double cff = 1.0 / std::cos((-70.0 + dy * ((double)j - 0.5)) * deg2rad);
      for (int i = i0; i <= i1; ++i) { wrkX(i, j) = val1 * cff; wrkY(i, j) = val2; }
    }
    for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) { M.pm(i, j) = wrkX(i, j); M.pn(i, j) = wrkY(i, j); }
    exchange_r2d(M, T, M.pm); exchange_r2d(M, T, M.pn);
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
      M.dndx(i, j) = 0.5 * ((1.0 / wrkY(i + 1, j)) - (1.0 / wrkY(i - 1, j)));
      M.dmde(i, j) = 0.5 * ((1.0 / wrkX(i, j + 1)) - (1.0 / wrkX(i, j - 1)));
    }
    exchange_r2d(M, T, M.dndx); exchange_r2d(M, T, M.dmde);
    for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) M.angler(i, j) = 0.0;
    exchange_r2d(M, T, M.angler);
    double v1 = 2.0 * (2.0 * pi * 366.25 / 365.25) / 86400.0;
    for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) M.f(i, j) = v1 * std::sin(M.latr(i, j) * deg2rad);
    exchange_r2d(M, T, M.f);
    for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i)
      M.h(i, j) = 500.0 + 1750.0 * (1.0 + std::tanh((68.0 + M.latr(i, j)) / dy));
    exchange_r2d(M, T, M.h);
  } else {
    double Xsize = 1000.0 * (double)Lm, Esize = 1000.0 * (double)Mm, depth = 150.0, f0 = -8.26e-05, beta = 0.0;
    double dx = Xsize / (double)Lm, dy = Esize / (double)Mm;
    for (int j = Jmin; j <= Jmax; ++j) for (int i = Imin; i <= Imax; ++i) {
      M.xr(i, j) = dx * ((double)(i - 1) + 0.5); M.yr(i, j) = dy * ((double)(j - 1) + 0.5);
    }
    for (int j = j0; j <= j1; ++j) for (int i = i0; i <= i1; ++i) { wrkX(i, j) = 1.0 / dx; wrkY(i, j) = 1.0 / dy; }
    for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) { M.pm(i, j) = wrkX(i, j); M.pn(i, j) = wrkY(i, j); }
    exchange_r2d(M, T, M.pm); exchange_r2d(M, T, M.pn);
    for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) M.angler(i, j) = 0.0;
    exchange_r2d(M, T, M.angler);
    if (beta == 0.0) { for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) M.f(i, j) = f0; }
    else { double v1 = 0.5 * Esize; for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) M.f(i, j) = f0 + beta * (M.yr(i, j) - v1); }
    exchange_r2d(M, T, M.f);
    for (int j = T.JstrT; j <= T.JendT; ++j) {       // EWperiodic branch, ana_grid.h:1058-1070
      double val1 = (j <= Mm / 2) ? (double)j : (double)(Mm + 1 - j);
      double val2 = std::min(depth, 84.5 + 66.526 * std::tanh((val1 - 10.0) / 7.0));
      for (int i = T.IstrT; i <= T.IendT; ++i) M.h(i, j) = val2;
    }
    exchange_r2d(M, T, M.h);
  }
}

// Utility/metrics.F (metrics_tile): derived metric arrays
void metrics(Model& M, const Tile& T) {
  F2 &pm = M.pm, &pn = M.pn;
  for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) {
    M.om_r(i, j) = 1.0 / pm(i, j); M.on_r(i, j) = 1.0 / pn(i, j);
    M.omn(i, j) = 1.0 / (pm(i, j) * pn(i, j)); M.fomn(i, j) = M.f(i, j) * M.omn(i, j);
  }
  exchange_r2d(M, T, M.om_r); exchange_r2d(M, T, M.on_r); exchange_r2d(M, T, M.omn); exchange_r2d(M, T, M.fomn);
  for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) {
    M.pnom_r(i, j) = pn(i, j) / pm(i, j); M.pmon_r(i, j) = pm(i, j) / pn(i, j);
  }
  exchange_r2d(M, T, M.pnom_r); exchange_r2d(M, T, M.pmon_r);
  for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrP; i <= T.IendT; ++i) {
    M.pmon_u(i, j) = (pm(i - 1, j) + pm(i, j)) / (pn(i - 1, j) + pn(i, j));
    M.pnom_u(i, j) = (pn(i - 1, j) + pn(i, j)) / (pm(i - 1, j) + pm(i, j));
    M.om_u(i, j) = 2.0 / (pm(i - 1, j) + pm(i, j));
    M.on_u(i, j) = 2.0 / (pn(i - 1, j) + pn(i, j));
  }
  exchange_u2d(M, T, M.pmon_u); exchange_u2d(M, T, M.pnom_u); exchange_u2d(M, T, M.om_u); exchange_u2d(M, T, M.on_u);
  for (int j = T.JstrP; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) {
    M.pmon_v(i, j) = (pm(i, j - 1) + pm(i, j)) / (pn(i, j - 1) + pn(i, j));
    M.pnom_v(i, j) = (pn(i, j - 1) + pn(i, j)) / (pm(i, j - 1) + pm(i, j));
    M.om_v(i, j) = 2.0 / (pm(i, j - 1) + pm(i, j));
    M.on_v(i, j) = 2.0 / (pn(i, j - 1) + pn(i, j));
  }
  exchange_v2d(M, T, M.pmon_v); exchange_v2d(M, T, M.pnom_v); exchange_v2d(M, T, M.om_v); exchange_v2d(M, T, M.on_v);
  for (int j = T.JstrP; j <= T.JendT; ++j) for (int i = T.IstrP; i <= T.IendT; ++i) {
    M.pnom_p(i, j) = (pn(i - 1, j - 1) + pn(i - 1, j) + pn(i, j - 1) + pn(i, j)) / (pm(i - 1, j - 1) + pm(i - 1, j) + pm(i, j - 1) + pm(i, j));
    M.pmon_p(i, j) = (pm(i - 1, j - 1) + pm(i - 1, j) + pm(i, j - 1) + pm(i, j)) / (pn(i - 1, j - 1) + pn(i - 1, j) + pn(i, j - 1) + pn(i, j));
    M.om_p(i, j) = 4.0 / (pm(i - 1, j - 1) + pm(i - 1, j) + pm(i, j - 1) + pm(i, j));
    M.on_p(i, j) = 4.0 / (pn(i - 1, j - 1) + pn(i - 1, j) + pn(i, j - 1) + pn(i, j));
  }
  exchange_p2d(M, T, M.pnom_p); exchange_p2d(M, T, M.pmon_p); exchange_p2d(M, T, M.om_p); exchange_p2d(M, T, M.on_p);
  // set_depth with zero free surface (A2d=0), metrics.F
  std::vector<double> zero((size_t)M.ni * M.nj, 0.0);
  set_depth(M, T, F2{zero.data(), M.LBi, M.ni, M.LBj, M.nj});
}

// Utility/ini_hmixcoef.F (constant coefficients; no sponge)
void ini_hmixcoef(Model& M, const Tile& T) {
  (void)T;
  for (int j = M.LBj; j <= M.UBj; ++j) for (int i = M.LBi; i <= M.UBi; ++i) {
    M.visc2_p(i, j) = M.c.visc2; M.visc2_r(i, j) = M.c.visc2;
    for (int it = 1; it <= M.NT; ++it) M.diff2(i, j, it) = M.c.tnu2[it - 1];
  }
}

// Functionals/ana_initial.h: BENCHMARK :545-558 ; UPWELLING :828-848 ; fluid at rest
void ana_initial(Model& M, const Tile& T) {
  const int N = M.N; const double rho0 = M.c.rho0, g = M.c.g;
  for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrP; i <= T.IendT; ++i) M.ubar(i, j, 1) = 0.0;
  for (int j = T.JstrP; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) M.vbar(i, j, 1) = 0.0;
  for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) M.zeta(i, j, 1) = 0.0;
  for (int k = 1; k <= N; ++k) {
    for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrP; i <= T.IendT; ++i) M.u(i, j, k, 1) = 0.0;
    for (int j = T.JstrP; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) M.v(i, j, k, 1) = 0.0;
  }
  if (M.c.app == BENCHMARK) {
    double val1 = (44.69 / 39.382) * (44.69 / 39.382);
    double val2 = val1 * (rho0 * 800.0 / g) * (5.0e-05 / ((42.689 / 44.69) * (42.689 / 44.69)));
    for (int k = 1; k <= N; ++k) for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) {
      M.t(i, j, k, 1, 1) = val2 * std::exp(M.z_r(i, j, k) / 800.0) * (0.6 - 0.4 * std::tanh(M.z_r(i, j, k) / 800.0));
      M.t(i, j, k, 1, 2) = 35.0;
    }
  } else {
    for (int k = 1; k <= N; ++k) for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) {
      M.t(i, j, k, 1, 1) = M.c.T0 + 8.0 * std::exp(M.z_r(i, j, k) / 50.0);
      M.t(i, j, k, 1, 2) = M.c.S0;
    }
  }
}

// Nonlinear/ini_fields.F (set_zeta_timeavg_tile)
void set_zeta_timeavg(Model& M, const Tile& T) {
  for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) M.Zt_avg1(i, j) = M.zeta(i, j, M.kstp);
  exchange_r2d(M, T, M.Zt_avg1);
}
// Nonlinear/ini_fields.F (ini_zeta_tile)
void ini_zeta(Model& M, const Tile& T) {
  zetabc(M, T, M.kstp);
  exchange_r2d(M, T, M.zeta.slab(M.kstp));
  set_zeta_timeavg(M, T);
}
// Nonlinear/ini_fields.F (ini_fields_tile)
void ini_fields(Model& M, const Tile& T) {
  const int N = M.N, nstp = M.nstp, kstp = M.kstp;
  u3dbc(M, T, nstp); v3dbc(M, T, nstp);
  exchange_u3d(M, T, M.u.vol(nstp)); exchange_v3d(M, T, M.v.vol(nstp));
  std::vector<double> DC((size_t)(T.ImaxS - T.IminS + 1) * (N + 1)), CF(DC.size());
  auto dc = [&](int i, int k) -> double& { return DC[(i - T.IminS) + (size_t)(T.ImaxS - T.IminS + 1) * k]; };
  auto cf = [&](int i, int k) -> double& { return CF[(i - T.IminS) + (size_t)(T.ImaxS - T.IminS + 1) * k]; };
  for (int j = T.JstrB; j <= T.JendB; ++j) {
    for (int i = T.IstrM; i <= T.IendB; ++i) { dc(i, 0) = 0.0; cf(i, 0) = 0.0; }
    for (int k = 1; k <= N; ++k) for (int i = T.IstrM; i <= T.IendB; ++i) {
      dc(i, k) = 0.5 * (M.Hz(i, j, k) + M.Hz(i - 1, j, k));
      dc(i, 0) = dc(i, 0) + dc(i, k);
      cf(i, 0) = cf(i, 0) + dc(i, k) * M.u(i, j, k, nstp);
    }
    for (int i = T.IstrM; i <= T.IendB; ++i) { double cff1 = 1.0 / dc(i, 0); double cff2 = cf(i, 0) * cff1; M.ubar(i, j, kstp) = cff2; }
    if (j >= T.JstrM) {
      for (int i = T.IstrB; i <= T.IendB; ++i) { dc(i, 0) = 0.0; cf(i, 0) = 0.0; }
      for (int k = 1; k <= N; ++k) for (int i = T.IstrB; i <= T.IendB; ++i) {
        dc(i, k) = 0.5 * (M.Hz(i, j, k) + M.Hz(i, j - 1, k));
        dc(i, 0) = dc(i, 0) + dc(i, k);
        cf(i, 0) = cf(i, 0) + dc(i, k) * M.v(i, j, k, nstp);
      }
      for (int i = T.IstrB; i <= T.IendB; ++i) { double cff1 = 1.0 / dc(i, 0); double cff2 = cf(i, 0) * cff1; M.vbar(i, j, kstp) = cff2; }
    }
  }
  u2dbc(M, T, kstp); v2dbc(M, T, kstp);
  exchange_u2d(M, T, M.ubar.slab(kstp)); exchange_v2d(M, T, M.vbar.slab(kstp));
  for (int it = 1; it <= M.NT; ++it) { t3dbc(M, T, nstp, it); }
  for (int it = 1; it <= M.NT; ++it) exchange_r3d(M, T, M.t.vol(nstp, it));
}

// ---------------------------------------------------------------------------
// Utility/dateclock.F: caldate/datevec/ROUND for time_ref = 0 (proleptic
// gregorian, reference 0001-01-01 => datenum 367); only yday and hour are used.
static double ufloor(double X) { return X - std::fmod(X, 1.0) - std::fmod(2.0 + std::copysign(1.0, X), 3.0); }
static double tfloor(double X, double CT) {
  double Q = 1.0; if (X < 0.0) Q = 1.0 - CT;
  double RMAX = Q / (2.0 - CT), EPS5 = CT / Q;
  double Y = ufloor(X + std::max(CT, std::min(RMAX, EPS5 * std::fabs(1.0 + ufloor(X)))));
  if (X <= 0.0 || (Y - X) < RMAX) return Y;
  return Y - 1.0;
}
static void caldate(double tdays, double& yday, double& hour) {
  double DateNumber = 367.0 + tdays;
  double DayFraction = std::fabs(DateNumber - std::trunc(DateNumber));
  double seconds = DayFraction * 86400.0;
  double CT = 3.0 * 2.220446049250313e-16;
  seconds = tfloor(seconds + 0.5, CT);
  hour = seconds / 3600.0;
  // yearday(0001,01,01+n): day-of-year for dates within year 1 (runs here are < 1 year)
  yday = (double)(1 + (int)std::floor(tdays)) + DayFraction;
}

// Nonlinear/set_data.F (analytical branches only); Functionals/ana_*.h
void set_data(Model& M, const Tile& T) {
  const Config& c = M.c;
  if (c.app == BENCHMARK) {
    // ana_cloud.h, ana_tair.h, ana_humid.h
    for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) { M.cloud(i, j) = 0.6; M.Tair(i, j) = 4.0; M.Hair(i, j) = 0.8; }
    exchange_r2d(M, T, M.cloud); exchange_r2d(M, T, M.Tair); exchange_r2d(M, T, M.Hair);
    // ana_srflux.h (ALBEDO; analytic zenith-angle formula)
    double yday, hour; caldate(M.tdays, yday, hour);
    double Dangle = 23.44 * std::cos((172.0 - yday) * 2.0 * pi / 365.2425);
    Dangle = Dangle * deg2rad;
    double Hangle = (12.0 - hour) * pi / 12.0;
    double Rsolar = Csolar / (c.rho0 * Cp);
    const double alb_w = 0.06;
    for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) {
      double LatRad = M.latr(i, j) * deg2rad;
      double cff1 = std::sin(LatRad) * std::sin(Dangle);
      double cff2 = std::cos(LatRad) * std::cos(Dangle);
      M.srflx(i, j) = 0.0;
      double zenith = cff1 + cff2 * std::cos(Hangle - M.lonr(i, j) * deg2rad);
      if (zenith > 0.0) {
        double cff = (0.7859 + 0.03477 * M.Tair(i, j)) / (1.0 + 0.00412 * M.Tair(i, j));
        double e_sat = std::pow(10.0, cff);
        double vap_p = e_sat * M.Hair(i, j);
        double cl = M.cloud(i, j);
        M.srflx(i, j) = Rsolar * zenith * zenith * (1.0 - 0.6 * (cl * cl * cl)) /
                        ((zenith + 2.7) * vap_p * 1.0e-3 + 1.085 * zenith + 0.1);
      }
      M.srflx(i, j) = (1.0 - alb_w) * M.srflx(i, j);
    }
    exchange_r2d(M, T, M.srflx);
    // ana_winds.h:118-126
    for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) {
      double cff = 0.2 * (60.0 + M.latr(i, j));
      M.Uwind(i, j) = 15.0 * std::exp(-cff * cff); M.Vwind(i, j) = 0.0;
    }
    exchange_r2d(M, T, M.Uwind); exchange_r2d(M, T, M.Vwind);
    // ana_rain.h, ana_btflux.h, ana_stflux.h (salt), ana_pair.h
    for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) {
      M.rain(i, j) = 0.0; M.btflux(i, j, 1) = 0.0; M.stflux(i, j, 2) = 0.0; M.btflux(i, j, 2) = 0.0; M.Pair(i, j) = 1025.0;
    }
    exchange_r2d(M, T, M.rain); exchange_r2d(M, T, M.stflux.slab(2)); exchange_r2d(M, T, M.Pair);
  } else {
    // ana_stflux.h / ana_btflux.h: zero ; ana_smflux.h:306-325 (EWperiodic branch)
    for (int it = 1; it <= 2; ++it) {
      for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) { M.stflux(i, j, it) = 0.0; M.btflux(i, j, it) = 0.0; }
      exchange_r2d(M, T, M.stflux.slab(it));
    }
    double windamp;
    if ((M.tdays - c.dstart) <= 2.0) windamp = -0.1 * std::sin(pi * (M.tdays - c.dstart) / 4.0) / c.rho0;
    else windamp = -0.1 / c.rho0;
    for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrP; i <= T.IendT; ++i) M.sustr(i, j) = windamp;
    for (int j = T.JstrP; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) M.svstr(i, j) = 0.0;
    exchange_u2d(M, T, M.sustr); exchange_v2d(M, T, M.svstr);
  }
}

// Nonlinear/initial.F:130-873 (analytical start, nrrec=0)
void initial(Model& M) {
  M.iif = 1; M.indx1 = 1; M.kstp = 1; M.krhs = 1; M.knew = 1; M.PREDICTOR_2D_STEP = false;
  M.iic = 0; M.nstp = 1; M.nrhs = 1; M.nnew = 1;
  M.tdays = M.c.dstart; M.time = M.tdays * 86400.0;
  M.ntstart = (int)((M.time - M.c.dstart * 86400.0) / M.c.dt) + 1; M.ntfirst = M.ntstart;
  set_scoord(M); set_weights(M);                         // Utility/set_grid.F:112,124
  for (auto& T : M.tiles) ana_grid(M, T);                // set_grid.F:135
  for (auto& T : M.tiles) metrics(M, T);
  for (auto& T : M.tiles) ini_hmixcoef(M, T);
  for (auto& T : M.tiles) set_depth(M, T, M.Zt_avg1);
  for (auto& T : M.tiles) ana_initial(M, T);
  for (auto& T : M.tiles) { set_zeta_timeavg(M, T); set_depth(M, T, M.Zt_avg1); }
  for (auto& T : M.tiles) set_massflux(M, T);
  for (auto& T : M.tiles) { omega(M, T); rho_eos(M, T); }
  M.iic = M.ntstart;
}

// Persistent worker pool for the tile loops (the reference's OpenMP team, nl_roms.h:304-310):
// worker w takes tiles w, w+nt, ... of the current loop; the loop end is a barrier.
namespace {
struct Pool {
  std::vector<std::thread> th; std::mutex mu; std::condition_variable cv_go, cv_done;
  const std::function<void(const Tile&)>* fn = nullptr; const std::vector<Tile>* tiles = nullptr;
  long gen = 0; int pending = 0, nt = 0; bool stop = false;
  void worker(int w, long seen) {
    for (;;) {
      std::unique_lock<std::mutex> lk(mu);
      cv_go.wait(lk, [&] { return stop || gen != seen; });
      if (stop) return;
      seen = gen;
      auto f = fn; auto tl = tiles; const int n = nt;
      lk.unlock();
      for (size_t t = w; t < tl->size(); t += n) (*f)((*tl)[t]);
      lk.lock();
      if (--pending == 0) cv_done.notify_one();
    }
  }
  void ensure(int n) {
    if ((int)th.size() == n) return;
    shutdown();
    stop = false; nt = n;
    for (int w = 0; w < n; ++w) th.emplace_back(&Pool::worker, this, w, gen);
  }
  void run(const std::vector<Tile>& tl, const std::function<void(const Tile&)>& f) {
    std::unique_lock<std::mutex> lk(mu);
    fn = &f; tiles = &tl; pending = nt; ++gen;
    cv_go.notify_all();
    cv_done.wait(lk, [&] { return pending == 0; });
  }
  void shutdown() {
    { std::lock_guard<std::mutex> lk(mu); stop = true; }
    cv_go.notify_all();
    for (auto& t : th) t.join();
    th.clear();
  }
  ~Pool() { shutdown(); }
};
Pool g_pool;
}  // namespace
static void pool_run(Model& M, const std::function<void(const Tile&)>& f) {
  g_pool.ensure(std::min<int>(M.nthreads, (int)M.tiles.size()));
  g_pool.run(M.tiles, f);
}

// ---------------------------------------------------------------------------
// One baroclinic step, Nonlinear/main3d.F:216-1148, as named phases so that a
// test can stop between any two reference tile loops.
void main3d_phase(Model& M, const std::string& ph) {
  // Tile loops.  With nthreads>1 the tiles of one loop run concurrently and the
  // loop end is the barrier -- the reference's shared-memory (OpenMP) mode,
  // Drivers/nl_roms.h:304-310 + main3d.F `!$OMP BARRIER` between tile loops.
  auto par = [&](auto fn) {
    std::function<void(const Tile&)> f = fn;
    pool_run(M, f);
  };
  auto fwd = [&](auto fn) { if (M.nthreads > 1) { par(fn); return; } for (size_t t = 0; t < M.tiles.size(); ++t) fn(M.tiles[t]); };
  auto rev = [&](auto fn) { if (M.nthreads > 1) { par(fn); return; } for (size_t t = M.tiles.size(); t-- > 0;) fn(M.tiles[t]); };
  if (ph == "begin") {
    M.nstp = 1 + ((M.iic - M.ntstart) % 2); M.nnew = 3 - M.nstp; M.nrhs = M.nstp;   // main3d.F:222-224
    M.tdays = M.time / 86400.0;
    fwd([&](const Tile& T) { set_data(M, T); });
    if (M.iic == M.ntstart) {                         // post_initial.F:56-66
      fwd([&](const Tile& T) { ini_zeta(M, T); set_depth(M, T, M.Zt_avg1); });
      rev([&](const Tile& T) { ini_fields(M, T); });
    }
  } else if (ph == "set_massflux") fwd([&](const Tile& T) { set_massflux(M, T); });
  else if (ph == "rho_eos") fwd([&](const Tile& T) { rho_eos(M, T); });
  else if (ph == "diag") diag(M);
  else if (ph == "bulk_flux") { if (M.c.app == BENCHMARK) fwd([&](const Tile& T) { bulk_flux(M, T); }); }
  else if (ph == "set_vbc") fwd([&](const Tile& T) { set_vbc(M, T); });
  else if (ph == "vmix") rev([&](const Tile& T) { if (M.c.app == BENCHMARK) lmd_vmix(M, T); else ana_vmix(M, T); });
  else if (ph == "omega") rev([&](const Tile& T) { omega(M, T); });
  else if (ph == "wvelocity") rev([&](const Tile& T) { wvelocity(M, T, M.nstp); });   // main3d.F:535, same tile loop as omega
  else if (ph == "set_zeta") fwd([&](const Tile& T) { set_zeta(M, T); });
  else if (ph == "pre_step3d") rev([&](const Tile& T) { pre_step3d(M, T); });
  else if (ph == "prsgrd") rev([&](const Tile& T) { prsgrd32(M, T); });
  else if (ph == "t3dmix2") rev([&](const Tile& T) { t3dmix2(M, T); });
  else if (ph == "rhs3d_tile") rev([&](const Tile& T) { rhs3d_tile(M, T); });
  else if (ph == "uv3dmix2") rev([&](const Tile& T) { uv3dmix2(M, T); });
  else if (ph == "step2d_loop") {
    // main3d.F:810-918 (LF-AM3 fast loop)
    for (int my_iif = 1; my_iif <= M.nfast + 1; ++my_iif) {
      int next_indx1 = 3 - M.indx1;
      if (!M.PREDICTOR_2D_STEP && my_iif <= (M.nfast + 1)) {
        M.PREDICTOR_2D_STEP = true; M.iif = my_iif;
        if (M.iif == 1) M.kstp = M.indx1; else M.kstp = 3 - M.indx1;
        M.knew = 3; M.krhs = M.indx1;
      }
      if (my_iif <= (M.nfast + 1)) rev([&](const Tile& T) { step2d(M, T); });
      if (M.PREDICTOR_2D_STEP) {
        M.PREDICTOR_2D_STEP = false; M.knew = next_indx1; M.kstp = 3 - M.knew; M.krhs = 3;
        if (M.iif < (M.nfast + 1)) M.indx1 = next_indx1;
      }
      if (M.iif < (M.nfast + 1)) fwd([&](const Tile& T) { step2d(M, T); });
    }
  } else if (ph == "set_depth") rev([&](const Tile& T) { set_depth(M, T, M.Zt_avg1); });
  else if (ph == "step3d_uv") rev([&](const Tile& T) { step3d_uv(M, T); });
  else if (ph == "omega2") fwd([&](const Tile& T) { omega(M, T); });
  else if (ph == "step3d_t") rev([&](const Tile& T) { step3d_t(M, T); });
  else if (ph == "end") { M.iic = M.iic + 1; M.time = M.time + M.c.dt; }
  else throw std::runtime_error("unknown main3d phase " + ph);
}

// NOTE: rhs3d's sub-calls are run as whole-domain phases (all tiles finish
// pre_step3d before any starts prsgrd).  With a 1x1 tiling this is the
// reference order exactly; for NtileI*NtileJ>1 it is the order the reference's
// OpenMP barriers would impose if each sub-call were its own parallel region,
// and the results are tile-independent either way because each sub-call only
// writes its own interior.
static const char* kPhases[] = {"begin", "set_massflux", "rho_eos", "diag", "bulk_flux", "set_vbc", "vmix", "omega", "wvelocity",
    "set_zeta", "pre_step3d", "prsgrd", "t3dmix2", "rhs3d_tile", "uv3dmix2", "step2d_loop", "set_depth",
    "step3d_uv", "omega2", "step3d_t", "end"};
void main3d_step(Model& M) { for (const char* p : kPhases) main3d_phase(M, p); }

}  // namespace orc
