// oracle/physics.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h header).
// In-loop physics that main3d runs for the two option sets: equation of state,
// vertical boundary conditions, vertical mixing (analytical / KPP), COARE bulk
// fluxes, harmonic horizontal mixing and the diag reductions.
#include "oracle.h"
#include <algorithm>

namespace orc {

void lmd_swfrac(Model& M, const Tile& T, double Zscale, S2& Z, S2& swdk);   // kernels3d.cpp

static const double pi = 3.14159265358979323846;
static const double vonKar = 0.41;                    // mod_scalars.F:469
static const double Cp = 3985.0, rhow = 1000.0;       // mod_scalars.F:456,462
static const double StefBo = 5.67e-8, emmiss = 0.97;  // mod_scalars.F:460-461

// ---------------------------------------------------------------------------
// Nonlinear/rho_eos.F:111-570 (NONLIN_EOS) and :576-886 (linear)
// Modules/mod_eoscoef.F:24-64
static const double A00 = +1.909256e+04, A01 = +2.098925e+02, A02 = -3.041638e+00, A03 = -1.852732e-03, A04 = -1.361629e-05;
static const double B00 = +1.044077e+02, B01 = -6.500517e+00, B02 = +1.553190e-01, B03 = +2.326469e-04;
static const double D00 = -5.587545e+00, D01 = +7.390729e-01, D02 = -1.909078e-02;
static const double E00 = +4.721788e-01, E01 = +1.028859e-02, E02 = -2.512549e-04, E03 = -5.939910e-07;
static const double F00 = -1.571896e-02, F01 = -2.598241e-04, F02 = +7.267926e-06;
static const double G00 = +2.042967e-03, G01 = +1.045941e-05, G02 = -5.782165e-10, G03 = +1.296821e-07;
static const double H00 = -2.595994e-07, H01 = -1.248266e-09, H02 = -3.508914e-09;
static const double Q00 = +9.99842594e+02, Q01 = +6.793952e-02, Q02 = -9.095290e-03, Q03 = +1.001685e-04, Q04 = -1.120083e-06, Q05 = +6.536332e-09;
static const double U00 = +8.24493e-01, U01 = -4.08990e-03, U02 = +7.64380e-05, U03 = -8.24670e-07, U04 = +5.38750e-09;
static const double V00 = -5.72466e-03, V01 = +1.02270e-04, V02 = -1.65460e-06;
static const double W00 = +4.8314e-04;

void rho_eos(Model& M, const Tile& T) {
  const int N = M.N, nrhs = M.nrhs; const Config& c = M.c; const double g = c.g, rho0 = c.rho0;
  F3 &Hz = M.Hz, &z_r = M.z_r, &z_w = M.z_w, &rho = M.rho, &pden = M.pden; F5& t = M.t; F2 &rhoA = M.rhoA, &rhoS = M.rhoS;
  if (c.app == UPWELLING) {
    for (int j = T.JstrT; j <= T.JendT; ++j) {
      for (int k = 1; k <= N; ++k) for (int i = T.IstrT; i <= T.IendT; ++i) {
        rho(i, j, k) = c.R0 - c.R0 * c.Tcoef * (t(i, j, k, nrhs, 1) - c.T0);
        rho(i, j, k) = rho(i, j, k) + c.R0 * c.Scoef * (t(i, j, k, nrhs, 2) - c.S0);
        rho(i, j, k) = rho(i, j, k) - 1000.0;
        pden(i, j, k) = rho(i, j, k);
      }
      for (int i = T.IstrT; i <= T.IendT; ++i) { double cff1 = rho(i, j, N) * Hz(i, j, N); rhoS(i, j) = 0.5 * cff1 * Hz(i, j, N); rhoA(i, j) = cff1; }
      for (int k = N - 1; k >= 1; --k) for (int i = T.IstrT; i <= T.IendT; ++i) {
        double cff1 = rho(i, j, k) * Hz(i, j, k);
        rhoS(i, j) = rhoS(i, j) + Hz(i, j, k) * (rhoA(i, j) + 0.5 * cff1);
        rhoA(i, j) = rhoA(i, j) + cff1;
      }
      double cff2 = 1.0 / rho0;
      for (int i = T.IstrT; i <= T.IendT; ++i) {
        double cff1 = 1.0 / (z_w(i, j, N) - z_w(i, j, 0));
        rhoA(i, j) = cff2 * cff1 * rhoA(i, j);
        rhoS(i, j) = 2.0 * cff1 * cff1 * cff2 * rhoS(i, j);
      }
    }
    exchange_r3d(M, T, rho); exchange_r3d(M, T, pden); exchange_r2d(M, T, rhoA); exchange_r2d(M, T, rhoS);
    return;
  }
  S2 DbulkDS(T.IminS, T.ImaxS, 1, N), DbulkDT(T.IminS, T.ImaxS, 1, N), Dden1DS(T.IminS, T.ImaxS, 1, N), Dden1DT(T.IminS, T.ImaxS, 1, N);
  S2 Scof(T.IminS, T.ImaxS, 1, N), Tcof(T.IminS, T.ImaxS, 1, N), wrk(T.IminS, T.ImaxS, 1, N), bulk(T.IminS, T.ImaxS, 1, N);
  S2 bulk0(T.IminS, T.ImaxS, 1, N), bulk1(T.IminS, T.ImaxS, 1, N), bulk2(T.IminS, T.ImaxS, 1, N), den(T.IminS, T.ImaxS, 1, N), den1(T.IminS, T.ImaxS, 1, N);
  double C[10], dCdT[10];
  for (int j = T.JstrT; j <= T.JendT; ++j) {
    for (int k = 1; k <= N; ++k) for (int i = T.IstrT; i <= T.IendT; ++i) {
      double Tt = std::max(-2.0, t(i, j, k, nrhs, 1));
      double Ts = std::max(0.0, t(i, j, k, nrhs, 2));
      double sqrtTs = std::sqrt(Ts);
      double Tp = z_r(i, j, k);
      double Tpr10 = 0.1 * Tp;
      C[0] = Q00 + Tt * (Q01 + Tt * (Q02 + Tt * (Q03 + Tt * (Q04 + Tt * Q05))));
      C[1] = U00 + Tt * (U01 + Tt * (U02 + Tt * (U03 + Tt * U04)));
      C[2] = V00 + Tt * (V01 + Tt * V02);
      dCdT[0] = Q01 + Tt * (2.0 * Q02 + Tt * (3.0 * Q03 + Tt * (4.0 * Q04 + Tt * 5.0 * Q05)));
      dCdT[1] = U01 + Tt * (2.0 * U02 + Tt * (3.0 * U03 + Tt * 4.0 * U04));
      dCdT[2] = V01 + Tt * 2.0 * V02;
      den1(i, k) = C[0] + Ts * (C[1] + sqrtTs * C[2] + Ts * W00);
      Dden1DS(i, k) = C[1] + 1.5 * C[2] * sqrtTs + 2.0 * W00 * Ts;
      Dden1DT(i, k) = dCdT[0] + Ts * (dCdT[1] + sqrtTs * dCdT[2]);
      C[3] = A00 + Tt * (A01 + Tt * (A02 + Tt * (A03 + Tt * A04)));
      C[4] = B00 + Tt * (B01 + Tt * (B02 + Tt * B03));
      C[5] = D00 + Tt * (D01 + Tt * D02);
      C[6] = E00 + Tt * (E01 + Tt * (E02 + Tt * E03));
      C[7] = F00 + Tt * (F01 + Tt * F02);
      C[8] = G01 + Tt * (G02 + Tt * G03);
      C[9] = H00 + Tt * (H01 + Tt * H02);
      dCdT[3] = A01 + Tt * (2.0 * A02 + Tt * (3.0 * A03 + Tt * 4.0 * A04));
      dCdT[4] = B01 + Tt * (2.0 * B02 + Tt * 3.0 * B03);
      dCdT[5] = D01 + Tt * 2.0 * D02;
      dCdT[6] = E01 + Tt * (2.0 * E02 + Tt * 3.0 * E03);
      dCdT[7] = F01 + Tt * 2.0 * F02;
      dCdT[8] = G02 + Tt * 2.0 * G03;
      dCdT[9] = H01 + Tt * 2.0 * H02;
      bulk0(i, k) = C[3] + Ts * (C[4] + sqrtTs * C[5]);
      bulk1(i, k) = C[6] + Ts * (C[7] + sqrtTs * G00);
      bulk2(i, k) = C[8] + Ts * C[9];
      bulk(i, k) = bulk0(i, k) - Tp * (bulk1(i, k) - Tp * bulk2(i, k));
      DbulkDS(i, k) = C[4] + sqrtTs * 1.5 * C[5] - Tp * (C[7] + sqrtTs * 1.5 * G00 - Tp * C[9]);
      DbulkDT(i, k) = dCdT[3] + Ts * (dCdT[4] + sqrtTs * dCdT[5]) - Tp * (dCdT[6] + Ts * dCdT[7] - Tp * (dCdT[8] + Ts * dCdT[9]));
      double cff = 1.0 / (bulk(i, k) + Tpr10);
      den(i, k) = den1(i, k) * bulk(i, k) * cff;
      den(i, k) = den(i, k) - 1000.0;
    }
    for (int i = T.IstrT; i <= T.IendT; ++i) { double cff1 = den(i, N) * Hz(i, j, N); rhoS(i, j) = 0.5 * cff1 * Hz(i, j, N); rhoA(i, j) = cff1; }
    for (int k = N - 1; k >= 1; --k) for (int i = T.IstrT; i <= T.IendT; ++i) {
      double cff1 = den(i, k) * Hz(i, j, k);
      rhoS(i, j) = rhoS(i, j) + Hz(i, j, k) * (rhoA(i, j) + 0.5 * cff1);
      rhoA(i, j) = rhoA(i, j) + cff1;
    }
    double cff2 = 1.0 / rho0;
    for (int i = T.IstrT; i <= T.IendT; ++i) {
      double cff1 = 1.0 / (z_w(i, j, N) - z_w(i, j, 0));
      rhoA(i, j) = cff2 * cff1 * rhoA(i, j);
      rhoS(i, j) = 2.0 * cff1 * cff1 * cff2 * rhoS(i, j);
    }
    for (int k = 1; k <= N - 1; ++k) for (int i = T.IstrT; i <= T.IendT; ++i) {
      double bulk_up = bulk0(i, k + 1) - z_w(i, j, k) * (bulk1(i, k + 1) - bulk2(i, k + 1) * z_w(i, j, k));
      double bulk_dn = bulk0(i, k) - z_w(i, j, k) * (bulk1(i, k) - bulk2(i, k) * z_w(i, j, k));
      double cff1 = 1.0 / (bulk_up + 0.1 * z_w(i, j, k));
      double cff2b = 1.0 / (bulk_dn + 0.1 * z_w(i, j, k));
      double den_up = cff1 * (den1(i, k + 1) * bulk_up);
      double den_dn = cff2b * (den1(i, k) * bulk_dn);
      M.bvf(i, j, k) = -g * (den_up - den_dn) / (0.5 * (den_up + den_dn) * (z_r(i, j, k + 1) - z_r(i, j, k)));
    }
    for (int i = T.IstrT; i <= T.IendT; ++i) { M.bvf(i, j, 0) = 0.0; M.bvf(i, j, N) = 0.0; }
    {
      const int k = N;
      for (int i = T.IstrT; i <= T.IendT; ++i) {
        double Tpr10 = 0.1 * z_r(i, j, k);
        double cff = bulk(i, k) + Tpr10;
        double cff1 = Tpr10 * den1(i, k);
        double cff2c = bulk(i, k) * cff;
        wrk(i, k) = (den(i, k) + 1000.0) * cff * cff;
        Tcof(i, k) = -(DbulkDT(i, k) * cff1 + Dden1DT(i, k) * cff2c);
        Scof(i, k) = (DbulkDS(i, k) * cff1 + Dden1DS(i, k) * cff2c);
      }
      for (int i = T.IstrT; i <= T.IendT; ++i) { double cff = 1.0 / wrk(i, N); M.alpha(i, j) = cff * Tcof(i, N); M.beta(i, j) = cff * Scof(i, N); }
    }
    for (int k = 1; k <= N; ++k) for (int i = T.IstrT; i <= T.IendT; ++i) { rho(i, j, k) = den(i, k); pden(i, j, k) = (den1(i, k) - 1000.0); }
  }
  exchange_r3d(M, T, rho); exchange_r3d(M, T, pden); exchange_r2d(M, T, M.alpha); exchange_r2d(M, T, M.beta);
  exchange_r2d(M, T, rhoA); exchange_r2d(M, T, rhoS); exchange_w3d(M, T, M.bvf);
}

// ---------------------------------------------------------------------------
// Nonlinear/set_vbc.F:620-720 (UV_QDRAG for BENCHMARK, UV_LDRAG for UPWELLING)
void set_vbc(Model& M, const Tile& T) {
  const int N = M.N, nrhs = M.nrhs; F4 &u = M.u, &v = M.v; F5& t = M.t;
  for (int j = T.JstrR; j <= T.JendR; ++j) for (int i = T.IstrR; i <= T.IendR; ++i) {
    M.stflx(i, j, 1) = M.stflux(i, j, 1); M.btflx(i, j, 1) = M.btflux(i, j, 1);
  }
  for (int j = T.JstrR; j <= T.JendR; ++j) for (int i = T.IstrR; i <= T.IendR; ++i) {
    double EmP = M.stflux(i, j, 2);
    M.stflx(i, j, 2) = EmP * t(i, j, N, nrhs, 2);
    M.btflx(i, j, 2) = M.btflx(i, j, 2) * t(i, j, 1, nrhs, 2);
  }
  if (M.c.app == BENCHMARK) {
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) {
      double cff1 = 0.25 * (v(i, j, 1, nrhs) + v(i, j + 1, 1, nrhs) + v(i - 1, j, 1, nrhs) + v(i - 1, j + 1, 1, nrhs));
      double cff2 = std::sqrt(u(i, j, 1, nrhs) * u(i, j, 1, nrhs) + cff1 * cff1);
      M.bustr(i, j) = 0.5 * (M.rdrag2(i - 1, j) + M.rdrag2(i, j)) * u(i, j, 1, nrhs) * cff2;
    }
    for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
      double cff1 = 0.25 * (u(i, j, 1, nrhs) + u(i + 1, j, 1, nrhs) + u(i, j - 1, 1, nrhs) + u(i + 1, j - 1, 1, nrhs));
      double cff2 = std::sqrt(cff1 * cff1 + v(i, j, 1, nrhs) * v(i, j, 1, nrhs));
      M.bvstr(i, j) = 0.5 * (M.rdrag2(i, j - 1) + M.rdrag2(i, j)) * v(i, j, 1, nrhs) * cff2;
    }
  } else {
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i)
      M.bustr(i, j) = 0.5 * (M.rdrag(i - 1, j) + M.rdrag(i, j)) * u(i, j, 1, nrhs);
    for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i)
      M.bvstr(i, j) = 0.5 * (M.rdrag(i, j - 1) + M.rdrag(i, j)) * v(i, j, 1, nrhs);
  }
  bc_u2d(M, T, M.bustr); bc_v2d(M, T, M.bvstr);
}

// Functionals/ana_vmix.h:200-206,327-337 (UPWELLING)
void ana_vmix(Model& M, const Tile& T) {
  const int N = M.N;
  for (int k = 1; k <= N - 1; ++k) for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i)
    M.Akv(i, j, k) = 2.0e-03 + 8.0e-03 * std::exp(M.z_w(i, j, k) / 150.0);
  exchange_w3d(M, T, M.Akv);
  for (int k = 1; k <= N - 1; ++k) for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i) {
    M.Akt(i, j, k, 1) = M.c.Akt_bak[0]; M.Akt(i, j, k, 2) = M.c.Akt_bak[1];
  }
  for (int it = 1; it <= M.NAT; ++it) exchange_w3d(M, T, M.Akt.vol(it));
}

// ---------------------------------------------------------------------------
// KPP: Nonlinear/lmd_vmix.F:99-434 (interior), lmd_skpp.F (surface boundary
// layer), lmd_vmix.F:437-660 (lmd_finish).  Constants mod_scalars.F:1635-1712.
static const double lmd_Ri0 = 0.7, lmd_bvfcon = -2.0e-5, lmd_nu0c = 0.01, lmd_nu0m = 10.0e-4, lmd_nu0s = 10.0e-4;
static const double lmd_Cstar = 10.0, lmd_Cv = 1.25, lmd_Ric = 0.3, lmd_am = 1.257, lmd_as = -28.86, lmd_betaT = -0.2;
static const double lmd_cekman = 0.7, lmd_cmonob = 1.0, lmd_cm = 8.36, lmd_cs = 98.96, lmd_epsilon = 0.1, lmd_zetam = -0.2, lmd_zetas = -1.0;

static void lmd_vmix_tile(Model& M, const Tile& T) {
  const int N = M.N, nstp = M.nstp; const double eps = 1.0e-14;
  F3 &Hz = M.Hz, &rho = M.rho, &bvf = M.bvf, &Akv = M.Akv; F4 &u = M.u, &v = M.v, &Akt = M.Akt;
  S3 Rig(T.IminS, T.ImaxS, T.JminS, T.JmaxS, 0, N);
  S2 FC(T.IminS, T.ImaxS, 0, N), dR(T.IminS, T.ImaxS, 0, N), dU(T.IminS, T.ImaxS, 0, N), dV(T.IminS, T.ImaxS, 0, N);
  const int i0 = std::max(1, T.Istr - 1), i1 = std::min(T.Iend + 1, M.Lm);
  for (int j = std::max(1, T.Jstr - 1); j <= std::min(T.Jend + 1, M.Mm); ++j) {
    for (int i = i0; i <= i1; ++i) { FC(i, 0) = 0.0; dR(i, 0) = 0.0; dU(i, 0) = 0.0; dV(i, 0) = 0.0; }
    for (int k = 1; k <= N - 1; ++k) for (int i = i0; i <= i1; ++i) {
      double cff = 1.0 / (2.0 * Hz(i, j, k + 1) + Hz(i, j, k) * (2.0 - FC(i, k - 1)));
      FC(i, k) = cff * Hz(i, j, k + 1);
      dR(i, k) = cff * (6.0 * (rho(i, j, k + 1) - rho(i, j, k)) - Hz(i, j, k) * dR(i, k - 1));
      dU(i, k) = cff * (3.0 * (u(i, j, k + 1, nstp) - u(i, j, k, nstp) + u(i + 1, j, k + 1, nstp) - u(i + 1, j, k, nstp)) - Hz(i, j, k) * dU(i, k - 1));
      dV(i, k) = cff * (3.0 * (v(i, j, k + 1, nstp) - v(i, j, k, nstp) + v(i, j + 1, k + 1, nstp) - v(i, j + 1, k, nstp)) - Hz(i, j, k) * dV(i, k - 1));
    }
    for (int i = i0; i <= i1; ++i) { dR(i, N) = 0.0; dU(i, N) = 0.0; dV(i, N) = 0.0; }
    for (int k = N - 1; k >= 1; --k) for (int i = i0; i <= i1; ++i) {
      dR(i, k) = dR(i, k) - FC(i, k) * dR(i, k + 1);
      dU(i, k) = dU(i, k) - FC(i, k) * dU(i, k + 1);
      dV(i, k) = dV(i, k) - FC(i, k) * dV(i, k + 1);
    }
    for (int k = 1; k <= N - 1; ++k) for (int i = i0; i <= i1; ++i) {
      double shear2 = dU(i, k) * dU(i, k) + dV(i, k) * dV(i, k);
      Rig(i, j, k) = bvf(i, j, k) / (shear2 + eps);
    }
  }
  for (int k = 1; k <= N - 1; ++k) for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
    double cff = std::min(1.0, std::max(0.0, Rig(i, j, k)) / lmd_Ri0);
    double nu_sx = 1.0 - cff * cff;
    nu_sx = nu_sx * nu_sx * nu_sx;
    double shear2 = bvf(i, j, k) / (Rig(i, j, k) + eps);
    cff = shear2 * shear2 / (shear2 * shear2 + 16.0e-10);
    nu_sx = cff * nu_sx;
    cff = 1.0 / std::sqrt(std::max(bvf(i, j, k), 1.0e-7));
    double lmd_iwm = 1.0e-6 * cff, lmd_iws = 1.0e-7 * cff;
    Akv(i, j, k) = lmd_iwm + lmd_nu0m * nu_sx;
    Akt(i, j, k, 1) = lmd_iws + lmd_nu0s * nu_sx;
    Akt(i, j, k, 2) = Akt(i, j, k, 1);
  }
}

// velocity scales, lmd_skpp.F (three inlined copies of the same block)
static inline void lmd_wscale(double Ustar, double Ustar3, double zetahat, double zetapar, double& wm, double& ws) {
  const double r3 = 1.0 / 3.0;
  if (zetahat >= 0.0) { wm = vonKar * Ustar / (1.0 + 5.0 * zetapar); ws = wm; }
  else {
    if (zetapar > lmd_zetam) wm = vonKar * Ustar * std::pow(1.0 - 16.0 * zetapar, 0.25);
    else wm = vonKar * std::pow(lmd_am * Ustar3 - lmd_cm * zetahat, r3);
    if (zetapar > lmd_zetas) ws = vonKar * Ustar * std::pow(1.0 - 16.0 * zetapar, 0.5);
    else ws = vonKar * std::pow(lmd_as * Ustar3 - lmd_cs * zetahat, r3);
  }
}

static void lmd_skpp_tile(Model& M, const Tile& T) {
  const int N = M.N, nstp = M.nstp; const double g = M.c.g, gorho0 = M.c.g / M.c.rho0;
  const double eps = 1.0e-10, small = 1.0e-20;
  const double lmd_Cg = lmd_Cstar * vonKar * std::pow(lmd_cs * vonKar * lmd_epsilon, 1.0 / 3.0);   // mod_scalars.F:4592
  F3 &Hz = M.Hz, &z_w = M.z_w, &pden = M.pden, &bvf = M.bvf, &Akv = M.Akv; F4 &u = M.u, &v = M.v, &Akt = M.Akt, &ghats = M.ghats;
  F2 &hsbl = M.hsbl, &srflx = M.srflx, &sustr = M.sustr, &svstr = M.svstr, &alpha = M.alpha, &beta = M.beta;
  auto ksbl = [&](int i, int j) -> int& { return M.ksbl[(i - M.LBi) + (size_t)M.ni * (j - M.LBj)]; };
#define SS(x) S2 x(T.IminS, T.ImaxS, T.JminS, T.JmaxS)
  SS(Bo); SS(Bosol); SS(Bfsfc); SS(Gm1); SS(Gt1); SS(Gs1); SS(Ustar); SS(dGm1dS); SS(dGt1dS); SS(dGs1dS); SS(f1); SS(sl_dpth); SS(swdk); SS(wm); SS(ws); SS(zgrid);
#undef SS
  S3 Bflux(T.IminS, T.ImaxS, T.JminS, T.JmaxS, 0, N);
  S2 FC(T.IminS, T.ImaxS, 0, N), dR(T.IminS, T.ImaxS, 0, N), dU(T.IminS, T.ImaxS, 0, N), dV(T.IminS, T.ImaxS, 0, N);
  std::vector<double> Rref(T.ImaxS - T.IminS + 1), Uref(Rref.size()), Vref(Rref.size());
  const double Vtc = lmd_Cv * std::sqrt(-lmd_betaT) / (std::sqrt(lmd_cs * lmd_epsilon) * lmd_Ric * vonKar * vonKar);
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) sl_dpth(i, j) = lmd_epsilon * (z_w(i, j, N) - hsbl(i, j));
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
    double a = 0.5 * (sustr(i, j) + sustr(i + 1, j)), b = 0.5 * (svstr(i, j) + svstr(i, j + 1));
    Ustar(i, j) = std::sqrt(std::sqrt(a * a + b * b));
  }
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
    Bo(i, j) = g * (alpha(i, j) * (M.stflx(i, j, 1) - srflx(i, j)) - beta(i, j) * M.stflx(i, j, 2));
    Bosol(i, j) = g * alpha(i, j) * srflx(i, j);
  }
  for (int k = 0; k <= N; ++k) {
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) zgrid(i, j) = z_w(i, j, N) - z_w(i, j, k);
    lmd_swfrac(M, T, -1.0, zgrid, swdk);
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
      Bflux(i, j, k) = (Bo(i, j) + Bosol(i, j) * (1.0 - swdk(i, j)));
      double cff = 1.0 - (0.5 + std::copysign(0.5, Bflux(i, j, k)));
      ghats(i, j, k, 1) = -cff * (M.stflx(i, j, 1) - srflx(i, j) + srflx(i, j) * (1.0 - swdk(i, j)));
      ghats(i, j, k, 2) = cff * M.stflx(i, j, 2);
    }
  }
  for (int j = T.Jstr; j <= T.Jend; ++j) {
    for (int i = T.Istr; i <= T.Iend; ++i) { FC(i, 0) = 0.0; dR(i, 0) = 0.0; dU(i, 0) = 0.0; dV(i, 0) = 0.0; }
    for (int k = 1; k <= N - 1; ++k) for (int i = T.Istr; i <= T.Iend; ++i) {
      double cff = 1.0 / (2.0 * Hz(i, j, k + 1) + Hz(i, j, k) * (2.0 - FC(i, k - 1)));
      FC(i, k) = cff * Hz(i, j, k + 1);
      dR(i, k) = cff * (6.0 * (pden(i, j, k + 1) - pden(i, j, k)) - Hz(i, j, k) * dR(i, k - 1));
      dU(i, k) = cff * (3.0 * (u(i, j, k + 1, nstp) - u(i, j, k, nstp) + u(i + 1, j, k + 1, nstp) - u(i + 1, j, k, nstp)) - Hz(i, j, k) * dU(i, k - 1));
      dV(i, k) = cff * (3.0 * (v(i, j, k + 1, nstp) - v(i, j, k, nstp) + v(i, j + 1, k + 1, nstp) - v(i, j + 1, k, nstp)) - Hz(i, j, k) * dV(i, k - 1));
    }
    for (int i = T.Istr; i <= T.Iend; ++i) { dR(i, N) = 0.0; dU(i, N) = 0.0; dV(i, N) = 0.0; }
    for (int k = N - 1; k >= 1; --k) for (int i = T.Istr; i <= T.Iend; ++i) {
      dR(i, k) = dR(i, k) - FC(i, k) * dR(i, k + 1);
      dU(i, k) = dU(i, k) - FC(i, k) * dU(i, k + 1);
      dV(i, k) = dV(i, k) - FC(i, k) * dV(i, k + 1);
    }
    const double cff1 = 1.0 / 3.0, cff2 = 1.0 / 6.0;
    for (int i = T.Istr; i <= T.Iend; ++i) {
      Rref[i - T.IminS] = pden(i, j, N) + Hz(i, j, N) * (cff1 * dR(i, N) + cff2 * dR(i, N - 1));
      Uref[i - T.IminS] = 0.5 * (u(i, j, N, nstp) + u(i + 1, j, N, nstp)) + Hz(i, j, N) * (cff1 * dU(i, N) + cff2 * dU(i, N - 1));
      Vref[i - T.IminS] = 0.5 * (v(i, j, N, nstp) + v(i, j + 1, N, nstp)) + Hz(i, j, N) * (cff1 * dV(i, N) + cff2 * dV(i, N - 1));
    }
    for (int i = T.Istr; i <= T.Iend; ++i) {
      FC(i, N) = 0.0;
      for (int k = N; k >= 1; --k) {
        double depth = z_w(i, j, N) - z_w(i, j, k - 1);
        double sigma = (Bflux(i, j, k - 1) < 0.0) ? std::min(sl_dpth(i, j), depth) : depth;
        double Ustar3 = Ustar(i, j) * Ustar(i, j) * Ustar(i, j);
        double zetahat = vonKar * sigma * Bflux(i, j, k - 1);
        double zetapar = zetahat / (Ustar3 + small);
        lmd_wscale(Ustar(i, j), Ustar3, zetahat, zetapar, wm(i, j), ws(i, j));
        double Rk = pden(i, j, k) - Hz(i, j, k) * (cff1 * dR(i, k - 1) + cff2 * dR(i, k));
        double Uk = 0.5 * (u(i, j, k, nstp) + u(i + 1, j, k, nstp)) - Hz(i, j, k) * (cff1 * dU(i, k - 1) + cff2 * dU(i, k));
        double Vk = 0.5 * (v(i, j, k, nstp) + v(i, j + 1, k, nstp)) - Hz(i, j, k) * (cff1 * dV(i, k - 1) + cff2 * dV(i, k));
        double Ritop = -gorho0 * (Rref[i - T.IminS] - Rk) * depth;
        double dUr = Uref[i - T.IminS] - Uk, dVr = Vref[i - T.IminS] - Vk;
        double Ribot = dUr * dUr + dVr * dVr + Vtc * depth * ws(i, j) * std::sqrt(std::fabs(bvf(i, j, k - 1)));
        FC(i, k - 1) = Ritop - lmd_Ric * Ribot;
      }
    }
    for (int i = T.Istr; i <= T.Iend; ++i) { ksbl(i, j) = 1; hsbl(i, j) = z_w(i, j, 1); }
    for (int k = N; k >= 2; --k) for (int i = T.Istr; i <= T.Iend; ++i)
      if (ksbl(i, j) == 1 && FC(i, k - 1) > 0.0) {
        hsbl(i, j) = (z_w(i, j, k) * FC(i, k - 1) - z_w(i, j, k - 1) * FC(i, k)) / (FC(i, k - 1) - FC(i, k));
        ksbl(i, j) = k;
      }
  }
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) zgrid(i, j) = z_w(i, j, N) - hsbl(i, j);
  lmd_swfrac(M, T, -1.0, zgrid, swdk);
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) Bfsfc(i, j) = (Bo(i, j) + Bosol(i, j) * (1.0 - swdk(i, j)));
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
    if (Ustar(i, j) > 0.0 && Bfsfc(i, j) > 0.0) {
      double hekman = lmd_cekman * Ustar(i, j) / std::max(std::fabs(M.f(i, j)), eps);
      double hmonob = lmd_cmonob * Ustar(i, j) * Ustar(i, j) * Ustar(i, j) / std::max(vonKar * Bfsfc(i, j), eps);
      hsbl(i, j) = (z_w(i, j, N) - std::min(std::min(hekman, hmonob), z_w(i, j, N) - hsbl(i, j)));
    }
    hsbl(i, j) = std::min(hsbl(i, j), z_w(i, j, N));
    hsbl(i, j) = std::max(hsbl(i, j), z_w(i, j, 0));
  }
  bc_r2d(M, T, hsbl);
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
    ksbl(i, j) = 1;
    for (int k = N; k >= 2; --k) if (ksbl(i, j) == 1 && z_w(i, j, k - 1) < hsbl(i, j)) ksbl(i, j) = k;
  }
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) zgrid(i, j) = z_w(i, j, N) - hsbl(i, j);
  lmd_swfrac(M, T, -1.0, zgrid, swdk);
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) Bfsfc(i, j) = (Bo(i, j) + Bosol(i, j) * (1.0 - swdk(i, j)));
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
    sl_dpth(i, j) = lmd_epsilon * (z_w(i, j, N) - hsbl(i, j));
    double cff = (Bfsfc(i, j) > 0.0) ? 1.0 : lmd_epsilon;
    double sigma = cff * (z_w(i, j, N) - hsbl(i, j));
    double Ustar3 = Ustar(i, j) * Ustar(i, j) * Ustar(i, j);
    double zetahat = vonKar * sigma * Bfsfc(i, j);
    double zetapar = zetahat / (Ustar3 + small);
    lmd_wscale(Ustar(i, j), Ustar3, zetahat, zetapar, wm(i, j), ws(i, j));
  }
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i)
    f1(i, j) = 5.0 * std::max(0.0, Bfsfc(i, j)) * vonKar / (Ustar(i, j) * Ustar(i, j) * Ustar(i, j) * Ustar(i, j) + eps);
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
    double zbl = z_w(i, j, N) - hsbl(i, j);
    if (hsbl(i, j) > z_w(i, j, 1)) {
      int k = ksbl(i, j);
      double cff = 1.0 / (z_w(i, j, k) - z_w(i, j, k - 1));
      double cff_dn = cff * (hsbl(i, j) - z_w(i, j, k - 1));
      double cff_up = cff * (z_w(i, j, k) - hsbl(i, j));
      double K_bl = cff_dn * Akv(i, j, k) + cff_up * Akv(i, j, k - 1);
      double dK_bl = cff * (Akv(i, j, k) - Akv(i, j, k - 1));
      Gm1(i, j) = K_bl / (zbl * wm(i, j) + eps);
      dGm1dS(i, j) = std::min(0.0, -dK_bl / (wm(i, j) + eps) - K_bl * f1(i, j));
      K_bl = cff_dn * Akt(i, j, k, 1) + cff_up * Akt(i, j, k - 1, 1);
      dK_bl = cff * (Akt(i, j, k, 1) - Akt(i, j, k - 1, 1));
      Gt1(i, j) = K_bl / (zbl * ws(i, j) + eps);
      dGt1dS(i, j) = std::min(0.0, -dK_bl / (ws(i, j) + eps) - K_bl * f1(i, j));
      K_bl = cff_dn * Akt(i, j, k, 2) + cff_up * Akt(i, j, k - 1, 2);
      dK_bl = cff * (Akt(i, j, k, 2) - Akt(i, j, k - 1, 2));
      Gs1(i, j) = K_bl / (zbl * ws(i, j) + eps);
      dGs1dS(i, j) = std::min(0.0, -dK_bl / (ws(i, j) + eps) - K_bl * f1(i, j));
    } else {
      ksbl(i, j) = 0;
      double a = 0.5 * (M.bustr(i, j) + M.bustr(i + 1, j)), b = 0.5 * (M.bvstr(i, j) + M.bvstr(i, j + 1));
      double Ustarb = std::sqrt(std::sqrt(a * a + b * b));
      double dK_bl = vonKar * Ustarb;
      double K_bl = dK_bl * (hsbl(i, j) - z_w(i, j, 0));
      Gm1(i, j) = K_bl / (zbl * wm(i, j) + eps);
      dGm1dS(i, j) = std::min(0.0, -dK_bl / (wm(i, j) + eps) - K_bl * f1(i, j));
      Gt1(i, j) = K_bl / (zbl * ws(i, j) + eps);
      dGt1dS(i, j) = std::min(0.0, -dK_bl / (ws(i, j) + eps) - K_bl * f1(i, j));
      Gs1(i, j) = Gt1(i, j);
      dGs1dS(i, j) = dGt1dS(i, j);
    }
  }
  for (int k = 1; k <= N - 1; ++k) for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
    double zbl = z_w(i, j, N) - hsbl(i, j);
    if (k > ksbl(i, j)) {
      double depth = z_w(i, j, N) - z_w(i, j, k);
      double sigma = (Bflux(i, j, k) < 0.0) ? std::min(sl_dpth(i, j), depth) : depth;
      double Ustar3 = Ustar(i, j) * Ustar(i, j) * Ustar(i, j);
      double zetahat = vonKar * sigma * Bflux(i, j, k);
      double zetapar = zetahat / (Ustar3 + small);
      lmd_wscale(Ustar(i, j), Ustar3, zetahat, zetapar, wm(i, j), ws(i, j));
      sigma = depth / (zbl + eps);
      double a1 = sigma - 2.0, a2 = 3.0 - 2.0 * sigma, a3 = sigma - 1.0;
      double Gm = a1 + a2 * Gm1(i, j) + a3 * dGm1dS(i, j);
      double Gt = a1 + a2 * Gt1(i, j) + a3 * dGt1dS(i, j);
      double Gs = a1 + a2 * Gs1(i, j) + a3 * dGs1dS(i, j);
      Akv(i, j, k) = depth * wm(i, j) * (1.0 + sigma * Gm);
      Akt(i, j, k, 1) = depth * ws(i, j) * (1.0 + sigma * Gt);
      Akt(i, j, k, 2) = depth * ws(i, j) * (1.0 + sigma * Gs);
      double cff = lmd_Cg * (1.0 - (0.5 + std::copysign(0.5, Bflux(i, j, k)))) / (zbl * ws(i, j) + eps);
      ghats(i, j, k, 1) = cff * ghats(i, j, k, 1);
      ghats(i, j, k, 2) = cff * ghats(i, j, k, 2);
    } else {
      ghats(i, j, k, 1) = 0.0; ghats(i, j, k, 2) = 0.0;
    }
  }
}

static void lmd_finish_tile(Model& M, const Tile& T) {
  const int N = M.N; F3 &bvf = M.bvf, &Akv = M.Akv; F4& Akt = M.Akt;
  for (int k = 1; k <= N - 1; ++k) for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
    double cff = std::max(bvf(i, j, k), lmd_bvfcon);
    cff = std::min(1.0, (lmd_bvfcon - cff) / lmd_bvfcon);
    double nu_sxc = 1.0 - cff * cff;
    nu_sxc = nu_sxc * nu_sxc * nu_sxc;
    Akv(i, j, k) = Akv(i, j, k) + lmd_nu0c * nu_sxc;
    Akt(i, j, k, 1) = Akt(i, j, k, 1) + lmd_nu0c * nu_sxc;
    Akt(i, j, k, 2) = Akt(i, j, k, 2) + lmd_nu0c * nu_sxc;
  }
  // lmd_vmix.F:563-640: edge copies.  The reference does the W/E copies even on a periodic axis; in its
  // serial (1x1) and MPI builds they are then overwritten by the periodic exchange inside bc_w3d_tile, so
  // the state after lmd_finish holds the periodic images.  With several shared-memory tiles the outcome
  // would depend on the tile order (a west tile running after its east neighbour clobbers A(0,j)); that
  // artifact is not restated: on a periodic axis the W/E copies are skipped.
  const bool ew_copy = !M.EWperiodic;
  for (int k = 0; k <= N; ++k) {
    if (ew_copy && T.W) for (int j = T.Jstr; j <= T.Jend; ++j) { for (int it = 1; it <= M.NAT; ++it) Akt(T.Istr - 1, j, k, it) = Akt(T.Istr, j, k, it); Akv(T.Istr - 1, j, k) = Akv(T.Istr, j, k); }
    if (ew_copy && T.E) for (int j = T.Jstr; j <= T.Jend; ++j) { for (int it = 1; it <= M.NAT; ++it) Akt(T.Iend + 1, j, k, it) = Akt(T.Iend, j, k, it); Akv(T.Iend + 1, j, k) = Akv(T.Iend, j, k); }
    if (T.S) for (int i = T.Istr; i <= T.Iend; ++i) { for (int it = 1; it <= M.NAT; ++it) Akt(i, T.Jstr - 1, k, it) = Akt(i, T.Jstr, k, it); Akv(i, T.Jstr - 1, k) = Akv(i, T.Jstr, k); }
    if (T.N) for (int i = T.Istr; i <= T.Iend; ++i) { for (int it = 1; it <= M.NAT; ++it) Akt(i, T.Jend + 1, k, it) = Akt(i, T.Jend, k, it); Akv(i, T.Jend + 1, k) = Akv(i, T.Jend, k); }
    auto corner = [&](int ic, int jc, int ia, int ja, int ib, int jb) {
      for (int it = 1; it <= M.NAT; ++it) Akt(ic, jc, k, it) = 0.5 * (Akt(ia, ja, k, it) + Akt(ib, jb, k, it));
      Akv(ic, jc, k) = 0.5 * (Akv(ia, ja, k) + Akv(ib, jb, k));
    };
    if (ew_copy && T.S && T.W) corner(T.Istr - 1, T.Jstr - 1, T.Istr, T.Jstr - 1, T.Istr - 1, T.Jstr);
    if (ew_copy && T.S && T.E) corner(T.Iend + 1, T.Jstr - 1, T.Iend, T.Jstr - 1, T.Iend + 1, T.Jstr);
    if (ew_copy && T.N && T.W) corner(T.Istr - 1, T.Jend + 1, T.Istr, T.Jend + 1, T.Istr - 1, T.Jend);
    if (ew_copy && T.N && T.E) corner(T.Iend + 1, T.Jend + 1, T.Iend, T.Jend + 1, T.Iend + 1, T.Jend);
  }
  bc_w3d(M, T, Akv);
  for (int it = 1; it <= M.NAT; ++it) bc_w3d(M, T, Akt.vol(it));
}

void lmd_vmix(Model& M, const Tile& T) { lmd_vmix_tile(M, T); lmd_skpp_tile(M, T); lmd_finish_tile(M, T); }

// ---------------------------------------------------------------------------
// Nonlinear/bulk_flux.F (COARE 3.0; LONGWAVE; no COOL_SKIN, no EMINUSP)
static const double blk_Cpa = 1004.67, blk_Cpw = 4000.0, blk_Rgas = 287.1, blk_Zabl = 600.0, blk_beta = 1.2;   // mod_scalars.F:1496-1500
static double bulk_psiu(double ZoL) {
  const double r3 = 1.0 / 3.0;
  if (ZoL < 0.0) {
    double x = std::pow(1.0 - 15.0 * ZoL, 0.25);
    double psik = 2.0 * std::log(0.5 * (1.0 + x)) + std::log(0.5 * (1.0 + x * x)) - 2.0 * std::atan(x) + 0.5 * pi;
    double cff = std::sqrt(3.0);
    double y = std::pow(1.0 - 10.15 * ZoL, r3);
    double psic = 1.5 * std::log(r3 * (1.0 + y + y * y)) - cff * std::atan((1.0 + 2.0 * y) / cff) + pi / cff;
    cff = ZoL * ZoL;
    double Fw = cff / (1.0 + cff);
    return (1.0 - Fw) * psik + Fw * psic;
  }
  double cff = std::min(50.0, 0.35 * ZoL);
  return -((1.0 + ZoL) + 0.6667 * (ZoL - 14.28) / std::exp(cff) + 8.525);
}
static double bulk_psit(double ZoL) {
  const double r3 = 1.0 / 3.0;
  if (ZoL < 0.0) {
    double x = std::pow(1.0 - 15.0 * ZoL, 0.5);
    double psik = 2.0 * std::log(0.5 * (1.0 + x));
    double cff = std::sqrt(3.0);
    double y = std::pow(1.0 - 34.15 * ZoL, r3);
    double psic = 1.5 * std::log(r3 * (1.0 + y + y * y)) - cff * std::atan((1.0 + 2.0 * y) / cff) + pi / cff;
    cff = ZoL * ZoL;
    double Fw = cff / (1.0 + cff);
    return (1.0 - Fw) * psik + Fw * psic;
  }
  double cff = std::min(50.0, 0.35 * ZoL);
  return -(std::pow(1.0 + 2.0 * ZoL, 1.5) + 0.6667 * (ZoL - 14.28) / std::exp(cff) + 8.525);
}

void bulk_flux(Model& M, const Tile& T) {
  const Config& c = M.c; const int N = M.N, nrhs = M.nrhs; const double g = c.g, rho0 = c.rho0;
  const double eps = 1.0e-20, r3 = 1.0 / 3.0; const int IterMax = 3;
  const double ZW = c.blk_ZW, ZT = c.blk_ZT, ZQ = c.blk_ZQ;
#define SS(x) S2 x(T.IminS, T.ImaxS, T.JminS, T.JmaxS)
  SS(Hlv); SS(LHeat); SS(LRad); SS(SHeat); SS(Taux); SS(Tauy); SS(Uair); SS(Vair);
#undef SS
  for (int j = T.Jstr - 1; j <= T.Jend + 1; ++j) for (int i = T.Istr - 1; i <= T.Iend + 1; ++i) { Uair(i, j) = M.Uwind(i, j); Vair(i, j) = M.Vwind(i, j); }
  double Hscale = rho0 * Cp;
  for (int j = T.Jstr - 1; j <= T.JendR; ++j) for (int i = T.Istr - 1; i <= T.IendR; ++i) {
    double Wmag = std::sqrt(Uair(i, j) * Uair(i, j) + Vair(i, j) * Vair(i, j));
    double PairM = M.Pair(i, j);
    double TairC = M.Tair(i, j), TairK = TairC + 273.16;
    double TseaC = M.t(i, j, N, nrhs, 1), TseaK = TseaC + 273.16;
    double RH = M.Hair(i, j);
    double delTc = 0.0, delQc = 0.0;
    LHeat(i, j) = M.lhflx(i, j) * Hscale; SHeat(i, j) = M.shflx(i, j) * Hscale;
    Taux(i, j) = 0.0; Tauy(i, j) = 0.0;
    double cff = (0.7859 + 0.03477 * TairC) / (1.0 + 0.00412 * TairC);
    double e_sat = std::pow(10.0, cff);
    double vap_p = e_sat * RH;
    double cff2 = TairK * TairK * TairK;
    double cff1 = cff2 * TairK;
    LRad(i, j) = -emmiss * StefBo * (cff1 * (0.39 - 0.05 * std::sqrt(vap_p)) * (1.0 - 0.6823 * M.cloud(i, j) * M.cloud(i, j)) + cff2 * 4.0 * (TseaK - TairK));
    cff = (1.0007 + 3.46e-6 * PairM) * 6.1121 * std::exp(17.502 * TairC / (240.97 + TairC));
    double Qair = 0.62197 * (cff / (PairM - 0.378 * cff + eps));
    double Q;
    if (RH < 2.0) { cff = cff * RH; Q = 0.62197 * (cff / (PairM - 0.378 * cff + eps)); } else Q = RH / 1000.0;
    cff = (1.0007 + 3.46e-6 * PairM) * 6.1121 * std::exp(17.502 * TseaC / (240.97 + TseaC));
    cff = cff * 0.98;
    double Qsea = 0.62197 * (cff / (PairM - 0.378 * cff));
    double rhoAir = PairM * 100.0 / (blk_Rgas * TairK * (1.0 + 0.61 * Q));
    double VisAir = 1.326e-5 * (1.0 + TairC * (6.542e-3 + TairC * (8.301e-6 - 4.84e-9 * TairC)));
    Hlv(i, j) = (2.501 - 0.00237 * TseaC) * 1.0e+6;
    double Wgus = 0.5;
    double delW = std::sqrt(Wmag * Wmag + Wgus * Wgus);
    double delQ = Qsea - Q, delT = TseaC - TairC;
    double ZoW = 0.0001;
    double u10 = delW * std::log(10.0 / ZoW) / std::log(ZW / ZoW);
    double Wstar = 0.035 * u10;
    double Zo10 = 0.011 * Wstar * Wstar / g + 0.11 * VisAir / Wstar;
    double t0 = vonKar / std::log(10.0 / Zo10);
    double Cd10 = t0 * t0;
    double Ch10 = 0.00115;
    double Ct10 = Ch10 / std::sqrt(Cd10);
    double ZoT10 = 10.0 / std::exp(vonKar / Ct10);
    double t1 = vonKar / std::log(ZW / Zo10);
    double Cd = t1 * t1;
    double Ct = vonKar / std::log(ZT / ZoT10);
    double CC = vonKar * Ct / Cd;
    delTc = 0.0;
    double Ribcu = -ZW / (blk_Zabl * 0.004 * (blk_beta * blk_beta * blk_beta));
    double Ri = -g * ZW * ((delT - delTc) + 0.61 * TairK * delQ) / (TairK * delW * delW + eps);
    double Zetu = (Ri < 0.0) ? CC * Ri / (1.0 + Ri / Ribcu) : CC * Ri / (1.0 + 3.0 * Ri / CC);
    double L10 = ZW / Zetu;
    Wstar = delW * vonKar / (std::log(ZW / Zo10) - bulk_psiu(ZW / L10));
    double Tstar = -(delT - delTc) * vonKar / (std::log(ZT / ZoT10) - bulk_psit(ZT / L10));
    double Qstar = -(delQ - delQc) * vonKar / (std::log(ZQ / ZoT10) - bulk_psit(ZQ / L10));
    double charn = std::min(0.028, -0.005 + 0.0017 * delW);
    for (int Iter = 1; Iter <= IterMax; ++Iter) {
      ZoW = charn * Wstar * Wstar / g + 0.11 * VisAir / (Wstar + eps);
      double Rr = ZoW * Wstar / VisAir;
      double ZoQ = std::min(1.6e-4, 5.8e-5 / std::pow(Rr, 0.72));
      double ZoT = ZoQ;
      double ZoL = vonKar * g * ZW * (Tstar * (1.0 + 0.61 * Q) + 0.61 * TairK * Qstar) / (TairK * Wstar * Wstar * (1.0 + 0.61 * Q) + eps);
      double L = ZW / (ZoL + eps);
      double Wpsi = bulk_psiu(ZoL), Tpsi = bulk_psit(ZT / L), Qpsi = bulk_psit(ZQ / L);
      Wstar = std::max(eps, delW * vonKar / (std::log(ZW / ZoW) - Wpsi));
      Tstar = -(delT - delTc) * vonKar / (std::log(ZT / ZoT) - Tpsi);
      Qstar = -(delQ - delQc) * vonKar / (std::log(ZQ / ZoQ) - Qpsi);
      double Bf = -g / TairK * Wstar * (Tstar + 0.61 * TairK * Qstar);
      if (Bf > 0.0) Wgus = blk_beta * std::pow(Bf * blk_Zabl, r3); else Wgus = 0.2;
      delW = std::sqrt(Wmag * Wmag + Wgus * Wgus);
    }
    double Hs = -blk_Cpa * rhoAir * Wstar * Tstar;
    double diffw = 2.11e-5 * std::pow(TairK / 273.16, 1.94);
    double diffh = 0.02411 * (1.0 + TairC * (3.309e-3 - 1.44e-6 * TairC)) / (rhoAir * blk_Cpa + eps);
    cff = Qair * Hlv(i, j) / (blk_Rgas * TairK * TairK);
    double wet_bulb = 1.0 / (1.0 + 0.622 * (cff * Hlv(i, j) * diffw) / (blk_Cpa * diffh));
    double Hsr = std::fabs(M.rain(i, j)) * wet_bulb * blk_Cpw * ((TseaC - TairC) + (Qsea - Q) * Hlv(i, j) / blk_Cpa);
    SHeat(i, j) = (Hs + Hsr);
    double Hl = -Hlv(i, j) * rhoAir * Wstar * Qstar;
    double upvel = -1.61 * Wstar * Qstar - (1.0 + 1.61 * Q) * Wstar * Tstar / TairK;
    double Hlw = rhoAir * Hlv(i, j) * upvel * Q;
    LHeat(i, j) = (Hl + Hlw);
    double Taur = 0.85 * std::fabs(M.rain(i, j)) * Wmag;
    cff = rhoAir * (Wstar * Wstar + Taur / rhoAir) / (Wmag + eps);
    Taux(i, j) = cff * Uair(i, j); Tauy(i, j) = cff * Vair(i, j);
  }
  Hscale = 1.0 / (rho0 * Cp);
  for (int j = T.JstrR; j <= T.JendR; ++j) for (int i = T.IstrR; i <= T.IendR; ++i) {
    M.lrflx(i, j) = LRad(i, j) * Hscale; M.lhflx(i, j) = -LHeat(i, j) * Hscale; M.shflx(i, j) = -SHeat(i, j) * Hscale;
    M.stflux(i, j, 1) = (M.srflx(i, j) + M.lrflx(i, j) + M.lhflx(i, j) + M.shflx(i, j));
  }
  double cff = 0.5 / rho0;
  for (int j = T.JstrR; j <= T.JendR; ++j) for (int i = T.Istr; i <= T.IendR; ++i) M.sustr(i, j) = cff * (Taux(i - 1, j) + Taux(i, j));
  for (int j = T.Jstr; j <= T.JendR; ++j) for (int i = T.IstrR; i <= T.IendR; ++i) M.svstr(i, j) = cff * (Tauy(i, j - 1) + Tauy(i, j));
  exchange_r2d(M, T, M.lrflx); exchange_r2d(M, T, M.lhflx); exchange_r2d(M, T, M.shflx); exchange_r2d(M, T, M.stflux.slab(1));
  exchange_u2d(M, T, M.sustr); exchange_v2d(M, T, M.svstr);
  (void)rhow;
}

// ---------------------------------------------------------------------------
// Nonlinear/t3dmix2_s.h:198-301 (UPWELLING) and t3dmix2_geo.h:219-419 (BENCHMARK)
void t3dmix2(Model& M, const Tile& T) {
  const int N = M.N, nrhs = M.nrhs, nnew = M.nnew; const double dt = M.c.dt;
  F3 &Hz = M.Hz, &z_r = M.z_r, &diff2 = M.diff2; F2 &pm = M.pm, &pn = M.pn; F5& t = M.t;
  S2 FE(T.IminS, T.ImaxS, T.JminS, T.JmaxS), FX(T.IminS, T.ImaxS, T.JminS, T.JmaxS);
  if (M.c.app == UPWELLING) {
    for (int itrc = 1; itrc <= M.NT; ++itrc) for (int k = 1; k <= N; ++k) {
      for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend + 1; ++i) {
        double cff = 0.25 * (diff2(i, j, itrc) + diff2(i - 1, j, itrc)) * M.pmon_u(i, j);
        FX(i, j) = cff * (Hz(i, j, k) + Hz(i - 1, j, k)) * (t(i, j, k, nrhs, itrc) - t(i - 1, j, k, nrhs, itrc));
      }
      for (int j = T.Jstr; j <= T.Jend + 1; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
        double cff = 0.25 * (diff2(i, j, itrc) + diff2(i, j - 1, itrc)) * M.pnom_v(i, j);
        FE(i, j) = cff * (Hz(i, j, k) + Hz(i, j - 1, k)) * (t(i, j, k, nrhs, itrc) - t(i, j - 1, k, nrhs, itrc));
      }
      for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
        double cff = dt * pm(i, j) * pn(i, j);
        double cff1 = cff * (FX(i + 1, j) - FX(i, j)), cff2 = cff * (FE(i, j + 1) - FE(i, j)), cff3 = cff1 + cff2;
        t(i, j, k, nnew, itrc) = t(i, j, k, nnew, itrc) + cff3;
      }
    }
    return;
  }
  S3 FS(T.IminS, T.ImaxS, T.JminS, T.JmaxS, 1, 2), dTdz(T.IminS, T.ImaxS, T.JminS, T.JmaxS, 1, 2), dTdx(T.IminS, T.ImaxS, T.JminS, T.JmaxS, 1, 2);
  S3 dTde(T.IminS, T.ImaxS, T.JminS, T.JmaxS, 1, 2), dZdx(T.IminS, T.ImaxS, T.JminS, T.JmaxS, 1, 2), dZde(T.IminS, T.ImaxS, T.JminS, T.JmaxS, 1, 2);
  for (int itrc = 1; itrc <= M.NT; ++itrc) {
    int k2 = 1, k1;
    for (int k = 0; k <= N; ++k) {
      k1 = k2; k2 = 3 - k1;
      if (k < N) {
        for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend + 1; ++i) {
          double cff = 0.5 * (pm(i, j) + pm(i - 1, j));
          dZdx(i, j, k2) = cff * (z_r(i, j, k + 1) - z_r(i - 1, j, k + 1));
          dTdx(i, j, k2) = cff * (t(i, j, k + 1, nrhs, itrc) - t(i - 1, j, k + 1, nrhs, itrc));
        }
        for (int j = T.Jstr; j <= T.Jend + 1; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
          double cff = 0.5 * (pn(i, j) + pn(i, j - 1));
          dZde(i, j, k2) = cff * (z_r(i, j, k + 1) - z_r(i, j - 1, k + 1));
          dTde(i, j, k2) = cff * (t(i, j, k + 1, nrhs, itrc) - t(i, j - 1, k + 1, nrhs, itrc));
        }
      }
      if (k == 0 || k == N) {
        for (int j = T.Jstr - 1; j <= T.Jend + 1; ++j) for (int i = T.Istr - 1; i <= T.Iend + 1; ++i) { dTdz(i, j, k2) = 0.0; FS(i, j, k2) = 0.0; }
      } else {
        for (int j = T.Jstr - 1; j <= T.Jend + 1; ++j) for (int i = T.Istr - 1; i <= T.Iend + 1; ++i) {
          double cff = 1.0 / (z_r(i, j, k + 1) - z_r(i, j, k));
          dTdz(i, j, k2) = cff * (t(i, j, k + 1, nrhs, itrc) - t(i, j, k, nrhs, itrc));
        }
      }
      if (k > 0) {
        for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend + 1; ++i) {
          double cff = 0.25 * (diff2(i, j, itrc) + diff2(i - 1, j, itrc)) * M.on_u(i, j);
          FX(i, j) = cff * (Hz(i, j, k) + Hz(i - 1, j, k)) *
                     (dTdx(i, j, k1) - 0.5 * (std::min(dZdx(i, j, k1), 0.0) * (dTdz(i - 1, j, k1) + dTdz(i, j, k2)) +
                                              std::max(dZdx(i, j, k1), 0.0) * (dTdz(i - 1, j, k2) + dTdz(i, j, k1))));
        }
        for (int j = T.Jstr; j <= T.Jend + 1; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
          double cff = 0.25 * (diff2(i, j, itrc) + diff2(i, j - 1, itrc)) * M.om_v(i, j);
          FE(i, j) = cff * (Hz(i, j, k) + Hz(i, j - 1, k)) *
                     (dTde(i, j, k1) - 0.5 * (std::min(dZde(i, j, k1), 0.0) * (dTdz(i, j - 1, k1) + dTdz(i, j, k2)) +
                                              std::max(dZde(i, j, k1), 0.0) * (dTdz(i, j - 1, k2) + dTdz(i, j, k1))));
        }
        if (k < N) {
          for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
            double cff = 0.5 * diff2(i, j, itrc);
            double cff1 = std::min(dZdx(i, j, k1), 0.0), cff2 = std::min(dZdx(i + 1, j, k2), 0.0);
            double cff3 = std::max(dZdx(i, j, k2), 0.0), cff4 = std::max(dZdx(i + 1, j, k1), 0.0);
            FS(i, j, k2) = cff * (cff1 * (cff1 * dTdz(i, j, k2) - dTdx(i, j, k1)) + cff2 * (cff2 * dTdz(i, j, k2) - dTdx(i + 1, j, k2)) +
                                  cff3 * (cff3 * dTdz(i, j, k2) - dTdx(i, j, k2)) + cff4 * (cff4 * dTdz(i, j, k2) - dTdx(i + 1, j, k1)));
            cff1 = std::min(dZde(i, j, k1), 0.0); cff2 = std::min(dZde(i, j + 1, k2), 0.0);
            cff3 = std::max(dZde(i, j, k2), 0.0); cff4 = std::max(dZde(i, j + 1, k1), 0.0);
            FS(i, j, k2) = FS(i, j, k2) + cff * (cff1 * (cff1 * dTdz(i, j, k2) - dTde(i, j, k1)) + cff2 * (cff2 * dTdz(i, j, k2) - dTde(i, j + 1, k2)) +
                                                 cff3 * (cff3 * dTdz(i, j, k2) - dTde(i, j, k2)) + cff4 * (cff4 * dTdz(i, j, k2) - dTde(i, j + 1, k1)));
          }
        }
        for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
          double cff = dt * pm(i, j) * pn(i, j);
          double cff1 = cff * (FX(i + 1, j) - FX(i, j)), cff2 = cff * (FE(i, j + 1) - FE(i, j));
          double cff3 = dt * (FS(i, j, k2) - FS(i, j, k1));
          double cff4 = cff1 + cff2 + cff3;
          t(i, j, k, nnew, itrc) = t(i, j, k, nnew, itrc) + cff4;
        }
      }
    }
  }
}

// Nonlinear/uv3dmix2_s.h:239-330
void uv3dmix2(Model& M, const Tile& T) {
  const int N = M.N, nrhs = M.nrhs, nnew = M.nnew; const double dt = M.c.dt;
  F3& Hz = M.Hz; F2 &pm = M.pm, &pn = M.pn; F4 &u = M.u, &v = M.v;
  S2 UFe(T.IminS, T.ImaxS, T.JminS, T.JmaxS), VFe(T.IminS, T.ImaxS, T.JminS, T.JmaxS), UFx(T.IminS, T.ImaxS, T.JminS, T.JmaxS), VFx(T.IminS, T.ImaxS, T.JminS, T.JmaxS);
  for (int k = 1; k <= N; ++k) {
    for (int j = T.JstrV - 1; j <= T.Jend; ++j) for (int i = T.IstrU - 1; i <= T.Iend; ++i) {
      double cff = Hz(i, j, k) * 0.5 *
                   (M.pmon_r(i, j) * ((pn(i, j) + pn(i + 1, j)) * u(i + 1, j, k, nrhs) - (pn(i - 1, j) + pn(i, j)) * u(i, j, k, nrhs)) -
                    M.pnom_r(i, j) * ((pm(i, j) + pm(i, j + 1)) * v(i, j + 1, k, nrhs) - (pm(i, j - 1) + pm(i, j)) * v(i, j, k, nrhs)));
      UFx(i, j) = M.on_r(i, j) * M.on_r(i, j) * M.visc2_r(i, j) * cff;
      VFe(i, j) = M.om_r(i, j) * M.om_r(i, j) * M.visc2_r(i, j) * cff;
    }
    for (int j = T.Jstr; j <= T.Jend + 1; ++j) for (int i = T.Istr; i <= T.Iend + 1; ++i) {
      double cff = 0.125 * (Hz(i - 1, j, k) + Hz(i, j, k) + Hz(i - 1, j - 1, k) + Hz(i, j - 1, k)) *
                   (M.pmon_p(i, j) * ((pn(i, j - 1) + pn(i, j)) * v(i, j, k, nrhs) - (pn(i - 1, j - 1) + pn(i - 1, j)) * v(i - 1, j, k, nrhs)) +
                    M.pnom_p(i, j) * ((pm(i - 1, j) + pm(i, j)) * u(i, j, k, nrhs) - (pm(i - 1, j - 1) + pm(i, j - 1)) * u(i, j - 1, k, nrhs)));
      UFe(i, j) = M.om_p(i, j) * M.om_p(i, j) * M.visc2_p(i, j) * cff;
      VFx(i, j) = M.on_p(i, j) * M.on_p(i, j) * M.visc2_p(i, j) * cff;
    }
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) {
      double cff = dt * 0.25 * (pm(i - 1, j) + pm(i, j)) * (pn(i - 1, j) + pn(i, j));
      double cff1 = 0.5 * (pn(i - 1, j) + pn(i, j)) * (UFx(i, j) - UFx(i - 1, j));
      double cff2 = 0.5 * (pm(i - 1, j) + pm(i, j)) * (UFe(i, j + 1) - UFe(i, j));
      double cff3 = cff * (cff1 + cff2);
      M.rufrc(i, j) = M.rufrc(i, j) + cff1 + cff2;
      u(i, j, k, nnew) = u(i, j, k, nnew) + cff3;
    }
    for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
      double cff = dt * 0.25 * (pm(i, j) + pm(i, j - 1)) * (pn(i, j) + pn(i, j - 1));
      double cff1 = 0.5 * (pn(i, j - 1) + pn(i, j)) * (VFx(i + 1, j) - VFx(i, j));
      double cff2 = 0.5 * (pm(i, j - 1) + pm(i, j)) * (VFe(i, j) - VFe(i, j - 1));
      double cff3 = cff * (cff1 - cff2);
      M.rvfrc(i, j) = M.rvfrc(i, j) + cff1 - cff2;
      v(i, j, k, nnew) = v(i, j, k, nnew) + cff3;
    }
  }
}

void rhs3d(Model& M, const Tile& T) { pre_step3d(M, T); prsgrd32(M, T); t3dmix2(M, T); rhs3d_tile(M, T); uv3dmix2(M, T); }

// Nonlinear/diag.F:209-411,512-542: volume-integrated kinetic / potential energy, volume, the maximum Courant number
// with its location, maximum speed and density anomaly, blow-up test.  Per tile the horizontal sums are taken in the
// reference's two stages (j collapsed per i, then along i, :296-322); tiles are combined in tile order (:363-384).
void diag(Model& M) {
  const int N = M.N, idia = M.nstp; const double g = M.c.g, dt = M.c.dt;
  const double spval = 1.0e37, Large = 1.0e35;     // mod_scalars.F
  double volume = 0.0, avgke = 0.0, avgpe = 0.0, maxspeed = -Large, maxrho = -Large;
  double max_C = 0.0, max_Cu = 0.0, max_Cv = 0.0, max_Cw = 0.0; int max_Ci = 0, max_Cj = 0, max_Ck = 0;
  for (const Tile& T : M.tiles) {
    S2 ke2d(T.IminS, T.ImaxS, T.JminS, T.JmaxS), pe2d(T.IminS, T.ImaxS, T.JminS, T.JmaxS);
    double my_max_C = 0.0, my_max_Cu = 0.0, my_max_Cv = 0.0, my_max_Cw = 0.0; int my_max_Ci = 0, my_max_Cj = 0, my_max_Ck = 0;
    double my_maxspeed = 0.0, my_maxrho = -spval;
    for (int j = T.Jstr; j <= T.Jend; ++j) {
      for (int i = T.Istr; i <= T.Iend; ++i) {
        ke2d(i, j) = 0.0;
        pe2d(i, j) = 0.5 * g * M.z_w(i, j, N) * M.z_w(i, j, N);
      }
      const double cff = g / M.c.rho0;
      for (int k = N; k >= 1; --k) for (int i = T.Istr; i <= T.Iend; ++i) {
        const double u2v2 = M.u(i, j, k, idia) * M.u(i, j, k, idia) + M.u(i + 1, j, k, idia) * M.u(i + 1, j, k, idia) +
                            M.v(i, j, k, idia) * M.v(i, j, k, idia) + M.v(i, j + 1, k, idia) * M.v(i, j + 1, k, idia);
        ke2d(i, j) = ke2d(i, j) + M.Hz(i, j, k) * 0.25 * u2v2;
        pe2d(i, j) = pe2d(i, j) + cff * M.Hz(i, j, k) * (M.rho(i, j, k) + 1000.0) * (M.z_r(i, j, k) - M.z_w(i, j, 0));
        const double my_Cu = 0.5 * std::fabs(M.u(i, j, k, idia) + M.u(i + 1, j, k, idia)) * dt * M.pm(i, j);
        const double my_Cv = 0.5 * std::fabs(M.v(i, j, k, idia) + M.v(i, j + 1, k, idia)) * dt * M.pn(i, j);
        const double my_Cw = 0.5 * std::fabs(M.wvel(i, j, k - 1) + M.wvel(i, j, k)) * dt / M.Hz(i, j, k);
        const double my_C = my_Cu + my_Cv + my_Cw;
        if (my_C > my_max_C) {
          my_max_C = my_C; my_max_Cu = my_Cu; my_max_Cv = my_Cv; my_max_Cw = my_Cw; my_max_Ci = i; my_max_Cj = j; my_max_Ck = k;
        }
        my_maxspeed = std::max(my_maxspeed, std::sqrt(0.5 * u2v2));
        my_maxrho = std::max(my_maxrho, M.rho(i, j, k));
      }
    }
    for (int i = T.Istr; i <= T.Iend; ++i) { pe2d(i, T.Jend + 1) = 0.0; pe2d(i, T.Jstr - 1) = 0.0; ke2d(i, T.Jstr - 1) = 0.0; }
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
      pe2d(i, T.Jend + 1) = pe2d(i, T.Jend + 1) + M.omn(i, j) * (M.z_w(i, j, N) - M.z_w(i, j, 0));
      pe2d(i, T.Jstr - 1) = pe2d(i, T.Jstr - 1) + M.omn(i, j) * pe2d(i, j);
      ke2d(i, T.Jstr - 1) = ke2d(i, T.Jstr - 1) + M.omn(i, j) * ke2d(i, j);
    }
    double my_volume = 0.0, my_avgpe = 0.0, my_avgke = 0.0;
    for (int i = T.Istr; i <= T.Iend; ++i) {
      my_volume = my_volume + pe2d(i, T.Jend + 1);
      my_avgpe = my_avgpe + pe2d(i, T.Jstr - 1);
      my_avgke = my_avgke + ke2d(i, T.Jstr - 1);
    }
    // global combination, diag.F:363-384
    volume = volume + my_volume; avgke = avgke + my_avgke; avgpe = avgpe + my_avgpe;
    maxspeed = std::max(maxspeed, my_maxspeed); maxrho = std::max(maxrho, my_maxrho);
    if (my_max_C == max_C) {
      max_Ci = std::min(max_Ci, my_max_Ci); max_Cj = std::min(max_Cj, my_max_Cj); max_Ck = std::min(max_Ck, my_max_Ck);
    } else if (my_max_C > max_C) {
      max_C = my_max_C; max_Cu = my_max_Cu; max_Cv = my_max_Cv; max_Cw = my_max_Cw;
      max_Ci = my_max_Ci; max_Cj = my_max_Cj; max_Ck = my_max_Ck;
    }
  }
  M.volume = volume; M.avgke = avgke / volume; M.avgpe = avgpe / volume;
  M.max_C = max_C; M.max_Cu = max_Cu; M.max_Cv = max_Cv; M.max_Cw = max_Cw; M.max_Ci = max_Ci; M.max_Cj = max_Cj; M.max_Ck = max_Ck;
  M.maxspeed = maxspeed; M.maxrho = maxrho;
  // diag.F:512-542: the reference tests the printed (1pe8.1) energies for NaN/Inf/overflow characters; restated as a
  // finiteness test.  max_speed = 20 m/s, max_rho = 200 kg/m3 (mod_scalars.F:573-574).
  if (!std::isfinite(M.avgke) || !std::isfinite(M.avgpe) || maxspeed > 20.0 || maxrho > 200.0) M.exit_flag = 1;
}

}  // namespace orc
