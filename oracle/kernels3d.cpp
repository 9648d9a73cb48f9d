// oracle/kernels3d.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h header).
// Restatement of the 3-D kernels on the main3d path.  Active advection
// options: U3 horizontal / C4 vertical for tracers (roms_*.in:133-137),
// third-order upstream horizontal / fourth-order centred vertical for momentum.
#include "oracle.h"
#include <algorithm>

namespace orc {

// Nonlinear/set_depth.F:192-245 (Vtransform=2) + exchanges :248-262
void set_depth(Model& M, const Tile& T, F2 Zt_avg1) {
  const int N = M.N; const double hc = M.hc;
  F2& h = M.h; F3 &z_w = M.z_w, &z_r = M.z_r, &Hz = M.Hz;
  for (int j = T.JstrT; j <= T.JendT; ++j) {
    for (int i = T.IstrT; i <= T.IendT; ++i) z_w(i, j, 0) = -h(i, j);
    for (int k = 1; k <= N; ++k) {
      double cff_r = hc * M.sc_r[k], cff_w = hc * M.sc_w[k];
      double cff1_r = M.Cs_r[k], cff1_w = M.Cs_w[k];
      for (int i = T.IstrT; i <= T.IendT; ++i) {
        double hwater = h(i, j);
        double hinv = 1.0 / (hc + hwater);
        double cff2_r = (cff_r + cff1_r * hwater) * hinv;
        double cff2_w = (cff_w + cff1_w * hwater) * hinv;
        z_w(i, j, k) = Zt_avg1(i, j) + (Zt_avg1(i, j) + hwater) * cff2_w;
        z_r(i, j, k) = Zt_avg1(i, j) + (Zt_avg1(i, j) + hwater) * cff2_r;
        Hz(i, j, k) = z_w(i, j, k) - z_w(i, j, k - 1);
      }
    }
  }
  exchange_r2d(M, T, h); exchange_w3d(M, T, z_w); exchange_r3d(M, T, z_r); exchange_r3d(M, T, Hz);
}

// Nonlinear/set_massflux.F:140-177
void set_massflux(Model& M, const Tile& T) {
  const int N = M.N, nrhs = M.nrhs;
  for (int k = 1; k <= N; ++k) {
    for (int j = T.JstrT; j <= T.JendT; ++j) for (int i = T.IstrP; i <= T.IendT; ++i)
      M.Huon(i, j, k) = 0.5 * (M.Hz(i, j, k) + M.Hz(i - 1, j, k)) * M.u(i, j, k, nrhs) * M.on_u(i, j);
    for (int j = T.JstrP; j <= T.JendT; ++j) for (int i = T.IstrT; i <= T.IendT; ++i)
      M.Hvom(i, j, k) = 0.5 * (M.Hz(i, j, k) + M.Hz(i, j - 1, k)) * M.v(i, j, k, nrhs) * M.om_v(i, j);
  }
  exchange_u3d(M, T, M.Huon); exchange_v3d(M, T, M.Hvom);
}

// Nonlinear/omega.F:215-355
void omega(Model& M, const Tile& T) {
  const int N = M.N; F3 &W = M.W, &Huon = M.Huon, &Hvom = M.Hvom, &z_w = M.z_w;
  std::vector<double> wrk(T.ImaxS - T.IminS + 1);
  for (int j = T.Jstr; j <= T.Jend; ++j) {
    for (int i = T.Istr; i <= T.Iend; ++i) W(i, j, 0) = 0.0;
    for (int k = 1; k <= N; ++k) for (int i = T.Istr; i <= T.Iend; ++i)
      W(i, j, k) = W(i, j, k - 1) - (Huon(i + 1, j, k) - Huon(i, j, k) + Hvom(i, j + 1, k) - Hvom(i, j, k));
    for (int i = T.Istr; i <= T.Iend; ++i) wrk[i - T.IminS] = W(i, j, N) / (z_w(i, j, N) - z_w(i, j, 0));
    for (int k = N - 1; k >= 1; --k) for (int i = T.Istr; i <= T.Iend; ++i)
      W(i, j, k) = W(i, j, k) - wrk[i - T.IminS] * (z_w(i, j, k) - z_w(i, j, 0));
    for (int i = T.Istr; i <= T.Iend; ++i) W(i, j, N) = 0.0;
  }
  bc_w3d(M, T, W);
}

// Nonlinear/wvelocity.F:151-283: "true" vertical velocity (m/s) at W-points from omega, called every step
// after omega (main3d.F:535) with Ninp=nstp; read by diag's Courant search (diag.F:246-247).
void wvelocity(Model& M, const Tile& T, int Ninp) {
  const int N = M.N; F3 &z_r = M.z_r, &z_w = M.z_w, &W = M.W, &wvel = M.wvel; F4 &u = M.u, &v = M.v; F2 &pm = M.pm, &pn = M.pn;
  // wvelocity.F:151-158 (periodic exchange of the time-averaged barotropic fluxes)
  exchange_u2d(M, T, M.DU_avg1); exchange_v2d(M, T, M.DV_avg1);
  S3 vert(T.IminS, T.ImaxS, T.JminS, T.JmaxS, 1, N);
  S2 wrk(T.IminS, T.ImaxS, T.JminS, T.JmaxS);
  // :171-192  (Ui + Vj)*GRADs(z)
  for (int k = 1; k <= N; ++k) {
    for (int j = T.Jstr; j <= T.Jend; ++j) {
      for (int i = T.Istr; i <= T.Iend + 1; ++i)
        wrk(i, j) = u(i, j, k, Ninp) * (z_r(i, j, k) - z_r(i - 1, j, k)) * (pm(i - 1, j) + pm(i, j));
      for (int i = T.Istr; i <= T.Iend; ++i) vert(i, j, k) = 0.25 * (wrk(i, j) + wrk(i + 1, j));
    }
    for (int j = T.Jstr; j <= T.Jend + 1; ++j) for (int i = T.Istr; i <= T.Iend; ++i)
      wrk(i, j) = v(i, j, k, Ninp) * (z_r(i, j, k) - z_r(i, j - 1, k)) * (pn(i, j - 1) + pn(i, j));
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i)
      vert(i, j, k) = vert(i, j, k) + 0.25 * (wrk(i, j) + wrk(i, j + 1));
  }
  // :203-268
  const double cff1 = 3.0 / 8.0, cff2 = 3.0 / 4.0, cff3 = 1.0 / 8.0, cff4 = 9.0 / 16.0, cff5 = 1.0 / 16.0;
  for (int j = T.Jstr; j <= T.Jend; ++j) {
    for (int i = T.Istr; i <= T.Iend; ++i)
      wrk(i, j) = (M.DU_avg1(i, j) - M.DU_avg1(i + 1, j) + M.DV_avg1(i, j) - M.DV_avg1(i, j + 1)) / (z_w(i, j, N) - z_w(i, j, 0));
    for (int i = T.Istr; i <= T.Iend; ++i) {
      const double slope = (z_r(i, j, 1) - z_w(i, j, 0)) / (z_r(i, j, 2) - z_r(i, j, 1));
      wvel(i, j, 0) = cff1 * (vert(i, j, 1) - slope * (vert(i, j, 2) - vert(i, j, 1))) + cff2 * vert(i, j, 1) - cff3 * vert(i, j, 2);
      wvel(i, j, 1) = pm(i, j) * pn(i, j) * (W(i, j, 1) + wrk(i, j) * (z_w(i, j, 1) - z_w(i, j, 0))) +
                      cff1 * vert(i, j, 1) + cff2 * vert(i, j, 2) - cff3 * vert(i, j, 3);
    }
    for (int k = 2; k <= N - 2; ++k) for (int i = T.Istr; i <= T.Iend; ++i)
      wvel(i, j, k) = pm(i, j) * pn(i, j) * (W(i, j, k) + wrk(i, j) * (z_w(i, j, k) - z_w(i, j, 0))) +
                      cff4 * (vert(i, j, k) + vert(i, j, k + 1)) - cff5 * (vert(i, j, k - 1) + vert(i, j, k + 2));
    for (int i = T.Istr; i <= T.Iend; ++i) {
      const double slope = (z_w(i, j, N) - z_r(i, j, N)) / (z_r(i, j, N) - z_r(i, j, N - 1));
      wvel(i, j, N) = pm(i, j) * pn(i, j) * wrk(i, j) * (z_w(i, j, N) - z_w(i, j, 0)) +
                      cff1 * (vert(i, j, N) + slope * (vert(i, j, N) - vert(i, j, N - 1))) + cff2 * vert(i, j, N) - cff3 * vert(i, j, N - 1);
      wvel(i, j, N - 1) = pm(i, j) * pn(i, j) * (W(i, j, N - 1) + wrk(i, j) * (z_w(i, j, N - 1) - z_w(i, j, 0))) +
                          cff1 * vert(i, j, N) + cff2 * vert(i, j, N - 1) - cff3 * vert(i, j, N - 2);
    }
  }
  bc_w3d(M, T, wvel);   // :272-274
}

// Nonlinear/set_zeta.F:101-118
void set_zeta(Model& M, const Tile& T) {
  for (int j = T.JstrR; j <= T.JendR; ++j) for (int i = T.IstrR; i <= T.IendR; ++i) {
    M.zeta(i, j, 1) = M.Zt_avg1(i, j); M.zeta(i, j, 2) = M.Zt_avg1(i, j);
  }
  exchange_r2d(M, T, M.zeta.slab(1)); exchange_r2d(M, T, M.zeta.slab(2));
}

// Nonlinear/lmd_swfrac.F (lmd_swfrac_tile): Jerlov two-band solar attenuation
static const double lmd_mu1[9] = {0.35, 0.6, 1.0, 1.5, 1.4, 0.42, 0.37, 0.33, 0.00468592};   // mod_scalars.F:1585
static const double lmd_mu2[9] = {23.0, 20.0, 17.0, 14.0, 7.9, 5.13, 3.54, 2.34, 1.51};      // mod_scalars.F:1589
static const double lmd_r1[9] = {0.58, 0.62, 0.67, 0.77, 0.78, 0.57, 0.57, 0.57, 0.55};      // mod_scalars.F:1593
void lmd_swfrac(Model& M, const Tile& T, double Zscale, S2& Z, S2& swdk) {
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
    int Jindex = (int)M.Jwtype(i, j);
    double fac1 = Zscale / lmd_mu1[Jindex - 1], fac2 = Zscale / lmd_mu2[Jindex - 1], fac3 = lmd_r1[Jindex - 1];
    swdk(i, j) = std::exp(Z(i, j) * fac1) * fac3 + std::exp(Z(i, j) * fac2) * (1.0 - fac3);
  }
}

// U3 horizontal tracer flux shared by pre_step3d.F:406-533 and step3d_t.F:641-767
static void tracer_hflux_u3(Model& M, const Tile& T, F3 tk /*tracer volume*/, int k, S2& FX, S2& FE, S2& curv) {
  F3 &Huon = M.Huon, &Hvom = M.Hvom;
  const double cff1 = 1.0 / 6.0;
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istrm1; i <= T.Iendp2; ++i) FX(i, j) = tk(i, j, k) - tk(i - 1, j, k);
  if (!M.EWperiodic) {
    if (T.W) for (int j = T.Jstr; j <= T.Jend; ++j) FX(T.Istr - 1, j) = FX(T.Istr, j);
    if (T.E) for (int j = T.Jstr; j <= T.Jend; ++j) FX(T.Iend + 2, j) = FX(T.Iend + 1, j);
  }
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr - 1; i <= T.Iend + 1; ++i) curv(i, j) = FX(i + 1, j) - FX(i, j);
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend + 1; ++i)
    FX(i, j) = Huon(i, j, k) * 0.5 * (tk(i - 1, j, k) + tk(i, j, k)) -
               cff1 * (curv(i - 1, j) * std::max(Huon(i, j, k), 0.0) + curv(i, j) * std::min(Huon(i, j, k), 0.0));
  for (int j = T.Jstrm1; j <= T.Jendp2; ++j) for (int i = T.Istr; i <= T.Iend; ++i) FE(i, j) = tk(i, j, k) - tk(i, j - 1, k);
  if (!M.NSperiodic) {
    if (T.S) for (int i = T.Istr; i <= T.Iend; ++i) FE(i, T.Jstr - 1) = FE(i, T.Jstr);
    if (T.N) for (int i = T.Istr; i <= T.Iend; ++i) FE(i, T.Jend + 2) = FE(i, T.Jend + 1);
  }
  for (int j = T.Jstr - 1; j <= T.Jend + 1; ++j) for (int i = T.Istr; i <= T.Iend; ++i) curv(i, j) = FE(i, j + 1) - FE(i, j);
  for (int j = T.Jstr; j <= T.Jend + 1; ++j) for (int i = T.Istr; i <= T.Iend; ++i)
    FE(i, j) = Hvom(i, j, k) * 0.5 * (tk(i, j - 1, k) + tk(i, j, k)) -
               cff1 * (curv(i, j - 1) * std::max(Hvom(i, j, k), 0.0) + curv(i, j) * std::min(Hvom(i, j, k), 0.0));
}

// C4 vertical tracer flux, pre_step3d.F:773-808 / step3d_t.F:1150-1185
static void tracer_vflux_c4(Model& M, const Tile& T, F3 tk, int j, S2& FC) {
  const int N = M.N; F3& W = M.W;
  const double cff1 = 0.5, cff2 = 7.0 / 12.0, cff3 = 1.0 / 12.0;
  for (int k = 2; k <= N - 2; ++k) for (int i = T.Istr; i <= T.Iend; ++i)
    FC(i, k) = W(i, j, k) * (cff2 * (tk(i, j, k) + tk(i, j, k + 1)) - cff3 * (tk(i, j, k - 1) + tk(i, j, k + 2)));
  for (int i = T.Istr; i <= T.Iend; ++i) {
    FC(i, 0) = 0.0;
    FC(i, 1) = W(i, j, 1) * (cff1 * tk(i, j, 1) + cff2 * tk(i, j, 2) - cff3 * tk(i, j, 3));
    FC(i, N - 1) = W(i, j, N - 1) * (cff1 * tk(i, j, N) + cff2 * tk(i, j, N - 1) - cff3 * tk(i, j, N - 2));
    FC(i, N) = 0.0;
  }
}

// Nonlinear/pre_step3d.F:329-1168
void pre_step3d(Model& M, const Tile& T) {
  const int N = M.N, NT = M.NT, NAT = M.NAT, nstp = M.nstp, nnew = M.nnew, nrhs = M.nrhs;
  const double dt = M.c.dt, lambda = 1.0;             // mod_scalars.F:750-753
  const bool first = (M.iic == M.ntfirst);
  F3 &Hz = M.Hz, &Huon = M.Huon, &Hvom = M.Hvom, &W = M.W, &z_r = M.z_r, &z_w = M.z_w, &Akv = M.Akv;
  F2 &pm = M.pm, &pn = M.pn; F5& t = M.t; F4 &u = M.u, &v = M.v, &ru = M.ru, &rv = M.rv;
  S2 FX(T.IminS, T.ImaxS, T.JminS, T.JmaxS), FE(T.IminS, T.ImaxS, T.JminS, T.JmaxS), curv(T.IminS, T.ImaxS, T.JminS, T.JmaxS);
  S2 CF(T.IminS, T.ImaxS, 0, N), DC(T.IminS, T.ImaxS, 0, N), FC(T.IminS, T.ImaxS, 0, N);
  S3 swdk(T.IminS, T.ImaxS, T.JminS, T.JmaxS, 0, N);
  if (M.c.app == BENCHMARK) {                           // SOLAR_SOURCE, pre_step3d.F:329-344
    for (int k = 1; k <= N - 1; ++k) {
      for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) FX(i, j) = z_w(i, j, N) - z_w(i, j, k);
      lmd_swfrac(M, T, -1.0, FX, FE);
      for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) swdk(i, j, k) = FE(i, j);
    }
  }
  // T_LOOP1 / K_LOOP: horizontal predictor
  for (int itrc = 1; itrc <= NT; ++itrc) for (int k = 1; k <= N; ++k) {
    tracer_hflux_u3(M, T, t.vol(nstp, itrc), k, FX, FE, curv);
    const double Gamma = 1.0 / 6.0;
    double cff, cff1, cff2;
    if (first) { cff = 0.5 * dt; cff1 = 1.0; cff2 = 0.0; }
    else { cff = (1.0 - Gamma) * dt; cff1 = 0.5 + Gamma; cff2 = 0.5 - Gamma; }
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i)
      t(i, j, k, 3, itrc) = Hz(i, j, k) * (cff1 * t(i, j, k, nstp, itrc) + cff2 * t(i, j, k, nnew, itrc)) -
                            cff * pm(i, j) * pn(i, j) * (FX(i + 1, j) - FX(i, j) + FE(i, j + 1) - FE(i, j));
  }
  // J_LOOP1 / T_LOOP2: vertical predictor with artificial continuity
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int itrc = 1; itrc <= NT; ++itrc) {
    tracer_vflux_c4(M, T, t.vol(nstp, itrc), j, FC);
    const double Gamma = 1.0 / 6.0;
    double cff = first ? 0.5 * dt : (1.0 - Gamma) * dt;
    for (int k = 1; k <= N; ++k) for (int i = T.Istr; i <= T.Iend; ++i)
      DC(i, k) = 1.0 / (Hz(i, j, k) - cff * pm(i, j) * pn(i, j) *
                 (Huon(i + 1, j, k) - Huon(i, j, k) + Hvom(i, j + 1, k) - Hvom(i, j, k) + (W(i, j, k) - W(i, j, k - 1))));
    for (int k = 1; k <= N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) {
      double cff1 = cff * pm(i, j) * pn(i, j);
      t(i, j, k, 3, itrc) = DC(i, k) * (t(i, j, k, 3, itrc) - cff1 * (FC(i, k) - FC(i, k - 1)));
    }
  }
  // start of the corrector for tracers: t(nnew)=Hz*t(nstp)+explicit vertical terms, pre_step3d.F:863-932
  for (int j = T.Jstr; j <= T.Jend; ++j) {
    double cff3 = dt * (1.0 - lambda);
    for (int itrc = 1; itrc <= NT; ++itrc) {
      int ltrc = std::min(NAT, itrc);
      for (int k = 1; k <= N - 1; ++k) for (int i = T.Istr; i <= T.Iend; ++i) {
        double cff = 1.0 / (z_r(i, j, k + 1) - z_r(i, j, k));
        FC(i, k) = cff3 * cff * M.Akt(i, j, k, ltrc) * (t(i, j, k + 1, nstp, itrc) - t(i, j, k, nstp, itrc));
      }
      if (M.c.app == BENCHMARK) {
        if (itrc <= NAT)                                   // LMD_NONLOCAL
          for (int k = 1; k <= N - 1; ++k) for (int i = T.Istr; i <= T.Iend; ++i)
            FC(i, k) = FC(i, k) - dt * M.Akt(i, j, k, itrc) * M.ghats(i, j, k, itrc);
        if (itrc == 1)                                     // SOLAR_SOURCE
          for (int k = 1; k <= N - 1; ++k) for (int i = T.Istr; i <= T.Iend; ++i)
            FC(i, k) = FC(i, k) + dt * M.srflx(i, j) * swdk(i, j, k);
      }
      for (int i = T.Istr; i <= T.Iend; ++i) { FC(i, 0) = dt * M.btflx(i, j, itrc); FC(i, N) = dt * M.stflx(i, j, itrc); }
      for (int k = 1; k <= N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) {
        double cff1 = Hz(i, j, k) * t(i, j, k, nstp, itrc);
        double cff2 = FC(i, k) - FC(i, k - 1);
        t(i, j, k, nnew, itrc) = cff1 + cff2;
      }
    }
  }
  // J_LOOP2: momentum pre-load, pre_step3d.F:943-1144
  for (int j = T.Jstr; j <= T.Jend; ++j) {
    double cff3 = dt * (1.0 - lambda);
    for (int k = 1; k <= N - 1; ++k) for (int i = T.IstrU; i <= T.Iend; ++i) {
      double cff = 1.0 / (z_r(i, j, k + 1) + z_r(i - 1, j, k + 1) - z_r(i, j, k) - z_r(i - 1, j, k));
      FC(i, k) = cff3 * cff * (u(i, j, k + 1, nstp) - u(i, j, k, nstp)) * (Akv(i, j, k) + Akv(i - 1, j, k));
    }
    for (int i = T.IstrU; i <= T.Iend; ++i) { FC(i, 0) = dt * M.bustr(i, j); FC(i, N) = dt * M.sustr(i, j); }
    double cff = dt * 0.25;
    for (int i = T.IstrU; i <= T.Iend; ++i) DC(i, 0) = cff * (pm(i, j) + pm(i - 1, j)) * (pn(i, j) + pn(i - 1, j));
    const int indx = 3 - nrhs;
    if (first) {
      for (int k = 1; k <= N; ++k) for (int i = T.IstrU; i <= T.Iend; ++i) {
        double cff1 = u(i, j, k, nstp) * 0.5 * (Hz(i, j, k) + Hz(i - 1, j, k));
        double cff2 = FC(i, k) - FC(i, k - 1);
        u(i, j, k, nnew) = cff1 + cff2;
      }
    } else if (M.iic == M.ntfirst + 1) {
      for (int k = 1; k <= N; ++k) for (int i = T.IstrU; i <= T.Iend; ++i) {
        double cff1 = u(i, j, k, nstp) * 0.5 * (Hz(i, j, k) + Hz(i - 1, j, k));
        double cff2 = FC(i, k) - FC(i, k - 1);
        double c3 = 0.5 * DC(i, 0);
        u(i, j, k, nnew) = cff1 - c3 * ru(i, j, k, indx) + cff2;
      }
    } else {
      const double cff1 = 5.0 / 12.0, cff2 = 16.0 / 12.0;
      for (int k = 1; k <= N; ++k) for (int i = T.IstrU; i <= T.Iend; ++i) {
        double c3 = u(i, j, k, nstp) * 0.5 * (Hz(i, j, k) + Hz(i - 1, j, k));
        double c4 = FC(i, k) - FC(i, k - 1);
        u(i, j, k, nnew) = c3 + DC(i, 0) * (cff1 * ru(i, j, k, nrhs) - cff2 * ru(i, j, k, indx)) + c4;
      }
    }
    if (j >= T.JstrV) {
      cff3 = dt * (1.0 - lambda);
      for (int k = 1; k <= N - 1; ++k) for (int i = T.Istr; i <= T.Iend; ++i) {
        double c = 1.0 / (z_r(i, j, k + 1) + z_r(i, j - 1, k + 1) - z_r(i, j, k) - z_r(i, j - 1, k));
        FC(i, k) = cff3 * c * (v(i, j, k + 1, nstp) - v(i, j, k, nstp)) * (Akv(i, j, k) + Akv(i, j - 1, k));
      }
      for (int i = T.Istr; i <= T.Iend; ++i) { FC(i, 0) = dt * M.bvstr(i, j); FC(i, N) = dt * M.svstr(i, j); }
      cff = dt * 0.25;
      for (int i = T.Istr; i <= T.Iend; ++i) DC(i, 0) = cff * (pm(i, j) + pm(i, j - 1)) * (pn(i, j) + pn(i, j - 1));
      if (first) {
        for (int k = 1; k <= N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) {
          double cff1 = v(i, j, k, nstp) * 0.5 * (Hz(i, j, k) + Hz(i, j - 1, k));
          double cff2 = FC(i, k) - FC(i, k - 1);
          v(i, j, k, nnew) = cff1 + cff2;
        }
      } else if (M.iic == M.ntfirst + 1) {
        for (int k = 1; k <= N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) {
          double cff1 = v(i, j, k, nstp) * 0.5 * (Hz(i, j, k) + Hz(i, j - 1, k));
          double cff2 = FC(i, k) - FC(i, k - 1);
          double c3 = 0.5 * DC(i, 0);
          v(i, j, k, nnew) = cff1 - c3 * rv(i, j, k, indx) + cff2;
        }
      } else {
        const double cff1 = 5.0 / 12.0, cff2 = 16.0 / 12.0;
        for (int k = 1; k <= N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) {
          double c3 = v(i, j, k, nstp) * 0.5 * (Hz(i, j, k) + Hz(i, j - 1, k));
          double c4 = FC(i, k) - FC(i, k - 1);
          v(i, j, k, nnew) = c3 + DC(i, 0) * (cff1 * rv(i, j, k, nrhs) - cff2 * rv(i, j, k, indx)) + c4;
        }
      }
    }
  }
  for (int itrc = 1; itrc <= NT; ++itrc) { t3dbc(M, T, 3, itrc); exchange_r3d(M, T, t.vol(3, itrc)); }
}

// Nonlinear/prsgrd32.h:238-433 (DJ_GRADPS, density Jacobian with cubic splines)
void prsgrd32(Model& M, const Tile& T) {
  const int N = M.N, nrhs = M.nrhs; const double g = M.c.g, rho0 = M.c.rho0;
  const double OneFifth = 0.2, OneTwelfth = 1.0 / 12.0, eps = 1.0e-10;
  F3 &rho = M.rho, &z_r = M.z_r, &z_w = M.z_w, &Hz = M.Hz; F4 &ru = M.ru, &rv = M.rv;
  S3 P(T.IminS, T.ImaxS, T.JminS, T.JmaxS, 1, N);
  S2 dR(T.IminS, T.ImaxS, 0, N), dZ(T.IminS, T.ImaxS, 0, N);
  S2 FC(T.IminS, T.ImaxS, T.JminS, T.JmaxS), aux(T.IminS, T.ImaxS, T.JminS, T.JmaxS), dRx(T.IminS, T.ImaxS, T.JminS, T.JmaxS), dZx(T.IminS, T.ImaxS, T.JminS, T.JmaxS);
  const double GRho = g / rho0, HalfGRho = 0.5 * GRho;
  for (int j = T.JstrV - 1; j <= T.Jend; ++j) {
    for (int k = 1; k <= N - 1; ++k) for (int i = T.IstrU - 1; i <= T.Iend; ++i) {
      dR(i, k) = rho(i, j, k + 1) - rho(i, j, k); dZ(i, k) = z_r(i, j, k + 1) - z_r(i, j, k);
    }
    for (int i = T.IstrU - 1; i <= T.Iend; ++i) { dR(i, N) = dR(i, N - 1); dZ(i, N) = dZ(i, N - 1); dR(i, 0) = dR(i, 1); dZ(i, 0) = dZ(i, 1); }
    for (int k = N; k >= 1; --k) for (int i = T.IstrU - 1; i <= T.Iend; ++i) {
      double cff = 2.0 * dR(i, k) * dR(i, k - 1);
      if (cff > eps) dR(i, k) = cff / (dR(i, k) + dR(i, k - 1)); else dR(i, k) = 0.0;
      dZ(i, k) = 2.0 * dZ(i, k) * dZ(i, k - 1) / (dZ(i, k) + dZ(i, k - 1));
    }
    for (int i = T.IstrU - 1; i <= T.Iend; ++i) {
      double cff1 = 1.0 / (z_r(i, j, N) - z_r(i, j, N - 1));
      double cff2 = 0.5 * (rho(i, j, N) - rho(i, j, N - 1)) * (z_w(i, j, N) - z_r(i, j, N)) * cff1;
      P(i, j, N) = g * z_w(i, j, N) + GRho * (rho(i, j, N) + cff2) * (z_w(i, j, N) - z_r(i, j, N));
    }
    for (int k = N - 1; k >= 1; --k) for (int i = T.IstrU - 1; i <= T.Iend; ++i)
      P(i, j, k) = P(i, j, k + 1) +
                   HalfGRho * ((rho(i, j, k + 1) + rho(i, j, k)) * (z_r(i, j, k + 1) - z_r(i, j, k)) -
                               OneFifth * ((dR(i, k + 1) - dR(i, k)) * (z_r(i, j, k + 1) - z_r(i, j, k) - OneTwelfth * (dZ(i, k + 1) + dZ(i, k))) -
                                           (dZ(i, k + 1) - dZ(i, k)) * (rho(i, j, k + 1) - rho(i, j, k) - OneTwelfth * (dR(i, k + 1) + dR(i, k)))));
  }
  for (int k = N; k >= 1; --k) {
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU - 1; i <= T.Iend + 1; ++i) {
      aux(i, j) = z_r(i, j, k) - z_r(i - 1, j, k); FC(i, j) = rho(i, j, k) - rho(i - 1, j, k);
    }
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU - 1; i <= T.Iend; ++i) {
      double cff = 2.0 * aux(i, j) * aux(i + 1, j);
      if (cff > eps) { double cff1 = 1.0 / (aux(i, j) + aux(i + 1, j)); dZx(i, j) = cff * cff1; } else dZx(i, j) = 0.0;
      double cff1 = 2.0 * FC(i, j) * FC(i + 1, j);
      if (cff1 > eps) { double cff2 = 1.0 / (FC(i, j) + FC(i + 1, j)); dRx(i, j) = cff1 * cff2; } else dRx(i, j) = 0.0;
    }
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i)
      ru(i, j, k, nrhs) = M.on_u(i, j) * 0.5 * (Hz(i, j, k) + Hz(i - 1, j, k)) *
                          (P(i - 1, j, k) - P(i, j, k) -
                           HalfGRho * ((rho(i, j, k) + rho(i - 1, j, k)) * (z_r(i, j, k) - z_r(i - 1, j, k)) -
                                       OneFifth * ((dRx(i, j) - dRx(i - 1, j)) * (z_r(i, j, k) - z_r(i - 1, j, k) - OneTwelfth * (dZx(i, j) + dZx(i - 1, j))) -
                                                   (dZx(i, j) - dZx(i - 1, j)) * (rho(i, j, k) - rho(i - 1, j, k) - OneTwelfth * (dRx(i, j) + dRx(i - 1, j))))));
  }
  for (int k = N; k >= 1; --k) {
    for (int j = T.JstrV - 1; j <= T.Jend + 1; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
      aux(i, j) = z_r(i, j, k) - z_r(i, j - 1, k); FC(i, j) = rho(i, j, k) - rho(i, j - 1, k);
    }
    for (int j = T.JstrV - 1; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
      double cff = 2.0 * aux(i, j) * aux(i, j + 1);
      if (cff > eps) { double cff1 = 1.0 / (aux(i, j) + aux(i, j + 1)); dZx(i, j) = cff * cff1; } else dZx(i, j) = 0.0;
      double cff1 = 2.0 * FC(i, j) * FC(i, j + 1);
      if (cff1 > eps) { double cff2 = 1.0 / (FC(i, j) + FC(i, j + 1)); dRx(i, j) = cff1 * cff2; } else dRx(i, j) = 0.0;
    }
    for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i)
      rv(i, j, k, nrhs) = M.om_v(i, j) * 0.5 * (Hz(i, j, k) + Hz(i, j - 1, k)) *
                          (P(i, j - 1, k) - P(i, j, k) -
                           HalfGRho * ((rho(i, j, k) + rho(i, j - 1, k)) * (z_r(i, j, k) - z_r(i, j - 1, k)) -
                                       OneFifth * ((dRx(i, j) - dRx(i, j - 1)) * (z_r(i, j, k) - z_r(i, j - 1, k) - OneTwelfth * (dZx(i, j) + dZx(i, j - 1))) -
                                                   (dZx(i, j) - dZx(i, j - 1)) * (rho(i, j, k) - rho(i, j - 1, k) - OneTwelfth * (dRx(i, j) + dRx(i, j - 1))))));
  }
}

// Nonlinear/rhs3d.F:498-1919 (rhs3d_tile)
void rhs3d_tile(Model& M, const Tile& T) {
  const int N = M.N, nrhs = M.nrhs; const double Gadv = -0.25;
  const bool curv = (M.c.app == BENCHMARK);             // CURVGRID, rhs3d.F:570-647
  F3 &Hz = M.Hz, &Huon = M.Huon, &Hvom = M.Hvom, &W = M.W; F4 &u = M.u, &v = M.v, &ru = M.ru, &rv = M.rv;
#define SS(x) S2 x(T.IminS, T.ImaxS, T.JminS, T.JmaxS)
  SS(Huee); SS(Huxx); SS(Hvee); SS(Hvxx); SS(UFx); SS(UFe); SS(VFx); SS(VFe); SS(uee); SS(uxx); SS(vee); SS(vxx);
#undef SS
  S2 FC(T.IminS, T.ImaxS, 0, N);
  for (int k = 1; k <= N; ++k) {
    // Coriolis, rhs3d.F:506-532
    for (int j = T.JstrV - 1; j <= T.Jend; ++j) for (int i = T.IstrU - 1; i <= T.Iend; ++i) {
      double cff = 0.5 * Hz(i, j, k) * M.fomn(i, j);
      UFx(i, j) = cff * (v(i, j, k, nrhs) + v(i, j + 1, k, nrhs));
      VFe(i, j) = cff * (u(i, j, k, nrhs) + u(i + 1, j, k, nrhs));
    }
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) { double cff1 = 0.5 * (UFx(i, j) + UFx(i - 1, j)); ru(i, j, k, nrhs) = ru(i, j, k, nrhs) + cff1; }
    for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) { double cff1 = 0.5 * (VFe(i, j) + VFe(i, j - 1)); rv(i, j, k, nrhs) = rv(i, j, k, nrhs) - cff1; }
    if (curv) {
      for (int j = T.JstrV - 1; j <= T.Jend; ++j) for (int i = T.IstrU - 1; i <= T.Iend; ++i) {
        double cff1 = 0.5 * (v(i, j, k, nrhs) + v(i, j + 1, k, nrhs));
        double cff2 = 0.5 * (u(i, j, k, nrhs) + u(i + 1, j, k, nrhs));
        double cff3 = cff1 * M.dndx(i, j), cff4 = cff2 * M.dmde(i, j);
        double cff = Hz(i, j, k) * (cff3 - cff4);
        UFx(i, j) = cff * cff1; VFe(i, j) = cff * cff2;
      }
      for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) { double cff1 = 0.5 * (UFx(i, j) + UFx(i - 1, j)); ru(i, j, k, nrhs) = ru(i, j, k, nrhs) + cff1; }
      for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) { double cff1 = 0.5 * (VFe(i, j) + VFe(i, j - 1)); rv(i, j, k, nrhs) = rv(i, j, k, nrhs) - cff1; }
    }
    // third-order upstream horizontal advection, rhs3d.F:725-1002
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrUm1; i <= T.Iendp1; ++i) {
      uxx(i, j) = u(i - 1, j, k, nrhs) - 2.0 * u(i, j, k, nrhs) + u(i + 1, j, k, nrhs);
      Huxx(i, j) = Huon(i - 1, j, k) - 2.0 * Huon(i, j, k) + Huon(i + 1, j, k);
    }
    if (!M.EWperiodic) {
      if (T.W) for (int j = T.Jstr; j <= T.Jend; ++j) { uxx(T.Istr, j) = uxx(T.Istr + 1, j); Huxx(T.Istr, j) = Huxx(T.Istr + 1, j); }
      if (T.E) for (int j = T.Jstr; j <= T.Jend; ++j) { uxx(T.Iend + 1, j) = uxx(T.Iend, j); Huxx(T.Iend + 1, j) = Huxx(T.Iend, j); }
    }
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU - 1; i <= T.Iend; ++i) {
      double cff1 = u(i, j, k, nrhs) + u(i + 1, j, k, nrhs);
      double cff = (cff1 > 0.0) ? uxx(i, j) : uxx(i + 1, j);
      UFx(i, j) = 0.25 * (cff1 + Gadv * cff) * (Huon(i, j, k) + Huon(i + 1, j, k) + Gadv * 0.5 * (Huxx(i, j) + Huxx(i + 1, j)));
    }
    for (int j = T.Jstrm1; j <= T.Jendp1; ++j) for (int i = T.IstrU; i <= T.Iend; ++i)
      uee(i, j) = u(i, j - 1, k, nrhs) - 2.0 * u(i, j, k, nrhs) + u(i, j + 1, k, nrhs);
    if (!M.NSperiodic) {
      if (T.S) for (int i = T.IstrU; i <= T.Iend; ++i) uee(i, T.Jstr - 1) = uee(i, T.Jstr);
      if (T.N) for (int i = T.IstrU; i <= T.Iend; ++i) uee(i, T.Jend + 1) = uee(i, T.Jend);
    }
    for (int j = T.Jstr; j <= T.Jend + 1; ++j) for (int i = T.IstrU - 1; i <= T.Iend; ++i)
      Hvxx(i, j) = Hvom(i - 1, j, k) - 2.0 * Hvom(i, j, k) + Hvom(i + 1, j, k);
    for (int j = T.Jstr; j <= T.Jend + 1; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) {
      double cff1 = u(i, j, k, nrhs) + u(i, j - 1, k, nrhs);
      double cff2 = Hvom(i, j, k) + Hvom(i - 1, j, k);
      double cff = (cff2 > 0.0) ? uee(i, j - 1) : uee(i, j);
      UFe(i, j) = 0.25 * (cff1 + Gadv * cff) * (cff2 + Gadv * 0.5 * (Hvxx(i, j) + Hvxx(i - 1, j)));
    }
    for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istrm1; i <= T.Iendp1; ++i)
      vxx(i, j) = v(i - 1, j, k, nrhs) - 2.0 * v(i, j, k, nrhs) + v(i + 1, j, k, nrhs);
    if (!M.EWperiodic) {
      if (T.W) for (int j = T.JstrV; j <= T.Jend; ++j) vxx(T.Istr - 1, j) = vxx(T.Istr, j);
      if (T.E) for (int j = T.JstrV; j <= T.Jend; ++j) vxx(T.Iend + 1, j) = vxx(T.Iend, j);
    }
    for (int j = T.JstrV - 1; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend + 1; ++i)
      Huee(i, j) = Huon(i, j - 1, k) - 2.0 * Huon(i, j, k) + Huon(i, j + 1, k);
    for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend + 1; ++i) {
      double cff1 = v(i, j, k, nrhs) + v(i - 1, j, k, nrhs);
      double cff2 = Huon(i, j, k) + Huon(i, j - 1, k);
      double cff = (cff2 > 0.0) ? vxx(i - 1, j) : vxx(i, j);
      VFx(i, j) = 0.25 * (cff1 + Gadv * cff) * (cff2 + Gadv * 0.5 * (Huee(i, j) + Huee(i, j - 1)));
    }
    for (int j = T.JstrVm1; j <= T.Jendp1; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
      vee(i, j) = v(i, j - 1, k, nrhs) - 2.0 * v(i, j, k, nrhs) + v(i, j + 1, k, nrhs);
      Hvee(i, j) = Hvom(i, j - 1, k) - 2.0 * Hvom(i, j, k) + Hvom(i, j + 1, k);
    }
    if (!M.NSperiodic) {
      if (T.S) for (int i = T.Istr; i <= T.Iend; ++i) { vee(i, T.Jstr) = vee(i, T.Jstr + 1); Hvee(i, T.Jstr) = Hvee(i, T.Jstr + 1); }
      if (T.N) for (int i = T.Istr; i <= T.Iend; ++i) { vee(i, T.Jend + 1) = vee(i, T.Jend); Hvee(i, T.Jend + 1) = Hvee(i, T.Jend); }
    }
    for (int j = T.JstrV - 1; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
      double cff1 = v(i, j, k, nrhs) + v(i, j + 1, k, nrhs);
      double cff = (cff1 > 0.0) ? vee(i, j) : vee(i, j + 1);
      VFe(i, j) = 0.25 * (cff1 + Gadv * cff) * (Hvom(i, j, k) + Hvom(i, j + 1, k) + Gadv * 0.5 * (Hvee(i, j) + Hvee(i, j + 1)));
    }
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.IstrU; i <= T.Iend; ++i) {
      double cff1 = UFx(i, j) - UFx(i - 1, j), cff2 = UFe(i, j + 1) - UFe(i, j), cff = cff1 + cff2;
      ru(i, j, k, nrhs) = ru(i, j, k, nrhs) - cff;
    }
    for (int j = T.JstrV; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
      double cff1 = VFx(i + 1, j) - VFx(i, j), cff2 = VFe(i, j) - VFe(i, j - 1), cff = cff1 + cff2;
      rv(i, j, k, nrhs) = rv(i, j, k, nrhs) - cff;
    }
  }
  // J_LOOP: vertical advection + vertical integrals, rhs3d.F:1133-1916
  for (int j = T.Jstr; j <= T.Jend; ++j) {
    const double cff1 = 9.0 / 16.0, cff2 = 1.0 / 16.0;
    for (int k = 2; k <= N - 2; ++k) for (int i = T.IstrU; i <= T.Iend; ++i)
      FC(i, k) = (cff1 * (u(i, j, k, nrhs) + u(i, j, k + 1, nrhs)) - cff2 * (u(i, j, k - 1, nrhs) + u(i, j, k + 2, nrhs))) *
                 (cff1 * (W(i, j, k) + W(i - 1, j, k)) - cff2 * (W(i + 1, j, k) + W(i - 2, j, k)));
    for (int i = T.IstrU; i <= T.Iend; ++i) {
      FC(i, N) = 0.0;
      FC(i, N - 1) = (cff1 * (u(i, j, N - 1, nrhs) + u(i, j, N, nrhs)) - cff2 * (u(i, j, N - 2, nrhs) + u(i, j, N, nrhs))) *
                     (cff1 * (W(i, j, N - 1) + W(i - 1, j, N - 1)) - cff2 * (W(i + 1, j, N - 1) + W(i - 2, j, N - 1)));
      FC(i, 1) = (cff1 * (u(i, j, 1, nrhs) + u(i, j, 2, nrhs)) - cff2 * (u(i, j, 1, nrhs) + u(i, j, 3, nrhs))) *
                 (cff1 * (W(i, j, 1) + W(i - 1, j, 1)) - cff2 * (W(i + 1, j, 1) + W(i - 2, j, 1)));
      FC(i, 0) = 0.0;
    }
    for (int k = 1; k <= N; ++k) for (int i = T.IstrU; i <= T.Iend; ++i) { double cff = FC(i, k) - FC(i, k - 1); ru(i, j, k, nrhs) = ru(i, j, k, nrhs) - cff; }
    if (j >= T.JstrV) {
      for (int k = 2; k <= N - 2; ++k) for (int i = T.Istr; i <= T.Iend; ++i)
        FC(i, k) = (cff1 * (v(i, j, k, nrhs) + v(i, j, k + 1, nrhs)) - cff2 * (v(i, j, k - 1, nrhs) + v(i, j, k + 2, nrhs))) *
                   (cff1 * (W(i, j, k) + W(i, j - 1, k)) - cff2 * (W(i, j + 1, k) + W(i, j - 2, k)));
      for (int i = T.Istr; i <= T.Iend; ++i) {
        FC(i, N) = 0.0;
        FC(i, N - 1) = (cff1 * (v(i, j, N - 1, nrhs) + v(i, j, N, nrhs)) - cff2 * (v(i, j, N - 2, nrhs) + v(i, j, N, nrhs))) *
                       (cff1 * (W(i, j, N - 1) + W(i, j - 1, N - 1)) - cff2 * (W(i, j + 1, N - 1) + W(i, j - 2, N - 1)));
        FC(i, 1) = (cff1 * (v(i, j, 1, nrhs) + v(i, j, 2, nrhs)) - cff2 * (v(i, j, 1, nrhs) + v(i, j, 3, nrhs))) *
                   (cff1 * (W(i, j, 1) + W(i, j - 1, 1)) - cff2 * (W(i, j + 1, 1) + W(i, j - 2, 1)));
        FC(i, 0) = 0.0;
      }
      for (int k = 1; k <= N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) { double cff = FC(i, k) - FC(i, k - 1); rv(i, j, k, nrhs) = rv(i, j, k, nrhs) - cff; }
    }
    for (int i = T.IstrU; i <= T.Iend; ++i) M.rufrc(i, j) = ru(i, j, 1, nrhs);
    for (int k = 2; k <= N; ++k) for (int i = T.IstrU; i <= T.Iend; ++i) M.rufrc(i, j) = M.rufrc(i, j) + ru(i, j, k, nrhs);
    for (int i = T.IstrU; i <= T.Iend; ++i) {
      double cff = M.om_u(i, j) * M.on_u(i, j);
      double c1 = M.sustr(i, j) * cff, c2 = -M.bustr(i, j) * cff;
      M.rufrc(i, j) = M.rufrc(i, j) + c1 + c2;
    }
    if (j >= T.JstrV) {
      for (int i = T.Istr; i <= T.Iend; ++i) M.rvfrc(i, j) = rv(i, j, 1, nrhs);
      for (int k = 2; k <= N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) M.rvfrc(i, j) = M.rvfrc(i, j) + rv(i, j, k, nrhs);
      for (int i = T.Istr; i <= T.Iend; ++i) {
        double cff = M.om_v(i, j) * M.on_v(i, j);
        double c1 = M.svstr(i, j) * cff, c2 = -M.bvstr(i, j) * cff;
        M.rvfrc(i, j) = M.rvfrc(i, j) + c1 + c2;
      }
    }
  }
}

// spline-form implicit vertical mixing shared by step3d_uv.F:392-438,859-905
static void spline_vvisc(const Tile& T, int N, double dt, int i0, int i1, S2& AK, S2& Hzk, S2& oHz, S2& FC, S2& CF, S2& BC, S2& DC, F3 q, int j) {
  (void)T;
  double cff1 = 1.0 / 6.0;
  for (int k = 1; k <= N - 1; ++k) for (int i = i0; i <= i1; ++i) {
    FC(i, k) = cff1 * Hzk(i, k) - dt * AK(i, k - 1) * oHz(i, k);
    CF(i, k) = cff1 * Hzk(i, k + 1) - dt * AK(i, k + 1) * oHz(i, k + 1);
  }
  for (int i = i0; i <= i1; ++i) { CF(i, 0) = 0.0; DC(i, 0) = 0.0; }
  cff1 = 1.0 / 3.0;
  for (int k = 1; k <= N - 1; ++k) for (int i = i0; i <= i1; ++i) {
    BC(i, k) = cff1 * (Hzk(i, k) + Hzk(i, k + 1)) + dt * AK(i, k) * (oHz(i, k) + oHz(i, k + 1));
    double cff = 1.0 / (BC(i, k) - FC(i, k) * CF(i, k - 1));
    CF(i, k) = cff * CF(i, k);
    DC(i, k) = cff * (q(i, j, k + 1) - q(i, j, k) - FC(i, k) * DC(i, k - 1));
  }
  for (int i = i0; i <= i1; ++i) DC(i, N) = 0.0;
  for (int k = N - 1; k >= 1; --k) for (int i = i0; i <= i1; ++i) DC(i, k) = DC(i, k) - CF(i, k) * DC(i, k + 1);
  for (int k = 1; k <= N; ++k) for (int i = i0; i <= i1; ++i) {
    DC(i, k) = DC(i, k) * AK(i, k);
    double cff = dt * oHz(i, k) * (DC(i, k) - DC(i, k - 1));
    q(i, j, k) = q(i, j, k) + cff;
  }
}

// Nonlinear/step3d_uv.F:330-1824
void step3d_uv(Model& M, const Tile& T) {
  const int N = M.N, nnew = M.nnew, nrhs = M.nrhs; const double dt = M.c.dt;
  F3 &Hz = M.Hz, &Akv = M.Akv, &Huon = M.Huon, &Hvom = M.Hvom; F2 &pm = M.pm, &pn = M.pn;
  F3 un = M.u.vol(nnew), vn = M.v.vol(nnew); F4 &ru = M.ru, &rv = M.rv;
  S2 AK(T.IminS, T.ImaxS, 0, N), BC(T.IminS, T.ImaxS, 0, N), CF(T.IminS, T.ImaxS, 0, N), DC(T.IminS, T.ImaxS, 0, N), FC(T.IminS, T.ImaxS, 0, N);
  S2 Hzk(T.IminS, T.ImaxS, 1, N), oHz(T.IminS, T.ImaxS, 1, N);
  double cffab;
  if (M.iic == M.ntfirst) cffab = 0.25 * dt;
  else if (M.iic == M.ntfirst + 1) cffab = 0.25 * dt * 3.0 / 2.0;
  else cffab = 0.25 * dt * 23.0 / 12.0;
  for (int j = T.Jstr; j <= T.Jend; ++j) {
    for (int i = T.IstrU; i <= T.Iend; ++i) {
      AK(i, 0) = 0.5 * (Akv(i - 1, j, 0) + Akv(i, j, 0));
      for (int k = 1; k <= N; ++k) {
        AK(i, k) = 0.5 * (Akv(i - 1, j, k) + Akv(i, j, k));
        Hzk(i, k) = 0.5 * (Hz(i - 1, j, k) + Hz(i, j, k));
        oHz(i, k) = 1.0 / Hzk(i, k);
      }
    }
    for (int i = T.IstrU; i <= T.Iend; ++i) DC(i, 0) = cffab * (pm(i, j) + pm(i - 1, j)) * (pn(i, j) + pn(i - 1, j));
    for (int k = 1; k <= N; ++k) for (int i = T.IstrU; i <= T.Iend; ++i) {
      un(i, j, k) = un(i, j, k) + DC(i, 0) * ru(i, j, k, nrhs);
      un(i, j, k) = un(i, j, k) * oHz(i, k);
    }
    spline_vvisc(T, N, dt, T.IstrU, T.Iend, AK, Hzk, oHz, FC, CF, BC, DC, un, j);
    // replace vertical mean by the barotropic transport, step3d_uv.F:597-715
    for (int i = T.IstrU; i <= T.Iend; ++i) { CF(i, 0) = Hzk(i, 1); DC(i, 0) = un(i, j, 1) * Hzk(i, 1); }
    for (int k = 2; k <= N; ++k) for (int i = T.IstrU; i <= T.Iend; ++i) { CF(i, 0) = CF(i, 0) + Hzk(i, k); DC(i, 0) = DC(i, 0) + un(i, j, k) * Hzk(i, k); }
    for (int i = T.IstrU; i <= T.Iend; ++i) {
      double cff1 = 1.0 / (CF(i, 0) * M.on_u(i, j));
      DC(i, 0) = (DC(i, 0) * M.on_u(i, j) - M.DU_avg1(i, j)) * cff1;
    }
    for (int k = 1; k <= N; ++k) for (int i = T.IstrU; i <= T.Iend; ++i) un(i, j, k) = un(i, j, k) - DC(i, 0);
    if (j >= T.JstrV) {
      for (int i = T.Istr; i <= T.Iend; ++i) {
        AK(i, 0) = 0.5 * (Akv(i, j - 1, 0) + Akv(i, j, 0));
        for (int k = 1; k <= N; ++k) {
          AK(i, k) = 0.5 * (Akv(i, j - 1, k) + Akv(i, j, k));
          Hzk(i, k) = 0.5 * (Hz(i, j - 1, k) + Hz(i, j, k));
          oHz(i, k) = 1.0 / Hzk(i, k);
        }
      }
      for (int i = T.Istr; i <= T.Iend; ++i) DC(i, 0) = cffab * (pm(i, j) + pm(i, j - 1)) * (pn(i, j) + pn(i, j - 1));
      for (int k = 1; k <= N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) {
        vn(i, j, k) = vn(i, j, k) + DC(i, 0) * rv(i, j, k, nrhs);
        vn(i, j, k) = vn(i, j, k) * oHz(i, k);
      }
      spline_vvisc(T, N, dt, T.Istr, T.Iend, AK, Hzk, oHz, FC, CF, BC, DC, vn, j);
      for (int i = T.Istr; i <= T.Iend; ++i) { CF(i, 0) = Hzk(i, 1); DC(i, 0) = vn(i, j, 1) * Hzk(i, 1); }
      for (int k = 2; k <= N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) { CF(i, 0) = CF(i, 0) + Hzk(i, k); DC(i, 0) = DC(i, 0) + vn(i, j, k) * Hzk(i, k); }
      for (int i = T.Istr; i <= T.Iend; ++i) {
        double cff1 = 1.0 / (CF(i, 0) * M.om_v(i, j));
        DC(i, 0) = (DC(i, 0) * M.om_v(i, j) - M.DV_avg1(i, j)) * cff1;
      }
      for (int k = 1; k <= N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) vn(i, j, k) = vn(i, j, k) - DC(i, 0);
    }
  }
  u3dbc(M, T, nnew); v3dbc(M, T, nnew);
  // 2D/3D coupling and time-centred mass fluxes, step3d_uv.F:1312-1756
  for (int j = T.JstrT; j <= T.JendT; ++j) {
    for (int i = T.IstrP; i <= T.IendT; ++i) { DC(i, 0) = 0.0; CF(i, 0) = 0.0; FC(i, 0) = 0.0; }
    for (int k = 1; k <= N; ++k) for (int i = T.IstrP; i <= T.IendT; ++i) {
      double cff = 0.5 * M.on_u(i, j);
      DC(i, k) = cff * (Hz(i, j, k) + Hz(i - 1, j, k));
      DC(i, 0) = DC(i, 0) + DC(i, k);
      CF(i, 0) = CF(i, 0) + DC(i, k) * un(i, j, k);
    }
    for (int i = T.IstrP; i <= T.IendT; ++i) {
      DC(i, 0) = 1.0 / DC(i, 0);
      CF(i, 0) = DC(i, 0) * (CF(i, 0) - M.DU_avg1(i, j));
      M.ubar(i, j, 1) = DC(i, 0) * M.DU_avg1(i, j);
      M.ubar(i, j, 2) = M.ubar(i, j, 1);
    }
    if (!M.EWperiodic) {
      if (T.W) for (int k = 1; k <= N; ++k) un(T.Istr, j, k) = un(T.Istr, j, k) - CF(T.Istr, 0);
      if (T.E) for (int k = 1; k <= N; ++k) un(T.Iend + 1, j, k) = un(T.Iend + 1, j, k) - CF(T.Iend + 1, 0);
    }
    if (!M.NSperiodic) {
      if (j == 0) for (int k = 1; k <= N; ++k) for (int i = T.IstrU; i <= T.Iend; ++i) un(i, j, k) = un(i, j, k) - CF(i, 0);
      if (j == M.Mm + 1) for (int k = 1; k <= N; ++k) for (int i = T.IstrU; i <= T.Iend; ++i) un(i, j, k) = un(i, j, k) - CF(i, 0);
    }
    for (int k = N; k >= 1; --k) for (int i = T.IstrP; i <= T.IendT; ++i) {
      Huon(i, j, k) = 0.5 * (Huon(i, j, k) + un(i, j, k) * DC(i, k));
      FC(i, 0) = FC(i, 0) + Huon(i, j, k);
    }
    for (int i = T.IstrP; i <= T.IendT; ++i) FC(i, 0) = DC(i, 0) * (FC(i, 0) - M.DU_avg2(i, j));
    for (int k = 1; k <= N; ++k) for (int i = T.IstrP; i <= T.IendT; ++i) Huon(i, j, k) = Huon(i, j, k) - DC(i, k) * FC(i, 0);
    if (j >= T.Jstr) {
      for (int i = T.IstrT; i <= T.IendT; ++i) { DC(i, 0) = 0.0; CF(i, 0) = 0.0; FC(i, 0) = 0.0; }
      for (int k = 1; k <= N; ++k) for (int i = T.IstrT; i <= T.IendT; ++i) {
        double cff = 0.5 * M.om_v(i, j);
        DC(i, k) = cff * (Hz(i, j, k) + Hz(i, j - 1, k));
        DC(i, 0) = DC(i, 0) + DC(i, k);
        CF(i, 0) = CF(i, 0) + DC(i, k) * vn(i, j, k);
      }
      for (int i = T.IstrT; i <= T.IendT; ++i) {
        DC(i, 0) = 1.0 / DC(i, 0);
        CF(i, 0) = DC(i, 0) * (CF(i, 0) - M.DV_avg1(i, j));
        M.vbar(i, j, 1) = DC(i, 0) * M.DV_avg1(i, j);
        M.vbar(i, j, 2) = M.vbar(i, j, 1);
      }
      if (!M.EWperiodic) {
        if (T.W) for (int k = 1; k <= N; ++k) vn(T.Istr - 1, j, k) = vn(T.Istr - 1, j, k) - CF(T.Istr - 1, 0);
        if (T.E) for (int k = 1; k <= N; ++k) vn(T.Iend + 1, j, k) = vn(T.Iend + 1, j, k) - CF(T.Iend + 1, 0);
      }
      if (!M.NSperiodic) {
        if (j == 1) for (int k = 1; k <= N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) vn(i, j, k) = vn(i, j, k) - CF(i, 0);
        if (j == M.Mm + 1) for (int k = 1; k <= N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) vn(i, j, k) = vn(i, j, k) - CF(i, 0);
      }
      for (int k = N; k >= 1; --k) for (int i = T.IstrT; i <= T.IendT; ++i) {
        Hvom(i, j, k) = 0.5 * (Hvom(i, j, k) + vn(i, j, k) * DC(i, k));
        FC(i, 0) = FC(i, 0) + Hvom(i, j, k);
      }
      for (int i = T.IstrT; i <= T.IendT; ++i) FC(i, 0) = DC(i, 0) * (FC(i, 0) - M.DV_avg2(i, j));
      for (int k = 1; k <= N; ++k) for (int i = T.IstrT; i <= T.IendT; ++i) Hvom(i, j, k) = Hvom(i, j, k) - DC(i, k) * FC(i, 0);
    }
  }
  exchange_u3d(M, T, un); exchange_v3d(M, T, vn); exchange_u3d(M, T, Huon); exchange_v3d(M, T, Hvom);
  for (int k = 1; k <= 2; ++k) { exchange_u2d(M, T, M.ubar.slab(k)); exchange_v2d(M, T, M.vbar.slab(k)); }
  // uv_C2A_grid (step3d_uv.F:1841) only fills ua,va for output: not on the prognostic path.
}

// Nonlinear/step3d_t.F:346-1924
void step3d_t(Model& M, const Tile& T) {
  const int N = M.N, NT = M.NT, NAT = M.NAT, nnew = M.nnew; const double dt = M.c.dt;
  F3& Hz = M.Hz; F2 &pm = M.pm, &pn = M.pn; F5& t = M.t;
  S2 FX(T.IminS, T.ImaxS, T.JminS, T.JmaxS), FE(T.IminS, T.ImaxS, T.JminS, T.JmaxS), curv(T.IminS, T.ImaxS, T.JminS, T.JmaxS);
  S2 CF(T.IminS, T.ImaxS, 0, N), BC(T.IminS, T.ImaxS, 0, N), DC(T.IminS, T.ImaxS, 0, N), FC(T.IminS, T.ImaxS, 0, N);
  S3 oHz(T.IminS, T.ImaxS, T.JminS, T.JmaxS, 1, N);
  for (int k = 1; k <= N; ++k) for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) oHz(i, j, k) = 1.0 / Hz(i, j, k);
  // T_LOOP1: horizontal advection of t(3), step3d_t.F:641-916
  for (int itrc = 1; itrc <= NT; ++itrc) for (int k = 1; k <= N; ++k) {
    tracer_hflux_u3(M, T, t.vol(3, itrc), k, FX, FE, curv);
    for (int j = T.Jstr; j <= T.Jend; ++j) for (int i = T.Istr; i <= T.Iend; ++i) {
      double cff = dt * pm(i, j) * pn(i, j);
      double cff1 = cff * (FX(i + 1, j) - FX(i, j));
      double cff2 = cff * (FE(i, j + 1) - FE(i, j));
      double cff3 = cff1 + cff2;
      t(i, j, k, nnew, itrc) = t(i, j, k, nnew, itrc) - cff3;
    }
  }
  // T_LOOP2: vertical advection, step3d_t.F:1150-1365
  for (int itrc = 1; itrc <= NT; ++itrc) for (int j = T.Jstr; j <= T.Jend; ++j) {
    tracer_vflux_c4(M, T, t.vol(3, itrc), j, FC);
    for (int i = T.Istr; i <= T.Iend; ++i) CF(i, 0) = dt * pm(i, j) * pn(i, j);
    for (int k = 1; k <= N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) {
      double cff1 = CF(i, 0) * (FC(i, k) - FC(i, k - 1));
      t(i, j, k, nnew, itrc) = t(i, j, k, nnew, itrc) - cff1;
      t(i, j, k, nnew, itrc) = t(i, j, k, nnew, itrc) * oHz(i, j, k);
    }
  }
  // J_LOOP2: spline implicit vertical diffusion, step3d_t.F:1672-1721
  for (int j = T.Jstr; j <= T.Jend; ++j) for (int itrc = 1; itrc <= NT; ++itrc) {
    int ltrc = std::min(NAT, itrc);
    double cff1 = 1.0 / 6.0;
    for (int k = 1; k <= N - 1; ++k) for (int i = T.Istr; i <= T.Iend; ++i) {
      FC(i, k) = cff1 * Hz(i, j, k) - dt * M.Akt(i, j, k - 1, ltrc) * oHz(i, j, k);
      CF(i, k) = cff1 * Hz(i, j, k + 1) - dt * M.Akt(i, j, k + 1, ltrc) * oHz(i, j, k + 1);
    }
    for (int i = T.Istr; i <= T.Iend; ++i) { CF(i, 0) = 0.0; DC(i, 0) = 0.0; }
    cff1 = 1.0 / 3.0;
    for (int k = 1; k <= N - 1; ++k) for (int i = T.Istr; i <= T.Iend; ++i) {
      BC(i, k) = cff1 * (Hz(i, j, k) + Hz(i, j, k + 1)) + dt * M.Akt(i, j, k, ltrc) * (oHz(i, j, k) + oHz(i, j, k + 1));
      double cff = 1.0 / (BC(i, k) - FC(i, k) * CF(i, k - 1));
      CF(i, k) = cff * CF(i, k);
      DC(i, k) = cff * (t(i, j, k + 1, nnew, itrc) - t(i, j, k, nnew, itrc) - FC(i, k) * DC(i, k - 1));
    }
    for (int i = T.Istr; i <= T.Iend; ++i) DC(i, N) = 0.0;
    for (int k = N - 1; k >= 1; --k) for (int i = T.Istr; i <= T.Iend; ++i) DC(i, k) = DC(i, k) - CF(i, k) * DC(i, k + 1);
    for (int k = 1; k <= N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) {
      DC(i, k) = DC(i, k) * M.Akt(i, j, k, ltrc);
      double c1 = dt * oHz(i, j, k) * (DC(i, k) - DC(i, k - 1));
      t(i, j, k, nnew, itrc) = t(i, j, k, nnew, itrc) + c1;
    }
  }
  for (int itrc = 1; itrc <= NT; ++itrc) { t3dbc(M, T, nnew, itrc); exchange_r3d(M, T, t.vol(nnew, itrc)); }
}

}  // namespace orc
