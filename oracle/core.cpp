// oracle/core.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h header).
// Model allocation, tile bounds, periodic exchanges and lateral boundary
// conditions for the E-W periodic / N-S closed channel used by UPWELLING and
// BENCHMARK.
#include "oracle.h"
#include <algorithm>
#include <cassert>
#include <stdexcept>

namespace orc {

Config config_upwelling() {
  // ROMS/External/roms_upwelling.in:94-96,231-233,466 ; upwelling.h:15-49
  Config c;
  c.app = UPWELLING; c.Lm = 41; c.Mm = 80; c.N = 16;
  c.dt = 300.0; c.ndtfast = 30;
  c.theta_s = 3.0; c.theta_b = 0.0; c.Tcline = 25.0;
  c.Akt_bak[0] = c.Akt_bak[1] = 1.0e-6; c.Akv_bak = 1.0e-5;
  c.tnu2[0] = c.tnu2[1] = 0.0; c.visc2 = 5.0;
  c.T0 = 14.0; c.S0 = 35.0; c.R0 = 1027.0; c.Tcoef = 1.7e-4; c.Scoef = 0.0;
  return c;
}

Config config_benchmark(int Lm, int Mm, int N) {
  // ROMS/External/roms_benchmark1.in:94-96,231-233 ; benchmark.h:17-57
  Config c;
  c.app = BENCHMARK; c.Lm = Lm; c.Mm = Mm; c.N = N;
  c.dt = 150.0; c.ndtfast = 20;
  c.theta_s = 0.0; c.theta_b = 0.0; c.Tcline = 400.0;
  c.Akt_bak[0] = c.Akt_bak[1] = 1.0e-5; c.Akv_bak = 1.0e-4;
  c.tnu2[0] = c.tnu2[1] = 500.0; c.visc2 = 5000.0;
  c.T0 = 10.0; c.S0 = 35.0; c.R0 = 1027.0; c.Tcoef = 1.7e-4; c.Scoef = 7.6e-4;
  return c;
}

Field& Model::add(const std::string& name, int kLB, int nk, int nl, int nm) {
  Field& fl = fields[name];
  fl.kLB = kLB; fl.nk = nk; fl.nl = nl; fl.nm = nm;
  fl.d.assign((size_t)ni * nj * nk * nl * nm, 0.0);   // mod_*.F initialise with IniVal=0
  return fl;
}
F2 Model::v2(const std::string& n) { Field& f = fields.at(n); return F2{f.d.data(), LBi, ni, LBj, nj}; }
F3 Model::v3(const std::string& n) { Field& f = fields.at(n); return F3{f.d.data(), LBi, ni, LBj, nj, f.kLB, f.nk}; }
F4 Model::v4(const std::string& n) { Field& f = fields.at(n); return F4{f.d.data(), LBi, ni, LBj, nj, f.kLB, f.nk, f.nl}; }
F5 Model::v5(const std::string& n) { Field& f = fields.at(n); return F5{f.d.data(), LBi, ni, LBj, nj, f.nk, f.nl, f.nm}; }

// Utility/get_bounds.F:1020-1039 (tile_bounds_2d) and :1044-1884 (var_bounds)
static Tile make_tile(const Model& M, int tile) {
  const Config& c = M.c;
  Tile T{};
  const int NtI = c.NtileI, NtJ = c.NtileJ, Lm = M.Lm, Mm = M.Mm;
  int ChunkSizeI = (Lm + NtI - 1) / NtI, ChunkSizeJ = (Mm + NtJ - 1) / NtJ;
  int MarginI = (NtI * ChunkSizeI - Lm) / 2, MarginJ = (NtJ * ChunkSizeJ - Mm) / 2;
  T.Jtile = tile / NtI; T.Itile = tile - T.Jtile * NtI;
  int my_Istr = 1 + T.Itile * ChunkSizeI - MarginI, my_Iend = my_Istr + ChunkSizeI - 1;
  my_Istr = std::max(my_Istr, 1); my_Iend = std::min(my_Iend, Lm);
  int my_Jstr = 1 + T.Jtile * ChunkSizeJ - MarginJ, my_Jend = my_Jstr + ChunkSizeJ - 1;
  my_Jstr = std::max(my_Jstr, 1); my_Jend = std::min(my_Jend, Mm);
  T.W = (T.Itile == 0); T.E = (T.Itile == NtI - 1); T.S = (T.Jtile == 0); T.N = (T.Jtile == NtJ - 1);
  const bool EW = M.EWperiodic, NS = M.NSperiodic;
  // --- I direction
  T.Istr = my_Istr; T.Iend = my_Iend;
  if (T.W && !EW) {
    T.IstrP = my_Istr; T.IstrR = my_Istr - 1; T.IstrT = T.IstrR; T.IstrU = my_Istr + 1;
    T.IstrB = T.IstrT + 1; T.IstrM = T.IstrP + 1;
    T.Istrm3 = std::max(0, my_Istr - 3); T.Istrm2 = std::max(0, my_Istr - 2);
    T.IstrUm2 = std::max(1, T.IstrU - 2); T.Istrm1 = std::max(1, my_Istr - 1);
    T.IstrUm1 = std::max(2, T.IstrU - 1);
  } else {
    T.IstrP = my_Istr; T.IstrR = my_Istr; T.IstrT = T.IstrR; T.IstrU = my_Istr;
    T.IstrB = my_Istr; T.IstrM = T.IstrU;
    T.Istrm3 = my_Istr - 3; T.Istrm2 = my_Istr - 2; T.IstrUm2 = T.IstrU - 2;
    T.Istrm1 = my_Istr - 1; T.IstrUm1 = T.IstrU - 1;
  }
  if (T.E && !EW) {
    T.IendR = my_Iend + 1; T.IendP = T.IendR; T.IendT = T.IendR; T.IendB = T.IendT - 1;
    T.Iendp1 = std::min(my_Iend + 1, Lm); T.Iendp2i = std::min(my_Iend + 2, Lm);
    T.Iendp2 = std::min(my_Iend + 2, Lm + 1); T.Iendp3 = std::min(my_Iend + 3, Lm + 1);
  } else {
    T.IendR = my_Iend; T.IendP = T.IendR; T.IendT = T.IendR; T.IendB = my_Iend;
    T.Iendp1 = my_Iend + 1; T.Iendp2i = my_Iend + 2; T.Iendp2 = my_Iend + 2; T.Iendp3 = my_Iend + 3;
  }
  // --- J direction
  T.Jstr = my_Jstr; T.Jend = my_Jend;
  if (T.S && !NS) {
    T.JstrP = my_Jstr; T.JstrR = my_Jstr - 1; T.JstrT = T.JstrR; T.JstrV = my_Jstr + 1;
    T.JstrB = T.JstrT + 1; T.JstrM = T.JstrP + 1;
    T.Jstrm3 = std::max(0, my_Jstr - 3); T.Jstrm2 = std::max(0, my_Jstr - 2);
    T.JstrVm2 = std::max(1, T.JstrV - 2); T.Jstrm1 = std::max(1, my_Jstr - 1);
    T.JstrVm1 = std::max(2, T.JstrV - 1);
  } else {
    T.JstrP = my_Jstr; T.JstrR = my_Jstr; T.JstrT = T.JstrR; T.JstrV = my_Jstr;
    T.JstrB = my_Jstr; T.JstrM = T.JstrV;
    T.Jstrm3 = my_Jstr - 3; T.Jstrm2 = my_Jstr - 2; T.JstrVm2 = T.JstrV - 2;
    T.Jstrm1 = my_Jstr - 1; T.JstrVm1 = T.JstrV - 1;
  }
  if (T.N && !NS) {
    T.JendR = my_Jend + 1; T.JendP = T.JendR; T.JendT = T.JendR; T.JendB = T.JendT - 1;
    T.Jendp1 = std::min(my_Jend + 1, Mm); T.Jendp2i = std::min(my_Jend + 2, Mm);
    T.Jendp2 = std::min(my_Jend + 2, Mm + 1); T.Jendp3 = std::min(my_Jend + 3, Mm + 1);
  } else {
    T.JendR = my_Jend; T.JendP = T.JendR; T.JendT = T.JendR; T.JendB = my_Jend;
    T.Jendp1 = my_Jend + 1; T.Jendp2i = my_Jend + 2; T.Jendp2 = my_Jend + 2; T.Jendp3 = my_Jend + 3;
  }
  // Include/tile.h:21-24
  T.IminS = T.Istr - 3; T.ImaxS = T.Iend + 3; T.JminS = T.Jstr - 3; T.JmaxS = T.Jend + 3;
  return T;
}

Model::Model(const Config& cfg) : c(cfg) {
  Lm = c.Lm; Mm = c.Mm; N = c.N; NT = c.NT; NAT = c.NAT;
  // Modules/mod_param.F:1633-1636
  int I_padd = (Lm + 2) / 2 - (Lm + 1) / 2, J_padd = (Mm + 2) / 2 - (Mm + 1) / 2;
  Im = Lm + I_padd; Jm = Mm + J_padd;
  EWperiodic = true; NSperiodic = false;   // roms_upwelling.in / roms_benchmark1.in :184-199
  // Utility/get_bounds.F:258-269 (serial / shared-memory allocation bounds, NghostPoints=2)
  LBi = -2; UBi = Im + 2; LBj = 0; UBj = Jm + 1;
  ni = UBi - LBi + 1; nj = UBj - LBj + 1;
  for (int t = 0; t < c.NtileI * c.NtileJ; ++t) tiles.push_back(make_tile(*this, t));

  static const char* two_d[] = {"h", "f", "fomn", "pm", "pn", "om_r", "on_r", "om_u", "on_u", "om_v", "on_v",
      "om_p", "on_p", "pmon_r", "pnom_r", "pmon_u", "pnom_u", "pmon_v", "pnom_v", "pmon_p", "pnom_p", "omn",
      "dndx", "dmde", "lonr", "latr", "xr", "yr", "angler", "rdrag", "rdrag2", "visc2_r", "visc2_p", "hsbl",
      "Jwtype", "Zt_avg1", "DU_avg1", "DU_avg2", "DV_avg1", "DV_avg2", "rufrc", "rvfrc", "rhoA", "rhoS",
      "alpha", "beta", "sustr", "svstr", "bustr", "bvstr", "srflx", "Uwind", "Vwind", "Tair", "Pair", "Hair",
      "cloud", "rain", "lrflx", "lhflx", "shflx"};
  for (const char* n : two_d) add(n, 1, 1);
  add("Hz", 1, N); add("z_r", 1, N); add("z_w", 0, N + 1); add("Huon", 1, N); add("Hvom", 1, N);
  add("diff2", 1, NT); add("Akv", 0, N + 1); add("bvf", 0, N + 1);
  add("Akt", 0, N + 1, NAT); add("ghats", 0, N + 1, NAT);
  add("zeta", 1, 3); add("ubar", 1, 3); add("vbar", 1, 3); add("rzeta", 1, 2); add("rubar", 1, 2); add("rvbar", 1, 2);
  add("rho", 1, N); add("pden", 1, N); add("W", 0, N + 1); add("wvel", 0, N + 1);
  add("u", 1, N, 2); add("v", 1, N, 2); add("ru", 0, N + 1, 2); add("rv", 0, N + 1, 2);
  add("t", 1, N, 3, NT);
  add("stflx", 1, NT); add("btflx", 1, NT); add("stflux", 1, NT); add("btflux", 1, NT);
  ksbl.assign((size_t)ni * nj, 0);

#define B2(x) x = v2(#x)
#define B3(x) x = v3(#x)
#define B4(x) x = v4(#x)
  B2(h); B2(f); B2(fomn); B2(pm); B2(pn); B2(om_r); B2(on_r); B2(om_u); B2(on_u); B2(om_v); B2(on_v); B2(om_p); B2(on_p);
  B2(pmon_r); B2(pnom_r); B2(pmon_u); B2(pnom_u); B2(pmon_v); B2(pnom_v); B2(pmon_p); B2(pnom_p); B2(omn);
  B2(dndx); B2(dmde); B2(lonr); B2(latr); B2(xr); B2(yr); B2(angler); B2(rdrag); B2(rdrag2);
  B3(Hz); B3(z_r); B3(z_w); B3(Huon); B3(Hvom);
  B2(visc2_r); B2(visc2_p); B2(hsbl); B2(Jwtype); B3(diff2); B3(Akv); B3(bvf); B4(Akt); B4(ghats);
  B2(Zt_avg1); B2(DU_avg1); B2(DU_avg2); B2(DV_avg1); B2(DV_avg2); B2(rufrc); B2(rvfrc); B2(rhoA); B2(rhoS);
  B3(zeta); B3(ubar); B3(vbar); B3(rzeta); B3(rubar); B3(rvbar); B3(rho); B3(pden); B3(W); B3(wvel);
  B4(u); B4(v); B4(ru); B4(rv); t = v5("t"); B2(alpha); B2(beta);
  B2(sustr); B2(svstr); B2(bustr); B2(bvstr); B2(srflx); B2(Uwind); B2(Vwind); B2(Tair); B2(Pair); B2(Hair);
  B2(cloud); B2(rain); B2(lrflx); B2(lhflx); B2(shflx); B3(stflx); B3(btflx); B3(stflux); B3(btflux);
#undef B2
#undef B3
#undef B4

  // Modules/mod_mixing.F:1430-1530 (initialize_mixing): background coefficients
  for (int k = 0; k <= N; ++k)
    for (int j = LBj; j <= UBj; ++j)
      for (int i = LBi; i <= UBi; ++i) {
        Akv(i, j, k) = c.Akv_bak;
        for (int it = 1; it <= NAT; ++it) Akt(i, j, k, it) = c.Akt_bak[it - 1];
      }
  for (int j = LBj; j <= UBj; ++j)
    for (int i = LBi; i <= UBi; ++i) {
      Jwtype(i, j) = (double)c.lmd_Jwt;                 // mod_mixing.F:1527
      rdrag(i, j) = c.rdrg; rdrag2(i, j) = c.rdrg2;     // mod_grid.F:1382-1384
    }
}

// ---------------------------------------------------------------------------
// Periodic exchanges, serial / shared-memory form.
// Nonlinear/exchange_2d.F:250-330 (r), :437-520 (u), :624-700 (v), p-type alike;
// Nonlinear/exchange_3d.F:280,492,704,917.  Only the E-W wrap can fire here.
// ---------------------------------------------------------------------------
static void ew_wrap(Model& M, const Tile& T, double* p, size_t plane, int nplanes, int Jmin, int Jmax) {
  if (!M.EWperiodic) return;
  assert(!M.NSperiodic);
  const int Lm = M.Lm, LBi = M.LBi, ni = M.ni, LBj = M.LBj;
  for (int k = 0; k < nplanes; ++k) {
    double* a = p + plane * k;
    auto A = [&](int i, int j) -> double& { return a[(i - LBi) + (size_t)ni * (j - LBj)]; };
    if (T.W)
      for (int j = Jmin; j <= Jmax; ++j) { A(Lm + 1, j) = A(1, j); A(Lm + 2, j) = A(2, j); }
    if (T.E)
      for (int j = Jmin; j <= Jmax; ++j) { A(-2, j) = A(Lm - 2, j); A(-1, j) = A(Lm - 1, j); A(0, j) = A(Lm, j); }
  }
}
void exchange_r2d(Model& M, const Tile& T, F2 A) { ew_wrap(M, T, A.p, 0, 1, T.JstrR, T.JendR); }
void exchange_u2d(Model& M, const Tile& T, F2 A) { ew_wrap(M, T, A.p, 0, 1, T.JstrR, T.JendR); }
void exchange_v2d(Model& M, const Tile& T, F2 A) { ew_wrap(M, T, A.p, 0, 1, T.Jstr, T.JendR); }
void exchange_p2d(Model& M, const Tile& T, F2 A) { ew_wrap(M, T, A.p, 0, 1, T.Jstr, T.JendR); }
void exchange_r3d(Model& M, const Tile& T, F3 A) { ew_wrap(M, T, A.p, (size_t)A.ni * A.nj, A.nk, T.JstrR, T.JendR); }
void exchange_u3d(Model& M, const Tile& T, F3 A) { ew_wrap(M, T, A.p, (size_t)A.ni * A.nj, A.nk, T.JstrR, T.JendR); }
void exchange_v3d(Model& M, const Tile& T, F3 A) { ew_wrap(M, T, A.p, (size_t)A.ni * A.nj, A.nk, T.Jstr, T.JendR); }
void exchange_w3d(Model& M, const Tile& T, F3 A) { ew_wrap(M, T, A.p, (size_t)A.ni * A.nj, A.nk, T.JstrR, T.JendR); }

// Nonlinear/bc_3d.F:588-723 (bc_w3d_tile): gradient condition on closed edges
void bc_w3d(Model& M, const Tile& T, F3 A) {
  if (!M.NSperiodic) {
    if (T.N) for (int k = A.LBk; k < A.LBk + A.nk; ++k) for (int i = T.Istr; i <= T.Iend; ++i) A(i, T.Jend + 1, k) = A(i, T.Jend, k);
    if (T.S) for (int k = A.LBk; k < A.LBk + A.nk; ++k) for (int i = T.Istr; i <= T.Iend; ++i) A(i, T.Jstr - 1, k) = A(i, T.Jstr, k);
  }
  exchange_w3d(M, T, A);
}
// Nonlinear/bc_2d.F (bc_r2d_tile): gradient condition
void bc_r2d(Model& M, const Tile& T, F2 A) {
  if (!M.NSperiodic) {
    if (T.N) for (int i = T.Istr; i <= T.Iend; ++i) A(i, T.Jend + 1) = A(i, T.Jend);
    if (T.S) for (int i = T.Istr; i <= T.Iend; ++i) A(i, T.Jstr - 1) = A(i, T.Jstr);
  }
  exchange_r2d(M, T, A);
}
// Nonlinear/bc_2d.F (bc_u2d_tile): closed walls -> gamma2 slip on tangential component
void bc_u2d(Model& M, const Tile& T, F2 A) {
  if (!M.NSperiodic) {
    if (T.N) for (int i = T.IstrU; i <= T.Iend; ++i) A(i, T.Jend + 1) = M.c.gamma2 * A(i, T.Jend);
    if (T.S) for (int i = T.IstrU; i <= T.Iend; ++i) A(i, T.Jstr - 1) = M.c.gamma2 * A(i, T.Jstr);
  }
  exchange_u2d(M, T, A);
}
// Nonlinear/bc_2d.F (bc_v2d_tile): closed walls -> zero normal component
void bc_v2d(Model& M, const Tile& T, F2 A) {
  if (!M.NSperiodic) {
    if (T.N) for (int i = T.Istr; i <= T.Iend; ++i) A(i, T.Jend + 1) = 0.0;
    if (T.S) for (int i = T.Istr; i <= T.Iend; ++i) A(i, T.Jstr) = 0.0;
  }
  exchange_v2d(M, T, A);
}
// Nonlinear/zetabc.F:353-358,437-442 (closed -> zero gradient)
void zetabc(Model& M, const Tile& T, int kout) {
  F3& zeta = M.zeta;
  if (T.S) for (int i = T.Istr; i <= T.Iend; ++i) zeta(i, T.Jstr - 1, kout) = zeta(i, T.Jstr, kout);
  if (T.N) for (int i = T.Istr; i <= T.Iend; ++i) zeta(i, T.Jend + 1, kout) = zeta(i, T.Jend, kout);
}
// Nonlinear/u2dbc_im.F:483-496 and northern analogue
void u2dbc(Model& M, const Tile& T, int kout) {
  F3& ubar = M.ubar;
  if (T.S) for (int i = T.IstrU; i <= T.Iend; ++i) ubar(i, T.Jstr - 1, kout) = M.c.gamma2 * ubar(i, T.Jstr, kout);
  if (T.N) for (int i = T.IstrU; i <= T.Iend; ++i) ubar(i, T.Jend + 1, kout) = M.c.gamma2 * ubar(i, T.Jend, kout);
}
// Nonlinear/v2dbc_im.F:253-258,395-400
void v2dbc(Model& M, const Tile& T, int kout) {
  F3& vbar = M.vbar;
  if (T.S) for (int i = T.Istr; i <= T.Iend; ++i) vbar(i, T.Jstr, kout) = 0.0;
  if (T.N) for (int i = T.Istr; i <= T.Iend; ++i) vbar(i, T.Jend + 1, kout) = 0.0;
}
// Nonlinear/t3dbc_im.F:334-341,415-422
void t3dbc(Model& M, const Tile& T, int nout, int itrc) {
  F5& t = M.t;
  if (T.S) for (int k = 1; k <= M.N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) t(i, T.Jstr - 1, k, nout, itrc) = t(i, T.Jstr, k, nout, itrc);
  if (T.N) for (int k = 1; k <= M.N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) t(i, T.Jend + 1, k, nout, itrc) = t(i, T.Jend, k, nout, itrc);
}
// Nonlinear/u3dbc_im.F:329-343,415-429
void u3dbc(Model& M, const Tile& T, int nout) {
  F4& u = M.u;
  if (T.S) for (int k = 1; k <= M.N; ++k) for (int i = T.IstrU; i <= T.Iend; ++i) u(i, T.Jstr - 1, k, nout) = M.c.gamma2 * u(i, T.Jstr, k, nout);
  if (T.N) for (int k = 1; k <= M.N; ++k) for (int i = T.IstrU; i <= T.Iend; ++i) u(i, T.Jend + 1, k, nout) = M.c.gamma2 * u(i, T.Jend, k, nout);
}
// Nonlinear/v3dbc_im.F:171-178,250-257
void v3dbc(Model& M, const Tile& T, int nout) {
  F4& v = M.v;
  if (T.S) for (int k = 1; k <= M.N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) v(i, T.Jstr, k, nout) = 0.0;
  if (T.N) for (int k = 1; k <= M.N; ++k) for (int i = T.Istr; i <= T.Iend; ++i) v(i, T.Jend + 1, k, nout) = 0.0;
}

}  // namespace orc
