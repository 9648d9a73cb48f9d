// oracle/oracle.h -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement (plain fp64 loops, no FMA contraction, no fast-math) of the
// ROMS nonlinear main3d hot path for the UPWELLING and BENCHMARK option sets.
// Every routine follows the cpp-active Fortran of /root/reference statement by
// statement (same loop bounds, same operation order); the file:line of the
// Fortran it follows is cited at each function.
//
// PARITY UNPINNED, with one exception: the reference ships no test data and no
// expected logs for this path, and no Fortran compiler
// exists in this image, so this restatement cannot be checked against reference
// output.  The exception is the only known-answer vector in the reference
// sources for this path, the "Check Values" of the equation of state in the
// header of ROMS/Nonlinear/rho_eos.F:21-29 (T=3, S=35.5, Z=-5000 m): rho_eos()
// reproduces den, den1, alpha and beta to the 14 digits printed there
// (tests/test_cpu.py).  Other pins: tiling invariance (1x1 == 2x2 == 4x2
// bit-for-bit, the reference's own verify.sh criterion), volume/tracer
// conservation and self-consistency tests under tests/.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
// legs may load this library.  The product (roms_b200/) never links it.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cmath>
#include <map>
#include <string>
#include <vector>

namespace orc {

// ---- Fortran-style array views (i fastest, explicit lower bounds) ----------
struct F2 {
  double* p = nullptr; int LBi = 0, ni = 0, LBj = 0, nj = 0;
  inline double& operator()(int i, int j) const {
    return p[(i - LBi) + (size_t)ni * (j - LBj)];
  }
};
struct F3 {  // (i,j,k) with k lower bound LBk ; also used for (i,j,timelevel)
  double* p = nullptr; int LBi = 0, ni = 0, LBj = 0, nj = 0, LBk = 1, nk = 0;
  inline double& operator()(int i, int j, int k) const {
    return p[(i - LBi) + (size_t)ni * ((j - LBj) + (size_t)nj * (k - LBk))];
  }
  inline F2 slab(int k) const {
    return F2{p + (size_t)ni * nj * (k - LBk), LBi, ni, LBj, nj};
  }
};
struct F4 {  // (i,j,k,l): l is 1-based (time level or tracer)
  double* p = nullptr; int LBi = 0, ni = 0, LBj = 0, nj = 0, LBk = 1, nk = 0, nl = 0;
  inline double& operator()(int i, int j, int k, int l) const {
    return p[(i - LBi) + (size_t)ni * ((j - LBj) + (size_t)nj * ((k - LBk) + (size_t)nk * (l - 1)))];
  }
  inline F3 vol(int l) const {
    return F3{p + (size_t)ni * nj * nk * (l - 1), LBi, ni, LBj, nj, LBk, nk};
  }
};
struct F5 {  // t(i,j,k,l,itrc)
  double* p = nullptr; int LBi = 0, ni = 0, LBj = 0, nj = 0, nk = 0, nl = 0, nm = 0;
  inline double& operator()(int i, int j, int k, int l, int m) const {
    return p[(i - LBi) + (size_t)ni * ((j - LBj) + (size_t)nj * ((k - 1) + (size_t)nk * ((l - 1) + (size_t)nl * (m - 1))))];
  }
  inline F3 vol(int l, int m) const {
    return F3{p + (size_t)ni * nj * nk * ((l - 1) + (size_t)nl * (m - 1)), LBi, ni, LBj, nj, 1, nk};
  }
};

// private per-call scratch, sized like the Fortran automatic arrays
struct S2 {
  std::vector<double> d; int lo1, n1, lo2;
  S2(int a0, int a1, int b0, int b1) : d((size_t)(a1 - a0 + 1) * (b1 - b0 + 1), 0.0), lo1(a0), n1(a1 - a0 + 1), lo2(b0) {}
  inline double& operator()(int i, int j) { return d[(i - lo1) + (size_t)n1 * (j - lo2)]; }
};
struct S3 {
  std::vector<double> d; int lo1, n1, lo2, n2, lo3;
  S3(int a0, int a1, int b0, int b1, int c0, int c1)
      : d((size_t)(a1 - a0 + 1) * (b1 - b0 + 1) * (c1 - c0 + 1), 0.0), lo1(a0), n1(a1 - a0 + 1), lo2(b0), n2(b1 - b0 + 1), lo3(c0) {}
  inline double& operator()(int i, int j, int k) { return d[(i - lo1) + (size_t)n1 * ((j - lo2) + (size_t)n2 * (k - lo3))]; }
};

// ---- run-time configuration (roms_*.in + application header) ---------------
enum App { UPWELLING = 0, BENCHMARK = 1 };

struct Config {
  int app = UPWELLING;
  int Lm = 41, Mm = 80, N = 16, NT = 2, NAT = 2;
  int NtileI = 1, NtileJ = 1;
  double dt = 300.0; int ndtfast = 30;
  // physical parameters
  double rho0 = 1025.0, g = 9.81;
  double theta_s = 3.0, theta_b = 0.0, Tcline = 25.0;
  double rdrg = 3.0e-4, rdrg2 = 3.0e-3;
  double Akt_bak[2] = {1.0e-6, 1.0e-6}, Akv_bak = 1.0e-5;
  double tnu2[2] = {0.0, 0.0}, visc2 = 5.0;
  double R0 = 1027.0, T0 = 14.0, S0 = 35.0, Tcoef = 1.7e-4, Scoef = 0.0;
  double gamma2 = 1.0;
  double blk_ZQ = 10.0, blk_ZT = 10.0, blk_ZW = 10.0;
  int lmd_Jwt = 1;
  double dstart = 0.0;
};
Config config_upwelling();
Config config_benchmark(int Lm, int Mm, int N);

// ---- per-tile loop bounds (Utility/get_bounds.F:777-1884) ------------------
struct Tile {
  int Itile, Jtile;
  bool W, E, S, N;  // DOMAIN%Western_Edge ... Northern_Edge
  int Istr, Iend, Jstr, Jend;
  int IstrR, IendR, JstrR, JendR;
  int IstrU, JstrV;
  int IstrP, IendP, JstrP, JendP;
  int IstrT, IendT, JstrT, JendT;
  int IstrB, IendB, JstrB, JendB;
  int IstrM, JstrM;
  int Istrm3, Istrm2, Istrm1, IstrUm2, IstrUm1, Iendp1, Iendp2, Iendp2i, Iendp3;
  int Jstrm3, Jstrm2, Jstrm1, JstrVm2, JstrVm1, Jendp1, Jendp2, Jendp2i, Jendp3;
  int IminS, ImaxS, JminS, JmaxS;
};

struct Field {
  std::vector<double> d;
  int kLB = 1, nk = 1, nl = 1, nm = 1;  // extents beyond (i,j)
};

struct Model {
  Config c;
  bool EWperiodic = true, NSperiodic = false;
  int Lm, Mm, N, NT, NAT, Im, Jm;
  int LBi, UBi, LBj, UBj, ni, nj;
  std::vector<Tile> tiles;
  std::map<std::string, Field> fields;

  // S-coordinate (Utility/set_scoord.F)
  double hc = 0.0;
  std::vector<double> sc_r, sc_w, Cs_r, Cs_w;  // index k (sc_r[0] unused)
  // fast-time filter (Utility/set_weights.F)
  int nfast = 0; double dtfast = 0.0;
  std::vector<double> weight1, weight2;  // 1-based index
  // stepping (Modules/mod_stepping.F, Nonlinear/initial.F:130-160)
  int iic = 0, ntstart = 1, ntfirst = 1;
  int nstp = 1, nnew = 1, nrhs = 1;
  int kstp = 1, knew = 1, krhs = 1, indx1 = 1, iif = 1;
  bool PREDICTOR_2D_STEP = false;
  double time = 0.0, tdays = 0.0;
  // diagnostics (Nonlinear/diag.F)
  double avgke = 0, avgpe = 0, volume = 0;
  double max_C = 0, max_Cu = 0, max_Cv = 0, max_Cw = 0, maxspeed = 0, maxrho = 0;   // diag.F:211-266
  int max_Ci = 0, max_Cj = 0, max_Ck = 0;
  int exit_flag = 0;                                                                  // mod_scalars.F:548-561 (1 = blow-up)
  int nthreads = 1;   // >1: tiles of one tile loop run concurrently (reference's OpenMP mode)

  // grid (mod_grid)
  F2 h, f, fomn, pm, pn, om_r, on_r, om_u, on_u, om_v, on_v, om_p, on_p;
  F2 pmon_r, pnom_r, pmon_u, pnom_u, pmon_v, pnom_v, pmon_p, pnom_p, omn;
  F2 dndx, dmde, lonr, latr, xr, yr, angler, rdrag, rdrag2;
  F3 Hz, z_r, z_w, Huon, Hvom;
  // mixing (mod_mixing)
  F2 visc2_r, visc2_p, hsbl, Jwtype;
  F3 diff2, Akv, bvf;
  F4 Akt, ghats;
  std::vector<int> ksbl;
  // coupling (mod_coupling)
  F2 Zt_avg1, DU_avg1, DU_avg2, DV_avg1, DV_avg2, rufrc, rvfrc, rhoA, rhoS;
  // ocean (mod_ocean)
  F3 zeta, ubar, vbar, rzeta, rubar, rvbar, rho, pden, W, wvel;
  F4 u, v, ru, rv;
  F5 t;
  F2 alpha, beta;
  // forces (mod_forces)
  F2 sustr, svstr, bustr, bvstr, srflx, Uwind, Vwind, Tair, Pair, Hair, cloud, rain;
  F2 lrflx, lhflx, shflx;
  F3 stflx, btflx, stflux, btflux;

  explicit Model(const Config& cfg);
  Field& add(const std::string& name, int kLB, int nk, int nl = 1, int nm = 1);
  F2 v2(const std::string& n); F3 v3(const std::string& n); F4 v4(const std::string& n); F5 v5(const std::string& n);
};

// ---- routines (one per reference _tile routine) ----------------------------
// index helpers / boundary + periodic exchanges
void exchange_r2d(Model& M, const Tile& T, F2 A);
void exchange_u2d(Model& M, const Tile& T, F2 A);
void exchange_v2d(Model& M, const Tile& T, F2 A);
void exchange_p2d(Model& M, const Tile& T, F2 A);
void exchange_r3d(Model& M, const Tile& T, F3 A);
void exchange_u3d(Model& M, const Tile& T, F3 A);
void exchange_v3d(Model& M, const Tile& T, F3 A);
void exchange_w3d(Model& M, const Tile& T, F3 A);
void bc_w3d(Model& M, const Tile& T, F3 A);
void bc_r2d(Model& M, const Tile& T, F2 A);
void bc_u2d(Model& M, const Tile& T, F2 A);
void bc_v2d(Model& M, const Tile& T, F2 A);
void zetabc(Model& M, const Tile& T, int kout);
void u2dbc(Model& M, const Tile& T, int kout);
void v2dbc(Model& M, const Tile& T, int kout);
void t3dbc(Model& M, const Tile& T, int nout, int itrc);
void u3dbc(Model& M, const Tile& T, int nout);
void v3dbc(Model& M, const Tile& T, int nout);

// initialisation
void set_scoord(Model& M);
void set_weights(Model& M);
void ana_grid(Model& M, const Tile& T);
void metrics(Model& M, const Tile& T);
void ini_hmixcoef(Model& M, const Tile& T);
void ana_initial(Model& M, const Tile& T);
void ini_zeta(Model& M, const Tile& T);
void ini_fields(Model& M, const Tile& T);
void set_zeta_timeavg(Model& M, const Tile& T);
void initial(Model& M);
void set_data(Model& M, const Tile& T);

// the hot path
void set_depth(Model& M, const Tile& T, F2 Zt);
void set_massflux(Model& M, const Tile& T);
void rho_eos(Model& M, const Tile& T);
void omega(Model& M, const Tile& T);
void wvelocity(Model& M, const Tile& T, int Ninp);
void set_zeta(Model& M, const Tile& T);
void bulk_flux(Model& M, const Tile& T);
void set_vbc(Model& M, const Tile& T);
void ana_vmix(Model& M, const Tile& T);
void lmd_vmix(Model& M, const Tile& T);   // lmd_vmix_tile + lmd_skpp_tile + lmd_finish_tile
void pre_step3d(Model& M, const Tile& T);
void prsgrd32(Model& M, const Tile& T);
void t3dmix2(Model& M, const Tile& T);
void rhs3d_tile(Model& M, const Tile& T);
void uv3dmix2(Model& M, const Tile& T);
void rhs3d(Model& M, const Tile& T);       // pre_step3d -> prsgrd -> t3dmix2 -> rhs3d_tile -> uv3dmix2
void step2d(Model& M, const Tile& T);
void step3d_uv(Model& M, const Tile& T);
void step3d_t(Model& M, const Tile& T);
void diag(Model& M);

// one baroclinic step (Nonlinear/main3d.F:189-1158)
void main3d_step(Model& M);
// named phases of main3d_step for per-kernel parity tests
void main3d_phase(Model& M, const std::string& phase);

}  // namespace orc
