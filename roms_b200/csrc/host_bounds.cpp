// roms_b200/csrc/host_bounds.cpp -- host-side index contract: the per-tile
// integers the reference keeps in BOUNDS(ng) (Utility/get_bounds.F:777-1884,
// Include/set_bounds.h:30-74).  In a Fortran build these come straight from
// BOUNDS(ng)%X(tile); this helper exists so C/C++/Python hosts (tests, bench,
// the C++ driver) can fill roms_b200_bounds the same way.
#include <algorithm>
#include <cstring>
#include "../../include/roms_b200.h"

extern "C" int roms_b200_tile_bounds(int Lm, int Mm, int N, int NT, int NAT, int NtileI, int NtileJ, int tile,
                                     int EWperiodic, int NSperiodic, int distributed, roms_b200_bounds* o) {
  if (!o || NtileI < 1 || NtileJ < 1 || tile < 0 || tile >= NtileI * NtileJ) return 1;
  std::memset(o, 0, sizeof(*o));
  o->Lm = Lm; o->Mm = Mm; o->N = N; o->NT = NT; o->NAT = NAT;
  o->EWperiodic = EWperiodic; o->NSperiodic = NSperiodic; o->NtileI = NtileI; o->NtileJ = NtileJ;
  // tile_bounds_2d, get_bounds.F:1020-1039
  const int ChunkI = (Lm + NtileI - 1) / NtileI, ChunkJ = (Mm + NtileJ - 1) / NtileJ;
  const int MarginI = (NtileI * ChunkI - Lm) / 2, MarginJ = (NtileJ * ChunkJ - Mm) / 2;
  const int Jt = tile / NtileI, It = tile - Jt * NtileI;
  o->Itile = It; o->Jtile = Jt;
  int Is = std::max(1 + It * ChunkI - MarginI, 1), Ie = std::min(1 + It * ChunkI - MarginI + ChunkI - 1, Lm);
  int Js = std::max(1 + Jt * ChunkJ - MarginJ, 1), Je = std::min(1 + Jt * ChunkJ - MarginJ + ChunkJ - 1, Mm);
  const bool W = (It == 0), E = (It == NtileI - 1), S = (Jt == 0), Nn = (Jt == NtileJ - 1);
  o->Western_Edge = W; o->Eastern_Edge = E; o->Southern_Edge = S; o->Northern_Edge = Nn;
  // array bounds: mod_param.F:1633-1636 padding, get_bounds.F:193-212 (distributed) / :258-269 (serial)
  const int Im = Lm + ((Lm + 2) / 2 - (Lm + 1) / 2), Jm = Mm + ((Mm + 2) / 2 - (Mm + 1) / 2);
  const int Ng = (distributed > 2) ? distributed : 2;     // NghostPoints (Utility/inp_par.F:211-226); >2 for a wider device mirror
  const int Imin = EWperiodic ? -Ng : 0, Imax = EWperiodic ? Im + Ng : Im + 1;
  const int Jmin = NSperiodic ? -Ng : 0, Jmax = NSperiodic ? Jm + Ng : Jm + 1;
  if (distributed) {
    o->LBi = (W && !EWperiodic) ? Imin : Is - Ng; o->UBi = (E && !EWperiodic) ? Imax : Ie + Ng;
    o->LBj = S ? Jmin : Js - Ng; o->UBj = Nn ? Jmax : Je + Ng;
  } else { o->LBi = Imin; o->UBi = Imax; o->LBj = Jmin; o->UBj = Jmax; }
  // var_bounds, get_bounds.F:1044-1884
  o->Istr = Is; o->Iend = Ie; o->Jstr = Js; o->Jend = Je;
  if (W && !EWperiodic) {
    o->IstrP = Is; o->IstrR = Is - 1; o->IstrT = o->IstrR; o->IstrU = Is + 1; o->IstrB = o->IstrT + 1; o->IstrM = o->IstrP + 1;
    o->Istrm3 = std::max(0, Is - 3); o->Istrm2 = std::max(0, Is - 2); o->IstrUm2 = std::max(1, o->IstrU - 2);
    o->Istrm1 = std::max(1, Is - 1); o->IstrUm1 = std::max(2, o->IstrU - 1);
  } else {
    o->IstrP = Is; o->IstrR = Is; o->IstrT = Is; o->IstrU = Is; o->IstrB = Is; o->IstrM = Is;
    o->Istrm3 = Is - 3; o->Istrm2 = Is - 2; o->IstrUm2 = Is - 2; o->Istrm1 = Is - 1; o->IstrUm1 = Is - 1;
  }
  if (E && !EWperiodic) {
    o->IendR = Ie + 1; o->IendP = o->IendR; o->IendT = o->IendR; o->IendB = o->IendT - 1;
    o->Iendp1 = std::min(Ie + 1, Lm); o->Iendp2i = std::min(Ie + 2, Lm); o->Iendp2 = std::min(Ie + 2, Lm + 1); o->Iendp3 = std::min(Ie + 3, Lm + 1);
  } else {
    o->IendR = Ie; o->IendP = Ie; o->IendT = Ie; o->IendB = Ie;
    o->Iendp1 = Ie + 1; o->Iendp2i = Ie + 2; o->Iendp2 = Ie + 2; o->Iendp3 = Ie + 3;
  }
  if (S && !NSperiodic) {
    o->JstrP = Js; o->JstrR = Js - 1; o->JstrT = o->JstrR; o->JstrV = Js + 1; o->JstrB = o->JstrT + 1; o->JstrM = o->JstrP + 1;
    o->Jstrm3 = std::max(0, Js - 3); o->Jstrm2 = std::max(0, Js - 2); o->JstrVm2 = std::max(1, o->JstrV - 2);
    o->Jstrm1 = std::max(1, Js - 1); o->JstrVm1 = std::max(2, o->JstrV - 1);
  } else {
    o->JstrP = Js; o->JstrR = Js; o->JstrT = Js; o->JstrV = Js; o->JstrB = Js; o->JstrM = Js;
    o->Jstrm3 = Js - 3; o->Jstrm2 = Js - 2; o->JstrVm2 = Js - 2; o->Jstrm1 = Js - 1; o->JstrVm1 = Js - 1;
  }
  if (Nn && !NSperiodic) {
    o->JendR = Je + 1; o->JendP = o->JendR; o->JendT = o->JendR; o->JendB = o->JendT - 1;
    o->Jendp1 = std::min(Je + 1, Mm); o->Jendp2i = std::min(Je + 2, Mm); o->Jendp2 = std::min(Je + 2, Mm + 1); o->Jendp3 = std::min(Je + 3, Mm + 1);
  } else {
    o->JendR = Je; o->JendP = Je; o->JendT = Je; o->JendB = Je;
    o->Jendp1 = Je + 1; o->Jendp2i = Je + 2; o->Jendp2 = Je + 2; o->Jendp3 = Je + 3;
  }
  return 0;
}

// tile_neighbors (Utility/mp_exchange.F:73-197): ranks of the W,E,S,N neighbours of this tile
// (rank = Jtile*NtileI+Itile); -1 where there is none.  E-W wraps periodically; an axis with a
// single tile has no neighbour (the periodic images are written locally).
extern "C" int roms_b200_tile_neighbors(const roms_b200_bounds* b, int* wesn) {
  if (!b || !wesn) return 1;
  const int It = b->Itile, Jt = b->Jtile, NI = b->NtileI, NJ = b->NtileJ;
  wesn[0] = (NI > 1 && (b->EWperiodic || It > 0)) ? Jt * NI + (It - 1 + NI) % NI : -1;
  wesn[1] = (NI > 1 && (b->EWperiodic || It < NI - 1)) ? Jt * NI + (It + 1) % NI : -1;
  wesn[2] = (NJ > 1 && (b->NSperiodic || Jt > 0)) ? ((Jt - 1 + NJ) % NJ) * NI + It : -1;
  wesn[3] = (NJ > 1 && (b->NSperiodic || Jt < NJ - 1)) ? ((Jt + 1) % NJ) * NI + It : -1;
  return 0;
}

// Single-phase eight-neighbour halo plan of the NVLink mailbox transport (k_halo.cu).  dir: 0 W, 1 E, 2 S, 3 N, 4 SW, 5 SE,
// 6 NW, 7 NE.  ranks8[d] = neighbour tile (or -1); snd/rcv[d] = {i0, i1, j0, j1} (Fortran indices) of what this tile sends
// towards d / where the block arriving from d goes.  W/E strips span Jstr..Jend extended to the array edge where the tile has
// no S/N neighbour (physical boundary rows), S/N strips likewise in i, corners are (w x w) blocks: after ONE exchange the halo
// frame holds what the two phases of mp_exchange (W/E, then S/N over the full i-range; mp_exchange.F:520-532,761-773) produce.
extern "C" int roms_b200_halo_plan(const roms_b200_bounds* b, int w, int* ranks8, int* snd, int* rcv) {
  if (!b || !ranks8 || !snd || !rcv) return 1;
  int nb[4];
  roms_b200_tile_neighbors(b, nb);
  const int NI = b->NtileI, me = b->Jtile * NI + b->Itile;
  for (int q = 0; q < 4; ++q) ranks8[q] = (nb[q] == me) ? -1 : nb[q];
  for (int q = 4; q < 8; ++q) ranks8[q] = -1;
  if (ranks8[2] >= 0 && ranks8[0] >= 0) ranks8[4] = (ranks8[2] / NI) * NI + ranks8[0] % NI;
  if (ranks8[2] >= 0 && ranks8[1] >= 0) ranks8[5] = (ranks8[2] / NI) * NI + ranks8[1] % NI;
  if (ranks8[3] >= 0 && ranks8[0] >= 0) ranks8[6] = (ranks8[3] / NI) * NI + ranks8[0] % NI;
  if (ranks8[3] >= 0 && ranks8[1] >= 0) ranks8[7] = (ranks8[3] / NI) * NI + ranks8[1] % NI;
  const bool hW = ranks8[0] >= 0, hE = ranks8[1] >= 0, hS = ranks8[2] >= 0, hN = ranks8[3] >= 0;
  const int ia = hW ? b->Istr : b->LBi, ib = hE ? b->Iend : b->UBi, ja = hS ? b->Jstr : b->LBj, jb = hN ? b->Jend : b->UBj;
  const int iW0 = b->Istr, iW1 = b->Istr + w - 1, iE0 = b->Iend - w + 1, iE1 = b->Iend;
  const int jS0 = b->Jstr, jS1 = b->Jstr + w - 1, jN0 = b->Jend - w + 1, jN1 = b->Jend;
  const int gW0 = b->Istr - w, gW1 = b->Istr - 1, gE0 = b->Iend + 1, gE1 = b->Iend + w;
  const int gS0 = b->Jstr - w, gS1 = b->Jstr - 1, gN0 = b->Jend + 1, gN1 = b->Jend + w;
  const int S[8][4] = {{iW0, iW1, ja, jb}, {iE0, iE1, ja, jb}, {ia, ib, jS0, jS1}, {ia, ib, jN0, jN1},
                       {iW0, iW1, jS0, jS1}, {iE0, iE1, jS0, jS1}, {iW0, iW1, jN0, jN1}, {iE0, iE1, jN0, jN1}};
  const int R[8][4] = {{gW0, gW1, ja, jb}, {gE0, gE1, ja, jb}, {ia, ib, gS0, gS1}, {ia, ib, gN0, gN1},
                       {gW0, gW1, gS0, gS1}, {gE0, gE1, gS0, gS1}, {gW0, gW1, gN0, gN1}, {gE0, gE1, gN0, gN1}};
  for (int d = 0; d < 8; ++d) for (int q = 0; q < 4; ++q) { snd[4 * d + q] = S[d][q]; rcv[4 * d + q] = R[d][q]; }
  return 0;
}
