// oracle/capi.cpp -- TEST INFRASTRUCTURE ONLY (see oracle.h header).
// Flat C entry points so tests/ and bench.py (cpu_baseline / --impl reference)
// can drive the restatement through ctypes.
#include "oracle.h"
#include <cstring>

using namespace orc;

extern "C" {

void* orc_create(int app, int Lm, int Mm, int N, int NtileI, int NtileJ) {
  Config c = (app == UPWELLING) ? config_upwelling() : config_benchmark(Lm, Mm, N);
  if (app == UPWELLING && Lm > 0) { c.Lm = Lm; c.Mm = Mm; c.N = N; }
  c.NtileI = NtileI; c.NtileJ = NtileJ;
  return new Model(c);
}
void orc_destroy(void* h) { delete (Model*)h; }
void orc_set_dt(void* h, double dt, int ndtfast) { Model* M = (Model*)h; M->c.dt = dt; M->c.ndtfast = ndtfast; }
void orc_set_threads(void* h, int n) { ((Model*)h)->nthreads = n; }
void orc_initial(void* h) { initial(*(Model*)h); }
void orc_step(void* h, int n) { for (int i = 0; i < n; ++i) main3d_step(*(Model*)h); }
int orc_phase(void* h, const char* ph) { try { main3d_phase(*(Model*)h, ph); } catch (...) { return 1; } return 0; }

long orc_field_size(void* h, const char* name) {
  Model* M = (Model*)h; auto it = M->fields.find(name);
  return it == M->fields.end() ? -1 : (long)it->second.d.size();
}
int orc_get(void* h, const char* name, double* out) {
  Model* M = (Model*)h; auto it = M->fields.find(name); if (it == M->fields.end()) return 1;
  std::memcpy(out, it->second.d.data(), it->second.d.size() * sizeof(double)); return 0;
}
int orc_set(void* h, const char* name, const double* in) {
  Model* M = (Model*)h; auto it = M->fields.find(name); if (it == M->fields.end()) return 1;
  std::memcpy(it->second.d.data(), in, it->second.d.size() * sizeof(double)); return 0;
}
// dims: LBi UBi LBj UBj N NT NAT Lm Mm nfast ndtfast
void orc_get_dims(void* h, int* o) {
  Model* M = (Model*)h;
  o[0] = M->LBi; o[1] = M->UBi; o[2] = M->LBj; o[3] = M->UBj; o[4] = M->N; o[5] = M->NT; o[6] = M->NAT;
  o[7] = M->Lm; o[8] = M->Mm; o[9] = M->nfast; o[10] = M->c.ndtfast;
}
// stepping state: iic ntfirst nstp nnew nrhs kstp knew krhs indx1 iif predictor
void orc_get_stepping(void* h, int* o) {
  Model* M = (Model*)h;
  o[0] = M->iic; o[1] = M->ntfirst; o[2] = M->nstp; o[3] = M->nnew; o[4] = M->nrhs; o[5] = M->kstp; o[6] = M->knew;
  o[7] = M->krhs; o[8] = M->indx1; o[9] = M->iif; o[10] = M->PREDICTOR_2D_STEP ? 1 : 0;
}
void orc_set_stepping(void* h, const int* o) {
  Model* M = (Model*)h;
  M->iic = o[0]; M->ntfirst = o[1]; M->nstp = o[2]; M->nnew = o[3]; M->nrhs = o[4]; M->kstp = o[5]; M->knew = o[6];
  M->krhs = o[7]; M->indx1 = o[8]; M->iif = o[9]; M->PREDICTOR_2D_STEP = (o[10] != 0);
}
// scalars: dt dtfast hc time tdays avgke avgpe volume
void orc_get_scalars(void* h, double* o) {
  Model* M = (Model*)h;
  o[0] = M->c.dt; o[1] = M->dtfast; o[2] = M->hc; o[3] = M->time; o[4] = M->tdays; o[5] = M->avgke; o[6] = M->avgpe; o[7] = M->volume;
}
// vectors: sc_r sc_w Cs_r Cs_w (N+1 each, index k) ; weight1 weight2 (2*ndtfast+2 each, 1-based)
int orc_get_vec(void* h, const char* name, double* out) {
  Model* M = (Model*)h; const std::vector<double>* v = nullptr; std::string n(name);
  if (n == "sc_r") v = &M->sc_r; else if (n == "sc_w") v = &M->sc_w; else if (n == "Cs_r") v = &M->Cs_r;
  else if (n == "Cs_w") v = &M->Cs_w; else if (n == "weight1") v = &M->weight1; else if (n == "weight2") v = &M->weight2;
  if (!v) return -1;
  std::memcpy(out, v->data(), v->size() * sizeof(double)); return (int)v->size();
}
// full diag: avgke avgpe volume max_C max_Cu max_Cv max_Cw max_Ci max_Cj max_Ck maxspeed maxrho exit_flag
void orc_get_diag(void* h, double* o) {
  Model* M = (Model*)h;
  o[0] = M->avgke; o[1] = M->avgpe; o[2] = M->volume; o[3] = M->max_C; o[4] = M->max_Cu; o[5] = M->max_Cv; o[6] = M->max_Cw;
  o[7] = M->max_Ci; o[8] = M->max_Cj; o[9] = M->max_Ck; o[10] = M->maxspeed; o[11] = M->maxrho; o[12] = M->exit_flag;
}
void orc_get_ksbl(void* h, int* out) { Model* M = (Model*)h; std::memcpy(out, M->ksbl.data(), M->ksbl.size() * sizeof(int)); }
// tile bounds as ints, in the order of include/roms_b200.h: roms_b200_bounds
// the ints of one tile in the order: Istr Iend Jstr Jend IstrR IendR JstrR JendR IstrU JstrV IstrP IendP JstrP JendP
// IstrT IendT JstrT JendT IstrB IendB JstrB JendB IstrM JstrM Istrm3 Istrm2 Istrm1 IstrUm2 IstrUm1 Iendp1 Iendp2 Iendp2i Iendp3
// Jstrm3 Jstrm2 Jstrm1 JstrVm2 JstrVm1 Jendp1 Jendp2 Jendp2i Jendp3 W E S N
void orc_get_tile(void* h, int tile, int* o) {
  const Tile& T = ((Model*)h)->tiles[tile];
  const int v[] = {T.Istr, T.Iend, T.Jstr, T.Jend, T.IstrR, T.IendR, T.JstrR, T.JendR, T.IstrU, T.JstrV, T.IstrP, T.IendP, T.JstrP, T.JendP,
                   T.IstrT, T.IendT, T.JstrT, T.JendT, T.IstrB, T.IendB, T.JstrB, T.JendB, T.IstrM, T.JstrM, T.Istrm3, T.Istrm2, T.Istrm1,
                   T.IstrUm2, T.IstrUm1, T.Iendp1, T.Iendp2, T.Iendp2i, T.Iendp3, T.Jstrm3, T.Jstrm2, T.Jstrm1, T.JstrVm2, T.JstrVm1,
                   T.Jendp1, T.Jendp2, T.Jendp2i, T.Jendp3, T.W, T.E, T.S, T.N};
  for (size_t q = 0; q < sizeof(v) / sizeof(int); ++q) o[q] = v[q];
}
int orc_ntiles(void* h) { return (int)((Model*)h)->tiles.size(); }

}  // extern "C"
