// roms_b200/csrc/roms_b200.cu -- C ABI of libroms_b200.so: device mirror,
// kernel entry points, the barotropic fast loop (CUDA graph) and the main3d
// sequencing.  See include/roms_b200.h for the contract.
#include "common.cuh"
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <string>
#include <vector>

namespace {
struct FieldInfo { const char* name; int kLB; const char* nk; const char* nl; const char* nm; };
#define X(name, kLB, nk, nl, nm) {#name, kLB, #nk, #nl, #nm},
const FieldInfo kFields[ROMS_B200_NFIELDS] = {ROMS_B200_FIELDS(X)};
#undef X
int resolve(const char* s, const roms_b200_bounds& b) {
  if (!strcmp(s, "N")) return b.N;
  if (!strcmp(s, "Np1")) return b.N + 1;
  if (!strcmp(s, "NT")) return b.NT;
  if (!strcmp(s, "NAT")) return b.NAT;
  return atoi(s);
}
template <typename T> int dev_alloc(T** p, size_t n) {
  CUDA_OK(cudaMalloc((void**)p, n * sizeof(T)));
  CUDA_OK(cudaMemset(*p, 0, n * sizeof(T)));
  return 0;
}
}  // namespace

extern "C" {

int roms_b200_field_id(const char* name) {
  for (int f = 0; f < ROMS_B200_NFIELDS; ++f) if (!strcmp(kFields[f].name, name)) return f;
  return -1;
}

static int create_impl(roms_b200_ctx* c, const roms_b200_bounds* b, const roms_b200_params* p, int device);
int roms_b200_create(const roms_b200_bounds* b, const roms_b200_params* p, int device, roms_b200_ctx** out) {
  if (!b || !p || !out) return 1;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "roms_b200: no CUDA device available; this library has no CPU path\n");
    return 2;
  }
  if (b->N > RB_MAXN) { fprintf(stderr, "roms_b200: N=%d exceeds RB_MAXN=%d\n", b->N, RB_MAXN); return 3; }
  if (!b->EWperiodic || b->NSperiodic) { fprintf(stderr, "roms_b200: only E-W periodic / N-S closed channels are supported\n"); return 3; }
  if (cudaSetDevice(device) != cudaSuccess) { fprintf(stderr, "roms_b200: cannot select CUDA device %d\n", device); return 2; }
  roms_b200_ctx* c = new roms_b200_ctx();
  memset(c, 0, sizeof(*c));
  c->device = device;
  const int rc = create_impl(c, b, p, device);
  if (rc) { roms_b200_destroy(c); return rc; }      // nothing of a half-built context survives a failed allocation
  *out = c;
  return 0;
}
static int create_impl(roms_b200_ctx* c, const roms_b200_bounds* b, const roms_b200_params* p, int device) {
  (void)device;
  Dev& D = c->D;
  D.b = *b; D.p = *p;
  D.ni = b->UBi - b->LBi + 1; D.nj = b->UBj - b->LBj + 1; D.nij = (size_t)D.ni * D.nj;
  D.wrapEW = (b->EWperiodic && b->NtileI == 1) ? 1 : 0;
  D.dist = (b->NtileI * b->NtileJ > 1) ? 1 : 0;
  D.halo = D.dist ? (b->Istr - b->LBi) : 2;
  if (D.dist && (D.halo < 3 || b->UBi - b->Iend < 3)) { fprintf(stderr, "roms_b200: distributed tiles need a mirror halo >= 3 (tile_bounds distributed=3)\n"); return 3; }
  {
    const int j0 = b->LBj > 0 ? b->LBj : 0, j1 = b->UBj < b->Mm + 1 ? b->UBj : b->Mm + 1;   // physical rows held by this tile
    if (D.dist) {
      const int i0 = D.wrapEW ? b->IstrT : b->LBi, i1 = D.wrapEW ? b->IendT : b->UBi;       // one tile in i: periodic images via st()
      D.rI0 = i0; D.rI1 = i1; D.rJ0 = j0; D.rJ1 = j1;
      D.uI0 = D.wrapEW ? b->IstrP : b->LBi + 1; D.vJ0 = (b->LBj + 1 > 1) ? b->LBj + 1 : 1;
      D.oI0 = D.wrapEW ? b->Istr : b->LBi + 1; D.oI1 = D.wrapEW ? b->Iend : b->UBi - 1;
      D.oJ0 = (b->LBj + 1 > 1) ? b->LBj + 1 : 1; D.oJ1 = (b->UBj - 1 < b->Mm) ? b->UBj - 1 : b->Mm;
    } else {
      D.rI0 = b->IstrT; D.rI1 = b->IendT; D.rJ0 = b->JstrT; D.rJ1 = b->JendT;
      D.uI0 = b->IstrP; D.vJ0 = b->JstrP;
      D.oI0 = b->Istr; D.oI1 = b->Iend; D.oJ0 = b->Jstr; D.oJ1 = b->Jend;
    }
  }
  CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUDA_OK(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
  CUDA_OK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  for (int q = 0; q < 6; ++q) CUDA_OK(cudaEventCreateWithFlags(&c->ev[q], cudaEventDisableTiming));
  c->two_streams = (getenv("ROMS_B200_ONE_STREAM") == nullptr);
  for (int f = 0; f < ROMS_B200_NFIELDS; ++f) {
    const int nk = resolve(kFields[f].nk, *b), nl = resolve(kFields[f].nl, *b), nm = resolve(kFields[f].nm, *b);
    D.kLB[f] = kFields[f].kLB; D.nk[f] = nk; D.nl[f] = nl;
    c->fsize[f] = D.nij * nk * nl * nm;
    if (dev_alloc(&D.f[f], c->fsize[f])) return 4;
  }
  double* v;
  if (dev_alloc(&v, (size_t)(b->N + 1) * 4)) return 4;
  D.sc_r = v; D.Cs_r = v + (b->N + 1); D.sc_w = v + 2 * (b->N + 1); D.Cs_w = v + 3 * (b->N + 1);
  if (dev_alloc(&v, (size_t)2 * (2 * p->ndtfast + 4))) return 4;
  D.w1 = v; D.w2 = v + (2 * p->ndtfast + 4);
  if (dev_alloc(&D.P, D.nij * b->N)) return 4;
  if (dev_alloc(&D.scratch2, D.nij * 12)) return 4;
  // four 3-D scratch volumes (ni,nj,0:N): KPP {dR, dU, dV, FC}, t3dmix2_geo dTdz per tracer, uv3dmix2's rufrc/rvfrc terms
  if (dev_alloc(&D.kpp4, D.nij * (size_t)(b->N + 1) * 4)) return 4;
  if (p->app == ROMS_B200_APP_BENCHMARK) {          // KPP surface buoyancy flux profile Bflux ; dTdz of t3dmix2_geo per tracer
    if (dev_alloc(&D.swdk, D.nij * (size_t)(b->N + 1))) return 4;
    if (dev_alloc(&D.dtdz, D.nij * (size_t)(b->N + 1) * b->NT)) return 4;
  }
  // diag (k_grid.cu): 3 sums per interior column i + 9 maxima + 9 per block of 128 columns of a row
  const size_t nred = (size_t)3 * D.ni + 16 + 12 * 64 + (size_t)9 * ((D.ni + 127) / 128) * D.nj;   // + one 12-double slot per tile (<= 64)
  if (dev_alloc(&D.red, nred)) return 4;
  if (dev_alloc(&D.ksbl, D.nij)) return 4;
  if (dev_alloc(&D.err, 1)) return 4;
  CUDA_OK(cudaMallocHost((void**)&c->h_red, sizeof(double) * (3 * D.ni + 16 + 12 * 64)));
  // initialise_mixing (mod_mixing.F:1430-1530) background values are the host's job (upload Akv,Akt,...)
  c->iic = 0; c->ntfirst = 1; c->nstp = 1; c->nnew = 1; c->nrhs = 1; c->indx1 = 1; c->time = 0.0;
  c->use_graph = (getenv("ROMS_B200_NO_GRAPH") == nullptr);
  return 0;
}

int roms_b200_destroy(roms_b200_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (int a = 0; a < 12; ++a) if (c->graph2d[a]) cudaGraphExecDestroy(c->graph2d[a]);
  for (int f = 0; f < ROMS_B200_NFIELDS; ++f) cudaFree(c->D.f[f]);
  cudaFree((void*)c->D.sc_r); cudaFree((void*)c->D.w1); cudaFree(c->D.P); cudaFree(c->D.scratch2); cudaFree(c->D.swdk); cudaFree(c->D.dtdz); cudaFree(c->D.kpp4); cudaFree(c->D.red); cudaFree(c->D.ksbl); cudaFree(c->D.err);
  if (c->h_red) cudaFreeHost(c->h_red);
  if (c->snap_stream) { cudaStreamSynchronize(c->snap_stream); cudaStreamDestroy(c->snap_stream); cudaEventDestroy(c->snap_ready); cudaEventDestroy(c->snap_done); }
  cudaFree(c->snap_buf); cudaFree(c->s2p_done);
  roms_b200_comm_destroy(c);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  for (int q = 0; q < 6; ++q) if (c->ev[q]) cudaEventDestroy(c->ev[q]);
  delete c;
  return 0;
}

int roms_b200_set_scoord(roms_b200_ctx* c, const double* sc_r, const double* Cs_r, const double* sc_w, const double* Cs_w) {
  // the reference allocates SCALARS(ng)%sc_r, Cs_r as (1:N) and sc_w, Cs_w as (0:N) (mod_scalars.F:1950-1968): the host passes the
  // arrays as they are; the kernels index all four by level k, so the rho-level vectors go to device offset 1
  const int N = c->D.b.N;
  CUDA_OK(cudaMemcpy((void*)(c->D.sc_r + 1), sc_r, N * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy((void*)(c->D.Cs_r + 1), Cs_r, N * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy((void*)c->D.sc_w, sc_w, (N + 1) * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy((void*)c->D.Cs_w, Cs_w, (N + 1) * sizeof(double), cudaMemcpyHostToDevice));
  return 0;
}
int roms_b200_set_weights(roms_b200_ctx* c, int nfast, const double* w1, const double* w2) {
  const int cap = 2 * c->D.p.ndtfast + 4;
  if (nfast + 3 > cap) return 1;
  std::vector<double> a(cap, 0.0), b(cap, 0.0);
  for (int i = 0; i <= nfast + 2 && i < cap; ++i) { a[i] = w1[i]; b[i] = w2[i]; }
  CUDA_OK(cudaMemcpy((void*)c->D.w1, a.data(), cap * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy((void*)c->D.w2, b.data(), cap * sizeof(double), cudaMemcpyHostToDevice));
  c->D.p.nfast = nfast;
  for (int x = 0; x < 12; ++x) if (c->graph2d[x]) { cudaGraphExecDestroy(c->graph2d[x]); c->graph2d[x] = nullptr; }
  return 0;
}

long roms_b200_field_size(const roms_b200_ctx* c, int f) { return (f < 0 || f >= ROMS_B200_NFIELDS) ? -1 : (long)c->fsize[f]; }
int roms_b200_upload(roms_b200_ctx* c, int f, const double* host) {
  if (f < 0 || f >= ROMS_B200_NFIELDS) return 1;
  CUDA_OK(cudaMemcpyAsync(c->D.f[f], host, c->fsize[f] * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}
int roms_b200_download(roms_b200_ctx* c, int f, double* host) {
  if (f < 0 || f >= ROMS_B200_NFIELDS) return 1;
  CUDA_OK(cudaMemcpyAsync(host, c->D.f[f], c->fsize[f] * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}
int roms_b200_download_interior(roms_b200_ctx* c, int f, int plane0, int nplanes, double* host) {
  if (f < 0 || f >= ROMS_B200_NFIELDS) return 1;
  const Dev& D = c->D; const roms_b200_bounds& b = D.b;
  const int nk = nplanes, wi = b.Iend - b.Istr + 1, wj = b.Jend - b.Jstr + 1;
  if ((size_t)(plane0 + nplanes) * D.nij > c->fsize[f]) return 1;
  const double* base = D.f[f] + D.nij * (size_t)plane0;
  for (int k = 0; k < nk; ++k) {
    const double* src = base + D.nij * k + (b.Istr - b.LBi) + (size_t)D.ni * (b.Jstr - b.LBj);
    CUDA_OK(cudaMemcpy2DAsync(host + (size_t)k * wi * wj, wi * sizeof(double), src, D.ni * sizeof(double), wi * sizeof(double), wj,
                              cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}
void* roms_b200_device_ptr(roms_b200_ctx* c, int f) { return (f < 0 || f >= ROMS_B200_NFIELDS) ? nullptr : (void*)c->D.f[f]; }
int roms_b200_sync(roms_b200_ctx* c) {
  CUDA_OK(cudaStreamSynchronize(c->stream)); CUDA_OK(cudaGetLastError());
  int err = 0;                                   // device-side error word
  CUDA_OK(cudaMemcpy(&err, c->D.err, sizeof(int), cudaMemcpyDeviceToHost));
  if (err & 2) {                                 // a halo message never arrived: exit_flag=2 as mp_exchange.F:544-553
    fprintf(stderr, "roms_b200: device error word 0x%x: a halo message from a neighbour tile did not arrive within ~2 s\n", err);
    return 2;
  }
  if (err) {                                     // "fatal algorithm result" -> exit_flag=8 (mod_scalars.F:548-561)
    fprintf(stderr, "roms_b200: device error word 0x%x (bit 0: non-finite or out-of-range reciprocal operand, a blown-up state; bit 2: "
                    "a pipeline barrier of step3d_t timed out)\n", err);
    return 8;
  }
  return 0;
}
long roms_b200_launch_count(const roms_b200_ctx* c) { return c->launches; }

#define ENTER(c) do { if (!(c)) return 1; CUDA_OK(cudaSetDevice((c)->device)); } while (0)
#define LEAVE() do { CUDA_OK(cudaGetLastError()); return 0; } while (0)

int roms_b200_set_massflux(roms_b200_ctx* c, int nrhs) { ENTER(c); k_set_massflux(c, nrhs); LEAVE(); }
int roms_b200_rho_eos(roms_b200_ctx* c, int nrhs) { ENTER(c); k_rho_eos(c, nrhs); LEAVE(); }
int roms_b200_omega(roms_b200_ctx* c) { ENTER(c); k_omega(c); LEAVE(); }
int roms_b200_wvelocity(roms_b200_ctx* c, int ninp) { ENTER(c); if (k_wvelocity(c, ninp)) return 1; LEAVE(); }
int roms_b200_set_zeta(roms_b200_ctx* c) { ENTER(c); k_set_zeta(c); LEAVE(); }
int roms_b200_set_depth(roms_b200_ctx* c) { ENTER(c); k_set_depth(c); LEAVE(); }
int roms_b200_bulk_flux(roms_b200_ctx* c, int nrhs) { ENTER(c); k_bulk_flux(c, nrhs); LEAVE(); }
int roms_b200_set_vbc(roms_b200_ctx* c, int nrhs) { ENTER(c); k_set_vbc(c, nrhs); LEAVE(); }
int roms_b200_ana_vmix(roms_b200_ctx* c) { ENTER(c); k_ana_vmix(c); LEAVE(); }
int roms_b200_lmd_vmix(roms_b200_ctx* c, int nstp) { ENTER(c); if (k_lmd_vmix(c, nstp)) return 1; LEAVE(); }
int roms_b200_pre_step3d(roms_b200_ctx* c, int nrhs, int nstp, int nnew, int iic, int ntfirst) { ENTER(c); k_pre_step3d(c, nrhs, nstp, nnew, iic, ntfirst); LEAVE(); }
int roms_b200_prsgrd(roms_b200_ctx* c, int nrhs) { ENTER(c); k_prsgrd(c, nrhs); LEAVE(); }
int roms_b200_t3dmix2(roms_b200_ctx* c, int nrhs, int nstp, int nnew) { ENTER(c); k_t3dmix2(c, nrhs, nstp, nnew); LEAVE(); }
int roms_b200_rhs3d_tile(roms_b200_ctx* c, int nrhs) { ENTER(c); k_rhs3d_tile(c, nrhs); LEAVE(); }
int roms_b200_uv3dmix2(roms_b200_ctx* c, int nrhs, int nnew) { ENTER(c); if (k_uv3dmix2(c, nrhs, nnew)) return 1; LEAVE(); }
int roms_b200_rhs3d(roms_b200_ctx* c, int nrhs, int nstp, int nnew, int iic, int ntfirst) {
  ENTER(c);
  k_pre_step3d(c, nrhs, nstp, nnew, iic, ntfirst); k_prsgrd(c, nrhs); k_t3dmix2(c, nrhs, nstp, nnew); k_rhs3d_tile(c, nrhs); k_uv3dmix2(c, nrhs, nnew);
  LEAVE();
}
int roms_b200_step2d(roms_b200_ctx* c, int krhs, int kstp, int knew, int nstp, int nnew, int iif, int pred, int iic, int ntfirst) {
  ENTER(c); k_step2d(c, krhs, kstp, knew, nstp, nnew, iif, pred, iic, ntfirst); if (k_step2d_join(c)) return 1; LEAVE();
}
int roms_b200_step3d_uv(roms_b200_ctx* c, int nrhs, int nstp, int nnew, int iic, int ntfirst) { ENTER(c); k_step3d_uv(c, nrhs, nstp, nnew, iic, ntfirst); LEAVE(); }
int roms_b200_step3d_t(roms_b200_ctx* c, int nrhs, int nstp, int nnew) { ENTER(c); k_step3d_t(c, nrhs, nstp, nnew); LEAVE(); }
int roms_b200_diag(roms_b200_ctx* c, int nstp, double* out3) { ENTER(c); if (k_diag(c, nstp, out3)) return 1; LEAVE(); }
int roms_b200_diag_full(roms_b200_ctx* c, int nstp, double* out13) {
  ENTER(c); double d3[3]; if (k_diag(c, nstp, d3)) return 1; memcpy(out13, c->last_diag, sizeof(c->last_diag)); LEAVE();
}
int roms_b200_diag_last(const roms_b200_ctx* c, double* out13) { if (!c || !out13) return 1; memcpy(out13, c->last_diag, sizeof(c->last_diag)); return 0; }
int roms_b200_diag_begin(roms_b200_ctx* c, int nstp) { ENTER(c); if (k_diag_begin(c, nstp)) return 1; LEAVE(); }
int roms_b200_diag_end(roms_b200_ctx* c, double* out3) { ENTER(c); if (k_diag_end(c, out3)) return 1; LEAVE(); }
int roms_b200_host_alloc(size_t bytes, void** p) { CUDA_OK(cudaMallocHost(p, bytes)); return 0; }
int roms_b200_host_free(void* p) { if (p) cudaFreeHost(p); return 0; }
int roms_b200_upload_async(roms_b200_ctx* c, int f, const double* pinned_host) {
  ENTER(c);
  if (f < 0 || f >= ROMS_B200_NFIELDS) return 1;
  CUDA_OK(cudaMemcpyAsync(c->D.f[f], pinned_host, c->fsize[f] * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  LEAVE();
}
int roms_b200_set_data(roms_b200_ctx* c, double tdays) { ENTER(c); k_set_data(c, tdays); LEAVE(); }
int roms_b200_ana_initial(roms_b200_ctx* c) { ENTER(c); k_ana_initial(c); LEAVE(); }
int roms_b200_ini_fields(roms_b200_ctx* c, int nstp, int kstp) { ENTER(c); k_ini_fields(c, nstp, kstp); LEAVE(); }
int roms_b200_fill(roms_b200_ctx* c, int f, double value) {
  ENTER(c);
  if (f < 0 || f >= ROMS_B200_NFIELDS) return 1;
  std::vector<double> h(c->fsize[f], value);
  CUDA_OK(cudaMemcpy(c->D.f[f], h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice));
  LEAVE();
}

// main3d.F:810-918: LF-AM3 fast loop.  The launch sequence depends only on
// (indx1 at entry, which of the three AB start-up forms the first predictor
// uses), so it is captured once per key and replayed as a CUDA graph.
struct FastPhase { int krhs, kstp, knew, iif, pred; };
static void fast_loop_sequence(int nfast, int* indx1_io, std::vector<FastPhase>& seq) {
  int indx1 = *indx1_io, kstp = 1, knew = 1, krhs = 1, iif = 1; bool PRED = false;
  for (int my_iif = 1; my_iif <= nfast + 1; ++my_iif) {
    const int next_indx1 = 3 - indx1;
    if (!PRED && my_iif <= nfast + 1) {
      PRED = true; iif = my_iif;
      kstp = (iif == 1) ? indx1 : 3 - indx1;
      knew = 3; krhs = indx1;
    }
    if (my_iif <= nfast + 1) seq.push_back(FastPhase{krhs, kstp, knew, iif, 1});
    if (PRED) {
      PRED = false; knew = next_indx1; kstp = 3 - knew; krhs = 3;
      if (iif < nfast + 1) indx1 = next_indx1;
    }
    if (iif < nfast + 1) seq.push_back(FastPhase{krhs, kstp, knew, iif, 0});
  }
  *indx1_io = indx1;
}
static int fast_loop_launch(roms_b200_ctx* c, int nstp, int nnew, int iic, int ntfirst, int* indx1_io) {
  const int nfast = c->D.p.nfast;
  std::vector<FastPhase> seq;
  fast_loop_sequence(nfast, indx1_io, seq);
  for (const FastPhase& f : seq) {
    k_step2d(c, f.krhs, f.kstp, f.knew, nstp, nnew, f.iif, f.pred, iic, ntfirst);
    if (!c->comm) continue;
    if (f.pred) {      // mp_exchange2d calls of step2d_LF_AM3.h:842,1013,1068,3043 aggregated into one message
      if (f.iif == nfast + 1) { const XF x[3] = {xf2(FID(Zt_avg1)), xf2(FID(DU_avg1)), xf2(FID(DV_avg1))}; if (xchg(c, x, 3)) return 1; }
      else if (c->deep) { }            // deep-halo predictor (k_step2d): its results exist on the 3 halo points the corrector reads
      else { const XF x[4] = {xf2(FID(zeta), f.knew), xf2(FID(ubar), f.knew), xf2(FID(vbar), f.knew), xf2(FID(rzeta), f.krhs)}; if (xchg(c, x, 4)) return 1; }
    } else { const XF x[3] = {xf2(FID(zeta), f.knew), xf2(FID(ubar), f.knew), xf2(FID(vbar), f.knew)}; if (xchg(c, x, 3)) return 1; }
    if (k_step2d_join(c)) return 1;
  }
  return 0;
}
// single tile: the whole loop as one persistent kernel (k_step2d.cu) when every block tile can be resident at once
static int fast_loop_persistent(roms_b200_ctx* c, int nstp, int nnew, int iic, int ntfirst, int* indx1_io) {
  if (c->comm) return 2;
  std::vector<FastPhase> seq; int ix = *indx1_io;
  fast_loop_sequence(c->D.p.nfast, &ix, seq);
  std::vector<int> ph(seq.size());
  for (size_t q = 0; q < seq.size(); ++q) ph[q] = seq[q].krhs | seq[q].kstp << 2 | seq[q].knew << 4 | seq[q].pred << 6 | seq[q].iif << 8;
  const int rc = k_step2d_persist(c, ph.data(), (int)ph.size(), nstp, nnew, iic, ntfirst);
  if (rc == 0) *indx1_io = ix;
  return rc;
}
int roms_b200_step2d_loop(roms_b200_ctx* c, int nstp, int nnew, int iic, int ntfirst, int* indx1) {
  ENTER(c);
  const int mode = (iic == ntfirst) ? 0 : (iic == ntfirst + 1 ? 1 : 2);
  { const int rc = fast_loop_persistent(c, nstp, nnew, iic, ntfirst, indx1); if (rc == 0) { LEAVE(); } if (rc != 2) return 1; }
  if (!c->use_graph) { if (fast_loop_launch(c, nstp, nnew, iic, ntfirst, indx1)) return 1; LEAVE(); }
  // the launch sequence is a pure function of (indx1 at entry, nstp, AB start-up mode)
  const int key = ((*indx1 - 1) & 1) * 6 + ((nstp - 1) & 1) * 3 + mode;
  if (!c->graph2d[key]) {
    // build this graph; for the steady-state form (mode 2) build all four (indx1,nstp) variants at once so
    // that no instantiation falls into a later (timed) step.  Capture records but does not execute.
    for (int ix = 1; ix <= 2; ++ix) for (int ns = 1; ns <= 2; ++ns) {
      const int kk = (ix - 1) * 6 + (ns - 1) * 3 + mode;
      if (c->graph2d[kk] || (mode != 2 && kk != key)) continue;
      cudaGraph_t gph;
      const long l0 = c->launches;
      int tmp = ix;
      CUDA_OK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
      const int rc = fast_loop_launch(c, ns, 3 - ns, iic, ntfirst, &tmp);
      CUDA_OK(cudaStreamEndCapture(c->stream, &gph));
      if (rc) return 1;
      CUDA_OK(cudaGraphInstantiate(&c->graph2d[kk], gph, 0));
      CUDA_OK(cudaGraphDestroy(gph));
      c->graph_launches[kk] = c->launches - l0;
      c->launches = l0;
    }
  }
  CUDA_OK(cudaGraphLaunch(c->graph2d[key], c->stream));
  c->launches += c->graph_launches[key];
  // advance indx1 exactly as the reference does: it flips once per completed sub-step pair
  int x = *indx1;
  for (int q = 0; q < c->D.p.nfast; ++q) x = 3 - x;
  *indx1 = x;
  LEAVE();
}

int roms_b200_get_stepping(const roms_b200_ctx* c, int* o, double* time) {
  o[0] = c->iic; o[1] = c->ntfirst; o[2] = c->nstp; o[3] = c->nnew; o[4] = c->nrhs; o[5] = c->indx1; if (time) *time = c->time; return 0;
}
int roms_b200_set_stepping(roms_b200_ctx* c, const int* in, double time) {
  c->iic = in[0]; c->ntfirst = in[1]; c->nstp = in[2]; c->nnew = in[3]; c->nrhs = in[4]; c->indx1 = in[5]; c->time = time; return 0;
}

// Independent branches of a step run side by side on a second stream (on benchmark-size tiles every kernel is latency-bound
// and leaves most of the GPU idle):
//   start of the step : [set_massflux -> omega -> wvelocity]  beside  [rho_eos -> diag -> bulk_flux -> set_vbc -> vertical mixing]
//                       (wvelocity waits for diag, which reads the old wvel)
//   after set_zeta    : the TRACER branch [pre_step3d (tracers) -> t3dmix2] beside the MOMENTUM branch [pre_step3d (momentum) ->
//                       prsgrd -> rhs3d -> uv3dmix2] AND the whole barotropic fast loop (which needs rufrc/rvfrc only); it joins
//                       before set_depth (which overwrites Hz, z_r, z_w).
// Read/write sets: SURVEY Appendix D; scratch volumes are private to a branch (D.dtdz vs D.kpp4).  Halo swaps stay on the launch
// stream in a fixed order (the mailbox protocol numbers them); the swap of t(3) therefore moves from behind pre_step3d
// (pre_step3d.F:1171) to the swap behind step3d_uv -- its first reader is step3d_t.  Every point is still advanced by the same
// operations on the same operands: results are bit-identical to the one-stream order (ROMS_B200_ONE_STREAM=1).
namespace {
struct OnStream2 {                        // launches of the k_* helpers go to c->stream: redirect them for a scope
  roms_b200_ctx* c; cudaStream_t save;
  explicit OnStream2(roms_b200_ctx* c_) : c(c_), save(c_->stream) { c->stream = c->stream2; }
  ~OnStream2() { c->stream = save; }
};
}  // namespace
// main3d.F:216-1148 on the device mirror (post_initial is the host's job: upload a state
// that has been through ini_zeta/ini_fields, i.e. what the reference holds when the first
// set_massflux is called).
int roms_b200_main3d(roms_b200_ctx* c, int nsteps, int analytic_forcing, int with_diag) {
  ENTER(c);
  const bool bench = (c->D.p.app == ROMS_B200_APP_BENCHMARK);
  const bool two = c->two_streams != 0;
  static const bool swap_vbc = (getenv("ROMS_B200_SWAP_VBC") != nullptr);
  for (int s = 0; s < nsteps; ++s) {
    c->nstp = 1 + ((c->iic - c->ntfirst) % 2); c->nnew = 3 - c->nstp; c->nrhs = c->nstp;
    const int nstp = c->nstp, nnew = c->nnew, nrhs = c->nrhs, iic = c->iic, ntf = c->ntfirst;
    if (analytic_forcing) k_set_data(c, c->time / 86400.0);
    // ---- branch A (stream2): mass fluxes, omega, diag, wvelocity ; branch B (launch stream): density, surface fluxes, vertical
    // mixing.  diag (main3d.F:300) reads the density of this step -> waits for rho_eos; wvelocity (main3d.F:535) overwrites the
    // wvel diag reads -> behind it on the same stream.  diag has its own scratch planes (8..10 of scratch2; KPP uses 0..5).
    if (two) {
      CUDA_OK(cudaEventRecord(c->ev[0], c->stream)); CUDA_OK(cudaStreamWaitEvent(c->stream2, c->ev[0], 0));
      OnStream2 on(c);
      k_set_massflux(c, nrhs); k_omega(c);
    } else k_set_massflux(c, nrhs);
    k_rho_eos(c, nrhs);
    if (two) {
      CUDA_OK(cudaEventRecord(c->ev[1], c->stream)); CUDA_OK(cudaStreamWaitEvent(c->stream2, c->ev[1], 0));
      { OnStream2 on(c);
        if (bench) { if (k_lmd_vmix_part(c, nstp, 1)) return 1; CUDA_OK(cudaEventRecord(c->ev[5], c->stream)); }   // KPP's spline derivatives: need the density only
        if (with_diag == 1) { double d[3]; if (k_diag(c, nstp, d)) return 1; }      // main3d.F:300 (diag with NINFO=1), host waits
        else if (with_diag) { if (k_diag_begin(c, nstp)) return 1; }                 // reductions + D2H stay asynchronous: the caller
                                                                                     // collects them with roms_b200_diag_end
        if (k_wvelocity(c, nstp)) return 1; }
      CUDA_OK(cudaEventRecord(c->ev[2], c->stream2));
    } else {
      if (with_diag == 1) { double d[3]; if (k_diag(c, nstp, d)) return 1; }
      else if (with_diag) { if (k_diag_begin(c, nstp)) return 1; }
    }
    if (bench) k_bulk_flux(c, nrhs);
    k_set_vbc(c, nrhs);
    // no halo swap of the stresses (bulk_flux, set_vbc evaluated two points into the halo) nor of Akv (KPP evaluated on the
    // first halo ring; ana_vmix on the whole mirror): ROMS_B200_SWAP_VBC=1 restores the two messages of the reference
    if (swap_vbc) { const XF x[4] = {xf2(FID(sustr)), xf2(FID(svstr)), xf2(FID(bustr)), xf2(FID(bvstr))}; if (xchg(c, x, 4)) return 1; }
    if (bench) {
      if (two) { CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev[5], 0)); if (k_lmd_vmix_part(c, nstp, 2)) return 1; }
      else if (k_lmd_vmix(c, nstp)) return 1;
      if (swap_vbc) { const XF x[1] = {xf3(c, FID(Akv))}; if (xchg(c, x, 1)) return 1; } } else k_ana_vmix(c);
    if (two) CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev[2], 0));              // join A
    else { k_omega(c); if (k_wvelocity(c, nstp)) return 1; }
    k_set_zeta(c);
    // ---- tracer branch (stream2) beside the momentum branch and the fast loop (launch stream)
    if (two) {
      CUDA_OK(cudaEventRecord(c->ev[3], c->stream)); CUDA_OK(cudaStreamWaitEvent(c->stream2, c->ev[3], 0));
      { OnStream2 on(c); k_pre_step3d_t(c, nrhs, nstp, nnew, iic, ntf); k_t3dmix2(c, nrhs, nstp, nnew); }
      CUDA_OK(cudaEventRecord(c->ev[4], c->stream2));
      k_pre_step3d_uv(c, nrhs, nstp, nnew, iic, ntf);
      k_prsgrd(c, nrhs); k_rhs3d_tile(c, nrhs); if (k_uv3dmix2(c, nrhs, nnew)) return 1;
    } else {
      k_pre_step3d(c, nrhs, nstp, nnew, iic, ntf);
      { XF x[HALO_MAXF]; int n = 0; for (int it = 1; it <= c->D.b.NT; ++it) x[n++] = xf3(c, FID(t), 3, it); if (xchg(c, x, n)) return 1; }   // pre_step3d.F:1171
      k_prsgrd(c, nrhs); k_t3dmix2(c, nrhs, nstp, nnew); k_rhs3d_tile(c, nrhs); if (k_uv3dmix2(c, nrhs, nnew)) return 1;
    }
    if (c->deep) { const XF x[2] = {xf2(FID(rufrc)), xf2(FID(rvfrc))}; if (xchg(c, x, 2)) return 1; }   // read by the deep-halo predictor
    if (roms_b200_step2d_loop(c, nstp, nnew, iic, ntf, &c->indx1)) return 1;
    if (two) CUDA_OK(cudaStreamWaitEvent(c->stream, c->ev[4], 0));              // join the tracer branch
    k_set_depth(c);
    k_step3d_uv(c, nrhs, nstp, nnew, iic, ntf);
    { XF x[HALO_MAXF]; int n = 0;                                                                      // step3d_uv.F:1805-1824
      x[n++] = xf3(c, FID(u), nnew); x[n++] = xf3(c, FID(v), nnew); x[n++] = xf3(c, FID(Huon)); x[n++] = xf3(c, FID(Hvom));
      x[n++] = xf2(FID(ubar), 1); x[n++] = xf2(FID(ubar), 2); x[n++] = xf2(FID(vbar), 1); x[n++] = xf2(FID(vbar), 2);
      if (two) for (int it = 1; it <= c->D.b.NT && n < HALO_MAXF; ++it) x[n++] = xf3(c, FID(t), 3, it);   // + t(3) (pre_step3d.F:1171)
      if (xchg(c, x, n)) return 1; }
    k_omega(c);
    k_step3d_t(c, nrhs, nstp, nnew);
    { XF x[HALO_MAXF]; int n = 0; for (int it = 1; it <= c->D.b.NT; ++it) x[n++] = xf3(c, FID(t), nnew, it); if (xchg(c, x, n)) return 1; }   // step3d_t.F:1920
    c->iic += 1; c->time += c->D.p.dt;
  }
  LEAVE();
}

// Output snapshots: mirror -> staging (launch stream, device-to-device) -> pinned host buffers (copy stream), see roms_b200.h
int roms_b200_snapshot_begin(roms_b200_ctx* c, int nfields, const int* fields, double* const* host) {
  ENTER(c);
  if (nfields <= 0 || !fields || !host || c->snap_pending) return 1;
  size_t total = 0;
  for (int q = 0; q < nfields; ++q) { if (fields[q] < 0 || fields[q] >= ROMS_B200_NFIELDS || !host[q]) return 1; total += c->fsize[fields[q]]; }
  if (!c->snap_stream) {
    CUDA_OK(cudaStreamCreateWithFlags(&c->snap_stream, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&c->snap_ready, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&c->snap_done, cudaEventDisableTiming));
  }
  if (total > c->snap_cap) {
    if (c->snap_buf) CUDA_OK(cudaFree(c->snap_buf));
    c->snap_buf = nullptr; c->snap_cap = 0;
    CUDA_OK(cudaMalloc((void**)&c->snap_buf, total * sizeof(double)));
    c->snap_cap = total;
  }
  size_t off = 0;
  for (int q = 0; q < nfields; ++q) {                 // the state as of this point of the launch stream
    const size_t n = c->fsize[fields[q]];
    CUDA_OK(cudaMemcpyAsync(c->snap_buf + off, c->D.f[fields[q]], n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    off += n;
  }
  CUDA_OK(cudaEventRecord(c->snap_ready, c->stream));
  CUDA_OK(cudaStreamWaitEvent(c->snap_stream, c->snap_ready, 0));
  off = 0;
  for (int q = 0; q < nfields; ++q) {
    const size_t n = c->fsize[fields[q]];
    CUDA_OK(cudaMemcpyAsync(host[q], c->snap_buf + off, n * sizeof(double), cudaMemcpyDeviceToHost, c->snap_stream));
    off += n;
  }
  CUDA_OK(cudaEventRecord(c->snap_done, c->snap_stream));
  c->snap_pending = 1;
  LEAVE();
}
int roms_b200_snapshot_end(roms_b200_ctx* c) {
  ENTER(c);
  if (!c->snap_pending) return 0;
  CUDA_OK(cudaEventSynchronize(c->snap_done));
  c->snap_pending = 0;
  LEAVE();
}

// ---- PERFECT_RESTART (Utility/wrt_rst.F:178-211 time indices; :345-900 fields): the state a bit-identical continuation needs.
// The list is the reference's for the UPWELLING / BENCHMARK option sets -- zeta(3), rzeta(2), ubar(3), rubar(2), vbar(3), rvbar(2),
// u(2), ru(0:N,2), v(2), rv(0:N,2), t(3,NT), rho, Hsbl, ghats, Akv, Akt(NAT) -- plus Zt_avg1, which the reference rebuilds from
// zeta in ini_zeta on restart and this library takes as stored.  Use with roms_b200_snapshot_begin/end (asynchronous) or
// roms_b200_download; after uploading the set into a freshly initialised context call roms_b200_set_stepping with the stored
// indices and roms_b200_restart_finish (the part of post_initial a restart repeats: set_depth from Zt_avg1).
int roms_b200_restart_fields(const roms_b200_ctx* c, int* ids, int cap) {
  if (!c || !ids) return -1;
  const bool bench = (c->D.p.app == ROMS_B200_APP_BENCHMARK);
  const int all[] = {FID(zeta), FID(rzeta), FID(ubar), FID(rubar), FID(vbar), FID(rvbar), FID(u), FID(ru), FID(v), FID(rv), FID(t), FID(rho),
                     FID(Akv), FID(Akt), FID(Zt_avg1), FID(hsbl), FID(ghats)};
  const int n = bench ? 17 : 15;                       // Hsbl (LMD_SKPP), ghats (LMD_NONLOCAL): BENCHMARK only
  if (cap < n) return -1;
  for (int q = 0; q < n; ++q) ids[q] = all[q];
  return n;
}
int roms_b200_restart_finish(roms_b200_ctx* c) { ENTER(c); k_set_depth(c); LEAVE(); }

// CUDA-event stopwatch on the context's launch stream (cudaEvent sees only that stream)
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
int roms_b200_timer_start(roms_b200_ctx* c) {
  ENTER(c);
  if (!g_ev0) { CUDA_OK(cudaEventCreate(&g_ev0)); CUDA_OK(cudaEventCreate(&g_ev1)); }
  CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaEventRecord(g_ev0, c->stream));
  LEAVE();
}
int roms_b200_timer_stop(roms_b200_ctx* c, float* ms) {
  ENTER(c);
  CUDA_OK(cudaEventRecord(g_ev1, c->stream));
  CUDA_OK(cudaEventSynchronize(g_ev1));
  CUDA_OK(cudaEventElapsedTime(ms, g_ev0, g_ev1));
  LEAVE();
}
// write `mbytes` MiB of scratch so that the next kernel starts with a cold L2
int roms_b200_flush_l2(roms_b200_ctx* c, int mbytes) {
  ENTER(c);
  static void* buf = nullptr; static size_t cap = 0;
  const size_t n = (size_t)mbytes << 20;
  if (cap < n) { if (buf) cudaFree(buf); CUDA_OK(cudaMalloc(&buf, n)); cap = n; }
  CUDA_OK(cudaMemsetAsync(buf, 1, n, c->stream));
  LEAVE();
}
int roms_b200_time_step3d_t(roms_b200_ctx* c, int nrhs, int nstp, int nnew, int reps, float* ms_avg) {
  ENTER(c);
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
  CUDA_OK(cudaEventRecord(e0, c->stream));
  for (int r = 0; r < reps; ++r) k_step3d_t(c, nrhs, nstp, nnew);
  CUDA_OK(cudaEventRecord(e1, c->stream));
  CUDA_OK(cudaEventSynchronize(e1));
  float ms = 0; CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
  *ms_avg = ms / reps;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  LEAVE();
}

}  // extern "C"
