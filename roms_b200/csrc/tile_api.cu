// roms_b200/csrc/tile_api.cu -- the boundary as a Fortran MPI host sees it.
//
// (1) Host arrays with THEIR OWN bounds.  The reference allocates every tile array as (LBi:UBi,LBj:UBj,...) with
//     NghostPoints = 2 (Utility/get_bounds.F:193-212: LBi = Imin on the western tile, else Istr-Nghost, ...), the device mirror of
//     a distributed tile carries a halo of 6 (deep-halo fast loop).  roms_b200_upload_bounds / roms_b200_download_bounds move the
//     intersection of the two index boxes plane by plane (cudaMemcpy2D), roms_b200_exchange_field refreshes the mirror's wider
//     halo from the neighbour tiles afterwards.  roms_b200_register_field remembers c_loc(array) and its bounds so that later
//     calls need only the field id (roms_b200_upload_registered / roms_b200_download_registered).
// (2) The `_tile` argument lists.  Every roms_b200_X_tile export below takes exactly the argument list of the reference's
//     X_tile for the UPWELLING / BENCHMARK cpp sets (file:line at each function; arguments that exist only under cpp options one
//     of the two applications lacks -- dndx/dmde (CURVGRID), rhoA/rhoS (VAR_RHO_2D), srflx (SOLAR_SOURCE), ghats (LMD_NONLOCAL) --
//     may be null), so the wrapper X(ng,tile) changes ONE token: CALL X_tile(...) -> rc = roms_b200_X_tile(ctx, ...).  The array
//     arguments are the HOST arrays: each one is checked against the registration of its field (same base address, same
//     bounds) -- a call with an array the mirror does not shadow is an error, not a silent no-op -- then the call forwards to
//     the mirror-resident entry point.  Hidden module inputs (iic, ntfirst, iif, PREDICTOR_2D_STEP, ...) cross the ABI through
//     roms_b200_set_stepping / roms_b200_set_fast_step before the call, as mod_stepping / mod_scalars hold them in the reference.
#include "common.cuh"
#include <cstring>

namespace {
int plane_count(const roms_b200_ctx* c, int f) { return (int)(c->fsize[f] / c->D.nij); }
// copy the intersection of the host box and the mirror box, all planes of field f; dir 0: host -> mirror, 1: mirror -> host
int copy_bounds(roms_b200_ctx* c, int f, double* host, int LBi, int UBi, int LBj, int UBj, int dir) {
  if (f < 0 || f >= ROMS_B200_NFIELDS || !host) return 1;
  const roms_b200_bounds& b = c->D.b;
  const int i0 = LBi > b.LBi ? LBi : b.LBi, i1 = UBi < b.UBi ? UBi : b.UBi, j0 = LBj > b.LBj ? LBj : b.LBj, j1 = UBj < b.UBj ? UBj : b.UBj;
  if (i0 > i1 || j0 > j1) return 1;
  const size_t hni = (size_t)(UBi - LBi + 1), hnij = hni * (size_t)(UBj - LBj + 1);
  const int np = plane_count(c, f);
  for (int p = 0; p < np; ++p) {
    double* h = host + hnij * p + (i0 - LBi) + hni * (size_t)(j0 - LBj);
    double* d = c->D.f[f] + c->D.nij * (size_t)p + (i0 - b.LBi) + (size_t)c->D.ni * (j0 - b.LBj);
    if (dir == 0) CUDA_OK(cudaMemcpy2DAsync(d, c->D.ni * sizeof(double), h, hni * sizeof(double), (size_t)(i1 - i0 + 1) * sizeof(double), j1 - j0 + 1, cudaMemcpyHostToDevice, c->stream));
    else CUDA_OK(cudaMemcpy2DAsync(h, hni * sizeof(double), d, c->D.ni * sizeof(double), (size_t)(i1 - i0 + 1) * sizeof(double), j1 - j0 + 1, cudaMemcpyDeviceToHost, c->stream));
  }
  CUDA_OK(cudaStreamSynchronize(c->stream));
  return 0;
}
// an array argument of a _tile call: null is allowed only for cpp-optional arguments; otherwise it must be the registered array
int chk(const roms_b200_ctx* c, const char* routine, int f, const void* p, int LBi, int UBi, int LBj, int UBj, bool optional = false) {
  if (!p) { if (optional) return 0; fprintf(stderr, "roms_b200: %s: array argument %d is null\n", routine, f); return 1; }
  if (!c->host_ptr[f]) { fprintf(stderr, "roms_b200: %s: field %d was never registered (roms_b200_register_field)\n", routine, f); return 1; }
  if ((const void*)c->host_ptr[f] != p) { fprintf(stderr, "roms_b200: %s: array argument of field %d is not the registered host array\n", routine, f); return 1; }
  const int* hb = c->host_b[f];
  if (hb[0] != LBi || hb[1] != UBi || hb[2] != LBj || hb[3] != UBj) { fprintf(stderr, "roms_b200: %s: bounds (%d:%d,%d:%d) differ from the registration of field %d\n", routine, LBi, UBi, LBj, UBj, f); return 1; }
  return 0;
}
// ng, tile and the scratch bounds of tile.h:21-24 (IminS = Istr-3 ... JmaxS = Jend+3)
int chk_tile(const roms_b200_ctx* c, const char* routine, int ng, int tile, int IminS, int ImaxS, int JminS, int JmaxS) {
  const roms_b200_bounds& b = c->D.b;
  if (ng != 1) { fprintf(stderr, "roms_b200: %s: ng=%d (one grid per context)\n", routine, ng); return 1; }
  if (tile != b.Jtile * b.NtileI + b.Itile) { fprintf(stderr, "roms_b200: %s: tile %d is not this context's tile %d\n", routine, tile, b.Jtile * b.NtileI + b.Itile); return 1; }
  if (IminS != b.Istr - 3 || ImaxS != b.Iend + 3 || JminS != b.Jstr - 3 || JmaxS != b.Jend + 3) { fprintf(stderr, "roms_b200: %s: IminS..JmaxS do not match BOUNDS(ng) of the tile\n", routine); return 1; }
  return 0;
}
#define F(name) ROMS_B200_F_##name
#define A(name) if (chk(c, R, F(name), name, LBi, UBi, LBj, UBj)) return 1
#define AO(name) if (chk(c, R, F(name), name, LBi, UBi, LBj, UBj, true)) return 1
#define T() if (!c) return 1; if (chk_tile(c, R, ng, tile, IminS, ImaxS, JminS, JmaxS)) return 1
}  // namespace

extern "C" {

int roms_b200_register_field(roms_b200_ctx* c, int f, const double* host, int LBi, int UBi, int LBj, int UBj) {
  if (!c || f < 0 || f >= ROMS_B200_NFIELDS || !host || UBi < LBi || UBj < LBj) return 1;
  c->host_ptr[f] = host; c->host_b[f][0] = LBi; c->host_b[f][1] = UBi; c->host_b[f][2] = LBj; c->host_b[f][3] = UBj;
  return 0;
}
int roms_b200_upload_bounds(roms_b200_ctx* c, int f, const double* host, int LBi, int UBi, int LBj, int UBj) {
  if (!c) return 1;
  CUDA_OK(cudaSetDevice(c->device));
  return copy_bounds(c, f, const_cast<double*>(host), LBi, UBi, LBj, UBj, 0);
}
int roms_b200_download_bounds(roms_b200_ctx* c, int f, double* host, int LBi, int UBi, int LBj, int UBj) {
  if (!c) return 1;
  CUDA_OK(cudaSetDevice(c->device));
  return copy_bounds(c, f, host, LBi, UBi, LBj, UBj, 1);
}
int roms_b200_upload_registered(roms_b200_ctx* c, int f) {
  if (!c || f < 0 || f >= ROMS_B200_NFIELDS || !c->host_ptr[f]) return 1;
  const int* hb = c->host_b[f];
  return roms_b200_upload_bounds(c, f, c->host_ptr[f], hb[0], hb[1], hb[2], hb[3]);
}
int roms_b200_download_registered(roms_b200_ctx* c, int f) {
  if (!c || f < 0 || f >= ROMS_B200_NFIELDS || !c->host_ptr[f]) return 1;
  const int* hb = c->host_b[f];
  return roms_b200_download_bounds(c, f, const_cast<double*>(c->host_ptr[f]), hb[0], hb[1], hb[2], hb[3]);
}
// refresh the halo of every plane of a field from the neighbour tiles (mp_exchange2d/3d/4d on the mirror); no-op on one tile
int roms_b200_exchange_field(roms_b200_ctx* c, int f) {
  if (!c || f < 0 || f >= ROMS_B200_NFIELDS) return 1;
  CUDA_OK(cudaSetDevice(c->device));
  const int np = plane_count(c, f);
  for (int p0 = 0; p0 < np; p0 += HALO_MAXPLANES) {
    double* base = c->D.f[f] + c->D.nij * (size_t)p0;
    const int n = (np - p0 < HALO_MAXPLANES) ? np - p0 : HALO_MAXPLANES;
    if (halo_exchange(c, &base, &n, 1)) return 1;
  }
  CUDA_OK(cudaGetLastError());
  return 0;
}
// before the first step2d_tile of a baroclinic step on several tiles: the deep-halo predictor reads rufrc, rvfrc three points into
// the halo (the reference needs no such swap: it swaps after every sub-step instead)
int roms_b200_fast_loop_begin(roms_b200_ctx* c) {
  if (!c) return 1;
  CUDA_OK(cudaSetDevice(c->device));
  if (!c->deep) return 0;
  const XF x[2] = {xf2(F(rufrc)), xf2(F(rvfrc))};
  return xchg(c, x, 2);
}
// iif(ng), PREDICTOR_2D_STEP(ng) of mod_scalars (main3d.F:820-880) for the next roms_b200_step2d_tile call
int roms_b200_set_fast_step(roms_b200_ctx* c, int iif, int predictor_2d_step) { if (!c) return 1; c->fast_iif = iif; c->fast_pred = predictor_2d_step; return 0; }

// ---- set_massflux_tile, set_massflux.F:73-82 (argument `model` = iNLM)
int roms_b200_set_massflux_tile(roms_b200_ctx* c, int ng, int tile, int model, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                                int nrhs, const double* u, const double* v, const double* Hz, const double* om_v, const double* on_u, double* Huon, double* Hvom) {
  const char* R = "set_massflux_tile"; (void)model; T();
  A(u); A(v); A(Hz); A(om_v); A(on_u); A(Huon); A(Hvom);
  return roms_b200_set_massflux(c, nrhs);
}
// ---- omega_tile, omega.F:96-114
int roms_b200_omega_tile(roms_b200_ctx* c, int ng, int tile, int model, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                         const double* Huon, const double* Hvom, const double* z_w, double* W) {
  const char* R = "omega_tile"; (void)model; T();
  A(Huon); A(Hvom); A(z_w); A(W);
  return roms_b200_omega(c);
}
// ---- set_zeta_tile, set_zeta.F:59-62
int roms_b200_set_zeta_tile(roms_b200_ctx* c, int ng, int tile, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                            const double* Zt_avg1, double* zeta) {
  const char* R = "set_zeta_tile"; T();
  A(Zt_avg1); A(zeta);
  return roms_b200_set_zeta(c);
}
// ---- set_depth_tile, set_depth.F:76-85
int roms_b200_set_depth_tile(roms_b200_ctx* c, int ng, int tile, int model, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                             int nstp, int nnew, const double* h, const double* Zt_avg1, double* Hz, double* z_r, double* z_w) {
  const char* R = "set_depth_tile"; (void)model; (void)nstp; (void)nnew; T();
  A(h); A(Zt_avg1); A(Hz); A(z_r); A(z_w);
  return roms_b200_set_depth(c);
}
// ---- pre_step3d_tile, pre_step3d.F:126-160 (srflx: SOLAR_SOURCE, ghats: LMD_NONLOCAL -- null in UPWELLING)
int roms_b200_pre_step3d_tile(roms_b200_ctx* c, int ng, int tile, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                              int nrhs, int nstp, int nnew, const double* pm, const double* pn, const double* Hz, const double* Huon, const double* Hvom,
                              const double* z_r, const double* z_w, const double* btflx, const double* bustr, const double* bvstr, const double* stflx,
                              const double* sustr, const double* svstr, const double* srflx, const double* Akt, const double* Akv, const double* ghats,
                              const double* W, const double* ru, const double* rv, double* t, double* u, double* v) {
  const char* R = "pre_step3d_tile"; T();
  A(pm); A(pn); A(Hz); A(Huon); A(Hvom); A(z_r); A(z_w); A(btflx); A(bustr); A(bvstr); A(stflx); A(sustr); A(svstr); AO(srflx); A(Akt); A(Akv); AO(ghats);
  A(W); A(ru); A(rv); A(t); A(u); A(v);
  if (roms_b200_pre_step3d(c, nrhs, nstp, nnew, c->iic, c->ntfirst)) return 1;
  XF x[HALO_MAXF]; int n = 0;                                             // mp_exchange4d of t(:,:,:,3,:), pre_step3d.F:1171
  for (int it = 1; it <= c->D.b.NT && n < HALO_MAXF; ++it) x[n++] = xf3(c, F(t), 3, it);
  return xchg(c, x, n);
}
// ---- prsgrd32_tile, prsgrd32.h:109-134
int roms_b200_prsgrd32_tile(roms_b200_ctx* c, int ng, int tile, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                            int nrhs, const double* om_v, const double* on_u, const double* Hz, const double* z_r, const double* z_w, const double* rho,
                            double* ru, double* rv) {
  const char* R = "prsgrd32_tile"; T();
  A(om_v); A(on_u); A(Hz); A(z_r); A(z_w); A(rho); A(ru); A(rv);
  return roms_b200_prsgrd(c, nrhs);
}
// ---- rhs3d_tile, rhs3d.F:196-221 (dmde, dndx: CURVGRID && UV_ADV -- null in UPWELLING)
int roms_b200_rhs3d_tile_tile(roms_b200_ctx* c, int ng, int tile, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                              int nrhs, const double* Hz, const double* Huon, const double* Hvom, const double* dmde, const double* dndx, const double* fomn,
                              const double* om_u, const double* om_v, const double* on_u, const double* on_v, const double* pm, const double* pn,
                              const double* bustr, const double* bvstr, const double* sustr, const double* svstr, const double* u, const double* v,
                              const double* W, double* rufrc, double* rvfrc, double* ru, double* rv) {
  const char* R = "rhs3d_tile"; T();
  A(Hz); A(Huon); A(Hvom); AO(dmde); AO(dndx); A(fomn); A(om_u); A(om_v); A(on_u); A(on_v); A(pm); A(pn); A(bustr); A(bvstr); A(sustr); A(svstr);
  A(u); A(v); A(W); A(rufrc); A(rvfrc); A(ru); A(rv);
  return roms_b200_rhs3d_tile(c, nrhs);         // rufrc, rvfrc are completed by uv3dmix2 and swapped before the fast loop (roms_b200_fast_loop_begin)
}
// ---- step2d_tile, step2d_LF_AM3.h:163-246 (dndx, dmde: CURVGRID; rhoA, rhoS: VAR_RHO_2D -- null in UPWELLING);
// iif and PREDICTOR_2D_STEP come from roms_b200_set_fast_step, iic/ntfirst from roms_b200_set_stepping
int roms_b200_step2d_tile(roms_b200_ctx* c, int ng, int tile, int LBi, int UBi, int LBj, int UBj, int UBk, int IminS, int ImaxS, int JminS, int JmaxS,
                          int krhs, int kstp, int knew, int nstp, int nnew, const double* fomn, const double* h, const double* om_u, const double* om_v,
                          const double* on_u, const double* on_v, const double* omn, const double* pm, const double* pn, const double* dndx, const double* dmde,
                          const double* pmon_r, const double* pnom_r, const double* pmon_p, const double* pnom_p, const double* om_r, const double* on_r,
                          const double* om_p, const double* on_p, const double* visc2_p, const double* visc2_r, const double* rhoA, const double* rhoS,
                          double* DU_avg1, double* DU_avg2, double* DV_avg1, double* DV_avg2, double* Zt_avg1, double* rufrc, double* rvfrc, double* ru, double* rv,
                          double* rubar, double* rvbar, double* rzeta, double* ubar, double* vbar, double* zeta) {
  const char* R = "step2d_tile"; T();
  if (UBk != c->D.b.N) { fprintf(stderr, "roms_b200: step2d_tile: UBk=%d, N=%d\n", UBk, c->D.b.N); return 1; }
  A(fomn); A(h); A(om_u); A(om_v); A(on_u); A(on_v); A(omn); A(pm); A(pn); AO(dndx); AO(dmde); A(pmon_r); A(pnom_r); A(pmon_p); A(pnom_p); A(om_r); A(on_r);
  A(om_p); A(on_p); A(visc2_p); A(visc2_r); AO(rhoA); AO(rhoS); A(DU_avg1); A(DU_avg2); A(DV_avg1); A(DV_avg2); A(Zt_avg1); A(rufrc); A(rvfrc); A(ru); A(rv);
  A(rubar); A(rvbar); A(rzeta); A(ubar); A(vbar); A(zeta);
  if (roms_b200_step2d(c, krhs, kstp, knew, nstp, nnew, c->fast_iif, c->fast_pred, c->iic, c->ntfirst)) return 1;
  // the mp_exchange2d calls of step2d_LF_AM3.h:842,1013,1068,3043, aggregated (same rules as the mirror-resident fast loop)
  const int nfast = c->D.p.nfast;
  if (c->fast_pred) {
    if (c->fast_iif == nfast + 1) { const XF x[3] = {xf2(F(Zt_avg1)), xf2(F(DU_avg1)), xf2(F(DV_avg1))}; return xchg(c, x, 3); }
    if (c->deep) return 0;                                                // deep-halo predictor: evaluated on the halo points the corrector reads
    const XF x[4] = {xf2(F(zeta), knew), xf2(F(ubar), knew), xf2(F(vbar), knew), xf2(F(rzeta), krhs)};
    return xchg(c, x, 4);
  }
  const XF x[3] = {xf2(F(zeta), knew), xf2(F(ubar), knew), xf2(F(vbar), knew)};
  return xchg(c, x, 3);
}
// ---- step3d_uv_tile, step3d_uv.F:134-172
int roms_b200_step3d_uv_tile(roms_b200_ctx* c, int ng, int tile, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                             int nrhs, int nstp, int nnew, const double* om_v, const double* on_u, const double* pm, const double* pn, const double* Hz,
                             const double* z_r, const double* z_w, const double* Akv, const double* DU_avg1, const double* DV_avg1, const double* DU_avg2,
                             const double* DV_avg2, double* ru, double* rv, double* u, double* v, double* ubar, double* vbar, double* Huon, double* Hvom) {
  const char* R = "step3d_uv_tile"; T();
  A(om_v); A(on_u); A(pm); A(pn); A(Hz); A(z_r); A(z_w); A(Akv); A(DU_avg1); A(DV_avg1); A(DU_avg2); A(DV_avg2); A(ru); A(rv); A(u); A(v); A(ubar); A(vbar);
  A(Huon); A(Hvom);
  if (roms_b200_step3d_uv(c, nrhs, nstp, nnew, c->iic, c->ntfirst)) return 1;
  const XF x[8] = {xf3(c, F(u), nnew), xf3(c, F(v), nnew), xf3(c, F(Huon)), xf3(c, F(Hvom)),                  // step3d_uv.F:1805-1824
                   xf2(F(ubar), 1), xf2(F(ubar), 2), xf2(F(vbar), 1), xf2(F(vbar), 2)};
  return xchg(c, x, 8);
}
// ---- step3d_t_tile, step3d_t.F:120-151
int roms_b200_step3d_t_tile(roms_b200_ctx* c, int ng, int tile, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                            int nrhs, int nstp, int nnew, const double* omn, const double* om_u, const double* om_v, const double* on_u, const double* on_v,
                            const double* pm, const double* pn, const double* Hz, const double* Huon, const double* Hvom, const double* z_r, const double* Akt,
                            const double* W, double* t) {
  const char* R = "step3d_t_tile"; T();
  A(omn); A(om_u); A(om_v); A(on_u); A(on_v); A(pm); A(pn); A(Hz); A(Huon); A(Hvom); A(z_r); A(Akt); A(W); A(t);
  if (roms_b200_step3d_t(c, nrhs, nstp, nnew)) return 1;
  XF x[HALO_MAXF]; int n = 0;                                             // mp_exchange4d of t(:,:,:,nnew,:), step3d_t.F:1920
  for (int it = 1; it <= c->D.b.NT && n < HALO_MAXF; ++it) x[n++] = xf3(c, F(t), nnew, it);
  return xchg(c, x, n);
}

// array bounds of a tile array as Utility/get_bounds.F:129-212 computes them for an r2dvar-type variable of an MPI build
// (NghostPoints = Nghost): what the reference host's arrays look like; a helper for C/Python hosts and tests
int roms_b200_mpi_array_bounds(int Lm, int Mm, int NtileI, int NtileJ, int tile, int EWperiodic, int NSperiodic, int Nghost, int* lbub4) {
  if (!lbub4 || NtileI < 1 || NtileJ < 1 || tile < 0 || tile >= NtileI * NtileJ) return 1;
  roms_b200_bounds b;
  if (roms_b200_tile_bounds(Lm, Mm, 1, 1, 1, NtileI, NtileJ, tile, EWperiodic, NSperiodic, 0, &b)) return 1;
  const int Im = Lm + ((Lm + 2) / 2 - (Lm + 1) / 2), Jm = Mm + ((Mm + 2) / 2 - (Mm + 1) / 2);   // mod_param.F:1633-1636
  const int Imin = EWperiodic ? -Nghost : 0, Imax = EWperiodic ? Im + Nghost : Im + 1;
  const int Jmin = NSperiodic ? -Nghost : 0, Jmax = NSperiodic ? Jm + Nghost : Jm + 1;
  lbub4[0] = (b.Itile == 0) ? Imin : b.Istr - Nghost;
  lbub4[1] = (b.Itile == NtileI - 1) ? Imax : b.Iend + Nghost;
  lbub4[2] = (b.Jtile == 0) ? Jmin : b.Jstr - Nghost;
  lbub4[3] = (b.Jtile == NtileJ - 1) ? Jmax : b.Jend + Nghost;
  return 0;
}

}  // extern "C"
