/* include/roms_b200.h -- C ABI of libroms_b200.so
 *
 * Drop-in boundary for the ROMS nonlinear main3d hot path.
 * Each kernel entry point replaces the `CALL X_tile(...)` line inside the public
 * wrapper `X(ng,tile)` of the reference; the Fortran side binds these symbols
 * through ISO_C_BINDING (roms_b200/fortran/roms_b200_mod.F90).
 *
 * Model: the host (Fortran) owns the mod_ocean/mod_grid/mod_coupling/mod_mixing/
 * mod_forces arrays; this library keeps a persistent DEVICE MIRROR of them with
 * the identical layout (column-major, i fastest, explicit lower bounds
 * (LBi:UBi,LBj:UBj[,k][,time][,tracer])).  `roms_b200_upload/download` move a
 * field between `c_loc(array)` and its mirror; the kernel entry points operate
 * on the mirror only and take the hidden module inputs of the reference
 * `_tile` routines (time indices, iic, iif, ...) explicitly.
 *
 * All functions return 0 on success; non-zero maps to exit_flag=8 on the
 * Fortran side (mod_scalars.F:548-561).  There is NO CPU fallback: if no CUDA
 * device is usable, roms_b200_create fails.
 */
#ifndef ROMS_B200_H
#define ROMS_B200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Application option sets (cpp headers ROMS/Include/upwelling.h, benchmark.h) */
#define ROMS_B200_APP_UPWELLING 0
#define ROMS_B200_APP_BENCHMARK 1

/* Per-tile integers: BOUNDS(ng)%X(tile) from ROMS/Include/set_bounds.h:30-74,
 * array bounds from ROMS/Include/tile.h / Utility/get_bounds.F:193-269, and the
 * DOMAIN(ng)%*_Edge(tile) flags.  Passed verbatim; never re-derived on device. */
typedef struct roms_b200_bounds {
  int Lm, Mm, N, NT, NAT;
  int LBi, UBi, LBj, UBj;
  int Istr, Iend, Jstr, Jend;
  int IstrR, IendR, JstrR, JendR;
  int IstrU, JstrV;
  int IstrP, IendP, JstrP, JendP;
  int IstrT, IendT, JstrT, JendT;
  int IstrB, IendB, JstrB, JendB;
  int IstrM, JstrM;
  int Istrm3, Istrm2, Istrm1, IstrUm2, IstrUm1, Iendp1, Iendp2, Iendp2i, Iendp3;
  int Jstrm3, Jstrm2, Jstrm1, JstrVm2, JstrVm1, Jendp1, Jendp2, Jendp2i, Jendp3;
  int Western_Edge, Eastern_Edge, Southern_Edge, Northern_Edge;
  int EWperiodic, NSperiodic;
  int NtileI, NtileJ, Itile, Jtile; /* Utility/get_bounds.F:1020-1039 */
} roms_b200_bounds;

/* Scalars the reference `_tile` routines read from mod_scalars / mod_param. */
typedef struct roms_b200_params {
  int app;               /* ROMS_B200_APP_*: selects the cpp option set */
  double dt, dtfast;     /* dt(ng), dtfast(ng) */
  int ndtfast, nfast;    /* set_weights.F */
  double rho0, g;        /* mod_scalars.F:466 ; roms_*.in RHO0 */
  double gamma2;         /* slipperiness, roms_*.in:466 */
  double hc;             /* set_scoord.F:176-177 */
  double R0, T0, S0, Tcoef, Scoef;   /* linear EOS */
  double Akt_bak[2], Akv_bak;        /* ana_vmix.h */
  double blk_ZQ, blk_ZT, blk_ZW;     /* bulk_flux.F heights */
  double dstart;
} roms_b200_params;

/* Field identifiers of the device mirror.  X(name, kLB, nk, nl, nm):
 * extents beyond (LBi:UBi,LBj:UBj) are (kLB:kLB+nk-1, 1:nl, 1:nm) with the
 * symbolic sizes N (levels), Np1 (0:N), NT, NAT resolved at create time. */
#define ROMS_B200_FIELDS(X) \
  X(h,1,1,1,1) X(f,1,1,1,1) X(fomn,1,1,1,1) X(pm,1,1,1,1) X(pn,1,1,1,1) \
  X(om_r,1,1,1,1) X(on_r,1,1,1,1) X(om_u,1,1,1,1) X(on_u,1,1,1,1) X(om_v,1,1,1,1) X(on_v,1,1,1,1) \
  X(om_p,1,1,1,1) X(on_p,1,1,1,1) X(pmon_r,1,1,1,1) X(pnom_r,1,1,1,1) X(pmon_u,1,1,1,1) X(pnom_u,1,1,1,1) \
  X(pmon_v,1,1,1,1) X(pnom_v,1,1,1,1) X(pmon_p,1,1,1,1) X(pnom_p,1,1,1,1) X(omn,1,1,1,1) \
  X(dndx,1,1,1,1) X(dmde,1,1,1,1) X(lonr,1,1,1,1) X(latr,1,1,1,1) X(xr,1,1,1,1) X(yr,1,1,1,1) X(angler,1,1,1,1) \
  X(rdrag,1,1,1,1) X(rdrag2,1,1,1,1) X(visc2_r,1,1,1,1) X(visc2_p,1,1,1,1) X(hsbl,1,1,1,1) X(Jwtype,1,1,1,1) \
  X(Zt_avg1,1,1,1,1) X(DU_avg1,1,1,1,1) X(DU_avg2,1,1,1,1) X(DV_avg1,1,1,1,1) X(DV_avg2,1,1,1,1) \
  X(rufrc,1,1,1,1) X(rvfrc,1,1,1,1) X(rhoA,1,1,1,1) X(rhoS,1,1,1,1) X(alpha,1,1,1,1) X(beta,1,1,1,1) \
  X(sustr,1,1,1,1) X(svstr,1,1,1,1) X(bustr,1,1,1,1) X(bvstr,1,1,1,1) X(srflx,1,1,1,1) \
  X(Uwind,1,1,1,1) X(Vwind,1,1,1,1) X(Tair,1,1,1,1) X(Pair,1,1,1,1) X(Hair,1,1,1,1) X(cloud,1,1,1,1) X(rain,1,1,1,1) \
  X(lrflx,1,1,1,1) X(lhflx,1,1,1,1) X(shflx,1,1,1,1) \
  X(Hz,1,N,1,1) X(z_r,1,N,1,1) X(z_w,0,Np1,1,1) X(Huon,1,N,1,1) X(Hvom,1,N,1,1) \
  X(diff2,1,NT,1,1) X(Akv,0,Np1,1,1) X(bvf,0,Np1,1,1) X(Akt,0,Np1,NAT,1) X(ghats,0,Np1,NAT,1) \
  X(zeta,1,3,1,1) X(ubar,1,3,1,1) X(vbar,1,3,1,1) X(rzeta,1,2,1,1) X(rubar,1,2,1,1) X(rvbar,1,2,1,1) \
  X(rho,1,N,1,1) X(pden,1,N,1,1) X(W,0,Np1,1,1) X(wvel,0,Np1,1,1) \
  X(u,1,N,2,1) X(v,1,N,2,1) X(ru,0,Np1,2,1) X(rv,0,Np1,2,1) X(t,1,N,3,NT) \
  X(stflx,1,NT,1,1) X(btflx,1,NT,1,1) X(stflux,1,NT,1,1) X(btflux,1,NT,1,1)

enum roms_b200_field {
#define X(name, kLB, nk, nl, nm) ROMS_B200_F_##name,
  ROMS_B200_FIELDS(X)
#undef X
  ROMS_B200_NFIELDS
};

typedef struct roms_b200_ctx roms_b200_ctx;

/* Host helper: fill roms_b200_bounds exactly as Utility/get_bounds.F does for `tile` of an
 * NtileI x NtileJ partition.  distributed!=0: MPI-style tile+halo array bounds (get_bounds.F:193-212),
 * else whole-domain bounds of the serial build (:258-269).  A Fortran host passes BOUNDS(ng) instead. */
int roms_b200_tile_bounds(int Lm, int Mm, int N, int NT, int NAT, int NtileI, int NtileJ, int tile,
                          int EWperiodic, int NSperiodic, int distributed, roms_b200_bounds* out);

/* W,E,S,N neighbour ranks of a tile (Utility/mp_exchange.F:73-197 tile_neighbors); -1 = none */
int roms_b200_tile_neighbors(const roms_b200_bounds* b, int* wesn);

/* Eight-neighbour single-phase halo plan of the NVLink mailbox transport: dir 0 W,1 E,2 S,3 N,4 SW,5 SE,6 NW,7 NE;
 * ranks8[d] = neighbour tile or -1; snd/rcv[4*d..] = {i0,i1,j0,j1}: block sent towards d / destination of the block
 * arriving from d (replaces the two phases of Utility/mp_exchange.F:520-532,761-773 by one exchange with corner messages) */
int roms_b200_halo_plan(const roms_b200_bounds* b, int halo, int* ranks8, int* snd, int* rcv);

/* ---- lifetime (replaces nothing in the reference; called from ROMS_initialize
 *      after ROMS_allocate_arrays, Drivers/nl_roms.h:180, and from ROMS_finalize) */
int roms_b200_create(const roms_b200_bounds* b, const roms_b200_params* p, int device, roms_b200_ctx** out);
int roms_b200_destroy(roms_b200_ctx* ctx);
/* S-coordinate vectors exactly as the reference allocates them (mod_scalars.F:1950-1968): sc_r, Cs_r hold N values (levels 1:N,
 * first element = level 1), sc_w, Cs_w hold N+1 values (levels 0:N): SCALARS(ng)%... (Utility/set_scoord.F) */
int roms_b200_set_scoord(roms_b200_ctx* ctx, const double* sc_r, const double* Cs_r, const double* sc_w, const double* Cs_w);
/* weight(1,:,ng), weight(2,:,ng) (Utility/set_weights.F); arrays are 1-based: w[0] unused, entries 1..nfast+2 are read
 * (length >= nfast+3) */
int roms_b200_set_weights(roms_b200_ctx* ctx, int nfast, const double* weight1, const double* weight2);

/* ---- host <-> device mirror (c_loc of the module array) */
int roms_b200_field_id(const char* name);
long roms_b200_field_size(const roms_b200_ctx* ctx, int field);     /* doubles */
int roms_b200_upload(roms_b200_ctx* ctx, int field, const double* host);
int roms_b200_download(roms_b200_ctx* ctx, int field, double* host);
void* roms_b200_device_ptr(roms_b200_ctx* ctx, int field);
int roms_b200_sync(roms_b200_ctx* ctx);

/* ---- per-tile kernels: each replaces `CALL X_tile(ng,tile,...)` in the wrapper `X(ng,tile)` */
int roms_b200_set_massflux(roms_b200_ctx* ctx, int nrhs);                      /* set_massflux.F:73   */
int roms_b200_rho_eos(roms_b200_ctx* ctx, int nrhs);                           /* rho_eos.F:111,576  */
int roms_b200_omega(roms_b200_ctx* ctx);                                       /* omega.F:96          */
int roms_b200_wvelocity(roms_b200_ctx* ctx, int ninp);                         /* wvelocity.F:63 (main3d.F:535, ninp=nstp) */
int roms_b200_set_zeta(roms_b200_ctx* ctx);                                    /* set_zeta.F:59       */
int roms_b200_set_depth(roms_b200_ctx* ctx);                                   /* set_depth.F:76      */
int roms_b200_bulk_flux(roms_b200_ctx* ctx, int nrhs);                         /* bulk_flux.F:111     */
int roms_b200_set_vbc(roms_b200_ctx* ctx, int nrhs);                           /* set_vbc.F           */
int roms_b200_ana_vmix(roms_b200_ctx* ctx);                                    /* ana_vmix.h          */
int roms_b200_lmd_vmix(roms_b200_ctx* ctx, int nstp);                          /* lmd_vmix.F:33       */
int roms_b200_pre_step3d(roms_b200_ctx* ctx, int nrhs, int nstp, int nnew, int iic, int ntfirst);   /* pre_step3d.F:126 */
int roms_b200_prsgrd(roms_b200_ctx* ctx, int nrhs);                            /* prsgrd32.h:109      */
int roms_b200_t3dmix2(roms_b200_ctx* ctx, int nrhs, int nstp, int nnew);       /* t3dmix2_s.h / _geo.h */
int roms_b200_rhs3d_tile(roms_b200_ctx* ctx, int nrhs);                        /* rhs3d.F:196         */
int roms_b200_uv3dmix2(roms_b200_ctx* ctx, int nrhs, int nnew);                /* uv3dmix2_s.h        */
/* rhs3d wrapper: pre_step3d -> prsgrd -> t3dmix2 -> rhs3d_tile -> uv3dmix2 (rhs3d.F:80-181) */
int roms_b200_rhs3d(roms_b200_ctx* ctx, int nrhs, int nstp, int nnew, int iic, int ntfirst);
int roms_b200_step2d(roms_b200_ctx* ctx, int krhs, int kstp, int knew, int nstp, int nnew,
                     int iif, int predictor_2d_step, int iic, int ntfirst);    /* step2d_LF_AM3.h:163 */
int roms_b200_step3d_uv(roms_b200_ctx* ctx, int nrhs, int nstp, int nnew, int iic, int ntfirst);   /* step3d_uv.F:134 */
int roms_b200_step3d_t(roms_b200_ctx* ctx, int nrhs, int nstp, int nnew);      /* step3d_t.F:120      */
/* diag_tile reductions: out[0]=avgke, out[1]=avgpe, out[2]=volume (diag.F:225-322) */
int roms_b200_diag(roms_b200_ctx* ctx, int nstp, double* out3);

/* Everything diag_tile reports (diag.F:209-411,512-542), of the last completed roms_b200_diag / roms_b200_diag_end:
 * out[0..2] as above, out[3..6] = max_C, max_Cu, max_Cv, max_Cw (largest Courant number and its components),
 * out[7..9] = max_Ci, max_Cj, max_Ck (its location; first point in the reference's scan order), out[10] = maxspeed,
 * out[11] = maxrho, out[12] = exit_flag the reference would set (0 = NoError, 1 = blow-up: non-finite energies,
 * maxspeed > max_speed = 20 m/s or maxrho > max_rho = 200 kg/m3, mod_scalars.F:573-574). */
#define ROMS_B200_NDIAG 13
int roms_b200_diag_full(roms_b200_ctx* ctx, int nstp, double* out13);
int roms_b200_diag_last(const roms_b200_ctx* ctx, double* out13);

/* diag in two halves (launch + asynchronous D2H of the partial sums ; wait + final sums + mp_reduce) so that the host can
 * evaluate the next step's set_data while the device works: diag.F:225-322,405 */
int roms_b200_diag_begin(roms_b200_ctx* ctx, int nstp);
int roms_b200_diag_end(roms_b200_ctx* ctx, double* out3);
/* pinned host staging buffers and an upload that does not block the host (the buffer must stay untouched until the next
 * roms_b200_sync / roms_b200_diag_end): the sync points of `set_data` (Nonlinear/set_data.F) */
int roms_b200_host_alloc(size_t bytes, void** p);
int roms_b200_host_free(void* p);
int roms_b200_upload_async(roms_b200_ctx* ctx, int field, const double* pinned_host);

/* ---- whole fast loop and whole baroclinic step on the device mirror.
 * roms_b200_step2d_loop runs main3d.F:810-918 (2*nfast+1 step2d calls) and
 * updates *indx1 as the reference does.  roms_b200_main3d runs nsteps of
 * main3d.F:216-1148 with forcing taken from the mirror (upload it beforehand,
 * or let analytic_forcing!=0 evaluate ana_* on the device each step).
 * with_diag: 0 = no diag; 1 = diag at the reference's place in every step (after rho_eos, main3d.F:300), the host waits for it;
 * 2 = same place, but only launched (reductions and their D2H copy stay asynchronous): roms_b200_diag_end after the call
 * returns the diag of the LAST step's start state -- the line the reference prints for that step. */
int roms_b200_step2d_loop(roms_b200_ctx* ctx, int nstp, int nnew, int iic, int ntfirst, int* indx1);
int roms_b200_main3d(roms_b200_ctx* ctx, int nsteps, int analytic_forcing, int with_diag);
/* stepping state of the mirror-resident loop: iic, ntfirst, nstp, nnew, nrhs, indx1 ; time (s) */
int roms_b200_get_stepping(const roms_b200_ctx* ctx, int* out6, double* time);
int roms_b200_set_stepping(roms_b200_ctx* ctx, const int* in6, double time);
/* analytic forcing on the device: set_data.F (ana_* branches) */
int roms_b200_set_data(roms_b200_ctx* ctx, double tdays);
/* number of kernel launches issued by this context so far */
long roms_b200_launch_count(const roms_b200_ctx* ctx);
/* average device time (ms) of the last `roms_b200_time_kernel` call */
int roms_b200_time_step3d_t(roms_b200_ctx* ctx, int nrhs, int nstp, int nnew, int reps, float* ms_avg);

/* ---- multi-GPU: one rank = one tile = one GPU (Drivers/nl_roms.h:145-157).  The halo swaps of
 * Utility/mp_exchange.F (mp_exchange2d/3d/4d) become NCCL send/recv inside roms_b200_main3d /
 * roms_b200_step2d_loop.  Rank 0 creates the id, the host broadcasts it (MPI_Bcast / torch.distributed),
 * every rank calls comm_init with rank == Jtile*NtileI+Itile of its bounds. */
int roms_b200_comm_unique_id(char* id128);
int roms_b200_comm_init(roms_b200_ctx* ctx, int rank, int nranks, const char* id128);
int roms_b200_comm_destroy(roms_b200_ctx* ctx);
/* NVLink peer mailboxes (default halo transport once connected; NCCL remains for the diag all-reduce and as
 * ROMS_B200_HALO_NCCL=1 fallback).  Every rank exports a 64-byte CUDA IPC handle; the host all-gathers them
 * (MPI_Allgather / torch.distributed) and passes the table (nranks x 64 bytes, indexed by rank = tile id). */
int roms_b200_p2p_handle(roms_b200_ctx* ctx, char* handle64);
int roms_b200_p2p_connect(roms_b200_ctx* ctx, const char* handles, int nranks);
/* copy the interior (Istr:Iend,Jstr:Jend) of `nplanes` consecutive (i,j) planes of a field, starting at
 * storage plane `plane0`, into a dense host buffer (nplanes, Jend-Jstr+1, Iend-Istr+1) */
int roms_b200_download_interior(roms_b200_ctx* ctx, int field, int plane0, int nplanes, double* host);


/* ---- the boundary as a Fortran MPI host sees it (roms_b200/csrc/tile_api.cu).
 * Host arrays keep THEIR bounds: (LBi:UBi,LBj:UBj) of Utility/get_bounds.F:193-212 with NghostPoints = 2, while a distributed
 * mirror carries a halo of 6.  upload/download_bounds move the intersection of the two index boxes, all planes of the field;
 * exchange_field then refreshes the mirror's wider halo from the neighbour tiles.  register_field remembers c_loc(array) and its
 * bounds: the *_registered calls and the argument checks of the *_tile entry points below use it. */
int roms_b200_register_field(roms_b200_ctx* ctx, int field, const double* host, int LBi, int UBi, int LBj, int UBj);
int roms_b200_upload_bounds(roms_b200_ctx* ctx, int field, const double* host, int LBi, int UBi, int LBj, int UBj);
int roms_b200_download_bounds(roms_b200_ctx* ctx, int field, double* host, int LBi, int UBi, int LBj, int UBj);
int roms_b200_upload_registered(roms_b200_ctx* ctx, int field);
int roms_b200_download_registered(roms_b200_ctx* ctx, int field);
int roms_b200_exchange_field(roms_b200_ctx* ctx, int field);
/* array bounds {LBi,UBi,LBj,UBj} of a tile array in an MPI build of the reference (get_bounds.F:129-212, r2dvar) */
int roms_b200_mpi_array_bounds(int Lm, int Mm, int NtileI, int NtileJ, int tile, int EWperiodic, int NSperiodic, int Nghost, int* lbub4);
/* iif(ng), PREDICTOR_2D_STEP(ng) (mod_scalars, set by main3d.F:820-880) for the next step2d_tile; the rufrc/rvfrc swap the
 * deep-halo predictor needs before the first sub-step of a baroclinic step (no-op on one tile) */
int roms_b200_set_fast_step(roms_b200_ctx* ctx, int iif, int predictor_2d_step);
int roms_b200_fast_loop_begin(roms_b200_ctx* ctx);

/* ---- `_tile` entry points: the argument lists of the reference's X_tile routines for the UPWELLING / BENCHMARK cpp sets, so that
 * the wrapper X(ng,tile) changes one token (CALL X_tile -> rc = roms_b200_ X _tile with ctx as first argument).  Array arguments are the HOST
 * arrays; each is checked against its registration (address and bounds), then the call runs on the mirror and performs the halo
 * swaps the reference routine ends with.  Arguments that exist only under cpp options one application lacks may be null:
 * dndx, dmde (CURVGRID), rhoA, rhoS (VAR_RHO_2D), srflx (SOLAR_SOURCE), ghats (LMD_NONLOCAL).  iic, ntfirst: roms_b200_set_stepping. */
int roms_b200_set_massflux_tile(roms_b200_ctx* ctx, int ng, int tile, int model, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                                int nrhs, const double* u, const double* v, const double* Hz, const double* om_v, const double* on_u,
                                double* Huon, double* Hvom);                                             /* set_massflux.F:73-82 */
int roms_b200_omega_tile(roms_b200_ctx* ctx, int ng, int tile, int model, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                         const double* Huon, const double* Hvom, const double* z_w, double* W);          /* omega.F:96-114 */
int roms_b200_set_zeta_tile(roms_b200_ctx* ctx, int ng, int tile, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                            const double* Zt_avg1, double* zeta);                                        /* set_zeta.F:59-62 */
int roms_b200_set_depth_tile(roms_b200_ctx* ctx, int ng, int tile, int model, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                             int nstp, int nnew, const double* h, const double* Zt_avg1, double* Hz, double* z_r, double* z_w);   /* set_depth.F:76-85 */
int roms_b200_pre_step3d_tile(roms_b200_ctx* ctx, int ng, int tile, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                              int nrhs, int nstp, int nnew, const double* pm, const double* pn, const double* Hz, const double* Huon, const double* Hvom,
                              const double* z_r, const double* z_w, const double* btflx, const double* bustr, const double* bvstr, const double* stflx,
                              const double* sustr, const double* svstr, const double* srflx, const double* Akt, const double* Akv, const double* ghats,
                              const double* W, const double* ru, const double* rv, double* t, double* u, double* v);   /* pre_step3d.F:126-160 */
int roms_b200_prsgrd32_tile(roms_b200_ctx* ctx, int ng, int tile, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                            int nrhs, const double* om_v, const double* on_u, const double* Hz, const double* z_r, const double* z_w, const double* rho,
                            double* ru, double* rv);                                                     /* prsgrd32.h:109-134 */
int roms_b200_rhs3d_tile_tile(roms_b200_ctx* ctx, int ng, int tile, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                              int nrhs, const double* Hz, const double* Huon, const double* Hvom, const double* dmde, const double* dndx, const double* fomn,
                              const double* om_u, const double* om_v, const double* on_u, const double* on_v, const double* pm, const double* pn,
                              const double* bustr, const double* bvstr, const double* sustr, const double* svstr, const double* u, const double* v,
                              const double* W, double* rufrc, double* rvfrc, double* ru, double* rv);    /* rhs3d.F:196-221 */
int roms_b200_step2d_tile(roms_b200_ctx* ctx, int ng, int tile, int LBi, int UBi, int LBj, int UBj, int UBk, int IminS, int ImaxS, int JminS, int JmaxS,
                          int krhs, int kstp, int knew, int nstp, int nnew, const double* fomn, const double* h, const double* om_u, const double* om_v,
                          const double* on_u, const double* on_v, const double* omn, const double* pm, const double* pn, const double* dndx, const double* dmde,
                          const double* pmon_r, const double* pnom_r, const double* pmon_p, const double* pnom_p, const double* om_r, const double* on_r,
                          const double* om_p, const double* on_p, const double* visc2_p, const double* visc2_r, const double* rhoA, const double* rhoS,
                          double* DU_avg1, double* DU_avg2, double* DV_avg1, double* DV_avg2, double* Zt_avg1, double* rufrc, double* rvfrc, double* ru, double* rv,
                          double* rubar, double* rvbar, double* rzeta, double* ubar, double* vbar, double* zeta);     /* step2d_LF_AM3.h:163-246 */
int roms_b200_step3d_uv_tile(roms_b200_ctx* ctx, int ng, int tile, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                             int nrhs, int nstp, int nnew, const double* om_v, const double* on_u, const double* pm, const double* pn, const double* Hz,
                             const double* z_r, const double* z_w, const double* Akv, const double* DU_avg1, const double* DV_avg1, const double* DU_avg2,
                             const double* DV_avg2, double* ru, double* rv, double* u, double* v, double* ubar, double* vbar, double* Huon, double* Hvom);
                                                                                                         /* step3d_uv.F:134-172 */
int roms_b200_step3d_t_tile(roms_b200_ctx* ctx, int ng, int tile, int LBi, int UBi, int LBj, int UBj, int IminS, int ImaxS, int JminS, int JmaxS,
                            int nrhs, int nstp, int nnew, const double* omn, const double* om_u, const double* om_v, const double* on_u, const double* on_v,
                            const double* pm, const double* pn, const double* Hz, const double* Huon, const double* Hvom, const double* z_r, const double* Akt,
                            const double* W, double* t);                                                 /* step3d_t.F:120-151 */

/* ---- output path without stalling the time loop (what `output` needs before wrt_his / wrt_rst / wrt_avg, output.F:217,703).
 * roms_b200_snapshot_begin copies the listed fields device-to-device into a staging area ON THE LAUNCH STREAM (it orders after
 * every kernel launched so far and costs microseconds), then device-to-host into the caller's PINNED buffers
 * (roms_b200_host_alloc; buffer q holds roms_b200_field_size(fields[q]) doubles) on a separate copy stream, and returns at
 * once: the following steps run while the snapshot drains over PCIe.  roms_b200_snapshot_end waits for the host copies; one
 * snapshot may be in flight at a time (a second begin before end returns 1). */
int roms_b200_snapshot_begin(roms_b200_ctx* ctx, int nfields, const int* fields, double* const* pinned_host);
int roms_b200_snapshot_end(roms_b200_ctx* ctx);

/* PERFECT_RESTART (Utility/wrt_rst.F:178-211,345-900): ids of the fields a bit-identical continuation needs (returns the count, -1
 * if `cap` is too small); write them with roms_b200_snapshot_begin/end or roms_b200_download together with the six stepping
 * integers and the time (roms_b200_get_stepping).  To restart: initialise a context as usual, upload the set, roms_b200_set_stepping,
 * roms_b200_restart_finish (what post_initial repeats on restart: depths from Zt_avg1). */
int roms_b200_restart_fields(const roms_b200_ctx* ctx, int* ids, int cap);
int roms_b200_restart_finish(roms_b200_ctx* ctx);

/* CUDA-event stopwatch on the context's launch stream, and an L2 flush (writes `mbytes` MiB) */
int roms_b200_timer_start(roms_b200_ctx* ctx);
int roms_b200_timer_stop(roms_b200_ctx* ctx, float* ms);
int roms_b200_flush_l2(roms_b200_ctx* ctx, int mbytes);

/* ---- start-up helpers on the device mirror */
int roms_b200_fill(roms_b200_ctx* ctx, int field, double value);       /* mod_*.F initialisation values */
int roms_b200_ana_initial(roms_b200_ctx* ctx);                         /* Functionals/ana_initial.h */
int roms_b200_ini_fields(roms_b200_ctx* ctx, int nstp, int kstp);      /* Nonlinear/ini_fields.F (ini_zeta + ini_fields) */

/* ---- driver surface of Master/roms_kernel.F -> Drivers/nl_roms.h for the analytical applications
 * (host side written in C++ because no Fortran compiler exists in the build image; a Fortran ROMS
 * keeps its own ROMS_initialize/run/finalize and only binds the kernel entry points above). */
typedef struct roms_b200_config {       /* roms_*.in values (Utility/read_phypar.F keywords) */
  int app, Lm, Mm, N, NT, NAT, NtileI, NtileJ;
  double dt; int ndtfast;
  double theta_s, theta_b, Tcline;
  double rho0, g, gamma2, rdrg, rdrg2;
  double Akt_bak[2], Akv_bak, tnu2[2], visc2;
  double R0, T0, S0, Tcoef, Scoef;
  double blk_ZQ, blk_ZT, blk_ZW; int lmd_Jwt;
} roms_b200_config;
typedef struct roms_b200_driver roms_b200_driver;
void roms_b200_default_config(int app, int Lm, int Mm, int N, roms_b200_config* cfg);   /* roms_upwelling.in / roms_benchmark1.in */
void roms_b200_host_scoord(int N, double theta_s, double theta_b, double* sc_r, double* Cs_r, double* sc_w, double* Cs_w); /* set_scoord.F */
int roms_b200_host_weights(int ndtfast, double* weight1, double* weight2);              /* set_weights.F; returns nfast */
int roms_b200_ROMS_initialize(const roms_b200_config* cfg, int tile, int distributed, int device, roms_b200_driver** out); /* nl_roms.h:61 */
int roms_b200_ROMS_run(roms_b200_driver* drv, int nsteps, int host_forcing, double* diag3);                                 /* nl_roms.h:247 */
int roms_b200_ROMS_finalize(roms_b200_driver* drv);                                                                         /* nl_roms.h:320 */
roms_b200_ctx* roms_b200_driver_ctx(roms_b200_driver* drv);
void roms_b200_driver_bounds(roms_b200_driver* drv, roms_b200_bounds* b);
int roms_b200_driver_nfast(roms_b200_driver* drv);

#ifdef __cplusplus
}
#endif
#endif /* ROMS_B200_H */
