!  roms_b200/fortran/roms_b200_mod.F90
!
!  ISO_C_BINDING interfaces of libroms_b200.so (include/roms_b200.h) for the
!  Fortran host (the unmodified myroms/roms sources).  NOT compiled in this
!  repository's image (no Fortran compiler there); INTEGRATION.md shows how it
!  is wired in with a new cpp option B200_KERNELS next to NONLINEAR.
!
!  Every kernel entry point replaces the `CALL X_tile (ng, tile, ...)` line of
!  the public wrapper `X (ng, tile)`; the wrapper keeps its wclock_on/off
!  profiling hooks (e.g. ROMS/Nonlinear/step3d_t.F:67,113).
!
      MODULE roms_b200_mod
      USE, INTRINSIC :: ISO_C_BINDING
      implicit none
      PUBLIC
!
!  Opaque device-mirror handle, one per (ng, tile) = one per MPI rank.
!
      type(c_ptr), save :: b200_ctx = c_null_ptr
!
!  BOUNDS(ng)%X(tile) integers, same order as `roms_b200_bounds`.
!
      TYPE, BIND(C) :: roms_b200_bounds
        integer(c_int) :: Lm, Mm, N, NT, NAT
        integer(c_int) :: LBi, UBi, LBj, UBj
        integer(c_int) :: Istr, Iend, Jstr, Jend
        integer(c_int) :: IstrR, IendR, JstrR, JendR
        integer(c_int) :: IstrU, JstrV
        integer(c_int) :: IstrP, IendP, JstrP, JendP
        integer(c_int) :: IstrT, IendT, JstrT, JendT
        integer(c_int) :: IstrB, IendB, JstrB, JendB
        integer(c_int) :: IstrM, JstrM
        integer(c_int) :: Istrm3, Istrm2, Istrm1, IstrUm2, IstrUm1
        integer(c_int) :: Iendp1, Iendp2, Iendp2i, Iendp3
        integer(c_int) :: Jstrm3, Jstrm2, Jstrm1, JstrVm2, JstrVm1
        integer(c_int) :: Jendp1, Jendp2, Jendp2i, Jendp3
        integer(c_int) :: Western_Edge, Eastern_Edge
        integer(c_int) :: Southern_Edge, Northern_Edge
        integer(c_int) :: EWperiodic, NSperiodic
        integer(c_int) :: NtileI, NtileJ, Itile, Jtile
      END TYPE roms_b200_bounds
!
!  mod_scalars / mod_param values the _tile routines read from modules.
!
      TYPE, BIND(C) :: roms_b200_params
        integer(c_int) :: app
        real(c_double) :: dt, dtfast
        integer(c_int) :: ndtfast, nfast
        real(c_double) :: rho0, g, gamma2, hc
        real(c_double) :: R0, T0, S0, Tcoef, Scoef
        real(c_double) :: Akt_bak(2), Akv_bak
        real(c_double) :: blk_ZQ, blk_ZT, blk_ZW, dstart
      END TYPE roms_b200_params

      INTERFACE
!
!  Lifetime and host <-> device mirror.
!
        integer(c_int) FUNCTION roms_b200_create (b, p, device, ctx)    &
     &                          BIND(C, name='roms_b200_create')
          IMPORT
          type(roms_b200_bounds), intent(in) :: b
          type(roms_b200_params), intent(in) :: p
          integer(c_int), value :: device
          type(c_ptr), intent(out) :: ctx
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_destroy (ctx)                 &
     &                          BIND(C, name='roms_b200_destroy')
          IMPORT
          type(c_ptr), value :: ctx
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_field_id (name)               &
     &                          BIND(C, name='roms_b200_field_id')
          IMPORT
          character(kind=c_char), intent(in) :: name(*)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_upload (ctx, field, host)     &
     &                          BIND(C, name='roms_b200_upload')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: field
          type(c_ptr), value :: host          ! c_loc(OCEAN(ng)%t) ...
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_download (ctx, field, host)   &
     &                          BIND(C, name='roms_b200_download')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: field
          type(c_ptr), value :: host
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_set_scoord (ctx, sc_r, Cs_r,  &
     &                                               sc_w, Cs_w)        &
     &                          BIND(C, name='roms_b200_set_scoord')
          IMPORT
          type(c_ptr), value :: ctx
          real(c_double), intent(in) :: sc_r(*), Cs_r(*), sc_w(*), Cs_w(*)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_set_weights (ctx, nfast, w1,  &
     &                                                w2)               &
     &                          BIND(C, name='roms_b200_set_weights')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nfast
          real(c_double), intent(in) :: w1(*), w2(*)
        END FUNCTION
!
!  Per-tile kernels (argument = the hidden module inputs of X_tile).
!
        integer(c_int) FUNCTION roms_b200_set_massflux (ctx, nrhs)      &
     &                          BIND(C, name='roms_b200_set_massflux')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nrhs
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_rho_eos (ctx, nrhs)           &
     &                          BIND(C, name='roms_b200_rho_eos')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nrhs
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_omega (ctx)                   &
     &                          BIND(C, name='roms_b200_omega')
          IMPORT
          type(c_ptr), value :: ctx
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_set_zeta (ctx)                &
     &                          BIND(C, name='roms_b200_set_zeta')
          IMPORT
          type(c_ptr), value :: ctx
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_set_depth (ctx)               &
     &                          BIND(C, name='roms_b200_set_depth')
          IMPORT
          type(c_ptr), value :: ctx
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_bulk_flux (ctx, nrhs)         &
     &                          BIND(C, name='roms_b200_bulk_flux')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nrhs
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_set_vbc (ctx, nrhs)           &
     &                          BIND(C, name='roms_b200_set_vbc')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nrhs
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_lmd_vmix (ctx, nstp)          &
     &                          BIND(C, name='roms_b200_lmd_vmix')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nstp
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_ana_vmix (ctx)                &
     &                          BIND(C, name='roms_b200_ana_vmix')
          IMPORT
          type(c_ptr), value :: ctx
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_rhs3d (ctx, nrhs, nstp, nnew, &
     &                                          iic, ntfirst)           &
     &                          BIND(C, name='roms_b200_rhs3d')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nrhs, nstp, nnew, iic, ntfirst
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_pre_step3d (ctx, nrhs, nstp,  &
     &                                   nnew, iic, ntfirst)            &
     &                          BIND(C, name='roms_b200_pre_step3d')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nrhs, nstp, nnew, iic, ntfirst
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_prsgrd (ctx, nrhs)            &
     &                          BIND(C, name='roms_b200_prsgrd')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nrhs
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_step2d (ctx, krhs, kstp, knew,&
     &                     nstp, nnew, iif, predictor, iic, ntfirst)    &
     &                          BIND(C, name='roms_b200_step2d')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: krhs, kstp, knew, nstp, nnew
          integer(c_int), value :: iif, predictor, iic, ntfirst
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_step3d_uv (ctx, nrhs, nstp,   &
     &                                   nnew, iic, ntfirst)            &
     &                          BIND(C, name='roms_b200_step3d_uv')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nrhs, nstp, nnew, iic, ntfirst
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_step3d_t (ctx, nrhs, nstp,    &
     &                                             nnew)                &
     &                          BIND(C, name='roms_b200_step3d_t')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nrhs, nstp, nnew
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_diag (ctx, nstp, out3)        &
     &                          BIND(C, name='roms_b200_diag')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nstp
          real(c_double), intent(out) :: out3(3)
        END FUNCTION
!
!  wvelocity.F:43 -> CALL roms_b200_wvelocity (ctx, Ninp)
!
        integer(c_int) FUNCTION roms_b200_wvelocity (ctx, ninp)         &
     &                          BIND(C, name='roms_b200_wvelocity')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: ninp
        END FUNCTION
!
!  diag.F:209-411,512-542 complete: avgke, avgpe, volume, max_C, max_Cu,
!  max_Cv, max_Cw, max_Ci, max_Cj, max_Ck, maxspeed, maxrho, exit_flag.
!
        integer(c_int) FUNCTION roms_b200_diag_full (ctx, nstp, out13)  &
     &                          BIND(C, name='roms_b200_diag_full')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nstp
          real(c_double), intent(out) :: out13(13)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_diag_last (ctx, out13)        &
     &                          BIND(C, name='roms_b200_diag_last')
          IMPORT
          type(c_ptr), value :: ctx
          real(c_double), intent(out) :: out13(13)
        END FUNCTION
!
!  Synchronisation + device error word (non-zero -> exit_flag=8), pieces of
!  rhs3d, the whole fast loop (main3d.F:810-918, indx1 updated as the
!  reference does) and the whole step on the mirror.
!
        integer(c_int) FUNCTION roms_b200_sync (ctx)                    &
     &                          BIND(C, name='roms_b200_sync')
          IMPORT
          type(c_ptr), value :: ctx
        END FUNCTION
        integer(c_long) FUNCTION roms_b200_field_size (ctx, field)      &
     &                          BIND(C, name='roms_b200_field_size')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: field
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_rhs3d_tile (ctx, nrhs)        &
     &                          BIND(C, name='roms_b200_rhs3d_tile')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nrhs
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_t3dmix2 (ctx, nrhs, nstp,     &
     &                                            nnew)                 &
     &                          BIND(C, name='roms_b200_t3dmix2')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nrhs, nstp, nnew
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_uv3dmix2 (ctx, nrhs, nnew)    &
     &                          BIND(C, name='roms_b200_uv3dmix2')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nrhs, nnew
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_step2d_loop (ctx, nstp, nnew, &
     &                                   iic, ntfirst, indx1)           &
     &                          BIND(C, name='roms_b200_step2d_loop')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nstp, nnew, iic, ntfirst
          integer(c_int), intent(inout) :: indx1
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_main3d (ctx, nsteps,          &
     &                                   analytic_forcing, with_diag)   &
     &                          BIND(C, name='roms_b200_main3d')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nsteps, analytic_forcing, with_diag
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_get_stepping (ctx, out6,      &
     &                                                 time)            &
     &                          BIND(C, name='roms_b200_get_stepping')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), intent(out) :: out6(6)   ! iic ntfirst nstp nnew nrhs indx1
          real(c_double), intent(out) :: time
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_set_stepping (ctx, in6, time) &
     &                          BIND(C, name='roms_b200_set_stepping')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), intent(in) :: in6(6)
          real(c_double), value :: time
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_set_data (ctx, tdays)         &
     &                          BIND(C, name='roms_b200_set_data')
          IMPORT
          type(c_ptr), value :: ctx
          real(c_double), value :: tdays
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_download_interior (ctx,       &
     &                             field, plane0, nplanes, host)        &
     &                    BIND(C, name='roms_b200_download_interior')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: field, plane0, nplanes
          type(c_ptr), value :: host
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_comm_destroy (ctx)            &
     &                          BIND(C, name='roms_b200_comm_destroy')
          IMPORT
          type(c_ptr), value :: ctx
        END FUNCTION
!
!  Output path (output.F:217,703): asynchronous snapshot of nfields mirror
!  fields into pinned host buffers (roms_b200_host_alloc); the time loop
!  continues, roms_b200_snapshot_end waits before wrt_his / wrt_rst read them.
!
        integer(c_int) FUNCTION roms_b200_snapshot_begin (ctx, nfields, &
     &                                   fields, pinned_host)           &
     &                          BIND(C, name='roms_b200_snapshot_begin')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nfields
          integer(c_int), intent(in) :: fields(*)
          type(c_ptr), intent(in) :: pinned_host(*)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_snapshot_end (ctx)            &
     &                          BIND(C, name='roms_b200_snapshot_end')
          IMPORT
          type(c_ptr), value :: ctx
        END FUNCTION
!
!  NCCL communicator (replaces mp_exchange2d/3d/4d): id from rank 0 via
!  mpi_bcast, then every rank calls comm_init.
!
        integer(c_int) FUNCTION roms_b200_comm_unique_id (id128)        &
     &                          BIND(C, name='roms_b200_comm_unique_id')
          IMPORT
          character(kind=c_char) :: id128(128)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_comm_init (ctx, rank, nranks, &
     &                                              id128)              &
     &                          BIND(C, name='roms_b200_comm_init')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: rank, nranks
          character(kind=c_char), intent(in) :: id128(128)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_p2p_handle (ctx, handle64)    &
     &                          BIND(C, name='roms_b200_p2p_handle')
          IMPORT
          type(c_ptr), value :: ctx
          character(kind=c_char) :: handle64(64)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_p2p_connect (ctx, handles,    &
     &                                                nranks)           &
     &                          BIND(C, name='roms_b200_p2p_connect')
          IMPORT
          type(c_ptr), value :: ctx
          character(kind=c_char), intent(in) :: handles(*)
          integer(c_int), value :: nranks
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_upload_async (ctx, field,     &
     &                                                 pinned_host)     &
     &                          BIND(C, name='roms_b200_upload_async')
          IMPORT
          type(c_ptr), value :: ctx, pinned_host
          integer(c_int), value :: field
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_host_alloc (bytes, p)         &
     &                          BIND(C, name='roms_b200_host_alloc')
          IMPORT
          integer(c_size_t), value :: bytes
          type(c_ptr) :: p
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_host_free (p)                 &
     &                          BIND(C, name='roms_b200_host_free')
          IMPORT
          type(c_ptr), value :: p
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_diag_begin (ctx, nstp)        &
     &                          BIND(C, name='roms_b200_diag_begin')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: nstp
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_diag_end (ctx, out3)          &
     &                          BIND(C, name='roms_b200_diag_end')
          IMPORT
          type(c_ptr), value :: ctx
          real(c_double) :: out3(3)
        END FUNCTION
!
!  The boundary as an MPI host sees it (roms_b200/csrc/tile_api.cu): host arrays keep their own bounds (LBi:UBi,LBj:UBj) with
!  NghostPoints = 2; register them once, then upload/download by field id.  The *_tile entry points take the argument lists of the
!  reference's X_tile routines (UPWELLING / BENCHMARK cpp sets), check every array against its registration, run on the mirror and
!  do the halo swaps the reference routine ends with.  Generated by tools/gen_fortran_iface.py from include/roms_b200.h.
!
        integer(c_int) FUNCTION roms_b200_register_field (ctx,        &
     &    field, host, LBi, UBi, LBj, UBj)                            &
     &                          BIND(C, name='roms_b200_register_field')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: field, LBi, UBi, LBj, UBj
          real(c_double), intent(in) :: host(*)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_upload_bounds (ctx,         &
     &    field, host, LBi, UBi, LBj, UBj)                            &
     &                          BIND(C, name='roms_b200_upload_bounds')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: field, LBi, UBi, LBj, UBj
          real(c_double), intent(in) :: host(*)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_download_bounds (ctx,       &
     &    field, host, LBi, UBi, LBj, UBj)                            &
     &                          BIND(C, name='roms_b200_download_bounds')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: field, LBi, UBi, LBj, UBj
          real(c_double), intent(inout) :: host(*)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_upload_registered (ctx,     &
     &    field)                                                      &
     &                          BIND(C, name='roms_b200_upload_registered')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: field
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_download_registered (       &
     &    ctx, field)                                                 &
     &                          BIND(C, name='roms_b200_download_registered')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: field
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_exchange_field (ctx,        &
     &    field)                                                      &
     &                          BIND(C, name='roms_b200_exchange_field')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: field
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_mpi_array_bounds (Lm,       &
     &    Mm, NtileI, NtileJ, tile, EWperiodic, NSperiodic,           &
     &    Nghost, lbub4)                                              &
     &                          BIND(C, name='roms_b200_mpi_array_bounds')
          IMPORT
          integer(c_int), value :: Lm, Mm, NtileI, NtileJ, tile,      &
     &      EWperiodic, NSperiodic, Nghost
          integer(c_int) :: lbub4(*)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_set_fast_step (ctx,         &
     &    iif, predictor_2d_step)                                     &
     &                          BIND(C, name='roms_b200_set_fast_step')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: iif, predictor_2d_step
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_fast_loop_begin (ctx)       &
     &                          BIND(C, name='roms_b200_fast_loop_begin')
          IMPORT
          type(c_ptr), value :: ctx
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_set_massflux_tile (ctx,     &
     &    ng, tile, model, LBi, UBi, LBj, UBj, IminS, ImaxS,          &
     &    JminS, JmaxS, nrhs, u, v, Hz, om_v, on_u, Huon, Hvom)       &
     &                          BIND(C, name='roms_b200_set_massflux_tile')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: ng, tile, model, LBi, UBi,         &
     &      LBj, UBj, IminS, ImaxS, JminS, JmaxS, nrhs
          real(c_double), intent(in) :: u(*), v(*), Hz(*),            &
     &      om_v(*), on_u(*)
          real(c_double), intent(inout) :: Huon(*), Hvom(*)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_omega_tile (ctx, ng,        &
     &    tile, model, LBi, UBi, LBj, UBj, IminS, ImaxS, JminS,       &
     &    JmaxS, Huon, Hvom, z_w, W)                                  &
     &                          BIND(C, name='roms_b200_omega_tile')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: ng, tile, model, LBi, UBi,         &
     &      LBj, UBj, IminS, ImaxS, JminS, JmaxS
          real(c_double), intent(in) :: Huon(*), Hvom(*), z_w(*)
          real(c_double), intent(inout) :: W(*)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_set_zeta_tile (ctx, ng,     &
     &    tile, LBi, UBi, LBj, UBj, IminS, ImaxS, JminS, JmaxS,       &
     &    Zt_avg1, zeta)                                              &
     &                          BIND(C, name='roms_b200_set_zeta_tile')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: ng, tile, LBi, UBi, LBj, UBj,      &
     &      IminS, ImaxS, JminS, JmaxS
          real(c_double), intent(in) :: Zt_avg1(*)
          real(c_double), intent(inout) :: zeta(*)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_set_depth_tile (ctx,        &
     &    ng, tile, model, LBi, UBi, LBj, UBj, IminS, ImaxS,          &
     &    JminS, JmaxS, nstp, nnew, h, Zt_avg1, Hz, z_r, z_w)         &
     &                          BIND(C, name='roms_b200_set_depth_tile')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: ng, tile, model, LBi, UBi,         &
     &      LBj, UBj, IminS, ImaxS, JminS, JmaxS, nstp, nnew
          real(c_double), intent(in) :: h(*), Zt_avg1(*)
          real(c_double), intent(inout) :: Hz(*), z_r(*), z_w(*)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_pre_step3d_tile (ctx,       &
     &    ng, tile, LBi, UBi, LBj, UBj, IminS, ImaxS, JminS,          &
     &    JmaxS, nrhs, nstp, nnew, pm, pn, Hz, Huon, Hvom, z_r,       &
     &    z_w, btflx, bustr, bvstr, stflx, sustr, svstr, srflx,       &
     &    Akt, Akv, ghats, W, ru, rv, t, u, v)                        &
     &                          BIND(C, name='roms_b200_pre_step3d_tile')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: ng, tile, LBi, UBi, LBj, UBj,      &
     &      IminS, ImaxS, JminS, JmaxS, nrhs, nstp, nnew
          real(c_double), intent(in) :: pm(*), pn(*), Hz(*),          &
     &      Huon(*), Hvom(*), z_r(*), z_w(*), btflx(*), bustr(*),     &
     &      bvstr(*), stflx(*), sustr(*), svstr(*), Akt(*),           &
     &      Akv(*), W(*), ru(*), rv(*)
          real(c_double), intent(in), optional :: srflx(*),           &
     &      ghats(*)
          real(c_double), intent(inout) :: t(*), u(*), v(*)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_prsgrd32_tile (ctx, ng,     &
     &    tile, LBi, UBi, LBj, UBj, IminS, ImaxS, JminS, JmaxS,       &
     &    nrhs, om_v, on_u, Hz, z_r, z_w, rho, ru, rv)                &
     &                          BIND(C, name='roms_b200_prsgrd32_tile')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: ng, tile, LBi, UBi, LBj, UBj,      &
     &      IminS, ImaxS, JminS, JmaxS, nrhs
          real(c_double), intent(in) :: om_v(*), on_u(*), Hz(*),      &
     &      z_r(*), z_w(*), rho(*)
          real(c_double), intent(inout) :: ru(*), rv(*)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_rhs3d_tile_tile (ctx,       &
     &    ng, tile, LBi, UBi, LBj, UBj, IminS, ImaxS, JminS,          &
     &    JmaxS, nrhs, Hz, Huon, Hvom, dmde, dndx, fomn, om_u,        &
     &    om_v, on_u, on_v, pm, pn, bustr, bvstr, sustr, svstr,       &
     &    u, v, W, rufrc, rvfrc, ru, rv)                              &
     &                          BIND(C, name='roms_b200_rhs3d_tile_tile')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: ng, tile, LBi, UBi, LBj, UBj,      &
     &      IminS, ImaxS, JminS, JmaxS, nrhs
          real(c_double), intent(in) :: Hz(*), Huon(*), Hvom(*),      &
     &      fomn(*), om_u(*), om_v(*), on_u(*), on_v(*), pm(*),       &
     &      pn(*), bustr(*), bvstr(*), sustr(*), svstr(*), u(*),      &
     &      v(*), W(*)
          real(c_double), intent(in), optional :: dmde(*), dndx(*)
          real(c_double), intent(inout) :: rufrc(*), rvfrc(*),        &
     &      ru(*), rv(*)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_step2d_tile (ctx, ng,       &
     &    tile, LBi, UBi, LBj, UBj, UBk, IminS, ImaxS, JminS,         &
     &    JmaxS, krhs, kstp, knew, nstp, nnew, fomn, h, om_u,         &
     &    om_v, on_u, on_v, omn, pm, pn, dndx, dmde, pmon_r,          &
     &    pnom_r, pmon_p, pnom_p, om_r, on_r, om_p, on_p,             &
     &    visc2_p, visc2_r, rhoA, rhoS, DU_avg1, DU_avg2,             &
     &    DV_avg1, DV_avg2, Zt_avg1, rufrc, rvfrc, ru, rv, rubar,     &
     &    rvbar, rzeta, ubar, vbar, zeta)                             &
     &                          BIND(C, name='roms_b200_step2d_tile')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: ng, tile, LBi, UBi, LBj, UBj,      &
     &      UBk, IminS, ImaxS, JminS, JmaxS, krhs, kstp, knew,        &
     &      nstp, nnew
          real(c_double), intent(in) :: fomn(*), h(*), om_u(*),       &
     &      om_v(*), on_u(*), on_v(*), omn(*), pm(*), pn(*),          &
     &      pmon_r(*), pnom_r(*), pmon_p(*), pnom_p(*), om_r(*),      &
     &      on_r(*), om_p(*), on_p(*), visc2_p(*), visc2_r(*)
          real(c_double), intent(in), optional :: dndx(*),            &
     &      dmde(*), rhoA(*), rhoS(*)
          real(c_double), intent(inout) :: DU_avg1(*),                &
     &      DU_avg2(*), DV_avg1(*), DV_avg2(*), Zt_avg1(*),           &
     &      rufrc(*), rvfrc(*), ru(*), rv(*), rubar(*), rvbar(*),     &
     &      rzeta(*), ubar(*), vbar(*), zeta(*)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_step3d_uv_tile (ctx,        &
     &    ng, tile, LBi, UBi, LBj, UBj, IminS, ImaxS, JminS,          &
     &    JmaxS, nrhs, nstp, nnew, om_v, on_u, pm, pn, Hz, z_r,       &
     &    z_w, Akv, DU_avg1, DV_avg1, DU_avg2, DV_avg2, ru, rv,       &
     &    u, v, ubar, vbar, Huon, Hvom)                               &
     &                          BIND(C, name='roms_b200_step3d_uv_tile')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: ng, tile, LBi, UBi, LBj, UBj,      &
     &      IminS, ImaxS, JminS, JmaxS, nrhs, nstp, nnew
          real(c_double), intent(in) :: om_v(*), on_u(*), pm(*),      &
     &      pn(*), Hz(*), z_r(*), z_w(*), Akv(*), DU_avg1(*),         &
     &      DV_avg1(*), DU_avg2(*), DV_avg2(*)
          real(c_double), intent(inout) :: ru(*), rv(*), u(*),        &
     &      v(*), ubar(*), vbar(*), Huon(*), Hvom(*)
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_step3d_t_tile (ctx, ng,     &
     &    tile, LBi, UBi, LBj, UBj, IminS, ImaxS, JminS, JmaxS,       &
     &    nrhs, nstp, nnew, omn, om_u, om_v, on_u, on_v, pm, pn,      &
     &    Hz, Huon, Hvom, z_r, Akt, W, t)                             &
     &                          BIND(C, name='roms_b200_step3d_t_tile')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int), value :: ng, tile, LBi, UBi, LBj, UBj,      &
     &      IminS, ImaxS, JminS, JmaxS, nrhs, nstp, nnew
          real(c_double), intent(in) :: omn(*), om_u(*), om_v(*),     &
     &      on_u(*), on_v(*), pm(*), pn(*), Hz(*), Huon(*),           &
     &      Hvom(*), z_r(*), Akt(*), W(*)
          real(c_double), intent(inout) :: t(*)
        END FUNCTION
!
!  PERFECT_RESTART state list (wrt_rst.F) and the restart epilogue.
!
        integer(c_int) FUNCTION roms_b200_restart_fields (ctx,        &
     &    ids, cap)                                                   &
     &                          BIND(C, name='roms_b200_restart_fields')
          IMPORT
          type(c_ptr), value :: ctx
          integer(c_int) :: ids(*)
          integer(c_int), value :: cap
        END FUNCTION
        integer(c_int) FUNCTION roms_b200_restart_finish (ctx)        &
     &                          BIND(C, name='roms_b200_restart_finish')
          IMPORT
          type(c_ptr), value :: ctx
        END FUNCTION
      END INTERFACE

      CONTAINS
!
!  Map a non-zero return code to the reference's error convention
!  (exit_flag=8, "fatal algorithm result", mod_scalars.F:548-561).
!
      SUBROUTINE b200_check (rc, line, file)
      USE mod_scalars, ONLY : exit_flag
      integer(c_int), intent(in) :: rc
      integer, intent(in) :: line
      character (len=*), intent(in) :: file
      IF (rc.ne.0) THEN
        exit_flag=8
        PRINT *, 'roms_b200 kernel failed, rc = ', rc, ' at ', file, line
      END IF
      END SUBROUTINE b200_check

      END MODULE roms_b200_mod
