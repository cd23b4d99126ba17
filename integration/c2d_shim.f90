! integration/c2d_shim.f90 -- replaces imcfield2d.o imcvol2d_para.o imcsurf2d_para.o imcredist.o update2d.o
!                             (subroutine update only) and the reductions of xec2d.f:325-394
module c2d_abi
  use iso_c_binding
  implicit none
  type, bind(C) :: c2d_dims
     integer(c_int) :: nz, nr, n_vol, num_nt, nphfield, n_gg, nphtotal, nph_lc, nmu
     integer(c_int) :: pair_switch, cr_sent, spec_switch, split1, split2, split3, spl3_trg
     integer(c_int64_t) :: census_capacity, event_capacity, collision_capacity
     integer(c_int32_t) :: rseed                      ! uint32_t
     integer(c_int) :: sigma_mode
     real(c_double) :: rmin, zmin
     type(c_ptr) :: z, r, E_ph, E_field, E_gg, gnt, hu, Elcmin, Elcmax, mu
     integer(c_int) :: n_ref                          ! COMMON /reflection/: c_loc(E_ref), c_loc(P_ref), c_loc(W_abs) as they lie
     type(c_ptr) :: E_ref, P_ref, W_abs
  end type
  type, bind(C) :: c2d_step_in
     integer(c_int) :: ncycle
     real(c_double) :: time, dt
     type(c_ptr) :: kappa_tot, eps_tot, eps_th, f_nt, Pnt, n_e, tea, f_pair, k_gg
     type(c_ptr) :: nsv
     type(c_ptr) :: ewsv, Eloss_th, Eloss_tot, zsurf, tbbl
  end type
  type, bind(C) :: c2d_tallies
     type(c_ptr) :: edep, prdep, ecens, npcen, n_field, n_ph, E_IC, nelectron, fout, edout
     type(c_ptr) :: erlki, erlko, erlku, erlkl, Ed_in, Ed_ref
     integer(c_int64_t) :: counters(16)
  end type
  ! c2d_fp_in / c2d_fp_out: field for field as in include/compton2d_b200.h
  type(c_ptr), save :: ctx = c_null_ptr
  interface
     integer(c_int) function c2d_create(ctx, device) bind(C, name='c2d_create')
       import; type(c_ptr) :: ctx; integer(c_int), value :: device
     end function
     integer(c_int) function c2d_init(ctx, dims) bind(C, name='c2d_init')
       import; type(c_ptr), value :: ctx; type(c2d_dims) :: dims
     end function
     integer(c_int) function c2d_step_begin(ctx, s) bind(C, name='c2d_step_begin')
       import; type(c_ptr), value :: ctx; type(c2d_step_in) :: s
     end function
     integer(c_int) function c2d_imcfield(ctx) bind(C, name='c2d_imcfield')
       import; type(c_ptr), value :: ctx
     end function
     integer(c_int) function c2d_imcvol(ctx, first, count, stride) bind(C, name='c2d_imcvol')
       import; type(c_ptr), value :: ctx; integer(c_int64_t), value :: first, count, stride
     end function
     integer(c_int) function c2d_sync(ctx) bind(C, name='c2d_sync')
       import; type(c_ptr), value :: ctx
     end function
     integer(c_int) function c2d_tally_read(ctx, t) bind(C, name='c2d_tally_read')
       import; type(c_ptr), value :: ctx; type(c2d_tallies) :: t
     end function
     integer(c_int) function c2d_census_export(ctx, dbuf, ibuf, sid, cap, n) bind(C, name='c2d_census_export')
       import; type(c_ptr), value :: ctx, dbuf, ibuf, sid; integer(c_int64_t), value :: cap; integer(c_int64_t) :: n
     end function
     integer(c_int) function c2d_events_drain(ctx, buf, cap, n) bind(C, name='c2d_events_drain')
       import; type(c_ptr), value :: ctx, buf; integer(c_int64_t), value :: cap; integer(c_int64_t) :: n
     end function
     integer(c_int) function c2d_events_write(ctx, path, append, n) bind(C, name='c2d_events_write')
       import; type(c_ptr), value :: ctx; character(kind=c_char) :: path(*); integer(c_int), value :: append
       integer(c_int64_t) :: n
     end function
     integer(c_int) function c2d_census_write(ctx, path, n) bind(C, name='c2d_census_write')
       import; type(c_ptr), value :: ctx; character(kind=c_char) :: path(*); integer(c_int64_t) :: n
     end function
     integer(c_int) function c2d_census_read(ctx, path, entries) bind(C, name='c2d_census_read')
       import; type(c_ptr), value :: ctx; character(kind=c_char) :: path(*); integer(c_int64_t), value :: entries
     end function
     integer(c_int) function c2d_lc_accumulate(ctx) bind(C, name='c2d_lc_accumulate')
       import; type(c_ptr), value :: ctx
     end function
     type(c_ptr) function c2d_last_error() bind(C, name='c2d_last_error')
       import
     end function
     subroutine c2d_pack_bin_first(common, nbin_pad, jmax, kmax, nbin, nz, nr, dense) bind(C, name='c2d_pack_bin_first')
       import; type(c_ptr), value :: common, dense; integer(c_int), value :: nbin_pad, jmax, kmax, nbin, nz, nr
     end subroutine
     subroutine c2d_pack_bin_last(common, jmax, kmax, nbin, nz, nr, dense) bind(C, name='c2d_pack_bin_last')
       import; type(c_ptr), value :: common, dense; integer(c_int), value :: jmax, kmax, nbin, nz, nr
     end subroutine
  end interface
contains
  subroutine c2d_check(rc)                       ! the reference's error behaviour: print and STOP
    integer(c_int), intent(in) :: rc
    character(kind=c_char), pointer :: msg(:)
    if (rc /= 0) then
       call c_f_pointer(c2d_last_error(), msg, [512])
       write(*,*) 'compton2d_b200 error ', rc, ': ', msg(1:index(transfer(msg, repeat(' ',512)), c_null_char) - 1)
       stop
    end if
  end subroutine
end module

subroutine imcfield2d                              ! same name as imcfield2d.f:1 -> xec2d.f is untouched
  use c2d_abi
  implicit none
  include 'general.pa'
  include 'commonblock.f'
  ! first call of a step: hand the step's tables to the device (what vol_bcast / z_surf_bcast shipped)
  call c2d_begin_step_from_common()
  call c2d_check(c2d_imcfield(ctx))
end subroutine

subroutine imcvol2d                                ! imcvol2d_para.f:1
  use c2d_abi
  implicit none
  include 'general.pa'
  include 'commonblock.f'
  ! rank myid tracks the volume histories h = myid, myid+numprocs, ... (global index = Philox stream)
  call c2d_check(c2d_imcvol(ctx, int(myid, c_int64_t), -1_c_int64_t, int(numprocs, c_int64_t)))
end subroutine

subroutine imcredist                               ! imcredist.f: nothing to do, histories are sharded statically
end subroutine

subroutine graphics_collect                        ! xec2d.f:371; also does xec_add / cens_add_up / E_add_up
  use c2d_abi
  implicit none
  include 'general.pa'
  include 'commonblock.f'
  type(c2d_tallies) :: t
  ! dense host copies, then scatter into COMMON: edep(j,k) = dense_edep(k + nr*(j-1)), n_field(:,j,k) via
  ! c2d_unpack_bin_first(dense, nphfield, jmax, kmax, nphfield, nz, nr, n_field), fout(1:nmu,1:nphtotal) ...
  call c2d_check(c2d_sync(ctx))
  ! escaping photons: either the reference's event file, byte for byte (imcleak2d.f:181) ...
  !   call c2d_check(c2d_events_write(ctx, 'p001_evb.dat'//c_null_char, 1_c_int, nev))
  ! ... or binned on the device with plcm's / pspt's rules (c2d_lc_begin was called after setup)
  call c2d_check(c2d_lc_accumulate(ctx))
  ! with more than one rank: MPI_Allreduce on the two device buffers of c2d_tally_device (CUDA-aware MPI / NCCL)
  call c2d_check(c2d_tally_read(ctx, t))
end subroutine

subroutine write_cens(entries, nunit_c)            ! census2d.f:1 -- called by write_record.f:435 for the worker's census
  use c2d_abi
  implicit none
  integer entries, nunit_c
  integer(c_int64_t) :: n
  character*35 localrecord, recordfilename
  include 'general.pa'
  include 'commonblock.f'
  close(nunit_c)                                   ! the library writes the same text (6e14.7 / 6i5 per photon)
  localrecord = recordfilename('census.dat', myid)
  call c2d_check(c2d_census_write(ctx, trim(localrecord)//c_null_char, n))
  open(unit=nunit_c, file=localrecord, position='append')
end subroutine

subroutine read_cens(entries, nunit_c)             ! census2d.f:41 -- called by read_record.f
  use c2d_abi
  implicit none
  integer entries, nunit_c
  character*35 localrecord, recordfilename
  include 'general.pa'
  include 'commonblock.f'
  localrecord = recordfilename('census.dat', myid)
  call c2d_check(c2d_census_read(ctx, trim(localrecord)//c_null_char, int(entries, c_int64_t)))
end subroutine
