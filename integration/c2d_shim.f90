! integration/c2d_shim.f90 -- the reference-side binding of libcompton2d_b200.so.
!
! Replaces, by NAME, the object files of the hot path in the reference's link line (Makefile:9-38):
!     imcfield2d.o  imcvol2d_para.o  imcsurf2d_para.o  imcredist.o  census2d.o
!     update2d.o (subroutine update; FP_calc & co. are no longer called)   and, in xec2d.f, xec_add / graphics_collect
! xec (xec2d.f:100-110 / 141-193), imcgen2d, reader, setup, write_record, read_record, graphics stay untouched.
! Run with one MPI rank per GPU; every rank executes the master branch of imcgen2d (set master = myid, or run
! numprocs = 1 per GPU and let c2d_comm_init connect the ranks).
!
! The shim does NO arithmetic.  Every subroutine hands the ADDRESSES of the COMMON members (c_loc) to the library in a
! c2d_common record; packing into the dense [zone][bin] layout, the device calls and the scatter back into COMMON happen
! in csrc/c2d_common.cu, which tests/test_common_bridge.py drives with fake COMMON blocks at the same padded extents
! (general.pa:7-31) -- this image has no Fortran compiler, so this file itself has not been compiled here.
module c2d_abi
  use iso_c_binding
  implicit none
  type, bind(C) :: c2d_dims                       ! include/compton2d_b200.h: c2d_dims
     integer(c_int) :: nz, nr, n_vol, num_nt, nphfield, n_gg, nphtotal, nph_lc, nmu
     integer(c_int) :: pair_switch, cr_sent, spec_switch, split1, split2, split3, spl3_trg
     integer(c_int64_t) :: census_capacity, event_capacity, collision_capacity
     integer(c_int32_t) :: rseed                  ! uint32_t
     integer(c_int) :: sigma_mode
     real(c_double) :: rmin, zmin
     type(c_ptr) :: z, r, E_ph, E_field, E_gg, gnt, hu, Elcmin, Elcmax, mu
     integer(c_int) :: n_ref
     type(c_ptr) :: E_ref, P_ref, W_abs
  end type
  type, bind(C) :: c2d_common                     ! include/compton2d_b200.h: c2d_common, member for member
     integer(c_int) :: jmax, kmax, ntmax, nmumax, nphomax, nphlcmax, ti
     type(c_ptr) :: kappa_tot, eps_tot, eps_th, f_nt, Pnt, n_e, tea, tna, B_field, f_pair, k_gg
     type(c_ptr) :: nsv
     type(c_ptr) :: ewsv, Eloss_th, Eloss_tot, zsurf, tbbl, edep, prdep, ecens
     type(c_ptr) :: npcen
     type(c_ptr) :: n_field, n_ph, E_IC
     type(c_ptr) :: nelectron
     type(c_ptr) :: fout, edout, erlki, erlko, erlku, erlkl, Ed_in, Ed_ref
     type(c_ptr) :: ndxout, cmcount1, cmcount2, lkcount
     type(c_ptr) :: vol, Eloss_sy, ec_old, turb_lev, gmin, gmax, p_nth, amxwl, Te_new
     type(c_ptr) :: dT_max, E_tot_old, E_tot_new, hr_total, hr_st_total, dr, dz, f_t_implicit
  end type
  type, bind(C) :: c2d_fp_in                      ! include/compton2d_b200.h: c2d_fp_in (pointer members unused by the bridge)
     integer(c_int) :: first_zone, n_zones, zone_stride, inj_switch, inj_dis, g2var_switch, pick_sw, cf_sentinel, ncycle
     real(c_double) :: time, dt, lnL, inj_g1, inj_g2, inj_p, inj_t, inj_gg, inj_sigma, inj_L, inj_v, pick_rate
     real(c_double) :: r_esc, r_acc, r_flare, z_flare, t_flare, sigma_r, sigma_z, sigma_t, flare_amp
     type(c_ptr) :: f_nt, Pnt, n_field, ecens, n_e, B_field, tea, tna, vol, f_pair, Eloss_sy, ec_old, turb_lev, F_IC
  end type
  type(c_ptr), save :: ctx = c_null_ptr
  logical, save :: step_begun = .false.
  interface
     integer(c_int) function c2d_create(ctx, device) bind(C, name='c2d_create')
       import; type(c_ptr) :: ctx; integer(c_int), value :: device
     end function
     integer(c_int) function c2d_init(ctx, dims) bind(C, name='c2d_init')
       import; type(c_ptr), value :: ctx; type(c2d_dims) :: dims
     end function
     integer(c_int) function c2d_comm_unique_id(id) bind(C, name='c2d_comm_unique_id')
       import; character(kind=c_char) :: id(128)
     end function
     integer(c_int) function c2d_comm_init(ctx, id, rank, world) bind(C, name='c2d_comm_init')
       import; type(c_ptr), value :: ctx; character(kind=c_char) :: id(128); integer(c_int), value :: rank, world
     end function
     integer(c_int) function c2d_common_step_begin(ctx, cm, ncycle, time, dt) bind(C, name='c2d_common_step_begin')
       import; type(c_ptr), value :: ctx; type(c2d_common) :: cm; integer(c_int), value :: ncycle; real(c_double), value :: time, dt
     end function
     integer(c_int) function c2d_common_collect(ctx, cm) bind(C, name='c2d_common_collect')
       import; type(c_ptr), value :: ctx; type(c2d_common) :: cm
     end function
     integer(c_int) function c2d_common_update(ctx, cm, scalars) bind(C, name='c2d_common_update')
       import; type(c_ptr), value :: ctx; type(c2d_common) :: cm; type(c2d_fp_in) :: scalars
     end function
     integer(c_int) function c2d_imcfield(ctx) bind(C, name='c2d_imcfield')
       import; type(c_ptr), value :: ctx
     end function
     integer(c_int) function c2d_imcvol(ctx, first, count, stride) bind(C, name='c2d_imcvol')
       import; type(c_ptr), value :: ctx; integer(c_int64_t), value :: first, count, stride
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
     type(c_ptr) function c2d_last_error() bind(C, name='c2d_last_error')
       import
     end function
  end interface
contains
  subroutine c2d_check(rc)                       ! the reference's error behaviour: print and STOP (imctrk2d.f:583-584)
    integer(c_int), intent(in) :: rc
    character(kind=c_char), pointer :: msg(:)
    integer :: n
    if (rc /= 0) then
       call c_f_pointer(c2d_last_error(), msg, [512])
       n = 1
       do while (n < 512 .and. msg(n) /= c_null_char)
          n = n + 1
       end do
       write(*,*) 'compton2d_b200 error ', rc, ': ', msg(1:n-1)
       stop
    end if
  end subroutine
end module

! ---- COMMON -> c2d_common: addresses only -------------------------------------------------------------------------
subroutine c2d_common_view(cm)
  use c2d_abi
  implicit none
  include 'general.pa'
  include 'commonblock.f'
  type(c2d_common), intent(out) :: cm
  cm%jmax = jmax; cm%kmax = kmax; cm%ntmax = ntmax; cm%nmumax = nmumax; cm%nphomax = nphomax; cm%nphlcmax = nphlcmax
  cm%ti = max(ti, 1)
  cm%kappa_tot = c_loc(kappa_tot); cm%eps_tot = c_loc(eps_tot); cm%eps_th = c_loc(eps_th)
  cm%f_nt = c_loc(f_nt); cm%Pnt = c_loc(Pnt); cm%n_e = c_loc(n_e); cm%tea = c_loc(tea); cm%tna = c_loc(tna)
  cm%B_field = c_loc(B_field); cm%f_pair = c_loc(f_pair); cm%k_gg = c_loc(k_gg); cm%nsv = c_loc(nsv)
  cm%ewsv = c_loc(ewsv); cm%Eloss_th = c_loc(Eloss_th); cm%Eloss_tot = c_loc(Eloss_tot); cm%zsurf = c_loc(zsurf)
  cm%tbbl = c_loc(tbbl)
  cm%edep = c_loc(edep); cm%prdep = c_loc(prdep); cm%ecens = c_loc(ecens); cm%npcen = c_loc(npcen)
  cm%n_field = c_loc(n_field); cm%n_ph = c_loc(n_ph); cm%E_IC = c_loc(E_IC); cm%nelectron = c_loc(nelectron)
  cm%fout = c_loc(fout); cm%edout = c_loc(edout); cm%erlki = c_loc(erlki); cm%erlko = c_loc(erlko)
  cm%erlku = c_loc(erlku); cm%erlkl = c_loc(erlkl); cm%Ed_in = c_loc(Ed_in); cm%Ed_ref = c_loc(Ed_ref)
  cm%ndxout = c_loc(ndxout); cm%cmcount1 = c_loc(cmcount1); cm%cmcount2 = c_loc(cmcount2); cm%lkcount = c_loc(lkcount)
  cm%vol = c_loc(vol); cm%Eloss_sy = c_loc(Eloss_sy); cm%ec_old = c_loc(ec_old); cm%turb_lev = c_loc(turb_lev)
  cm%gmin = c_loc(gmin); cm%gmax = c_loc(gmax); cm%p_nth = c_loc(p_nth); cm%amxwl = c_loc(amxwl); cm%Te_new = c_loc(Te_new)
  cm%dT_max = c_loc(dT_max); cm%E_tot_old = c_loc(E_tot_old); cm%E_tot_new = c_loc(E_tot_new)
  cm%hr_total = c_loc(hr_total); cm%hr_st_total = c_loc(hr_st_total); cm%dr = c_loc(dr); cm%dz = c_loc(dz)
  cm%f_t_implicit = c_loc(f_t_implicit)
end subroutine

! ---- once, at the first tracking call: what setup_bcast (fp_mpi.f:12-319) shipped to the workers -----------------------
subroutine c2d_shim_init
  use c2d_abi
  implicit none
  include 'mpif.h'
  include 'general.pa'
  include 'commonblock.f'
  type(c2d_dims) :: d
  character(kind=c_char) :: id(128)
  integer :: ngpu_per_node
  if (c_associated(ctx)) return
  ngpu_per_node = 8
  call c2d_check(c2d_create(ctx, int(mod(myid, ngpu_per_node), c_int)))
  d%nz = nz; d%nr = nr; d%n_vol = n_vol; d%num_nt = num_nt; d%nphfield = nphfield; d%n_gg = n_gg
  d%nphtotal = nphtotal; d%nph_lc = nph_lc; d%nmu = nmu
  d%pair_switch = pair_switch; d%cr_sent = cr_sent; d%spec_switch = spec_switch
  d%split1 = split1; d%split2 = split2; d%split3 = split3; d%spl3_trg = spl3_trg
  d%census_capacity = ucens; d%event_capacity = ucens; d%collision_capacity = 262144
  d%rseed = rseed; d%sigma_mode = 1
  d%rmin = rmin; d%zmin = zmin
  d%z = c_loc(z); d%r = c_loc(r); d%E_ph = c_loc(E_ph); d%E_field = c_loc(E_field); d%E_gg = c_loc(E_gg)
  d%gnt = c_loc(gnt); d%hu = c_loc(hu); d%Elcmin = c_loc(Elcmin); d%Elcmax = c_loc(Elcmax); d%mu = c_loc(mu)
  if (cr_sent /= 0) then
     d%n_ref = n_ref; d%E_ref = c_loc(E_ref); d%P_ref = c_loc(P_ref); d%W_abs = c_loc(W_abs)
  else
     d%n_ref = 0; d%E_ref = c_null_ptr; d%P_ref = c_null_ptr; d%W_abs = c_null_ptr
  end if
  call c2d_check(c2d_init(ctx, d))
  if (numprocs > 1) then                          ! NCCL inside the library; its id travels once over the host's MPI
     if (myid == 0) call c2d_check(c2d_comm_unique_id(id))
     call MPI_BCAST(id, 128, MPI_CHARACTER, 0, MPI_COMM_WORLD, ierr)
     call c2d_check(c2d_comm_init(ctx, id, int(myid, c_int), int(numprocs, c_int)))
  end if
end subroutine

! ---- the subroutines xec calls, under their own names --------------------------------------------------------------
subroutine imcfield2d                             ! imcfield2d.f:1 -- field_calc on the resident bank
  use c2d_abi
  implicit none
  include 'general.pa'
  include 'commonblock.f'
  type(c2d_common) :: cm
  call c2d_shim_init
  ! the first tracking call of the cycle hands over what imcgen2d left in COMMON (vol_bcast, z_surf_bcast, the job messages)
  call c2d_common_view(cm)
  call c2d_check(c2d_common_step_begin(ctx, cm, int(ncycle, c_int), time, dt(1)))
  step_begun = .true.
  call c2d_check(c2d_imcfield(ctx))
end subroutine

subroutine imcvol2d                               ! imcvol2d_para.f:1
  use c2d_abi
  implicit none
  include 'general.pa'
  include 'commonblock.f'
  ! rank myid tracks the volume histories h = myid, myid + numprocs, ... (global index = Philox stream)
  call c2d_check(c2d_imcvol(ctx, int(myid, c_int64_t), -1_c_int64_t, int(numprocs, c_int64_t)))
end subroutine

subroutine imcsurf2d                              ! imcsurf2d_para.f:1
  ! The shipped input has no surface source (every tbb = -0.d0: nsurf* = 0, imcgen2d.f:424-443).  With sources: fill a
  ! c2d_surf_in from nsurf*, ewsurf*, tbb*(:, ti) and the file_sp tables of /insp/ and call c2d_surf_begin + c2d_imcsurf,
  ! exactly as compton2d_b200/host.py:imcsurf2d does.
end subroutine

subroutine imcredist                              ! imcredist.f: histories are sharded statically, nothing moves
end subroutine

subroutine xec_add                                ! xec2d.f:325 -- its sums were delivered by the last graphics_collect
end subroutine

subroutine update                                 ! update2d.f:7 (called when T_const = 0 and ncycle > 0, xec2d.f:87)
  use c2d_abi
  implicit none
  include 'general.pa'
  include 'commonblock.f'
  type(c2d_common) :: cm
  type(c2d_fp_in) :: s
  lnL = 20.d0                                     ! update2d.f:139
  call c2d_common_view(cm)
  ! cens_add_up: the FP update needs the REDUCED edep / ecens / n_field -- collect first (the library reduces once per step)
  call c2d_check(c2d_common_collect(ctx, cm))
  s%first_zone = 0; s%n_zones = -1; s%zone_stride = 1
  s%inj_switch = inj_switch; s%inj_dis = inj_dis; s%g2var_switch = g2var_switch; s%pick_sw = pick_sw
  s%cf_sentinel = cf_sentinel; s%ncycle = ncycle
  s%time = time; s%dt = dt(1); s%lnL = lnL
  s%inj_g1 = inj_g1; s%inj_g2 = inj_g2; s%inj_p = inj_p; s%inj_t = inj_t; s%inj_gg = inj_gg; s%inj_sigma = inj_sigma
  s%inj_L = inj_L; s%inj_v = inj_v; s%pick_rate = pick_rate; s%r_esc = r_esc; s%r_acc = r_acc
  s%r_flare = r_flare; s%z_flare = z_flare; s%t_flare = t_flare; s%sigma_r = sigma_r; s%sigma_z = sigma_z
  s%sigma_t = sigma_t; s%flare_amp = flare_amp
  s%f_nt = c_null_ptr; s%Pnt = c_null_ptr; s%n_field = c_null_ptr; s%ecens = c_null_ptr; s%n_e = c_null_ptr
  s%B_field = c_null_ptr; s%tea = c_null_ptr; s%tna = c_null_ptr; s%vol = c_null_ptr; s%f_pair = c_null_ptr
  s%Eloss_sy = c_null_ptr; s%ec_old = c_null_ptr; s%turb_lev = c_null_ptr; s%F_IC = c_null_ptr
  call c2d_check(c2d_common_update(ctx, cm, s))
  if (ncycle <= 1) dT_max = dmax1(dT_max, df_T)   ! photon_fill leaves dT_max = df_T (update2d.f:1917)
  ! the time-step controller of update2d.f:233-243 only logs dt_new (l.246-261): kept in the host's update if wanted
end subroutine

subroutine graphics_collect                       ! xec2d.f:371; also stands in for cens_add_up / E_add_up / xec_add
  use c2d_abi
  implicit none
  include 'general.pa'
  include 'commonblock.f'
  type(c2d_common) :: cm
  integer(c_int64_t) :: nev
  character*35 localeventfile, evlfilename
  call c2d_common_view(cm)
  call c2d_check(c2d_common_collect(ctx, cm))     ! one all-reduce per packed buffer inside the library, then COMMON
  ! escaping photons: the reference's event file, byte for byte (imcleak2d.f:181), appended per cycle
  localeventfile = evlfilename(eventfile, myid)
  call c2d_check(c2d_events_write(ctx, trim(localeventfile)//c_null_char, 1_c_int, nev))
  step_begun = .false.
end subroutine

subroutine write_cens(entries, nunit_c)           ! census2d.f:1 -- called by write_record.f:435 for this rank's census
  use c2d_abi
  implicit none
  integer entries, nunit_c
  integer(c_int64_t) :: n
  character*35 localrecord, recordfilename
  include 'general.pa'
  include 'commonblock.f'
  close(nunit_c)                                  ! the library writes the same text (6e14.7 / 6i5 per photon)
  localrecord = recordfilename('census.dat', myid)
  call c2d_check(c2d_census_write(ctx, trim(localrecord)//c_null_char, n))
  open(unit=nunit_c, file=localrecord, position='append')
end subroutine

subroutine read_cens(entries, nunit_c)            ! census2d.f:41 -- called by read_record.f
  use c2d_abi
  implicit none
  integer entries, nunit_c
  character*35 localrecord, recordfilename
  include 'general.pa'
  include 'commonblock.f'
  call c2d_shim_init
  localrecord = recordfilename('census.dat', myid)
  call c2d_check(c2d_census_read(ctx, trim(localrecord)//c_null_char, int(entries, c_int64_t)))
end subroutine
