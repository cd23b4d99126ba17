"""Host-side mirror of the reference's per-step interface, on top of the C ABI.

The reference's "API" is the sequence of argument-less subroutine calls in `xec`
(src_20121026/xec2d.f:141-193) operating on COMMON blocks.  `Compton2d` keeps the same
names and order -- imcgen2d -> imcfield2d -> imcvol2d -> imcsurf2d -> imcredist -> update ->
graphics_collect -- with the COMMON members passed as dense numpy arrays
([zone][bin], zone = (k-1) + nr*(j-1), fp_mpi.f:333-352).  Every method calls straight into
libcompton2d_b200.so; nothing here computes physics.

Multi-GPU: one process per GPU.  Volume histories are sharded by global history index
(h mod world == rank), census photons stay on the GPU that wrote them, and the per-step
reductions of the reference (cens_add_up, E_add_up, xec_add, graphics_collect) become one
all-reduce of the packed FP64 tally buffer and one of the int64 buffer.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _l

_F = np.float64


def _d(a, n=None, name="array"):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=_F).ravel()
    if n is not None and a.size != n:
        raise ValueError(f"{name}: expected {n} elements, got {a.size}")
    return a


def _p(a):
    return a.ctypes.data_as(_l.dp) if a is not None else None


class _CudaView:
    """__cuda_array_interface__ wrapper so torch can all-reduce the library's device buffer in place."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def pref_calc(n_ref=500, device=0):
    """== Pref_calc (ref_matrix.f:7-89) on the device -> (E_ref [n_ref], P_ref [n_in][n_out], i.e. COMMON /reflection/ P_ref(n_out,
    n_in) as it lies in memory): the Compton-reflection Green's function tables Compton2d(cr_sent=..., E_ref=, P_ref=) takes."""
    L = _l.load()
    h = C.c_void_p()
    _l.check(L.c2d_create(C.byref(h), int(device)))
    try:
        E, P = np.zeros(n_ref), np.zeros((n_ref, n_ref))
        _l.check(L.c2d_pref_calc(h, int(n_ref), E.ctypes.data, P.ctypes.data))
    finally:
        L.c2d_destroy(h)
    return E, P


class Compton2d:
    def __init__(self, *, nz, nr, z, r, rmin=0.0, zmin=0.0, E_ph, E_field, E_gg, gnt, hu, Elcmin=(), Elcmax=(),
                 mu=(1.0,), pair_switch=0, cr_sent=0, spec_switch=0, E_ref=None, P_ref=None, W_abs=None, split=(1000, 1000, 300, 10), rseed=9857,
                 census_capacity=5_000_000, event_capacity=1_000_000, collision_capacity=65536, sigma_mode=0,
                 device=0, rank=0, world=1):
        self.L = _l.load()
        self.rank, self.world = int(rank), int(world)
        self.nz, self.nr, self.nzone = int(nz), int(nr), int(nz) * int(nr)
        self._keep = dict(z=_d(z, nz, "z"), r=_d(r, nr, "r"), E_ph=_d(E_ph), E_field=_d(E_field), E_gg=_d(E_gg),
                          gnt=_d(gnt), hu=_d(hu), Elcmin=_d(Elcmin), Elcmax=_d(Elcmax), mu=_d(mu))
        k = self._keep
        self.n_vol, self.num_nt, self.nphfield, self.n_gg = k["E_ph"].size, k["gnt"].size, k["E_field"].size, k["E_gg"].size
        self.nphtotal, self.nph_lc, self.nmu = k["hu"].size - 1, k["Elcmin"].size, k["mu"].size
        d = _l.Dims(nz=nz, nr=nr, n_vol=self.n_vol, num_nt=self.num_nt, nphfield=self.nphfield, n_gg=self.n_gg,
                    nphtotal=self.nphtotal, nph_lc=self.nph_lc, nmu=self.nmu, pair_switch=pair_switch,
                    cr_sent=cr_sent, spec_switch=spec_switch, split1=split[0], split2=split[1], split3=split[2],
                    spl3_trg=split[3], census_capacity=census_capacity, event_capacity=event_capacity,
                    collision_capacity=collision_capacity, rseed=rseed, sigma_mode=sigma_mode, rmin=rmin, zmin=zmin)
        for name in ("z", "r", "E_ph", "E_field", "E_gg", "gnt", "hu", "Elcmin", "Elcmax", "mu"):
            setattr(d, name, _p(k[name]) if k[name].size else None)
        if E_ref is not None:            # Compton-reflection tables of ref_matrix.f, P_ref / W_abs in COMMON (Fortran) order
            k["E_ref"] = _d(E_ref); k["P_ref"] = _d(P_ref, k["E_ref"].size ** 2, "P_ref"); k["W_abs"] = _d(W_abs, k["E_ref"].size ** 2, "W_abs")
            d.n_ref = k["E_ref"].size
            d.E_ref, d.P_ref, d.W_abs = _p(k["E_ref"]), _p(k["P_ref"]), _p(k["W_abs"])
        self.census_capacity, self.event_capacity = int(census_capacity), int(event_capacity)
        self.h = C.c_void_p()
        _l.check(self.L.c2d_create(C.byref(self.h), int(device)))
        try:
            _l.check(self.L.c2d_init(self.h, C.byref(d)))
        except Exception:
            self.L.c2d_destroy(self.h)
            self.h = None
            raise
        self.ncycle = 0
        self.time = 0.0
        self._comm_ready = self.world == 1
        self._ext_keep = None

    # ------------------------------------------------------------------ ranks
    def comm_init(self, group=None):
        """Join this context to the job's ranks.  With torch.distributed on NCCL the library opens its own NCCL
        communicator (rank 0 makes the id, the 128 bytes travel through torch.distributed once) and every later collective --
        the tally all-reduce, the FP all-gather, the 3rd-split exchange -- runs inside libcompton2d_b200.so on its stream.
        On any other backend (gloo in the CPU tests of the host logic, where "device" memory is host memory) the library
        is handed that transport through c2d_comm_init_ext."""
        if self._comm_ready:
            return
        import torch
        import torch.distributed as dist
        if dist.get_backend(group) == "nccl":
            idb = (C.c_ubyte * _l.COMM_ID_BYTES)()
            if self.rank == 0:
                _l.check(self.L.c2d_comm_unique_id(idb))
            t = torch.tensor(list(bytes(idb)), dtype=torch.uint8, device="cuda")
            dist.broadcast(t, src=0, group=group)
            idb = (C.c_ubyte * _l.COMM_ID_BYTES)(*t.cpu().tolist())
            _l.check(self.L.c2d_comm_init(self.h, idb, self.rank, self.world))
        else:
            world = self.world

            def view(ptr, n, dt):
                return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(dt)), shape=(int(n),))

            def allreduce(user, send, recv, n, kind, stream):
                try:
                    dt = C.c_int64 if kind == 1 else C.c_double
                    t = torch.from_numpy(view(send, n, dt).copy())
                    dist.all_reduce(t, op=dist.ReduceOp.MAX if kind == 2 else dist.ReduceOp.SUM, group=group)
                    view(recv, n, dt)[:] = t.numpy()
                    return 0
                except Exception:                      # never raise through the C frame
                    return 1

            def allgather(user, send, recv, nbytes, stream):
                try:
                    t = torch.from_numpy(view(send, nbytes, C.c_ubyte).copy())
                    out = [torch.empty_like(t) for _ in range(world)]
                    dist.all_gather(out, t, group=group)
                    view(recv, nbytes * world, C.c_ubyte)[:] = torch.cat(out).numpy()
                    return 0
                except Exception:
                    return 1

            vt = _l.CommExt(None, _l.ALLREDUCE_FN(allreduce), _l.ALLGATHER_FN(allgather))
            self._ext_keep = vt                        # the callbacks must outlive the context
            _l.check(self.L.c2d_comm_init_ext(self.h, self.rank, self.world, C.byref(vt)))
        self._comm_ready = True

    def close(self):
        if getattr(self, "h", None):
            self.L.c2d_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ one time step
    def imcgen2d(self, *, ncycle, time, dt, kappa_tot, eps_tot, eps_th, f_nt, Pnt, n_e, tea=None, f_pair=None,
                 k_gg=None, nsv=None, ewsv=None, Eloss_th=None, Eloss_tot=None, zsurf=None, tbbl=None):
        """Hand the step's inputs -- what the reference's imcgen2d leaves in COMMON and broadcasts
        (vol_mpi.f:39-45,100-119; surf_mpi.f:18-98) -- to the device and zero the per-step tallies."""
        nzn = self.nzone
        a = dict(kappa_tot=_d(kappa_tot, nzn * self.n_vol, "kappa_tot"), eps_tot=_d(eps_tot, nzn * self.n_vol, "eps_tot"),
                 eps_th=_d(eps_th, nzn * self.n_vol, "eps_th"), f_nt=_d(f_nt, nzn * self.num_nt, "f_nt"),
                 Pnt=_d(Pnt, nzn * self.num_nt, "Pnt"), n_e=_d(n_e, nzn, "n_e"), tea=_d(tea, nzn, "tea"),
                 f_pair=_d(f_pair, nzn, "f_pair"), k_gg=_d(k_gg, nzn * self.n_gg, "k_gg"), ewsv=_d(ewsv, nzn, "ewsv"),
                 Eloss_th=_d(Eloss_th, nzn, "Eloss_th"), Eloss_tot=_d(Eloss_tot, nzn, "Eloss_tot"),
                 zsurf=_d(zsurf, nzn, "zsurf"), tbbl=_d(tbbl, self.nr, "tbbl"))
        nsv_a = None if nsv is None else np.ascontiguousarray(nsv, dtype=np.int32).ravel()
        s = _l.StepIn(ncycle=int(ncycle), time=float(time), dt=float(dt))
        for k, v in a.items():
            setattr(s, k, _p(v))
        s.nsv = nsv_a.ctypes.data_as(_l.ip) if nsv_a is not None else None
        _l.check(self.L.c2d_step_begin(self.h, C.byref(s)))
        self.ncycle = int(ncycle)
        self.n_volume = int(nsv_a.sum()) if nsv_a is not None else 0

    def imcgen2d_device(self, *, ncycle, time, dt, nst, f_nt, Pnt, n_e, tea, B_field, vol, zsurf, f_pair=None, k_gg=None,
                        tbbl=None):
        """== imcgen2d for the volume sources with volume_em on the device (volume2d.f:10-394, imcgen2d.f:203-459): the
        opacity / emissivity tables are built from the live electron spectrum in HBM and never visit the host.  Returns
        dict(nsv, ewsv, Eloss_sy, Eloss_th, Eloss_tot, Emiss_tot) -- what the reference leaves in COMMON."""
        nzn = self.nzone
        a = dict(f_nt=_d(f_nt, nzn * self.num_nt, "f_nt"), Pnt=_d(Pnt, nzn * self.num_nt, "Pnt"), n_e=_d(n_e, nzn, "n_e"),
                 tea=_d(tea, nzn, "tea"), B_field=_d(B_field, nzn, "B_field"), vol=_d(vol, nzn, "vol"), zsurf=_d(zsurf, nzn, "zsurf"),
                 f_pair=_d(f_pair, nzn, "f_pair"), k_gg=_d(k_gg, nzn * self.n_gg, "k_gg"), tbbl=_d(tbbl, self.nr, "tbbl"))
        gi = _l.GenIn(ncycle=int(ncycle), time=float(time), dt=float(dt), nst=int(nst))
        for k, v in a.items():
            setattr(gi, k, _p(v))
        out = dict(nsv=np.zeros(nzn, np.int32), ewsv=np.zeros(nzn), Eloss_sy=np.zeros(nzn), Eloss_th=np.zeros(nzn),
                   Eloss_tot=np.zeros(nzn))
        go = _l.GenOut()
        go.nsv = out["nsv"].ctypes.data_as(_l.ip)
        for k in ("ewsv", "Eloss_sy", "Eloss_th", "Eloss_tot"):
            setattr(go, k, _p(out[k]))
        _l.check(self.L.c2d_step_begin_gen(self.h, C.byref(gi), C.byref(go)))
        self.ncycle = int(ncycle)
        self.n_volume = int(go.n_volume)
        out["Emiss_tot"] = go.Emiss_tot
        return out

    def step_tables(self):
        """-> dict(kappa_tot, eps_tot, eps_th) [zone][n_vol]: the step tables as the device holds them."""
        t = {k: np.zeros((self.nzone, self.n_vol)) for k in ("kappa_tot", "eps_tot", "eps_th")}
        _l.check(self.L.c2d_step_tables_read(self.h, t["kappa_tot"].ctypes.data, t["eps_tot"].ctypes.data, t["eps_th"].ctypes.data))
        return t

    def imcfield2d(self):
        """== field_calc (imcfield2d.f:57-148): re-track this GPU's census bank."""
        _l.check(self.L.c2d_imcfield(self.h))

    def imcvol2d(self):
        """== sum over zones of vol_calc (imcvol2d_para.f:90-414); this rank's share of the histories."""
        _l.check(self.L.c2d_imcvol(self.h, self.rank, -1, self.world))

    def imcsurf2d(self, *, nsurfu=None, nsurfl=None, nsurfi=None, nsurfo=None, ewsurfu=None, ewsurfl=None, ewsurfi=None,
                  ewsurfo=None, tbbu=None, tbbl=None, tbbi=None, tbbo=None, spectra=(), spec_u=None, spec_l=None,
                  spec_i=None, spec_o=None, star_switch=0):
        """== imcsurf2d: sum over boundary rings of z_surf_calc + r_surf_calc (imcsurf2d_para.f:228-534); this rank's share
        of the histories.  Arguments are the COMMON members the reference's master sends with the ring jobs
        (surf_mpi.f:18-98,136,337): photon counts and weights per ring, the boundary temperatures of the current time
        interval (tbb > 0: Planck spectrum; otherwise the ring samples spectrum table spec_*[ring] of `spectra`, a list of
        dicts from compton2d_b200.spectrum.file_sp).  Without sources (the shipped input: tbb = -0.d0) it does nothing."""
        def iv(a, n):
            return np.zeros(n, np.int32) if a is None else np.ascontiguousarray(a, dtype=np.int32).ravel()

        def dv(a, n):
            return np.zeros(n) if a is None else _d(a, n)
        nz, nr = self.nz, self.nr
        cnt = dict(nsurfi=iv(nsurfi, nz), nsurfo=iv(nsurfo, nz), nsurfu=iv(nsurfu, nr), nsurfl=iv(nsurfl, nr))
        self.n_surface = int(sum(int(v.sum()) for v in cnt.values()))
        if self.n_surface == 0:
            return None
        dbl = dict(ewsurfi=dv(ewsurfi, nz), ewsurfo=dv(ewsurfo, nz), ewsurfu=dv(ewsurfu, nr), ewsurfl=dv(ewsurfl, nr),
                   tbbi=dv(tbbi, nz), tbbo=dv(tbbo, nz), tbbu=dv(tbbu, nr), tbbl=dv(tbbl, nr))
        sp = dict(spec_i=iv(spec_i, nz), spec_o=iv(spec_o, nz), spec_u=iv(spec_u, nr), spec_l=iv(spec_l, nr))
        nfmax = max([len(t["E_file"]) for t in spectra], default=0)
        tab = {k: np.zeros((max(len(spectra), 1), max(nfmax, 1))) for k in ("E_file", "F_file", "P_file", "a1", "I_file")}
        nfile = np.zeros(max(len(spectra), 1), np.int32)
        for q, t in enumerate(spectra):
            for k in tab:
                tab[k][q, : len(t[k])] = t[k]
            nfile[q] = t["nfile"]
        si = _l.SurfIn(star_switch=int(star_switch), n_spec=len(spectra), nfmax=nfmax)
        for k, v in {**cnt, **sp}.items():
            setattr(si, k, v.ctypes.data_as(_l.ip))
        for k, v in {**dbl, **tab}.items():
            setattr(si, k, _p(np.ascontiguousarray(v).ravel()))
        keep = [np.ascontiguousarray(v).ravel() for v in tab.values()]          # pointers above must outlive the call
        for k, v in zip(tab, keep):
            setattr(si, k, _p(v))
        si.nfile = nfile.ctypes.data_as(_l.ip)
        _l.check(self.L.c2d_surf_begin(self.h, C.byref(si)))
        _l.check(self.L.c2d_imcsurf(self.h, self.rank, -1, self.world))

    def imcredist(self):
        """Census rebalancing through the master (imcredist.f) is not needed: histories are sharded
        statically and census photons never move between GPUs."""
        return None

    def sync(self):
        _l.check(self.L.c2d_sync(self.h))

    def graphics_collect(self, group=None):
        """cens_add_up + E_add_up + xec_add + graphics_collect (update2d.f:1931-2083, xec2d.f:325-394): c2d_tally_reduce --
        one all-reduce(sum) per packed tally buffer, in place on the device, inside the library (NCCL on its stream); the
        run-long escaping spectrum fout_total advances by the reduced per-step one."""
        self.comm_init(group)
        _l.check(self.L.c2d_tally_reduce(self.h, None))

    def tally_device_views(self):
        pf, pi = C.c_void_p(), C.c_void_p()
        nf, ni = C.c_int64(), C.c_int64()
        _l.check(self.L.c2d_tally_device(self.h, C.byref(pf), C.byref(nf), C.byref(pi), C.byref(ni)))
        return _CudaView(pf.value, nf.value, "<f8"), _CudaView(pi.value, ni.value, "<i8")

    def comtot(self, zone, xnu, mode=0):
        """== comtot(j,k,xnuc,1,comac,enexc) (comtot2d.f:1-334) with the current step's tables: macroscopic Compton
        opacity [1/cm] for arrays of 0-based zone ids and photon energies [keV]."""
        z = np.ascontiguousarray(zone, dtype=np.int32).ravel()
        x = np.ascontiguousarray(xnu, dtype=_F).ravel()
        out = np.zeros(x.size)
        _l.check(self.L.c2d_comtot(self.h, x.size, z.ctypes.data, x.ctypes.data, int(mode), out.ctypes.data))
        return out

    # ------------------------------------------------------------------ Fokker-Planck
    def IC_loss(self):
        """== IC_loss (icloss2d.f:1-125): build F_IC[num_nt][nphfield] on the device and return it."""
        F = np.zeros((self.num_nt, self.nphfield))
        _l.check(self.L.c2d_ic_loss(self.h, F.ctypes.data))
        return F

    def gamma_bar(self, Theta):
        """== gamma_bar(Theta) (volume2d.f:572-594) -> (quadrature values, predictor-table values; negative beyond the table)."""
        th = np.ascontiguousarray(Theta, dtype=_F).ravel()
        g, gt = np.zeros(th.size), np.zeros(th.size)
        _l.check(self.L.c2d_gamma_bar(self.h, th.size, th.ctypes.data, g.ctypes.data, gt.ctypes.data))
        return g, gt

    def selftest_gap(self, n, seed=1):
        """Candidate gaps of the trial uniforms: fast path vs the double-precision expression over 2 n uniforms
        -> (mismatches, uniforms that took the slow path)."""
        bad, slow = C.c_int64(0), C.c_int64(0)
        _l.check(self.L.c2d_selftest_gap(self.h, int(n), int(seed), C.byref(bad), C.byref(slow)))
        return bad.value, slow.value

    def update(self, *, time, dt, f_nt, Pnt, n_e, B_field, tea, tna, vol, Eloss_sy, ec_old, turb_lev, gmin, gmax,
               p_nth, amxwl, f_pair=None, n_field=None, ecens=None, F_IC=None, lnL=20.0, r_esc=0.3, r_acc=1.0,
               inj=None, flare=None, group=None):
        """== update minus MPI (update2d.f:7-334): FP_calc for this rank's zones (zone % world == rank), then
        the exchange of the updated electron state (FP_recv_result + FP_end_bcast, fp_mpi.f:926,1017-1079).
        `n_field` / `ecens` default to this step's tallies on the device (after graphics_collect they are
        the reduced ones).  Returns a dict of the FP outputs for all zones."""
        nzn, N = self.nzone, self.num_nt
        inj = dict(inj or {})
        flare = dict(flare or {})
        keep = dict(f_nt=_d(f_nt, nzn * N, "f_nt"), Pnt=_d(Pnt, nzn * N, "Pnt"), n_field=_d(n_field, nzn * self.nphfield, "n_field"),
                    ecens=_d(ecens, nzn, "ecens"), n_e=_d(n_e, nzn, "n_e"), B_field=_d(B_field, nzn, "B_field"),
                    tea=_d(tea, nzn, "tea"), tna=_d(tna, nzn, "tna"), vol=_d(vol, nzn, "vol"), f_pair=_d(f_pair, nzn, "f_pair"),
                    Eloss_sy=_d(Eloss_sy, nzn, "Eloss_sy"), ec_old=_d(ec_old, nzn, "ec_old"), turb_lev=_d(turb_lev, nzn, "turb_lev"),
                    F_IC=_d(F_IC, N * self.nphfield, "F_IC"))
        fi = _l.FpIn(first_zone=self.rank, n_zones=-1, zone_stride=self.world, inj_switch=int(inj.get("inj_switch", 0)),
                     inj_dis=int(inj.get("inj_dis", 2)), g2var_switch=int(inj.get("g2var_switch", 0)),
                     pick_sw=int(inj.get("pick_sw", 0)), cf_sentinel=int(flare.get("cf_sentinel", 0)), ncycle=self.ncycle,
                     time=float(time), dt=float(dt), lnL=float(lnL), inj_g1=float(inj.get("inj_g1", 1e2)),
                     inj_g2=float(inj.get("inj_g2", 3e4)), inj_p=float(inj.get("inj_p", 1.0)), inj_t=float(inj.get("inj_t", 1.2e6)),
                     inj_gg=float(inj.get("inj_gg", 1e2)), inj_sigma=float(inj.get("inj_sigma", 1e1)),
                     inj_L=float(inj.get("inj_L", 5e40)), inj_v=float(inj.get("inj_v", 2.9965e10)),
                     pick_rate=float(inj.get("pick_rate", 0.8e-3)), r_esc=float(r_esc), r_acc=float(r_acc),
                     r_flare=float(flare.get("r_flare", 0.0)), z_flare=float(flare.get("z_flare", 0.0)),
                     t_flare=float(flare.get("t_flare", 1e6)), sigma_r=float(flare.get("sigma_r", 1e15)),
                     sigma_z=float(flare.get("sigma_z", 1e15)), sigma_t=float(flare.get("sigma_t", 1e6)),
                     flare_amp=float(flare.get("flare_amp", 10.0)))
        for k, v in keep.items():
            setattr(fi, k, _p(v))
        out = dict(f_nt=np.zeros((nzn, N)), Pnt=np.zeros((nzn, N)), n_e=np.zeros(nzn), Te_new=np.zeros(nzn),
                   gmin=np.array(gmin, dtype=_F).ravel().copy(), gmax=np.array(gmax, dtype=_F).ravel().copy(),
                   p_nth=np.array(p_nth, dtype=_F).ravel().copy(), amxwl=np.array(amxwl, dtype=_F).ravel().copy(),
                   f_pair=np.zeros(nzn), zone_E_old=np.zeros(nzn), zone_E_new=np.zeros(nzn), zone_hr=np.zeros(nzn),
                   zone_hr_st=np.zeros(nzn), zone_dT=np.zeros(nzn))
        steps = np.zeros(nzn, np.int32)
        fo = _l.FpOut()
        for k, v in out.items():
            setattr(fo, k, _p(v))
        fo.fp_steps = steps.ctypes.data_as(_l.ip)
        _l.check(self.L.c2d_update_fp(self.h, C.byref(fi), C.byref(fo)))
        out["fp_steps"] = steps
        if self.world > 1:
            self._exchange_fp(out, group)
        # totals over all zones in zone order (E_add_up, update2d.f:2002-2083)
        out.update(E_tot_old=float(np.sum(out["zone_E_old"])), E_tot_new=float(np.sum(out["zone_E_new"])),
                   hr_total=float(np.sum(out["zone_hr"])), hr_st_total=float(np.sum(out["zone_hr_st"])),
                   dT_max=float(np.max(out["zone_dT"])), dr=fo.dr, dz=fo.dz, f_t_implicit=fo.f_t_implicit)
        # update2d.f:268-274: the temperature carried to the next step is clamped to [5, 1000] keV
        out["tea"] = np.minimum(1.0e3, np.maximum(5.0, out["Te_new"]))
        return out

    def _exchange_fp(self, out, group=None):
        """Every rank owns zones rank::world; a sum over ranks of the zero-padded outputs is the all-gather
        of the updated electron state (the reference's FP_recv_result + FP_end_bcast)."""
        import torch
        import torch.distributed as dist
        own = np.zeros(self.nzone, bool)
        own[self.rank::self.world] = True
        keys = ["f_nt", "Pnt", "n_e", "Te_new", "gmin", "gmax", "p_nth", "amxwl", "f_pair", "zone_E_old", "zone_E_new",
                "zone_hr", "zone_hr_st", "zone_dT"]
        parts = []
        for k in keys:
            a = out[k].reshape(self.nzone, -1).astype(_F).copy()
            a[~own] = 0.0
            parts.append(a.ravel())
        parts.append(np.where(own, out["fp_steps"], 0).astype(_F))
        flat = np.concatenate(parts)
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        t = torch.from_numpy(flat).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        flat = t.cpu().numpy()
        off = 0
        for k in keys:
            n = out[k].size
            out[k] = flat[off: off + n].reshape(out[k].shape).copy()
            off += n
        out["fp_steps"] = flat[off: off + self.nzone].astype(np.int32)

    # ------------------------------------------------------------------ results
    def tallies(self):
        nzn = self.nzone
        out = dict(edep=np.zeros(nzn), prdep=np.zeros(nzn), ecens=np.zeros(nzn), npcen=np.zeros(nzn, np.int64),
                   n_field=np.zeros((nzn, self.nphfield)), n_ph=np.zeros((nzn, self.n_gg)), E_IC=np.zeros(self.num_nt),
                   nelectron=np.zeros(self.num_nt + 2, np.int64), fout=np.zeros((_l.NMUMAX, self.nphtotal)),
                   edout=np.zeros((_l.NMUMAX, _l.NLCMAX)), erlki=np.zeros(self.nz), erlko=np.zeros(self.nz),
                   erlku=np.zeros(self.nr), erlkl=np.zeros(self.nr), Ed_in=np.zeros(self.nr), Ed_ref=np.zeros(self.nr))
        out["fout_total"] = np.zeros((_l.NMUMAX, self.nphtotal))
        t = _l.Tallies()
        for k, v in out.items():
            setattr(t, k, v.ctypes.data_as(_l.lp if v.dtype == np.int64 else _l.dp))
        _l.check(self.L.c2d_tally_read(self.h, C.byref(t)))
        out["counters"] = {n: int(t.counters[i]) for i, n in enumerate(_l.COUNTERS)}
        out["counters_global"] = {n: int(t.counters_global[i]) for i, n in enumerate(_l.COUNTERS)}
        return out

    # ------------------------------------------------------------------ closed time step on the resident plasma
    def plasma_set(self, **kw):
        """Upload (parts of) the per-zone plasma state: f_nt, Pnt [zone][num_nt]; n_e, tea, tna, B_field, f_pair, gmin, gmax,
        p_nth, amxwl, turb_lev, vol, zsurf [zone]; k_gg [zone][n_gg].  The first call needs all but f_pair, turb_lev, k_gg."""
        nzn, N = self.nzone, self.num_nt
        size = dict(f_nt=nzn * N, Pnt=nzn * N, k_gg=nzn * self.n_gg)
        p = _l.Plasma()
        keep = {}
        for name, _ in _l.Plasma._fields_:
            v = kw.pop(name, None)
            keep[name] = _d(v, size.get(name, nzn), name)
            setattr(p, name, _p(keep[name]))
        if kw:
            raise TypeError(f"unknown plasma members {sorted(kw)}")
        _l.check(self.L.c2d_plasma_set(self.h, C.byref(p)))

    def plasma_get(self):
        nzn, N = self.nzone, self.num_nt
        out = {n: np.zeros((nzn, N) if n in ("f_nt", "Pnt") else nzn) for n, _ in _l.PlasmaOut._fields_}
        p = _l.PlasmaOut()
        for k, v in out.items():
            setattr(p, k, _p(v))
        _l.check(self.L.c2d_plasma_get(self.h, C.byref(p)))
        return out

    def run_config(self, *, nst, T_const=0, quirk_cycle0=0, inj=None, flare=None, r_esc=0.3, r_acc=1.0, surface=None):
        """The run constants of input.dat the step needs (reader.f:157-597); `surface` = dict(tbbi/tbbo [nz], tbbu/tbbl [nr],
        spectra=[file_sp dicts], spec_* per ring, t0, star_switch, Rstar, dist_star) or None."""
        inj, flare = dict(inj or {}), dict(flare or {})
        r = _l.Run(nst=int(nst), T_const=int(T_const), quirk_cycle0=int(quirk_cycle0), inj_switch=int(inj.get("inj_switch", 0)),
                   inj_dis=int(inj.get("inj_dis", 2)), g2var_switch=int(inj.get("g2var_switch", 0)), pick_sw=int(inj.get("pick_sw", 0)),
                   cf_sentinel=int(flare.get("cf_sentinel", 0)), inj_g1=float(inj.get("inj_g1", 1e2)), inj_g2=float(inj.get("inj_g2", 3e4)),
                   inj_p=float(inj.get("inj_p", 1.0)), inj_t=float(inj.get("inj_t", 1.2e6)), inj_gg=float(inj.get("inj_gg", 1e2)),
                   inj_sigma=float(inj.get("inj_sigma", 1e1)), inj_L=float(inj.get("inj_L", 5e40)), inj_v=float(inj.get("inj_v", 2.9965e10)),
                   pick_rate=float(inj.get("pick_rate", 0.8e-3)), r_esc=float(r_esc), r_acc=float(r_acc),
                   r_flare=float(flare.get("r_flare", 0.0)), z_flare=float(flare.get("z_flare", 0.0)), t_flare=float(flare.get("t_flare", 1e6)),
                   sigma_r=float(flare.get("sigma_r", 1e15)), sigma_z=float(flare.get("sigma_z", 1e15)), sigma_t=float(flare.get("sigma_t", 1e6)),
                   flare_amp=float(flare.get("flare_amp", 10.0)))
        keep = [r]
        if surface:
            nz, nr = self.nz, self.nr
            spectra = list(surface.get("spectra", ()))
            nfmax = max([len(t["E_file"]) for t in spectra], default=1)
            sr = _l.SurfRun(star_switch=int(surface.get("star_switch", 0)), Rstar=float(surface.get("Rstar", 1.0)),
                            dist_star=float(surface.get("dist_star", 1.0)), t0=float(surface.get("t0", 0.0)), n_spec=len(spectra), nfmax=nfmax)
            for name, n in (("tbbi", nz), ("tbbo", nz), ("tbbu", nr), ("tbbl", nr)):
                a = _d(surface.get(name), n, name); keep.append(a); setattr(sr, name, _p(a))
            for name, n in (("spec_i", nz), ("spec_o", nz), ("spec_u", nr), ("spec_l", nr)):
                v = surface.get(name)
                a = None if v is None else np.ascontiguousarray(v, dtype=np.int32).ravel()
                keep.append(a); setattr(sr, name, a.ctypes.data_as(_l.ip) if a is not None else None)
            tab = {k: np.zeros((max(len(spectra), 1), nfmax)) for k in ("E_file", "F_file", "P_file", "a1", "I_file")}
            nfile = np.zeros(max(len(spectra), 1), np.int32); intf = np.zeros(max(len(spectra), 1))
            for q, t in enumerate(spectra):
                for k in tab:
                    tab[k][q, : len(t[k])] = t[k]
                nfile[q] = t["nfile"]; intf[q] = t["int_file"]
            for k, v in tab.items():
                keep.append(v); setattr(sr, k, _p(v.ravel()))
            keep += [nfile, intf]
            sr.nfile = nfile.ctypes.data_as(_l.ip); sr.int_file = _p(intf)
            keep.append(sr)
            r.surf = C.pointer(sr)
        self._run = (r, keep)
        return r

    def step(self, dt, group=None):
        """== one iteration of xec's loop (xec2d.f:100-110 / 141-193) on the resident plasma: imcgen2d -> imcfield2d ->
        imcvol2d -> imcsurf2d -> update -> graphics_collect, entirely inside the library (c2d_step).  Advances self.time and
        self.ncycle as xec does (l.94-103).  -> dict of the scalars the reference logs for the cycle."""
        self.comm_init(group)
        rep = _l.StepReport()
        _l.check(self.L.c2d_step(self.h, C.byref(self._run[0]), int(self.ncycle), float(self.time), float(dt), C.byref(rep)))
        out = {n: getattr(rep, n) for n, _ in _l.StepReport._fields_ if n != "ms"}
        out["ms"] = dict(zip(_l.FAMILIES, list(rep.ms)))
        if self.ncycle > 0:
            self.time += float(dt)
        self.ncycle += 1
        return out

    def census_count(self):
        n = C.c_int64()
        _l.check(self.L.c2d_census_count(self.h, C.byref(n)))
        return n.value

    def census(self):
        """-> (dbuf[n,6], ibuf[n,6], sid[n]) in the record layout of imctrk2d.f:566-580."""
        n = self.census_count()
        d = np.zeros((max(n, 1), 6)); i = np.zeros((max(n, 1), 6), np.int32); s = np.zeros(max(n, 1), np.uint64)
        m = C.c_int64()
        _l.check(self.L.c2d_census_export(self.h, d.ctypes.data, i.ctypes.data, s.ctypes.data, max(n, 1), C.byref(m)))
        return d[:n], i[:n], s[:n]

    def census_import(self, dbuf, ibuf, sid=None):
        d = np.ascontiguousarray(dbuf, dtype=_F).reshape(-1, 6)
        i = np.ascontiguousarray(ibuf, dtype=np.int32).reshape(-1, 6)
        s = None if sid is None else np.ascontiguousarray(sid, dtype=np.uint64)
        _l.check(self.L.c2d_census_import(self.h, d.ctypes.data, i.ctypes.data, s.ctypes.data if s is not None else None,
                                          d.shape[0]))

    def events(self):
        """-> [n,7] escape records (t_bound,xnu,ew,rpre,zpre,wmu,phi), imcleak2d.f:171,181."""
        buf = np.zeros((self.event_capacity, 7))
        n = C.c_int64()
        _l.check(self.L.c2d_events_drain(self.h, buf.ctypes.data, self.event_capacity, C.byref(n)))
        return buf[: min(n.value, self.event_capacity)].copy()

    # ------------------------------------------------------------------ reference file formats
    def write_cens(self, path):
        """== write_cens (census2d.f:1-36) for this rank's census: p###_census.dat as write_record leaves it
        (write_record.f:432-437).  Returns the number of photons written (ndxout of p###_misc.dat)."""
        n = C.c_int64()
        _l.check(self.L.c2d_census_write(self.h, str(path).encode(), C.byref(n)))
        return n.value

    def read_cens(self, path, entries):
        """== read_cens (census2d.f:41-73): load `entries` photons of a p###_census.dat into the resident bank."""
        _l.check(self.L.c2d_census_read(self.h, str(path).encode(), int(entries)))

    def write_events(self, path, append=True):
        """This step's escape records in the reference's event-file format (format 105, imcleak2d.f:181); the reference
        appends to unit 1000+myid for the whole run.  Returns the number of records written."""
        n = C.c_int64()
        _l.check(self.L.c2d_events_write(self.h, str(path).encode(), 1 if append else 0, C.byref(n)))
        return n.value

    # ------------------------------------------------------------------ event-list reduction
    def lc_begin(self, red, slot=0):
        """Set the bin tables of the light-curve / SED reduction (`lightcurve.lc_reduction` = postprocessing/plcm.c's
        rules, `lightcurve.sed_reduction` = pspt.c's) and zero the accumulators.  `slot` 0..3: independent reducers fed by
        the same lc_accumulate (the reference runs plcm and pspt on the same event files)."""
        _l.check(self.L.c2d_lc_select(self.h, int(slot)))
        a = {k: _d(red[k]) for k in ("t_lo", "t_hi", "mu0", "mu1", "E0", "E1")}
        if a["t_lo"].size != a["t_hi"].size or a["mu0"].size != a["mu1"].size or a["E0"].size != a["E1"].size:
            raise ValueError("bin tables: lower and upper edges differ in length")
        cfg = _l.LcConfig(gam_bulk=float(red["gam_bulk"]), rmax=float(red["rmax"]), t_offset=float(red["t_offset"]),
                          n_t=a["t_lo"].size, n_mu=a["mu0"].size, n_bands=a["E0"].size, mu_closed=int(red["mu_closed"]),
                          first_band_only=int(red["first_band_only"]))
        for k, v in a.items():
            setattr(cfg, k, _p(v))
        _l.check(self.L.c2d_lc_begin(self.h, C.byref(cfg)))
        if not hasattr(self, "_lc_shapes"):
            self._lc_shapes = {}
        self._lc_shapes[int(slot)] = (a["t_lo"].size, a["mu0"].size, a["E0"].size)

    def lc_accumulate(self):
        """Bin this step's escape records where they lie in HBM (instead of `write(1000+myid, 105)`, imcleak2d.f:181,
        followed by plcm / pspt on the files).  Once per step, after imcfield2d / imcvol2d / imcsurf2d."""
        _l.check(self.L.c2d_lc_accumulate(self.h))

    def lc_accumulate_events(self, events):
        """The same for an [n,7] array of records in host memory (an event file read back)."""
        ev = np.ascontiguousarray(events, dtype=_F).reshape(-1, 7)
        _l.check(self.L.c2d_lc_accumulate_events(self.h, ev.ctypes.data, ev.shape[0]))

    def lc_read(self, group=None, slot=0):
        """-> (sum_ew, sum_ew2, count) [time][mu][band], summed over ranks (every rank wrote its own p###_evb.dat in the
        reference and plcm read them all, plcm.c:370-383,452-466); feed to lightcurve.finalize / finalize_sed."""
        shp = self._lc_shapes[int(slot)]
        _l.check(self.L.c2d_lc_select(self.h, int(slot)))
        s1, s2, cnt = np.zeros(shp), np.zeros(shp), np.zeros(shp, np.int64)
        _l.check(self.L.c2d_lc_read(self.h, s1.ctypes.data, s2.ctypes.data, cnt.ctypes.data))
        if self.world == 1:
            return s1, s2, cnt
        import torch
        import torch.distributed as dist
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"     # a few kB: reduced from the host copies
        tf = torch.from_numpy(np.stack([s1, s2])).to(dev)
        ti = torch.from_numpy(cnt).to(dev)
        dist.all_reduce(tf, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(ti, op=dist.ReduceOp.SUM, group=group)
        f = tf.cpu().numpy()
        return f[0], f[1], ti.cpu().numpy()

    def lc_bench(self, n_events):
        """-> (ms, algorithmic bytes) of k_lc_reduce on n generated records with the current tables."""
        ms, nb = C.c_float(), C.c_int64()
        _l.check(self.L.c2d_lc_bench(self.h, int(n_events), C.byref(ms), C.byref(nb)))
        return ms.value, nb.value

    # ------------------------------------------------------------------ measurement
    def timing(self, enable=True):
        _l.check(self.L.c2d_kernel_time_ms(self.h, -1 if enable else -5, None, None))

    def mark(self, stop=False):
        """Record the start / stop CUDA event of a timed bracket on the library's stream (kernel_time(7))."""
        _l.check(self.L.c2d_kernel_time_ms(self.h, -4 if stop else -3, None, None))

    def timing_reset(self):
        _l.check(self.L.c2d_kernel_time_ms(self.h, -2, None, None))

    def kernel_time(self, which):
        ms, n = C.c_float(), C.c_int64()
        _l.check(self.L.c2d_kernel_time_ms(self.h, int(which), C.byref(ms), C.byref(n)))
        return ms.value, n.value
