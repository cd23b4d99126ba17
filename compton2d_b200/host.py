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


class Compton2d:
    def __init__(self, *, nz, nr, z, r, rmin=0.0, zmin=0.0, E_ph, E_field, E_gg, gnt, hu, Elcmin=(), Elcmax=(),
                 mu=(1.0,), pair_switch=0, cr_sent=0, spec_switch=0, split=(1000, 1000, 300, 10), rseed=9857,
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
        self.census_capacity, self.event_capacity = int(census_capacity), int(event_capacity)
        self.h = C.c_void_p()
        _l.check(self.L.c2d_create(C.byref(self.h), int(device)))
        try:
            _l.check(self.L.c2d_init(self.h, C.byref(d)))
        except Exception:
            self.L.c2d_destroy(self.h)
            self.h = None
            raise
        self.fout_total = np.zeros((_l.NMUMAX, self.nphtotal))      # cumulative over the run (xec2d.f:389-394)
        self.ncycle = 0

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

    def imcfield2d(self):
        """== field_calc (imcfield2d.f:57-148): re-track this GPU's census bank."""
        _l.check(self.L.c2d_imcfield(self.h))

    def imcvol2d(self):
        """== sum over zones of vol_calc (imcvol2d_para.f:90-414); this rank's share of the histories."""
        _l.check(self.L.c2d_imcvol(self.h, self.rank, -1, self.world))

    def imcsurf2d(self):
        """Surface sources (imcsurf2d_para.f); the shipped inputs have none (tbb = -0.d0)."""
        return None

    def imcredist(self):
        """Census rebalancing through the master (imcredist.f) is not needed: histories are sharded
        statically and census photons never move between GPUs."""
        return None

    def sync(self):
        _l.check(self.L.c2d_sync(self.h))

    def graphics_collect(self, group=None):
        """cens_add_up + E_add_up + xec_add + graphics_collect (update2d.f:1931-2083, xec2d.f:325-394):
        one all-reduce(sum) per packed tally buffer, in place on the device."""
        self.sync()
        if self.world == 1:
            return
        import torch
        import torch.distributed as dist
        f, i = self.tally_device_views()
        tf = torch.as_tensor(f, device="cuda")
        ti = torch.as_tensor(i, device="cuda")
        # counters[0] = ndxout and counters[4] = nevents index this rank's census bank / event buffer:
        # they are per-rank state and must survive the reduction
        keep = ti[[-16, -12]].clone()
        dist.all_reduce(tf, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(ti, op=dist.ReduceOp.SUM, group=group)
        self.ndxout_global, self.nevents_global = (int(x) for x in ti[[-16, -12]].tolist())
        ti[[-16, -12]] = keep
        torch.cuda.synchronize()

    def tally_device_views(self):
        pf, pi = C.c_void_p(), C.c_void_p()
        nf, ni = C.c_int64(), C.c_int64()
        _l.check(self.L.c2d_tally_device(self.h, C.byref(pf), C.byref(nf), C.byref(pi), C.byref(ni)))
        return _CudaView(pf.value, nf.value, "<f8"), _CudaView(pi.value, ni.value, "<i8")

    # ------------------------------------------------------------------ results
    def tallies(self):
        nzn = self.nzone
        out = dict(edep=np.zeros(nzn), prdep=np.zeros(nzn), ecens=np.zeros(nzn), npcen=np.zeros(nzn, np.int64),
                   n_field=np.zeros((nzn, self.nphfield)), n_ph=np.zeros((nzn, self.n_gg)), E_IC=np.zeros(self.num_nt),
                   nelectron=np.zeros(self.num_nt + 2, np.int64), fout=np.zeros((_l.NMUMAX, self.nphtotal)),
                   edout=np.zeros((_l.NMUMAX, _l.NLCMAX)), erlki=np.zeros(self.nz), erlko=np.zeros(self.nz),
                   erlku=np.zeros(self.nr), erlkl=np.zeros(self.nr), Ed_in=np.zeros(self.nr), Ed_ref=np.zeros(self.nr))
        t = _l.Tallies()
        for k, v in out.items():
            setattr(t, k, v.ctypes.data_as(_l.lp if v.dtype == np.int64 else _l.dp))
        _l.check(self.L.c2d_tally_read(self.h, C.byref(t)))
        out["counters"] = {n: int(t.counters[i]) for i, n in enumerate(_l.COUNTERS)}
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

    # ------------------------------------------------------------------ measurement
    def timing(self, enable=True):
        _l.check(self.L.c2d_kernel_time_ms(self.h, -1 if enable else -2, None, None))

    def timing_reset(self):
        _l.check(self.L.c2d_kernel_time_ms(self.h, -2, None, None))

    def kernel_time(self, which):
        ms, n = C.c_float(), C.c_int64()
        _l.check(self.L.c2d_kernel_time_ms(self.h, int(which), C.byref(ms), C.byref(n)))
        return ms.value, n.value
