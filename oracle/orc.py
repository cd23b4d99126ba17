"""ctypes harness around oracle/libc2d_oracle.so (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; nothing under compton2d_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "libc2d_oracle.so"


def build(force: bool = False) -> Path:
    srcs = [HERE / "c2d_oracle.c", HERE / "c2d_oracle_fp.c", HERE / "c2d_oracle.h"]
    if force or not LIB.exists() or LIB.stat().st_mtime < max(p.stat().st_mtime for p in srcs):
        subprocess.check_call(["make", "-s", "-C", str(HERE)])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int] * 9 + [C.c_int64, C.c_int64]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_array.restype = C.c_void_p
        L.orc_array.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        L.orc_set_scalar.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
        L.orc_get_scalar.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_get_scalar.restype = C.c_double
        L.orc_ran1_seed.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_ran1_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_philox_key.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_seed_zone.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_knuth_get.argtypes = [C.c_void_p] * 4
        L.orc_knuth_set.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_mix64.restype = C.c_uint64
        L.orc_mix64.argtypes = [C.c_uint64, C.c_uint64]
        L.orc_philox_uniform.restype = C.c_double
        L.orc_philox_uniform.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        L.orc_philox4x32.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        for fn in ("orc_dilog", "orc_intg_v", "orc_gammln", "orc_gamma_bar"):
            getattr(L, fn).restype = C.c_double
            getattr(L, fn).argtypes = [C.c_double]
        L.orc_McDonald.restype = C.c_double
        L.orc_McDonald.argtypes = [C.c_double, C.c_double]
        L.orc_comtot.restype = C.c_double
        L.orc_comtot.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double]
        L.orc_vol_calc.argtypes = [C.c_void_p, C.c_int]
        L.orc_field_calc.argtypes = [C.c_void_p]
        L.orc_volume_em.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.orc_imcgen_volume.argtypes = [C.c_void_p, C.c_int]
        L.orc_z_surf_calc.argtypes = [C.c_void_p, C.c_int]
        L.orc_r_surf_calc.argtypes = [C.c_void_p, C.c_int]
        L.orc_file_sp.restype = C.c_int
        L.orc_file_sp.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_double] * 7 + [C.c_void_p] * 6
        L.orc_step_zero_worker.argtypes = [C.c_void_p]
        L.orc_imctrk2d.argtypes = [C.c_void_p, C.c_int]
        L.orc_IC_loss.argtypes = [C.c_void_p]
        L.orc_fp_calc.argtypes = [C.c_void_p, C.c_int]
        L.orc_tridag.argtypes = [C.c_int] + [C.c_void_p] * 5
        _lib = L
    return _lib


class Oracle:
    """One orc_state.  Arrays are numpy views into the C allocation: `o.a.edep[...]`;
    scalars via `o["dt"] = ...` / `o["ndxout"]`."""

    def __init__(self, nz, nr, n_vol=400, num_nt=200, nphfield=400, n_gg=100, nphtotal=128,
                 nph_lc=5, nmu=1, cens_cap=200000, ev_cap=200000):
        self.L = lib()
        self.dims = dict(nz=nz, nr=nr, nzone=nz * nr, n_vol=n_vol, num_nt=num_nt, nphfield=nphfield,
                         n_gg=n_gg, nphtotal=nphtotal, nph_lc=nph_lc, nmu=nmu)
        self.p = self.L.orc_create(nz, nr, n_vol, num_nt, nphfield, n_gg, nphtotal, nph_lc, nmu,
                                   cens_cap, ev_cap)
        self.cens_cap, self.ev_cap = cens_cap, ev_cap
        self.a = _Arrays(self)

    def __del__(self):
        try:
            self.L.orc_destroy(self.p)
        except Exception:
            pass

    def array(self, name):
        n, es = C.c_int64(), C.c_int()
        ptr = self.L.orc_array(self.p, name.encode(), C.byref(n), C.byref(es))
        if not ptr:
            raise KeyError(name)
        ints64 = ("npcen", "nelectron", "sidout", "sidin")
        dt = {8: np.float64, 4: np.int32}[es.value]
        if name in ints64:
            dt = np.uint64 if name.startswith("sid") else np.int64
        buf = (C.c_char * (n.value * es.value)).from_address(ptr)
        return np.frombuffer(buf, dtype=dt)

    def __setitem__(self, k, v):
        if self.L.orc_set_scalar(self.p, k.encode(), float(v)) != 0:
            raise KeyError(k)

    def __getitem__(self, k):
        v = self.L.orc_get_scalar(self.p, k.encode())
        if v != v:
            raise KeyError(k)
        return v

    # -- RNG
    def ran1_seed(self, idum, iv, iy):
        iv = np.ascontiguousarray(iv, dtype=np.int32)
        self.L.orc_ran1_seed(self.p, int(idum), iv.ctypes.data, int(iy))

    def ran1_get(self):
        idum, iy = C.c_int(), C.c_int()
        iv = np.zeros(32, np.int32)
        self.L.orc_ran1_get(self.p, C.byref(idum), iv.ctypes.data, C.byref(iy))
        return idum.value, iv, iy.value

    def seed_zone(self, seed):
        """rand.f:9-88: fills o.a.seeds / zseeds / rseeds from `seed`; -> the seed of the next step (seeds(1,1))."""
        v = C.c_int(int(seed))
        self.L.orc_seed_zone(self.p, C.byref(v))
        return v.value

    def knuth_get(self):
        ranx, lst, cnt = np.zeros(100), np.zeros(10000), C.c_int()
        self.L.orc_knuth_get(self.p, ranx.ctypes.data, lst.ctypes.data, C.byref(cnt))
        return ranx, lst, cnt.value

    def knuth_set(self, ranx, randlist, randcounter):
        ranx, lst = np.ascontiguousarray(ranx, dtype=np.float64), np.ascontiguousarray(randlist, dtype=np.float64)
        self.L.orc_knuth_set(self.p, ranx.ctypes.data, lst.ctypes.data, int(randcounter))

    def philox_key(self, seed, cycle):
        self.L.orc_philox_key(self.p, int(seed), int(cycle))

    # -- phases
    def field_calc(self):
        self.L.orc_field_calc(self.p)

    def vol_calc(self, zone):
        self.L.orc_vol_calc(self.p, int(zone))

    def vol_all(self):
        for zone in range(1, self.dims["nzone"] + 1):
            self.vol_calc(zone)

    def imcgen_volume(self, nst):
        self.L.orc_imcgen_volume(self.p, int(nst))

    def surf_all(self):
        """imcsurf2d (imcsurf2d_para.f:1-225): all z rings, then all r rings."""
        for js in range(1, self.dims["nz"] + 1):
            self.L.orc_z_surf_calc(self.p, js)
        for ks in range(1, self.dims["nr"] + 1):
            self.L.orc_r_surf_calc(self.p, ks)

    def set_spectrum(self, slot, E_file, F_file, P_file, a1, I_file, nfile):
        a = self.a
        for name, v in (("E_file", E_file), ("F_file", F_file), ("P_file", P_file), ("a1_file", a1), ("I_file", I_file)):
            arr = getattr(a, name)
            arr[slot * 500: slot * 500 + 500] = 0.0
            arr[slot * 500: slot * 500 + len(v)] = v
        a.nfile[slot] = int(nfile)

    def fp_calc(self, zone):
        self.L.orc_fp_calc(self.p, int(zone))

    def step_zero(self):
        self.L.orc_step_zero_worker(self.p)

    def comtot(self, j, k, xnu):
        return self.L.orc_comtot(self.p, j, k, float(xnu))

    def census(self):
        n = int(self["ndxout"])
        return (self.a.dbufout[: 6 * n].reshape(n, 6).copy(), self.a.ibufout[: 6 * n].reshape(n, 6).copy(),
                self.a.sidout[:n].copy())

    def events(self):
        n = min(int(self["nev"]), self.ev_cap)
        return self.a.events[: 7 * n].reshape(n, 7).copy()


class _Arrays:
    def __init__(self, o):
        object.__setattr__(self, "_o", o)
        object.__setattr__(self, "_c", {})

    def __getattr__(self, name):
        c = self._c
        if name not in c:
            c[name] = self._o.array(name)
        return c[name]

    def __setattr__(self, name, val):
        a = getattr(self, name)
        v = np.asarray(val).ravel()
        a[: v.size] = v


def file_sp(rows, g_bulk, R_blr, fr_blr, R_ir, fr_ir, R_disk, d_jet):
    """orc_file_sp on an [n,>=4] table (E, L_disk, F_blr, F_ir) -> dict(E_file, F_file, P_file, a1, I_file, int_file, nfile)."""
    rows = np.ascontiguousarray(np.asarray(rows, float)[:, :4].T)
    out = [np.zeros(500) for _ in range(5)]
    integ = C.c_double()
    n = lib().orc_file_sp(rows.shape[1], rows[0].ctypes.data, rows[1].ctypes.data, rows[2].ctypes.data, rows[3].ctypes.data,
                          g_bulk, R_blr, fr_blr, R_ir, fr_ir, R_disk, d_jet, *[o.ctypes.data for o in out], C.byref(integ))
    return dict(E_file=out[0], F_file=out[1], P_file=out[2], a1=out[3], I_file=out[4], int_file=integ.value, nfile=n)


def events_reduce(events, red):
    """orc_events_reduce on an [n,7] event list with the bin tables of compton2d_b200.lightcurve.lc_reduction /
    sed_reduction.  -> (sum_ew, sum_ew2, count) [n_t][n_mu][n_bands]."""
    ev = np.ascontiguousarray(events, dtype=np.float64).reshape(-1, 7)
    arr = {k: np.ascontiguousarray(red[k], float) for k in ("t_lo", "t_hi", "mu0", "mu1", "E0", "E1")}
    n_t, n_mu, nb = len(arr["t_lo"]), len(arr["mu0"]), len(arr["E0"])
    s1 = np.zeros((n_t, n_mu, nb)); s2 = np.zeros((n_t, n_mu, nb)); cnt = np.zeros((n_t, n_mu, nb), np.int64)
    L = lib()
    L.orc_events_reduce.argtypes = [C.c_int64, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 3
    L.orc_events_reduce(len(ev), ev.ctypes.data, red["gam_bulk"], red["rmax"], red["t_offset"], n_t, arr["t_lo"].ctypes.data,
                        arr["t_hi"].ctypes.data, n_mu, arr["mu0"].ctypes.data, arr["mu1"].ctypes.data, int(red["mu_closed"]), nb,
                        arr["E0"].ctypes.data, arr["E1"].ctypes.data, int(red["first_band_only"]),
                        s1.ctypes.data, s2.ctypes.data, cnt.ctypes.data)
    return s1, s2, cnt
