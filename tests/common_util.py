"""Fake COMMON blocks for the bridge tests: numpy arrays in Fortran order at the reference's padded compile-time extents
(general.pa:7-31: jmax = kmax = 99, nmumax = 32, nphomax = 128, nphlcmax = 10), i.e. byte for byte what `c_loc(edep)` etc. hand
over from integration/c2d_shim.f90."""
import ctypes as C

import numpy as np

from compton2d_b200 import lib as L

JMAX = KMAX = 99
NTMAX, NMUMAX, NPHOMAX, NPHLCMAX = 100, 32, 128, 10


class FakeCommon:
    def __init__(self, P, ti=1, jmax=JMAX, kmax=KMAX):
        self.P = P
        nz, nr = P["nz"], P["nr"]
        n_vol, N, nf, ngg = P["E_ph"].size, P["gnt"].size, P["E_field"].size, P["E_gg"].size
        f = lambda *shape: np.zeros(shape, order="F")
        fi = lambda *shape: np.zeros(shape, np.int32, order="F")
        a = self.a = {}
        for k in ("kappa_tot", "eps_tot", "eps_th"):
            a[k] = f(n_vol, jmax, kmax)
        a["k_gg"] = f(ngg, jmax, kmax); a["n_field"] = f(nf, jmax, kmax); a["n_ph"] = f(ngg, jmax, kmax)
        for k in ("f_nt", "Pnt"):
            a[k] = f(jmax, kmax, N)
        for k in ("n_e", "tea", "tna", "B_field", "f_pair", "ewsv", "Eloss_th", "Eloss_tot", "zsurf", "edep", "prdep", "ecens", "vol",
                  "Eloss_sy", "ec_old", "turb_lev", "gmin", "gmax", "p_nth", "amxwl", "Te_new"):
            a[k] = f(jmax, kmax)
        a["nsv"] = fi(jmax, kmax); a["npcen"] = fi(jmax, kmax); a["nelectron"] = fi(N)
        a["tbbl"] = f(kmax, NTMAX); a["E_IC"] = f(N); a["fout"] = f(NMUMAX, NPHOMAX); a["edout"] = f(NMUMAX, NPHLCMAX)
        a["erlki"] = f(jmax); a["erlko"] = f(jmax)
        for k in ("erlku", "erlkl", "Ed_in", "Ed_ref"):
            a[k] = f(kmax)
        for k in ("ndxout", "cmcount1", "cmcount2", "lkcount"):
            a[k] = fi(1)
        for k in ("dT_max", "E_tot_old", "E_tot_new", "hr_total", "hr_st_total", "dr", "dz", "f_t_implicit"):
            a[k] = f(1)
        # garbage outside the used extent: a shim bug that reads or writes padding must show
        rng = np.random.default_rng(5)
        for k, v in a.items():
            if v.dtype == np.float64 and v.ndim >= 2:
                v[...] = rng.random(v.shape) * 1e30
        self.jmax, self.kmax, self.ti = jmax, kmax, ti
        self.fill_step(P)

    def zone2(self, dense):
        """dense [zone] -> (nz, nr) block of a (jmax, kmax) array"""
        return np.asarray(dense).reshape(self.P["nz"], self.P["nr"])

    def fill_step(self, P):
        nz, nr = P["nz"], P["nr"]
        a = self.a
        for k in ("kappa_tot", "eps_tot", "eps_th", "k_gg"):
            a[k][:, :nz, :nr] = np.transpose(np.asarray(P[k]).reshape(nz, nr, -1), (2, 0, 1))
        for k in ("f_nt", "Pnt"):
            a[k][:nz, :nr, :] = np.asarray(P[k]).reshape(nz, nr, -1)
        for k in ("n_e", "tea", "f_pair", "ewsv", "Eloss_th", "Eloss_tot", "zsurf"):
            a[k][:nz, :nr] = self.zone2(P[k])
        a["nsv"][:nz, :nr] = self.zone2(P["nsv"])
        a["tbbl"][:nr, self.ti - 1] = P["tbbl"]

    def struct(self):
        cm = L.Common(jmax=self.jmax, kmax=self.kmax, ntmax=NTMAX, nmumax=NMUMAX, nphomax=NPHOMAX, nphlcmax=NPHLCMAX, ti=self.ti)
        for name, typ in L.Common._fields_:
            if name in self.a:
                setattr(cm, name, self.a[name].ctypes.data_as(typ))
        return cm
