"""The COMMON-block side of the boundary without a Fortran compiler: fake COMMON blocks at the padded
extents of general.pa:7-31 (jmax = kmax = 99, column-major), a full tracked step through pack -> C ABI -> unpack
(c2d_common_step_begin / c2d_common_collect, the two calls integration/c2d_shim.f90 makes), compared bit for bit with the
same step through the dense arrays.  Runs on the emu build here; the same test with the FP update runs on the GPU."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "emu"))


def _run_pair(P, cycles=(1, 2)):
    from common_util import FakeCommon
    from compton2d_b200 import lib as L
    from parity_util import gpu_step, make_gpu
    dense, bridged = make_gpu(P), make_gpu(P)
    fc = FakeCommon(P)
    for cyc in cycles:
        t = gpu_step(dense, P, cyc, field=cyc > cycles[0])
        dense.graphics_collect()
        t = dense.tallies()
        cm = fc.struct()
        L.check(bridged.L.c2d_common_step_begin(bridged.h, C.byref(cm), cyc, P["time"], P["dt"]))
        if cyc > cycles[0]:
            bridged.imcfield2d()
        bridged.imcvol2d()
        L.check(bridged.L.c2d_common_collect(bridged.h, C.byref(cm)))
    return dense, bridged, fc, t


def _check_tallies(P, fc, t, exact=True):
    """exact: the serial emu build sums in one order; on the device the two contexts commit their FP64 atomics in different
    orders, so sums agree to rounding (1e-12) and only the integers are identical."""
    nz, nr = P["nz"], P["nr"]
    a = fc.a
    same = np.array_equal if exact else (lambda x, y: np.allclose(x, y, rtol=1e-12, atol=0.0))
    for k in ("edep", "prdep", "ecens"):
        assert same(a[k][:nz, :nr].ravel(), t[k]), k
    assert np.array_equal(a["npcen"][:nz, :nr].ravel(), t["npcen"])
    assert same(np.transpose(a["n_field"][:, :nz, :nr], (1, 2, 0)).reshape(nz * nr, -1), t["n_field"])
    assert same(a["E_IC"], t["E_IC"]) and np.array_equal(a["nelectron"], t["nelectron"][1:-1])
    npt, nlc, nmu = P["hu"].size - 1, P["Elcmin"].size, P["mu"].size
    assert same(a["fout"][:nmu, :npt], t["fout_total"][:nmu]) and t["fout_total"].sum() > 0
    assert same(a["edout"][:nmu, :nlc], t["edout"][:nmu, :nlc])
    for k, n in (("erlki", nz), ("erlko", nz), ("erlku", nr), ("erlkl", nr)):
        assert same(a[k][:n], t[k]), k
    assert int(a["ndxout"][0]) == t["counters"]["ndxout"] and int(a["cmcount2"][0]) == t["counters"]["cmcount2"]
    # nothing outside the used extent was touched
    assert a["edep"][nz:, :].min() > 0 and a["edep"][:, nr:].min() > 0 and a["n_field"][:, nz:, :].min() > 0


@pytest.fixture()
def emu(monkeypatch):
    import build_emu
    from compton2d_b200 import lib as L
    so = build_emu.build()
    monkeypatch.setattr(L, "LIBPATH", so)
    monkeypatch.setattr(L, "_lib", None)
    yield
    monkeypatch.setattr(L, "_lib", None)


def test_common_bridge_step_cpu(emu):
    from synth import make_problem
    P = make_problem(nz=5, nr=4, nst=1500, ngroups=64, split=(40, 20, 6, 10), n_e=3.0e5)
    dense, bridged, fc, t = _run_pair(P)
    try:
        _check_tallies(P, fc, t)
        assert t["counters"]["cmcount2"] > 0 and t["counters"]["nevents"] > 0
    finally:
        dense.close(); bridged.close()


@pytest.mark.gpu
def test_common_bridge_step_and_update_gpu():
    """The same on the device, plus the shim's `update`: c2d_common_update against Compton2d.update on the dense arrays."""
    from common_util import FakeCommon
    from compton2d_b200 import lib as L
    from golden_util import Golden
    from loop_util import initial_state
    o, st, P, scal = initial_state(Golden("small"))
    del o
    dense, bridged, fc, t = _run_pair(P, cycles=(1, 2))
    try:
        _check_tallies(P, fc, t, exact=False)
        nz, nr, nzone = P["nz"], P["nr"], P["nzone"]
        N = P["gnt"].size
        Eloss_sy = np.linspace(1.0, 2.0, nzone) * 1e40
        ec_old = t["ecens"] * 0.9
        kw = dict(time=2 * scal["dt"], dt=scal["dt"], f_nt=st["f_nt"], Pnt=st["Pnt"], n_e=st["n_e"], B_field=st["B_field"], tea=st["tea"],
                  tna=st["tna"], vol=st["vol"], Eloss_sy=Eloss_sy, ec_old=ec_old, turb_lev=st["turb_lev"], gmin=st["gmin"], gmax=st["gmax"],
                  p_nth=st["p_nth"], amxwl=st["amxwl"], f_pair=st["f_pair"], lnL=20.0, r_esc=scal["r_esc"], r_acc=scal["r_acc"], inj=scal["inj"])
        ref = dense.update(**kw)
        a = fc.a
        for k in ("f_nt", "Pnt"):
            a[k][:nz, :nr, :] = np.asarray(st[k]).reshape(nz, nr, N)
        for k in ("n_e", "tea", "tna", "B_field", "f_pair", "vol", "turb_lev", "gmin", "gmax", "p_nth", "amxwl"):
            a[k][:nz, :nr] = np.asarray(st[k]).reshape(nz, nr)
        a["Eloss_sy"][:nz, :nr] = Eloss_sy.reshape(nz, nr); a["ec_old"][:nz, :nr] = ec_old.reshape(nz, nr)
        inj = scal["inj"]
        fi = L.FpIn(inj_switch=int(inj["inj_switch"]), inj_dis=int(inj["inj_dis"]), g2var_switch=int(inj["g2var_switch"]), pick_sw=int(inj["pick_sw"]),
                    ncycle=2, time=2 * scal["dt"], dt=scal["dt"], lnL=20.0, inj_g1=inj["inj_g1"], inj_g2=inj["inj_g2"], inj_p=inj["inj_p"],
                    inj_t=inj["inj_t"], inj_gg=inj["inj_gg"], inj_sigma=inj["inj_sigma"], inj_L=inj["inj_L"], inj_v=inj["inj_v"],
                    pick_rate=inj["pick_rate"], r_esc=scal["r_esc"], r_acc=scal["r_acc"], t_flare=1e6, sigma_r=1e15, sigma_z=1e15, sigma_t=1e6,
                    flare_amp=10.0)
        cm = fc.struct()
        L.check(bridged.L.c2d_common_update(bridged.h, C.byref(cm), C.byref(fi)))
        # (each context cools its electrons with its OWN photon-field tally, equal to the other's up to summation order)
        for k in ("f_nt", "Pnt"):
            assert np.allclose(a[k][:nz, :nr, :].reshape(nzone, N), ref[k], rtol=1e-9, atol=1e-300), k
        for k in ("gmin", "gmax", "p_nth", "tea"):
            assert np.array_equal(a[k][:nz, :nr].ravel(), ref[k]), k
        for k in ("n_e", "amxwl", "Te_new"):
            assert np.allclose(a[k][:nz, :nr].ravel(), ref[k], rtol=1e-10, atol=0), k
        assert abs(float(a["dT_max"][0]) / ref["dT_max"] - 1) < 1e-10 and abs(float(a["E_tot_new"][0]) / ref["E_tot_new"] - 1) < 1e-10
        assert (a["tea"][:nz, :nr] <= 1.0e3).all() and a["f_nt"][nz:, :, :].min() > 0
    finally:
        dense.close(); bridged.close()
