"""The step prologue on the device (volume_em + the volume-source budget of imcgen2d, csrc/c2d_gen.cu)
against the tables the UNMODIFIED reference binary computed (tests/golden/refbin_*.npz), and against host-supplied tables
for the step that follows.  Tolerance 1e-11: the device's pow / exp / sqrt differ from the binary's Intel libm by ulps."""
import numpy as np
import pytest

from gen_util import NST, check_against_binary, plasma_from_snapshot
from golden_util import Golden
from parity_util import gpu_step, make_gpu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,cyc", [(n, c) for n in ("small", "annulus", "dense") for c in (1, 2)])
def test_step_prologue_matches_reference_binary(name, cyc):
    g = Golden(name)
    P, inp, ref = plasma_from_snapshot(g, cyc)
    gpu = make_gpu(P, sigma_mode=1)
    try:
        out = gpu.imcgen2d_device(ncycle=cyc, nst=NST[name], f_pair=P["f_pair"], k_gg=P["k_gg"], tbbl=P["tbbl"], **inp)
        check_against_binary(out, gpu.step_tables(), ref)
        assert abs(out["Emiss_tot"] - ref["Eloss_tot"].sum()) <= 1e-12 * out["Emiss_tot"]
        gpu.imcvol2d(); gpu.sync()
        t_dev = gpu.tallies()
    finally:
        gpu.close()
    P2 = dict(P, kappa_tot=ref["kappa_tot"], eps_tot=ref["eps_tot"], eps_th=ref["eps_th"], nsv=ref["nsv"], ewsv=ref["ewsv"],
              Eloss_th=ref["Eloss_th"], Eloss_tot=ref["Eloss_tot"], f_nt=inp["f_nt"], Pnt=inp["Pnt"], n_e=inp["n_e"], tea=inp["tea"],
              dt=inp["dt"], time=inp["time"])
    gpu = make_gpu(P2, sigma_mode=1)
    try:
        t_host = gpu_step(gpu, P2, cyc, field=False)
    finally:
        gpu.close()
    assert np.array_equal(t_dev["npcen"], t_host["npcen"]) and t_dev["counters"]["ndxout"] == t_host["counters"]["ndxout"]
    assert np.abs(t_dev["ecens"] - t_host["ecens"]).max() <= 1e-9 * np.abs(t_host["ecens"]).max()


def test_step_prologue_rejects_foreign_energy_grid():
    from compton2d_b200 import C2dError
    g = Golden("small")
    P, inp, ref = plasma_from_snapshot(g, 1)
    P = dict(P, E_ph=P["E_ph"] * 1.001)
    gpu = make_gpu(P)
    try:
        with pytest.raises(C2dError) as e:
            gpu.imcgen2d_device(ncycle=1, nst=1000, **inp)
        assert e.value.code == 2
    finally:
        gpu.close()


@pytest.mark.gpu
def test_pref_calc_on_the_device_matches_reference_binary():
    """k_pref_calc on the B200 against the binary's COMMON /reflection/ (see tests/test_host_logic_cpu.py for the g++ build)."""
    from pathlib import Path
    from compton2d_b200.host import pref_calc
    g = np.load(Path(__file__).resolve().parent / "golden" / "refbin_reflect.npz")
    E, P = pref_calc(500)
    ref = g["c0_A/p_ref"]
    ref = ref if np.all(ref[0, 1:] == 1.0) else ref.T
    assert np.abs(E / g["c0_A/e_ref"] - 1.0).max() < 1e-13
    assert np.array_equal(P == 1.0, ref == 1.0) and np.array_equal(P == 0.0, ref == 0.0)
    assert np.abs(P - ref).max() < 1e-12
