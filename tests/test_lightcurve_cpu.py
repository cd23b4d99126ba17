"""Event-list reduction, CPU side: the oracle's restatement of plcm.c's per-photon transform and binning,
and the host-side parser / finalizer of compton2d_b200.lightcurve, against what the reference's own plcm wrote."""
import numpy as np
import pytest

import orc
from compton2d_b200.lightcurve import lc_reduction, sed_reduction
from lc_util import check_against_plcm, check_against_pspt, load_case, load_sed_case


@pytest.mark.parametrize("name", ["small", "dense"])
def test_oracle_lightcurve_matches_reference_plcm(name):
    g, cfg, events = load_case(name)
    assert cfg["n_mu"] == 2 and len(cfg["E0"]) == 9 and cfg["gam_bulk"] == 33.0
    # plcm reads the event file the binary wrote with e14.7: feed the oracle the same 8-digit values
    s1, s2, cnt = orc.events_reduce(events, lc_reduction(cfg))
    check_against_plcm(g, cfg, s1, s2, cnt)


@pytest.mark.parametrize("name", ["small", "dense"])
def test_oracle_sed_matches_reference_pspt(name):
    g, cfg, events = load_sed_case(name)
    assert cfg["n_t"] == 7 and len(cfg["E0"]) == 25
    s1, s2, cnt = orc.events_reduce(events, sed_reduction(cfg))
    check_against_pspt(g, cfg, s1, cnt)
    # the regions overlap (1e-3 .. 1e-2 keV is in both): pspt counts a photon once, in the first band that holds it
    assert int(cnt.sum()) <= len(events)


def test_parse_reference_input_files():
    """The parser reads plcm's own answer files (their content is reproduced here; postprocessing/mrk421_lc.input)."""
    from compton2d_b200.lightcurve import parse_lc_input
    text = "p001_evb.dat\n33\n1e16\nlc07_ev0.dat\n1  \nlc07_ev0.dat\n0.99944 \n0.99964 \n7e2 \n0 \n7e4 \n7 \n" + \
        "".join(f"{a}\n{b}\n1\n0\n" for a, b in ((1e-3, 3e-3), (2, 4), (9, 15), (15, 20), (20, 60), (5e5, 5e7), (1e9, 1e10)))
    c = parse_lc_input(text)
    assert c["n_mu"] == 1 and c["dt"] == 700.0 and c["t_max"] == 7e4 and len(c["E0"]) == 7
    assert np.allclose(c["E1"], [3e-3, 4, 15, 20, 60, 5e7, 1e10]) and c["mu0"][0] == 0.99944
    # README-style file with comments after '!'
    c = parse_lc_input("p001_evb.dat\n15  ! Gamma\n1e17 ! r_max\nx.dat\n1\nx.dat\n0.9974 ! mu_min\n0.9982\n2e4\n0\n1.5e6\n1\n1e-6\n3e-6\n2\n1\n")
    assert c["gam_bulk"] == 15 and np.allclose(c["E0"], [1e-6, 2e-6]) and np.allclose(c["E1"], [2e-6, 3e-6])
    # pspt's answer file (postprocessing/mrk421_sed.input)
    from compton2d_b200.lightcurve import parse_sed_input
    c = parse_sed_input("p001_evb.dat\n33\n1e16\nsed30.dat\n30\n1.6e4\n6e4\n0.99944\n0.99964\n1\n1e-7\n1e10\n100\n0\nn\n")
    assert c["n_t"] == 30 and c["t_end"] == 6e4 and len(c["E0"]) == 100 and c["E0"][0] == 1e-7 and abs(c["E1"][-1] / 1e10 - 1) < 1e-12


def test_surface_budget_matches_reference_binary():
    """compton2d_b200.spectrum.surface_budget (imcgen2d.f:120-190,424-476) against what the binary's master shipped to its
    worker in the `surface` golden run (upper and lower disk fed by a spectrum file, nst = 1200)."""
    from compton2d_b200.spectrum import surface_budget
    from golden_util import Golden
    s = Golden("surface").snap("c1_D")
    b = surface_budget(nst=1200, time=float(s.time), dt=float(s.dt[0]), z=s.z, r=s.r, rmin=float(s.rmin), tbbi=s.tbbi[:, 0],
                       tbbo=s.tbbo[:, 0], tbbu=s.tbbu[:, 0], tbbl=s.tbbl[:, 0], t0=0.0, int_file=float(s.int_file))
    for f in "ioul":
        assert np.array_equal(b["nsurf" + f], np.asarray(s["nsurf" + f])), f
        assert np.allclose(b["ewsurf" + f], np.asarray(s["ewsurf" + f]), rtol=1e-14, atol=0), f
    assert b["nsurfu"].sum() == 1200 and b["fbias"] == 1.0
    # the > 10 nst cap
    c = surface_budget(nst=100, time=0.0, dt=1.0, z=s.z, r=s.r, tbbu=-np.ones(4), tbbl=-np.ones(4), int_file=1.0, n_volume=2000)
    n_new = 2000 + 2 * int((100 * np.diff(np.concatenate(([0.0], np.asarray(s.r) ** 2))) / s.r[-1] ** 2).astype(int).sum())
    assert c["fbias"] == 1000.0 / n_new and c["nsurfu"].sum() <= 50
