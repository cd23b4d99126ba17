"""Event-list reduction, CPU side: the oracle's restatement of plcm.c's per-photon transform and binning,
and the host-side parser / finalizer of compton2d_b200.lightcurve, against what the reference's own plcm wrote."""
import numpy as np
import pytest

import orc
from lc_util import check_against_plcm, load_case


@pytest.mark.parametrize("name", ["small", "dense"])
def test_oracle_lightcurve_matches_reference_plcm(name):
    g, cfg, events = load_case(name)
    assert cfg["n_mu"] == 2 and len(cfg["E0"]) == 9 and cfg["gam_bulk"] == 33.0
    # plcm reads the event file the binary wrote with e14.7: feed the oracle the same 8-digit values
    s1, s2, cnt = orc.lightcurve(events, cfg)
    check_against_plcm(g, cfg, s1, s2, cnt)


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
