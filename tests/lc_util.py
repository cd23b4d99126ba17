"""Shared by the light-curve tests: compare finalized accumulators with the outputs the reference's own
plcm (oracle/_ref/plcm, compiled unmodified from postprocessing/plcm.c) wrote -- tests/golden/plcm_*.npz."""
from pathlib import Path

import numpy as np

from compton2d_b200.lightcurve import finalize, finalize_sed, parse_lc_input, parse_sed_input

GOLD = Path(__file__).resolve().parent / "golden"


def load_case(name):
    g = np.load(GOLD / f"plcm_{name}.npz")
    cfg = parse_lc_input(str(g["answers"]))
    events = np.load(GOLD / f"refbin_{name}.npz")["events"]
    return g, cfg, events


def check_against_plcm(g, cfg, s1, s2, cnt, tol=2e-6):
    """plcm prints %e (7 significant digits) and stops after the first time bin that ends beyond t_max."""
    res = finalize(s1, s2, cnt, cfg["dt"], cfg["mu0"], cfg["mu1"])
    for n in range(cfg["n_mu"]):
        rows = len(g[f"time{n}"])
        assert np.array_equal(res["particle_sum"][:rows, n, :], g[f"count{n}"]), n
        for key, arr in (("F", f"F{n}"), ("F1", f"F1_{n}"), ("F2", f"F2_{n}")):
            ours = np.maximum(res[key][:rows, n, :], 1e-20) if key == "F" else res[key][:rows, n, :]
            ref = g[arr]
            assert np.allclose(ours, ref, rtol=tol, atol=1e-26), (key, n, np.abs(ours / np.maximum(ref, 1e-300) - 1).max())
        # sF = sqrt(<ew^2> - <ew>^2) suffers cancellation when all weights in a bin are nearly equal
        sf, ref = res["sF"][:rows, n, :], g[f"sF{n}"]
        scale = np.sqrt(res["F2"][:rows, n, :])
        assert np.all(np.abs(sf - ref) <= 2e-3 * scale + 1e-26), n
    assert int(cnt.sum()) > 1000


def load_sed_case(name):
    g = np.load(GOLD / f"pspt_{name}.npz")
    cfg = parse_sed_input(str(g["answers"]))
    events = np.load(GOLD / f"refbin_{name}.npz")["events"]
    return g, cfg, events


def check_against_pspt(g, cfg, s1, cnt, tol=2e-6):
    """pspt prints sqrt(E0 E1), F per time bin (%e) and the photon count of the last time bin only."""
    res = finalize_sed(s1, cnt, cfg)
    assert np.allclose(res["E"], g["E"], rtol=tol)
    assert np.array_equal(res["particle_sum"], g["count_last"])
    assert np.allclose(res["F"], g["F"], rtol=tol, atol=1e-26), np.abs(res["F"] / g["F"] - 1).max()
    assert (g["F"] > 1e-20).sum() > 50
