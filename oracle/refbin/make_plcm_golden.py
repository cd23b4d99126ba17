"""Generate tests/golden/plcm_*.npz and pspt_*.npz: outputs of the reference's own post-processors (postprocessing/plcm.c
light curves, postprocessing/pspt.c SEDs; compiled unmodified into oracle/_ref/ by refrun.build_plcm) on the event lists of
the reference-binary golden runs.
Runs ONLY in the authoring container (needs /root/reference).

usage: python oracle/refbin/make_plcm_golden.py"""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import refrun  # noqa: E402

REPO = Path(__file__).resolve().parent.parent.parent
GOLD = REPO / "tests" / "golden"

# answers in the order of plcm.c's prompts (postprocessing/README.lc.input): two angular bins wide enough to catch the
# ~1000 escape records of a golden run, 40 time bins with an offset, three regions with linear and logarithmic, partly overlapping bands
ANSWERS = """p001_evb.dat
33
0.75e16
lc07_ev0.dat
2
lc07_ev0.dat
-1.0
0.9995
lc07_ev1.dat
0.9995
1.0001
7.3e2
1.5e2
2.9e4
3
1e-6
1e-2
4
0
1e-3
10
2
1
1e2
1e8
3
0
"""


# pspt.c:98-196: 7 time bins over the arrival times of the golden runs' escape records, a closed mu window, two energy
# regions (20 logarithmic + 5 linear bins, overlapping: only the first matching band counts)
SED_ANSWERS = """p001_evb.dat
33
0.75e16
sed07.dat
7
1.1e2
2.1e4
-1.0
1.0
2
1e-8
1e-2
20
0
1e-3
10
5
1
n
"""


def parse_sed_output(text, n_t):
    """sedXX.dat rows: sqrt(E0 E1), F[n_t], particle_sum of the last time bin."""
    rows = np.array([[float(x) for x in l.split()] for l in text.splitlines() if l and l[0] != "#"])
    return dict(E=rows[:, 0], F=rows[:, 1: 1 + n_t], count_last=rows[:, 1 + n_t].astype(np.int64))


def parse_outputs(files, n_mu, n_bins):
    """lcXX.dat rows: time, F[n_bins], particle_sum[n_bins], sF/F1[n_bins]; lcXX_aux.dat rows: time, F1, F2, sF."""
    out = {}
    for n in range(n_mu):
        main = np.array([[float(x) for x in l.split()] for l in files[f"lc07_ev{n}.dat"].splitlines() if l and l[0] != "#"])
        aux = np.array([[float(x) for x in l.split()] for l in files[f"lc07_ev{n}_aux.dat"].splitlines() if l and l[0] != "#"])
        out[f"time{n}"] = main[:, 0]
        out[f"F{n}"] = main[:, 1: 1 + n_bins]
        out[f"count{n}"] = main[:, 1 + n_bins: 1 + 2 * n_bins].astype(np.int64)
        out[f"ratio{n}"] = main[:, 1 + 2 * n_bins: 1 + 3 * n_bins]
        out[f"F1_{n}"] = aux[:, 1: 1 + n_bins]
        out[f"F2_{n}"] = aux[:, 1 + n_bins: 1 + 2 * n_bins]
        out[f"sF{n}"] = aux[:, 1 + 2 * n_bins: 1 + 3 * n_bins]
    return out


def main():
    assert refrun.build_plcm() is not None and refrun.build_plcm(prog="pspt") is not None, "reference tree not present"
    for name in ("small", "dense"):
        ev = np.load(GOLD / f"refbin_{name}.npz")["events"]
        wd = Path("/tmp") / f"c2d_plcm_{name}"
        shutil.rmtree(wd, ignore_errors=True)
        files = refrun.run_plcm(ev, ANSWERS, wd)
        out = parse_outputs(files, 2, 9)
        out["answers"] = np.array(ANSWERS)
        np.savez_compressed(GOLD / f"plcm_{name}.npz", **out)
        wd = Path("/tmp") / f"c2d_pspt_{name}"
        shutil.rmtree(wd, ignore_errors=True)
        files = refrun.run_plcm(ev, SED_ANSWERS, wd, prog="pspt")
        sed = parse_sed_output(files["sed07.dat"], 7)
        sed["answers"] = np.array(SED_ANSWERS)
        np.savez_compressed(GOLD / f"pspt_{name}.npz", **sed)
        print(name, "pspt:", sed["F"].shape, "photons in the last time bin", int(sed["count_last"].sum()), "nonzero F", int((sed["F"] > 1e-20).sum()))
        print(name, len(ev), "events ->", {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim}, "counted", int(sum(out[f"count{n}"].sum() for n in range(2))))


if __name__ == "__main__":
    main()
