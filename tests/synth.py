"""Synthetic problems for the parity tests: compton2d_b200.synth (shared with bench.py) plus the surface-source builder."""
from pathlib import Path

import numpy as np

from compton2d_b200.synth import C_LIGHT, PI, hu_regions, make_problem, zone_tables  # noqa: F401

GOLD = Path(__file__).resolve().parent / "golden"


def make_surface(P, nst, kind="file", tbb=0.01, seed=3):
    """Surface sources for problem P in the reference's own budget (imcgen2d.f:431-484): upper and lower disk rings with
    nsurf(k) = nst (r_k^2 - r_{k-1}^2) / r_nr^2 photons each.  kind = "file": energies from the reference's
    disk/blackbody_G15_4spectra.in table (tests/golden/spectrum_G15.npy) through compton2d_b200.spectrum.file_sp with the
    shipped external-field geometry (input.dat:106-112); kind = "planck": Planckian at `tbb` keV on all four boundaries
    (a path the reference never feeds with photons -- nsurf = 0 unless tbb < 0 -- but keeps in its source)."""
    from compton2d_b200.spectrum import file_sp
    nz, nr = P["nz"], P["nr"]
    r = P["r"]; rlo = np.concatenate(([P["rmin"]], r[:-1]))
    nring = np.floor(nst * (r ** 2 - rlo ** 2) / r[-1] ** 2).astype(np.int32)
    S = dict(star_switch=0)
    if kind == "file":
        t = file_sp(np.load(GOLD / "spectrum_G15.npy"), g_bulk=33.0, R_blr=2.18e38, fr_blr=0.1, R_ir=0.78e19, fr_ir=0.5)
        area = PI * (r ** 2 - rlo ** 2)
        ew = P["dt"] * area * t["int_file"] / np.maximum(nring, 1)          # erin = dt Asurf int_file (imcgen2d.f:159,177)
        S.update(nsurfu=nring, nsurfl=nring, ewsurfu=ew, ewsurfl=ew, tbbu=-np.ones(nr), tbbl=-np.ones(nr), spectra=[t])
    else:
        nzr = np.full(nz, max(int(nst) // max(nz, 1), 1), np.int32)
        S.update(nsurfu=nring, nsurfl=nring, nsurfi=nzr if P["rmin"] > 0 else None, nsurfo=nzr,
                 ewsurfu=np.full(nr, 1.0e40), ewsurfl=np.full(nr, 2.0e40), ewsurfi=np.full(nz, 3.0e40), ewsurfo=np.full(nz, 4.0e40),
                 tbbu=np.full(nr, tbb), tbbl=np.full(nr, 2 * tbb), tbbi=np.full(nz, 0.5 * tbb), tbbo=np.full(nz, tbb), spectra=[])
    return S
