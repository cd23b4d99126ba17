"""Synthetic problems of the shapes BASELINE.json names, for bench.py and the parity tests.

The per-zone plasma tables (synchrotron emissivity CDFs eps_tot/eps_th, opacity kappa_tot, electron
distribution f_nt/Pnt on the reference's grids) are those the UNMODIFIED reference binary computed for its
shipped inputm.dat plasma (tea = tna = 100 keV, n_e = 80, B = 0.13 G, gmin = 100, gmax = 1e5, p = 2.3) --
extracted once into compton2d_b200/data/zone_tables.npz by oracle/refbin/make_golden.py's snapshots (zone (1,1) of
refbin_small, label c1_C).  Every zone gets the same plasma (as in the shipped run), the grid, photon budget
and energy groups follow setup2d.f / imcgen2d.f:
  z(j), r(k) uniform (setup2d.f:60-72); dt = mcdt*min(r_nr/nr, z_nz/nz)/inj_v (setup2d.f:50-52);
  vol = pi (r_k^2 - r_{k-1}^2) dz, zsurf = 2 pi ((r_k + r_{k-1}) dz + (r_k^2 - r_{k-1}^2)) (setup2d.f:80-114);
  Eloss_tot = emissivity*vol*dt; nsv = 0.5*nst*Eloss/sum(Eloss) (imcgen2d.f:448); ewsv = Eloss_tot/nsv.
"""
from pathlib import Path

import numpy as np

DATA = Path(__file__).resolve().parent / "data"
PI = 3.1415926536
C_LIGHT = 2.9979245620e10


def zone_tables():
    return dict(np.load(DATA / "zone_tables.npz"))


def hu_regions(regions):
    """setup2d.f:163-173: log-uniform group edges per region [(Emin, Emax, nbins), ...]."""
    hu = [regions[0][0]]
    for lo, hi, n in regions:
        q = (hi / lo) ** (1.0 / n)
        for m in range(1, n + 1):
            hu.append(lo * q ** m)
    return np.array(hu)


def make_problem(nz=16, nr=16, nst=20000, ngroups=100, split=(1000, 1000, 300, 10), z_max=1.0e16, r_max=0.75e16,
                 rmin=0.0, mcdt=1.0, n_e=None, pair_switch=0, jitter=0.0, seed=1, num_nt=None):
    T = zone_tables()
    nzone = nz * nr
    z = z_max * np.arange(1, nz + 1) / nz
    r = rmin + (r_max - rmin) * np.arange(1, nr + 1) / nr
    g_bulk = float(T["g_bulk"])
    inj_v = C_LIGHT * np.sqrt(1.0 - 1.0 / g_bulk ** 2)
    dt = mcdt * min((r_max - rmin) / nr, z_max / nz) / inj_v
    rlo = np.concatenate(([rmin], r[:-1]))
    dz = z_max / nz
    vol_k = PI * (r ** 2 - rlo ** 2) * dz
    zsurf_k = 2 * PI * ((r + rlo) * dz + (r ** 2 - rlo ** 2))
    vol = np.tile(vol_k, nz)
    zsurf = np.tile(zsurf_k, nz)
    rng = np.random.default_rng(seed)
    ne = np.full(nzone, float(T["n_e"]) if n_e is None else float(n_e))
    if jitter:
        ne = ne * (1.0 + jitter * (rng.random(nzone) - 0.5))
    Eloss_tot = float(T["emis_tot"]) * vol * dt
    Eloss_th = float(T["emis_th"]) * vol * dt
    nsv = np.floor(0.5 * nst * Eloss_tot / Eloss_tot.sum() + 0.5).astype(np.int32)
    ewsv = Eloss_tot / np.maximum(nsv, 1)
    tile = lambda a: np.ascontiguousarray(np.tile(np.asarray(a, float), (nzone, 1)))
    E_gg = 50.0 * (100.0 ** (1.0 / 100.0)) ** np.arange(100)          # setup2d.f:199-202
    P = dict(nz=nz, nr=nr, nzone=nzone, rmin=float(rmin), zmin=0.0, z=z, r=r, E_ph=T["E_ph"], E_field=T["E_field"], E_gg=E_gg,
             gnt=T["gnt"], hu=hu_regions([(1e-7, 1e11, ngroups)]),
             Elcmin=np.array([1e-7, 1e-3, 1.0, 1e2, 1e6]), Elcmax=np.array([1e-3, 1.0, 1e2, 1e6, 1e10]),
             mu=np.array([1.0]), pair_switch=pair_switch, spec_switch=0, split=tuple(split), dt=float(dt), time=0.0,
             kappa_tot=tile(T["kappa_tot"]), eps_tot=tile(T["eps_tot"]), eps_th=tile(T["eps_th"]),
             k_gg=np.zeros((nzone, 100)), f_nt=tile(T["f_nt"]), Pnt=tile(T["Pnt"]), n_e=ne,
             tea=np.full(nzone, float(T["tea"])), f_pair=np.zeros(nzone), nsv=nsv, ewsv=ewsv, Eloss_th=Eloss_th,
             Eloss_tot=Eloss_tot, zsurf=zsurf, tbbl=np.zeros(nr), vol=vol,
             tna=np.full(nzone, float(T["tna"])), B_field=np.full(nzone, float(T["B_field"])))
    return P


def context_for(P, rseed=9857, sigma_mode=0, census_capacity=400000, event_capacity=400000, device=0, rank=0, world=1,
                collision_capacity=65536):
    """A Compton2d context for problem dict P (make_problem or a reference snapshot in the same layout)."""
    from .host import Compton2d
    return Compton2d(nz=P["nz"], nr=P["nr"], z=P["z"], r=P["r"], rmin=P["rmin"], zmin=P["zmin"], E_ph=P["E_ph"],
                     E_field=P["E_field"], E_gg=P["E_gg"], gnt=P["gnt"], hu=P["hu"], Elcmin=P["Elcmin"],
                     Elcmax=P["Elcmax"], mu=P["mu"], pair_switch=P["pair_switch"], spec_switch=P["spec_switch"], cr_sent=P.get("cr_sent", 0),
                     E_ref=P.get("E_ref"), P_ref=P.get("P_ref"), W_abs=P.get("W_abs"),
                     split=P["split"], rseed=rseed, census_capacity=census_capacity, event_capacity=event_capacity,
                     collision_capacity=collision_capacity, sigma_mode=sigma_mode, device=device, rank=rank, world=world)
