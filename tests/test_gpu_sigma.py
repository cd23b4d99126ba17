"""sigma_mode 1 (per-step comtot table, c2d_sigma.cu) against the bin-by-bin evaluation of comtot2d.f:219-247
(sigma_mode 0 / c2d_comtot mode 0) and, through the replay tests, against the CPU oracle.

Tolerance: 2e-10 relative on comtot itself.  The table interpolates the smooth part of the exact-branch sum and adds
the rounding noise of the reference's log(1 + y) bin by bin (kn_eta); what is left is the ~1e-11 cancellation noise
of va - vb in intg_v, which no interpolant can follow."""
import numpy as np
import pytest

from parity_util import compare, gpu_step, make_gpu, make_oracle, oracle_step
from synth import make_problem

pytestmark = pytest.mark.gpu


def _begin(g, P, cyc=1):
    g.imcgen2d(ncycle=cyc, time=0.0, dt=P["dt"], kappa_tot=P["kappa_tot"], eps_tot=P["eps_tot"], eps_th=P["eps_th"],
               f_nt=P["f_nt"], Pnt=P["Pnt"], n_e=P["n_e"], tea=P["tea"], f_pair=P["f_pair"], k_gg=P["k_gg"], nsv=P["nsv"],
               ewsv=P["ewsv"], Eloss_th=P["Eloss_th"], Eloss_tot=P["Eloss_tot"], zsurf=P["zsurf"], tbbl=P["tbbl"])


def test_table_matches_bin_by_bin():
    P = make_problem(nz=8, nr=8, nst=1000, ngroups=100, jitter=0.4)
    rng = np.random.default_rng(5)
    P["f_nt"] = P["f_nt"] * (1.0 + 0.8 * rng.random(P["f_nt"].shape))     # every zone its own electron spectrum
    g = make_gpu(P, sigma_mode=1)
    try:
        _begin(g, P)
        n = 400_000
        x = 10.0 ** rng.uniform(-12.0, 13.9, n)
        z = rng.integers(0, P["nzone"], n)
        a, b = g.comtot(z, x, 0), g.comtot(z, x, 1)
        assert np.all(a > 0) and np.all(np.isfinite(b))
        r = np.abs(b / a - 1.0)
        assert r.max() < 2e-10, (r.max(), x[r.argmax()])
        assert np.median(r) < 1e-11
        # beyond the table (x > 1e14 keV) the table mode falls back to the bin-by-bin sum: identical bits
        xb = np.array([2e14, 1e15, 1e17]); zb = np.zeros(3, int)
        assert np.array_equal(g.comtot(zb, xb, 0), g.comtot(zb, xb, 1))
        # low-energy limit: every bin on the Thomson branch, comac -> n_e sigma_T sum(f dgamma)
        lo = g.comtot(np.zeros(1, int), np.array([1e-12]), 1)[0]
        w = (P["f_nt"][0, :-1] * np.diff(P["gnt"])).sum()
        assert abs(lo / (P["n_e"][0] * 6.6516e-25 * w) - 1.0) < 1e-9
    finally:
        g.close()


def test_table_follows_the_step_tables():
    """The table is rebuilt by every c2d_step_begin from that step's f_nt (the FP update changes it every step)."""
    P = make_problem(nz=3, nr=3, nst=100, ngroups=32)
    g = make_gpu(P, sigma_mode=1)
    try:
        rng = np.random.default_rng(1)
        x = 10.0 ** rng.uniform(-8, 8, 5000); z = rng.integers(0, 9, 5000)
        for cyc in (1, 2):
            P["f_nt"] = P["f_nt"] * (1.0 + 0.5 * rng.random(P["f_nt"].shape))
            _begin(g, P, cyc)
            a, b = g.comtot(z, x, 0), g.comtot(z, x, 1)
            assert np.abs(b / a - 1.0).max() < 2e-10
    finally:
        g.close()


@pytest.mark.parametrize("cfg", [dict(nz=16, nr=16, nst=6000, ngroups=100),
                                 dict(nz=3, nr=3, nst=600, ngroups=64, split=(20, 10, 4, 10), n_e=4.0e7),
                                 dict(nz=5, nr=4, nst=8000, ngroups=100, split=(1, 1, 1, 10), n_e=2.0e6)])
def test_replay_parity_with_table(cfg):
    """Philox replay against the CPU oracle with sigma_mode 1: same bar as tests/test_gpu_parity.py."""
    P = make_problem(**cfg)
    o, g = make_oracle(P), make_gpu(P, sigma_mode=1)
    try:
        for cyc in (1, 2):
            ro = oracle_step(o, P, cyc, field=cyc > 1)
            rg = gpu_step(g, P, cyc, field=cyc > 1)
            bad = compare(rg, ro, tol=1e-10, census_tol=1e-4)
            assert not bad, (cyc, bad)
    finally:
        g.close()


def test_third_split_block_teams(monkeypatch):
    """C2D_SPLIT3_ROUNDS=1: the warp phase of the 3rd split gives up after 32 passes, so most sub-children are
    finished by the 512-thread teams of k_split3_hard; replay parity (incl. the nelectron tally of every pass the
    sequential rejection loop makes) must not change."""
    monkeypatch.setenv("C2D_SPLIT3_ROUNDS", "1")
    P = make_problem(nz=3, nr=3, nst=600, ngroups=64, split=(20, 10, 4, 1000), n_e=4.0e7)
    o, g = make_oracle(P), make_gpu(P)
    try:
        for cyc in (1, 2):
            ro = oracle_step(o, P, cyc, field=cyc > 1)
            rg = gpu_step(g, P, cyc, field=cyc > 1)
            bad = compare(rg, ro, tol=1e-10, census_tol=1e-4)
            assert not bad, (cyc, bad)
        assert rg["counters"]["cmcount2"] > rg["counters"]["cmcount1"] * P["split"][1]      # 3rd split happened
    finally:
        g.close()


@pytest.mark.parametrize("hard_rounds,grid_blocks", [(0, 0), (1, 0), (0, 1)])
def test_third_split_grid_team(monkeypatch, hard_rounds, grid_blocks):
    """The last rung of the 3rd-split ladder: after 32 warp passes and hard_rounds * 512 block-team passes a sub-child is
    handed to k_split3_ultra, where the whole (cooperatively launched) grid tries one pass per thread between grid-wide
    barriers.  Replay parity must not change: the lowest accepted pass wins, the tallies of the passes before it count.
    grid_blocks = 1 shrinks the grid team to one block so that sub-children need several super-rounds (batches of 1, 2, 4 ..
    passes per thread, tallies of the passes beyond the winner taken back)."""
    if grid_blocks:
        monkeypatch.setenv("C2D_GRID_TEAM_BLOCKS", str(grid_blocks))
    monkeypatch.setenv("C2D_SPLIT3_ROUNDS", "1")
    monkeypatch.setenv("C2D_SPLIT3_HARD_ROUNDS", str(hard_rounds))
    monkeypatch.setenv("C2D_KN_CAP", "1000000000")                  # no child goes to a team: counter 9 counts sub-children only
    P = make_problem(nz=3, nr=3, nst=600, ngroups=64, split=(20, 10, 4, 1000), n_e=4.0e7)
    o, g = make_oracle(P), make_gpu(P)
    try:
        for cyc in (1, 2):
            ro = oracle_step(o, P, cyc, field=cyc > 1)
            rg = gpu_step(g, P, cyc, field=cyc > 1)
            bad = compare(rg, ro, tol=1e-10, census_tol=1e-4)
            assert not bad, (cyc, bad)
            c = rg["counters"]
            if hard_rounds == 0:                                    # the grid team finished what the warps left
                assert c["teams"] == c["s3hard"]
            else:
                assert c["teams"] <= c["s3hard"]
        assert c["s3hard"] > 0
    finally:
        g.close()


@pytest.mark.parametrize("kn_cap,hard_rounds,grid_blocks", [(1, 32, 0), (8, 0, 0), (1, 1, 0), (1, 0, 1)])
def test_children_block_teams(monkeypatch, kn_cap, hard_rounds, grid_blocks):
    """C2D_KN_CAP: a child whose first kn_cap trips through the electron / Klein-Nishina rejection loop (compb_2d.f label 100)
    fail, and whose photon has h nu >= 0.02 m c^2, is finished by a 256-thread team of k_children_hard (thread t makes trip
    base + t on draws 5 k .. 5 k + 4 of the child's stream), and after hard_rounds rounds of it by the whole grid
    (k_children_ultra, cooperative launch).  Hard photons: 100 keV Planckian boundaries (imcsurf2d), ~12 trips
    per child on average.  Replay parity incl. the nelectron tally of every rejected trip."""
    from synth import make_surface
    monkeypatch.setenv("C2D_KN_CAP", str(kn_cap))
    monkeypatch.setenv("C2D_KN_HARD_ROUNDS", str(hard_rounds))
    if grid_blocks:
        monkeypatch.setenv("C2D_GRID_TEAM_BLOCKS", str(grid_blocks))
    P = make_problem(nz=6, nr=6, nst=600, ngroups=100, split=(100, 10, 4, 10), n_e=2.0e7, rmin=0.1e16)
    S = make_surface(P, 3000, "planck", tbb=100.0)
    o, g = make_oracle(P), make_gpu(P, sigma_mode=1)
    try:
        for cyc in (1, 2):
            ro = oracle_step(o, P, cyc, field=cyc > 1, surf=S)
            rg = gpu_step(g, P, cyc, field=cyc > 1, surf=S)
            bad = compare(rg, ro, tol=1e-10, census_tol=1e-5)
            assert not bad, (cyc, bad)
            assert rg["counters"]["teams"] > 100                    # children that went to the block teams
            assert ro["nelectron"].sum() > 5 * ro["counters"]["cmcount2"]
    finally:
        g.close()
