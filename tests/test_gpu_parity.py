"""GPU parity: the CUDA library (through its C ABI) against the CPU oracle in Philox replay mode.

Bar (BASELINE.json north_star): integer census / escape / collision counts and every census integer
bit-exact under the fixed Philox streams; FP64 tallies within 1e-10 of each array's maximum (sums are
accumulated in a different order and the device's log/exp/cos/acos differ from glibc's by <= 2 ulp);
census doubles within 1e-9 relative except scattered photons, where compb2d's cancellations amplify
ulp-level differences (same effect as between the oracle and the ifort binary, see
tests/test_oracle_vs_refbin.py) -- those are held to 1e-5."""
import numpy as np
import pytest

from golden_util import Golden
from parity_util import compare, gpu_step, make_gpu, make_oracle, oracle_step, problem_from_snapshot
from synth import make_problem

pytestmark = pytest.mark.gpu


def _run(P, cycles=(1, 2, 3), tol=1e-10, census_tol=1e-5, **kw):
    o = make_oracle(P)
    g = make_gpu(P, **kw)
    try:
        for cyc in cycles:
            ro = oracle_step(o, P, cyc, field=cyc > cycles[0])
            rg = gpu_step(g, P, cyc, field=cyc > cycles[0])
            bad = compare(rg, ro, tol=tol, census_tol=census_tol)
            assert not bad, (cyc, bad, rg["counters"], ro["counters"])
            d_g, d_o = rg["census"][0], ro["census"][0]
            if len(d_g):
                so, sg = np.argsort(ro["census"][2]), np.argsort(rg["census"][2])
                r = np.abs(d_g[sg] - d_o[so]) / np.maximum(np.abs(d_o[so]), 1e-300)
                assert (r.max(axis=1) > 1e-9).mean() < 0.02
    finally:
        g.close()
    return rg


@pytest.mark.parametrize("name,scale", [("small", 0.3), ("annulus", 0.3), ("dense", 1.0)])
def test_replay_parity_reference_tables(name, scale):
    """Tables and photon budgets of the reference binary's own runs (shipped physics; `dense` has
    n_e = 3e5 so that split2/split3 and scattered-photon tracking are exercised)."""
    g = Golden(name)
    P = problem_from_snapshot(g.snap("c1_B"), g.snap("c1_C"), nsv_scale=scale)
    _run(P)


def test_replay_parity_c1_shape():
    """BASELINE configs[1] shape (16x16 zones, 100 groups, split 1000/1000/300) at 3000 histories/step."""
    P = make_problem(nz=16, nr=16, nst=6000, ngroups=100)
    _run(P, cycles=(1, 2))


def test_replay_parity_multiple_scattering():
    """Dense plasma + small splits: collisions of scattered photons ('multiple scattering')."""
    P = make_problem(nz=3, nr=3, nst=600, ngroups=64, split=(20, 10, 4, 10), n_e=4.0e7)
    rg = _run(P, cycles=(1, 2), census_tol=1e-4)
    assert rg["counters"]["multiple"] > 0


def test_replay_parity_analog():
    """split = (1,1,1): analog Monte Carlo."""
    P = make_problem(nz=5, nr=4, nst=8000, ngroups=100, split=(1, 1, 1, 10), n_e=2.0e6)
    _run(P, cycles=(1, 2))


def test_long_rays_slow_path():
    """mcdt = 30: rays cross more zones than the fast path holds -> literal per-trial flights."""
    P = make_problem(nz=24, nr=20, nst=400, ngroups=32, split=(8, 4, 2, 10), mcdt=30.0, n_e=3.0e5)
    rg = _run(P, cycles=(1,))
    # (round 1's history kernel kept <= 12 crossings per ray and flew longer ones trial by trial; the streaming kernel has no
    # such limit -- these rays must simply replay)
    c = rg["counters"]
    assert c["segments"] > 8 * c["histories"]          # 8.6 crossings per ray on average, the longest > 30


def test_energy_conservation_and_determinism_full_size():
    """Size-independent properties at 2e6 histories (the oracle would need ~15 min): energy book-keeping
    E_in = ecens + leaks + edep - 2 sum(E_IC) and run-to-run identical integer tallies."""
    P = make_problem(nz=16, nr=16, nst=4_000_000, ngroups=100)
    res = []
    for rep in range(2):
        g = make_gpu(P, census_capacity=3_000_000, event_capacity=600_000)
        try:
            res.append(gpu_step(g, P, 1, field=False))
        finally:
            g.close()
    a, b = res
    assert a["counters"]["histories"] == int(P["nsv"].sum())
    assert np.array_equal(a["npcen"], b["npcen"]) and a["counters"] == b["counters"]
    assert int(a["npcen"].sum()) == a["counters"]["ndxout"] == len(a["census"][2])
    e_in = float((P["ewsv"] * P["nsv"]).sum())
    leak = a["erlki"].sum() + a["erlko"].sum() + a["erlku"].sum() + a["erlkl"].sum()
    # lower-face leaks are tallied in erlkl only when tbbl > 0 (imcleak2d.f:95-97): take them from the events
    ev = a["events"]
    e_out = a["ecens"].sum() + a["edep"].sum() - 2.0 * a["E_IC"].sum()
    lower = ev[ev[:, 4] <= P["zmin"] * (1 + 1e-12) + 1e-3][:, 2].sum()
    assert abs(e_in - (e_out + leak + lower)) / e_in < 1e-6
    assert abs(a["census"][0][:, 4].sum() - a["ecens"].sum()) / a["ecens"].sum() < 1e-12


def test_sharded_equals_whole():
    """Two shards (h mod 2) on one GPU reproduce the unsharded integer tallies: results do not depend on
    how histories are distributed (the multi-GPU contract)."""
    P = make_problem(nz=6, nr=5, nst=40000, ngroups=100, split=(50, 20, 5, 10), n_e=3.0e5)
    g = make_gpu(P)
    whole = gpu_step(g, P, 1, field=False)
    g.close()
    parts = []
    for rank in range(2):
        g = make_gpu(P, rank=rank, world=2)
        parts.append(gpu_step(g, P, 1, field=False))
        g.close()
    assert np.array_equal(parts[0]["npcen"] + parts[1]["npcen"], whole["npcen"])
    assert np.array_equal(parts[0]["nelectron"] + parts[1]["nelectron"], whole["nelectron"])
    for k in ("ndxout", "cmcount1", "cmcount2", "lkcount", "nevents"):
        assert parts[0]["counters"][k] + parts[1]["counters"][k] == whole["counters"][k]
    for k in ("edep", "ecens", "n_field", "fout", "E_IC"):
        s = parts[0][k] + parts[1][k]
        assert np.abs(s - whole[k]).max() <= 1e-11 * np.abs(whole[k]).max()
    sid = np.sort(np.concatenate([parts[0]["census"][2], parts[1]["census"][2]]))
    assert np.array_equal(sid, np.sort(whole["census"][2]))


@pytest.mark.parametrize("kind,sigma_mode", [("file", 0), ("file", 1), ("planck", 1)])
def test_replay_parity_surface_sources(kind, sigma_mode):
    """imcsurf2d (imcsurf2d_para.f:228-534): BASELINE configs[2] -- upper and lower disk rings sampling the reference's
    disk/blackbody_G15_4spectra.in through file_sp / file_sample (lower disk beamed at wmu = 0.99999999) -- and Planckian
    boundaries on all four faces (planck2d.f).  The oracle's restatement of this phase is pinned to the reference binary
    in tests/test_oracle_vs_refbin.py::test_surface_sources_match_reference_binary."""
    from synth import make_surface
    P = make_problem(nz=16, nr=16, nst=3000, ngroups=100, split=(100, 50, 10, 10), n_e=3.0e5,
                     rmin=0.1e16 if kind == "planck" else 0.0)
    S = make_surface(P, 6000, kind)
    o, g = make_oracle(P), make_gpu(P, sigma_mode=sigma_mode)
    try:
        for cyc in (1, 2):
            ro = oracle_step(o, P, cyc, field=cyc > 1, surf=S)
            rg = gpu_step(g, P, cyc, field=cyc > 1, surf=S)
            bad = compare(rg, ro, tol=1e-10, census_tol=1e-5)
            assert not bad, (cyc, bad)
        assert rg["counters"]["histories"] > 5000
    finally:
        g.close()


@pytest.mark.parametrize("sigma_mode", [0, 1])
def test_replay_parity_thin_rays_with_collisions(sigma_mode):
    """tau ~ 1e-3 < 1/256 with split1 = 1000: trial draws visited candidate by candidate (TrialGen), ~0.8 collisions per
    history so the candidate positions, the fine words and the literal trial flights are all exercised."""
    P = make_problem(nz=6, nr=6, nst=6000, ngroups=64, split=(1000, 4, 2, 10), n_e=2.0e6)
    rg = _run(P, cycles=(1, 2), census_tol=1e-4, sigma_mode=sigma_mode)
    c = rg["counters"]
    assert 0.1 < c["cmcount1"] / c["histories"] < 1000 / 256


@pytest.mark.parametrize("cr_sent,spec_switch,sigma_mode", [(1, 0, 0), (4, 1, 1)])
def test_replay_parity_mirror_at_the_lower_surface(cr_sent, spec_switch, sigma_mode):
    """SURVEY 8f N4, the live part: specular reflection at z = zmin (imcleak2d.f:107-122; the oracle's restatement is pinned to the
    reference binary run with cr_sent = 1, tests/golden/refbin_mirror.npz).  Tables of that run; recombined flights and scattered
    children bounce and go on in row 1, so nothing leaves downwards."""
    g = Golden("mirror")
    P = problem_from_snapshot(g.snap("c1_B"), g.snap("c1_C"))
    P["cr_sent"] = cr_sent; P["spec_switch"] = spec_switch
    rg = _run(P, sigma_mode=sigma_mode)
    ev = rg["events"]
    assert len(ev) > 100 and not np.any((ev[:, 4] <= P["zmin"]) & (ev[:, 5] < 0)) and rg["counters"]["cmcount2"] > 0


@pytest.mark.parametrize("cr_sent,sigma_mode", [(3, 0), (2, 1), (1, 1)])
def test_replay_parity_compton_reflection(cr_sent, sigma_mode):
    """Compton reflection through the tables of ref_matrix.f (static inputs; here the reference binary's own, from the `reflect`
    golden, whose run pins the oracle's restatement): hot lower boundary (imcleak2d.f:124-165; the flight goes on in row 1 with a
    new energy, so the opacity look-ups of the flight are redone) and outer cylinder (l.216-273)."""
    from parity_util import reflection_tables
    g = Golden("reflect")
    P = problem_from_snapshot(g.snap("c1_B"), g.snap("c1_C"))
    P["cr_sent"] = cr_sent
    reflection_tables(P)
    assert P["tbbl"].min() > 0
    rg = _run(P, sigma_mode=sigma_mode)
    if cr_sent != 2:
        assert rg["Ed_ref"].sum() > 0 and (rg["census"][0][:, 5] >= 1.0).sum() > 50
    if cr_sent != 1:
        ev = rg["events"]
        assert ((ev[:, 4] == 0) & (ev[:, 5] > 0) & (ev[:, 1] >= 1.0)).sum() > 10      # reflected at the outer disk


def test_candidate_gap_fast_path_is_exact():
    """TrialGen's gaps floor(ln U / ln(255/256)) come from the hardware's single-precision logarithm whenever the product is
    5e-3 away from an integer (error bound 2.3e-3, c2d_device.cuh:trial_gap); the double-precision expression decides the
    rest.  1e9 Philox uniforms: no disagreement, about 1 % on the slow path."""
    P = make_problem(nz=3, nr=3, nst=600, ngroups=64)
    g = make_gpu(P)
    try:
        bad, slow = g.selftest_gap(500_000_000, seed=12345)
        assert bad == 0
        assert 0.005 < slow / 1.0e9 < 0.02
    finally:
        g.close()
