"""Pin the CPU oracle (oracle/c2d_oracle.c) to the UNMODIFIED reference binary.

The fixtures tests/golden/refbin_*.npz are COMMON-block snapshots of
/root/reference/src_20121026/compton2d (run under oracle/refbin/fakempi.c by
oracle/refbin/make_golden.py) at the phase boundaries of cycles 0-2.  Each test loads the
state before a phase into the oracle, runs the oracle's restatement of that phase with the
reference's own ran1 stream (rand_switch = 2), and compares with the snapshot after it.

Bar: every integer (census count, census bins/zones/seeds, npcen, nelectron, event count,
cmcount1/2, lkcount, in/outcounter) and the RNG end state (= the exact number of draws
consumed) bit-exact; FP64 tallies to 1e-11 relative; census doubles to 1e-11 for >= 99 % of the
records and 1e-4 for all (the binary was built with ifort's default fast FP model, whose 1-ulp
differences are amplified by the cancellations in compb2d / acos near +-1).
"""
import numpy as np
import pytest

from oracle_golden import Golden, cname, compare_tracking, load_snapshot, new_oracle

CASES = [(n, c, ph) for n in ("small", "annulus", "dense", "mirror", "reflect") for c in (1, 2) for ph in ("field", "vol")]


def _run_phase(g, cyc, phase):
    l0, l1 = (f"c{cyc}_A", f"c{cyc}_B") if phase == "field" else (f"c{cyc}_B", f"c{cyc}_C")
    s0, s1 = g.snap(l0), g.snap(l1)
    o = new_oracle(s0)
    load_snapshot(o, s0)
    if phase == "vol":
        # these reach the worker with vol_bcast / the per-zone job message (vol_mpi.f:39-45,100-119),
        # i.e. after snapshot B was taken
        a = o.a
        for v in ("nsv", "zsurf", "eloss_th", "eloss_tot", "ewsv"):
            setattr(a, cname(v), s1[v])
        a.E_ph = s1.e_ph
        for v in ("eps_tot", "eps_th"):
            setattr(a, v, np.transpose(s1[v], (1, 2, 0)))
        o.vol_all()
    else:
        o.field_calc()
    return o, s1


@pytest.mark.parametrize("name,cyc,phase", CASES)
def test_phase_matches_reference_binary(name, cyc, phase):
    g = Golden(name)
    o, s1 = _run_phase(g, cyc, phase)
    assert int(o["fatal"]) == 0
    bad = compare_tracking(o, s1)
    d = bad.pop("dbufout", None)
    assert not bad, bad
    if d is not None:
        n = int(s1.ndxout)
        d0 = o.a.dbufout[: 6 * n].reshape(n, 6)
        d1 = s1.dbufout[: 6 * n].reshape(n, 6)
        r = np.abs(d0 - d1) / np.maximum(np.abs(d1), 1e-300)
        assert r.max() < 1e-4
        assert (r.max(axis=1) > 1e-11).mean() < 0.15      # scattered children (dense case)
        assert (r.max(axis=1) > 1e-7).mean() < 0.01


@pytest.mark.parametrize("name", ["small", "annulus", "dense", "mirror", "reflect"])
def test_event_file_matches_reference_binary(name):
    """The worker's event file (imcleak2d.f:171-181) over cycles 1-2, in order."""
    g = Golden(name)
    got = []
    for cyc in (1, 2):
        for phase in ("field", "vol"):
            o, _ = _run_phase(g, cyc, phase)
            got.append(o.events())
    got = np.concatenate(got)
    ref = g.events
    assert got.shape == ref.shape
    # columns: t_bound xnu ew rpre zpre wmu phi, printed as e14.7.  t_bound at the lower face is
    # the stale value of the previous escape (imcleak2d.f:170-171), so it depends on the state
    # carried across phases; compare it only where the oracle has it
    r = np.abs(got[:, 1:] - ref[:, 1:]) / np.maximum(np.abs(ref[:, 1:]), 1e-30)
    assert (r.max(axis=1) > 6e-7).mean() < 0.01       # e14.7: 7 digits after "0."
    assert r.max() < 1e-3
    tb = np.abs(got[:, 0] - ref[:, 0]) / np.maximum(np.abs(ref[:, 0]), 1e-30)
    assert (tb > 6e-7).mean() < 0.05


def test_ran1_lattice():
    """rand.f:354: ran1 returns 7-digit numbers in [0, 1)."""
    import orc
    o = orc.Oracle(1, 1)
    o.ran1_seed(-9857, np.zeros(32, np.int32), 0)
    u = np.array([o.L.orc_fibran(o.p) for _ in range(1000)]) if False else None
    L = o.L
    L.orc_fibran.restype = __import__("ctypes").c_double
    L.orc_fibran.argtypes = [__import__("ctypes").c_void_p]
    u = np.array([L.orc_fibran(o.p) for _ in range(2000)])
    assert (u >= 0).all() and (u < 1).all()
    assert np.allclose(u * 1e7, np.round(u * 1e7), atol=1e-6)
    assert 0.45 < u.mean() < 0.55


# ------------------------------------------------------------------------------------------ Fokker-Planck
@pytest.mark.parametrize("name,cyc", [("small", 1), ("small", 2), ("annulus", 1), ("annulus", 2)])
def test_fp_calc_matches_reference_binary(name, cyc):
    """FP_calc (update2d.f:337-1741) over all zones: snapshot E (before the FP farm) -> F (after).
    The per-zone job message (fp_mpi.f:612-700) delivers the master's tea, tna, B_field, Eloss_*, ecens,
    ec_old, n_field, ... so those inputs are read from F; what FP_calc itself overwrites (f_nt, Pnt, n_e,
    amxwl, gmin, gmax, p_nth, Te_new, f_pair) starts from E."""
    from oracle_golden import rel
    g = Golden(name)
    E, F = g.snap(f"c{cyc}_E"), g.snap(f"c{cyc}_F")
    o = new_oracle(E)
    load_snapshot(o, F)
    a = o.a
    for v in ("f_nt", "pnt", "n_e", "amxwl", "gmin", "gmax", "p_nth", "te_new", "f_pair"):
        setattr(a, cname(v), E[v])
    for s_ in ("E_tot_old", "E_tot_new", "hr_total", "hr_st_total", "dT_max"):
        o[s_] = 0.0
    nz, nr = o.dims["nz"], o.dims["nr"]
    for zone in range(1, nz * nr + 1):
        o.fp_calc(zone)
    assert (a.fp_steps[: nz * nr] == 5).all()
    for v, tol in (("f_nt", 1e-12), ("pnt", 1e-12)):
        assert rel(getattr(a, cname(v))[: nz * nr * 200].reshape(nz, nr, 200), F[v]) < tol, v
    for v, tol in (("n_e", 1e-14), ("amxwl", 1e-11), ("gmin", 0.0), ("gmax", 0.0), ("p_nth", 0.0), ("te_new", 1e-14)):
        assert rel(getattr(a, cname(v))[: nz * nr].reshape(nz, nr), F[v]) <= tol, v
    for gk, ck in (("e_tot_old", "E_tot_old"), ("e_tot_new", "E_tot_new"), ("hr_total", "hr_total"),
                   ("hr_st_total", "hr_st_total"), ("dt_max", "dT_max"), ("f_t_implicit", "f_t_implicit")):
        assert abs(o[ck] - float(F[gk])) <= 1e-12 * abs(float(F[gk])), ck


def test_ic_loss_matches_reference_binary():
    """IC_loss (icloss2d.f) vs the F_IC table in the reference binary's COMMON /icl/."""
    g = Golden("small")
    F = g.snap("c1_F")
    o = new_oracle(F)
    load_snapshot(o, F)
    o.a.F_IC[:] = 0
    o.L.orc_IC_loss(o.p)
    fic = o.a.F_IC[: 200 * 400].reshape(200, 400)
    assert np.abs(fic / F.f_ic - 1).max() < 1e-8      # f_Li series stops at |term/sum| < 1e-10


# ------------------------------------------------------------------------------------------ surface sources
def _load_surface_inputs(o, s1):
    """What z_surf_bcast / r_surf_send_job deliver after snapshot C was taken (surf_mpi.f:18-98,337) plus the
    file_sp tables the worker builds per ring (imcsurf2d_para.f:441-442)."""
    a = o.a
    for v in ("nsurfu", "nsurfl", "nsurfi", "nsurfo", "ewsurfu", "ewsurfl", "ewsurfi", "ewsurfo"):
        setattr(a, v, s1[v])
    a.tbbu = s1.tbbu[:, 0]; a.tbbl = s1.tbbl[:, 0]; a.tbbi = s1.tbbi[:, 0]; a.tbbo = s1.tbbo[:, 0]
    o["star_switch"] = int(s1.star_switch)
    o.set_spectrum(0, s1.e_file, s1.f_file, s1.p_file, s1.a1, s1.i_file, int(s1.nfile))
    a.spec_u[:] = 0; a.spec_l[:] = 0; a.spec_i[:] = 0; a.spec_o[:] = 0
    # tables that reach the worker with z_surf_bcast: the current kappa_tot, dt, hu ...
    a.kappa_tot = np.transpose(s1.kappa_tot, (1, 2, 0))
    o["dt"] = float(s1.dt[0])


@pytest.mark.parametrize("cyc", [1, 2])
def test_surface_sources_match_reference_binary(cyc):
    """imcsurf2d (upper + lower disk, file_sample spectrum; BASELINE configs[2]) between snapshots C and D."""
    g = Golden("surface")
    s0, s1 = g.snap(f"c{cyc}_C"), g.snap(f"c{cyc}_D")
    assert int(s1.ndxout) - int(s0.ndxout) > 1000
    o = new_oracle(s0)
    load_snapshot(o, s0)
    _load_surface_inputs(o, s1)
    o.surf_all()
    bad = compare_tracking(o, s1)
    assert not bad, bad


def test_file_sp_matches_reference_binary():
    """file_sp (imcsurf2d_para.f:544-685): the spectrum tables the binary built from disk/blackbody_G15_4spectra.in.
    The input rows are stored in the golden (tests/golden/spectrum_G15.npy) so the test does not need /root/reference."""
    from pathlib import Path
    g = Golden("surface")
    s1 = g.snap("c1_D")
    rows = np.load(Path(__file__).resolve().parent / "golden" / "spectrum_G15.npy")
    import orc
    t = orc.file_sp(rows, float(s1.g_bulk), float(s1.r_blr), float(s1.fr_blr), float(s1.r_ir), float(s1.fr_ir),
                    float(s1.r_disk), float(s1.d_jet))
    assert t["nfile"] == int(s1.nfile)
    n = t["nfile"]
    assert np.array_equal(t["E_file"][:n], s1.e_file[:n])
    for k, ref in (("F_file", s1.f_file), ("P_file", s1.p_file), ("a1", s1.a1), ("I_file", s1.i_file)):
        r = np.abs(t[k][: n - 1] - ref[: n - 1]) / np.maximum(np.abs(ref[: n - 1]), 1e-300)
        assert r.max() < 1e-11, (k, r.max())
    assert abs(t["int_file"] / float(s1.int_file) - 1.0) < 1e-12


def test_host_file_sp_matches_reference_binary():
    """compton2d_b200.spectrum.file_sp (the product's host-side parser) against the binary's tables."""
    from pathlib import Path
    from compton2d_b200.spectrum import file_sp
    s1 = Golden("surface").snap("c1_D")
    rows = np.load(Path(__file__).resolve().parent / "golden" / "spectrum_G15.npy")
    t = file_sp(rows, g_bulk=float(s1.g_bulk), R_blr=float(s1.r_blr), fr_blr=float(s1.fr_blr), R_ir=float(s1.r_ir),
                fr_ir=float(s1.fr_ir))
    n = t["nfile"]
    assert n == int(s1.nfile) and np.array_equal(t["E_file"][:n], s1.e_file[:n])
    for k, ref in (("F_file", s1.f_file), ("P_file", s1.p_file), ("a1", s1.a1), ("I_file", s1.i_file)):
        r = np.abs(t[k][: n - 1] - ref[: n - 1]) / np.maximum(np.abs(ref[: n - 1]), 1e-300)
        assert r.max() < 1e-11, (k, r.max())


# ------------------------------------------------------------------------------------------ step prologue
@pytest.mark.parametrize("name,cyc", [(n, c) for n in ("small", "annulus", "dense") for c in (1, 2)])
def test_volume_em_and_budget_match_reference_binary(name, cyc):
    """volume_em (volume2d.f:10-394) + the volume-source budget of imcgen2d (imcgen2d.f:203-459) from the plasma state the
    master held at the top of the cycle: emission CDFs, opacity, energy loss and photon counts as the binary computed them."""
    from gen_util import NST, plasma_from_snapshot, rel
    g = Golden(name)
    P, inp, ref = plasma_from_snapshot(g, cyc)
    A = g.snap(f"c{cyc}_A")
    o = new_oracle(A)
    load_snapshot(o, A)
    o.a.zsurf = inp["zsurf"]; o.a.vol = inp["vol"]
    o["dt"] = inp["dt"]
    o.imcgen_volume(NST[name])
    nzone = P["nzone"]
    a = o.a
    assert np.array_equal(a.E_ph[:400], P["E_ph"])
    assert np.array_equal(a.nsv[:nzone], ref["nsv"])
    for k in ("ewsv", "Eloss_tot", "Eloss_th"):
        assert rel(getattr(a, k)[:nzone], ref[k]) < 1e-13, k
    for k in ("kappa_tot", "eps_tot", "eps_th"):
        assert rel(getattr(a, k)[: nzone * 400], ref[k]) < 1e-13, k


# ---- rand_switch = 1: Knuth's lagged-Fibonacci generator (rand.f:9-316), tests/golden/refbin_knuth.npz ----
def test_seed_zone_matches_reference_binary():
    """seed_zone (rand.f:9-88) on the master deals one seed per zone every step and leaves seeds(1,1) as the next step's
    rseed; the worker's COMMON /random/ seeds after the volume phase of cycle c are the master's seeds of that cycle
    (vol_mpi.f:107,188)."""
    g = Golden("knuth")
    s = g.snap("c0_C")
    o = new_oracle(s)
    o["rand_switch"] = 1
    seed = 9857                                   # input.dat:77; the master is rank 0 (fp_mpi.f:68: rseed + 84725 myid)
    for cyc in (0, 1, 2):
        seed = o.seed_zone(seed)
        got = g.snap(f"c{cyc}_C").seeds
        assert np.array_equal(o.a.seeds.reshape(got.shape), got), cyc
        assert seed == int(got[0, 0])
    assert (o.a.rseeds == o.dims["nz"] * o.dims["nr"] + o.dims["nz"] + 1).all()       # rand.f:81-84


@pytest.mark.parametrize("cyc,phase", [(c, ph) for c in (1, 2) for ph in ("field", "vol")])
def test_phase_matches_reference_binary_knuth(cyc, phase):
    """The tracking phases with rand_switch = 1: a fresh generator per zone job (initialize_rand, imcvol2d_para.f:115) and per
    census photon, seeded by the integer the photon carries in ibuf(6) (imcfield2d.f:115-116).  Same bar as the ran1 cases;
    the generator's end state (RANX, randlist, randcounter) is compared instead of ran1's."""
    g = Golden("knuth")
    l0, l1 = (f"c{cyc}_A", f"c{cyc}_B") if phase == "field" else (f"c{cyc}_B", f"c{cyc}_C")
    s0, s1 = g.snap(l0), g.snap(l1)
    assert int(s0.rand_switch) == 1
    o = new_oracle(s0)
    load_snapshot(o, s0)
    o.knuth_set(s0.knuth_ranx, s0.randlist, int(s0.randcounter))
    o.a.seeds = s0.seeds
    if phase == "vol":
        a = o.a
        for v in ("nsv", "zsurf", "eloss_th", "eloss_tot", "ewsv"):
            setattr(a, cname(v), s1[v])
        a.E_ph = s1.e_ph
        for v in ("eps_tot", "eps_th"):
            setattr(a, v, np.transpose(s1[v], (1, 2, 0)))
        a.seeds = s1.seeds                         # arrive with the zone jobs (vol_mpi.f:188)
        o.vol_all()
    else:
        o.field_calc()
    assert int(o["fatal"]) == 0
    bad = compare_tracking(o, s1)
    d = bad.pop("dbufout", None)
    assert not bad, bad
    ranx, lst, cnt = o.knuth_get()
    assert cnt == int(s1.randcounter) and np.array_equal(ranx, s1.knuth_ranx) and np.array_equal(lst, s1.randlist)
    assert np.array_equal(o.a.seeds.reshape(s1.seeds.shape), s1.seeds)
    if d is not None:
        n = int(s1.ndxout)
        r = np.abs(o.a.dbufout[: 6 * n].reshape(n, 6) - s1.dbufout[: 6 * n].reshape(n, 6)) / np.maximum(np.abs(s1.dbufout[: 6 * n].reshape(n, 6)), 1e-300)
        assert r.max() < 1e-4 and (r.max(axis=1) > 1e-7).mean() < 0.01
