"""Host-logic tests on the GPU-less box.  The kernel sources of compton2d_b200/csrc are compiled with g++
against tests/emu/fake/cuda_runtime.h (serial "launches", TEST INFRASTRUCTURE ONLY -- the product never
loads this build) so that the C-ABI driver, the queue draining, the bank ping-pong, the sharding arithmetic
and the Python host mirror can be exercised against the CPU oracle without a device.  The real parity gate
is tests/test_gpu_parity.py on the B200."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "emu"))


@pytest.fixture()
def emu(monkeypatch):
    import build_emu
    from compton2d_b200 import lib as L
    so = build_emu.build()
    monkeypatch.setattr(L, "LIBPATH", so)
    monkeypatch.setattr(L, "_lib", None)
    yield
    monkeypatch.setattr(L, "_lib", None)


def test_step_sequence_matches_oracle(emu):
    from parity_util import compare, gpu_step, make_gpu, make_oracle, oracle_step
    from synth import make_problem
    P = make_problem(nz=4, nr=3, nst=1500, ngroups=64, split=(40, 20, 6, 10), n_e=3.0e5)
    o, g = make_oracle(P), make_gpu(P)
    for cyc in (1, 2, 3):
        ro = oracle_step(o, P, cyc, field=cyc > 1)
        rg = gpu_step(g, P, cyc, field=cyc > 1)
        bad = compare(rg, ro, tol=1e-11, census_tol=1e-6)   # scattered photons: compb2d amplifies ulp-level differences
        assert not bad, (cyc, bad)
    assert rg["counters"]["cmcount2"] > 0
    g.close()


def test_multiple_scattering_and_third_split(emu):
    """Dense plasma, small splits: scattered photons collide again ('multiple scattering', imctrk2d.f:614) and most
    children trigger the 3rd split (l.639-669)."""
    from parity_util import compare, gpu_step, make_gpu, make_oracle, oracle_step
    from synth import make_problem
    P = make_problem(nz=3, nr=3, nst=600, ngroups=64, split=(20, 10, 4, 10), n_e=4.0e7)
    o, g = make_oracle(P), make_gpu(P)
    for cyc in (1, 2):
        ro = oracle_step(o, P, cyc, field=cyc > 1)
        rg = gpu_step(g, P, cyc, field=cyc > 1)
        bad = compare(rg, ro, tol=1e-10, census_tol=1e-4)
        assert not bad, (cyc, bad)
    assert rg["counters"]["multiple"] > 0
    g.close()


def test_thin_rays_with_collisions(emu):
    """tau ~ 1e-3 < 1/256 with split1 = 1000: the trial draws go candidate by candidate (TrialGen gaps), and there are
    enough collisions (~0.8 per history) for the candidate positions to matter."""
    from parity_util import compare, gpu_step, make_gpu, make_oracle, oracle_step
    from synth import make_problem
    P = make_problem(nz=6, nr=6, nst=1200, ngroups=64, split=(1000, 4, 2, 10), n_e=2.0e6)
    o, g = make_oracle(P), make_gpu(P, sigma_mode=1)
    for cyc in (1, 2):
        ro = oracle_step(o, P, cyc, field=cyc > 1)
        rg = gpu_step(g, P, cyc, field=cyc > 1)
        bad = compare(rg, ro, tol=1e-10, census_tol=1e-4)
        assert not bad, (cyc, bad)
    c = rg["counters"]
    assert 0.1 < c["cmcount1"] / c["histories"] < 1000 / 256
    g.close()


@pytest.mark.parametrize("kind", ["file", "planck"])
def test_surface_sources(emu, kind):
    """imcsurf2d: disk sources sampled from an external spectrum (BASELINE configs[2]) and Planckian boundaries."""
    from parity_util import compare, gpu_step, make_gpu, make_oracle, oracle_step
    from synth import make_problem, make_surface
    P = make_problem(nz=4, nr=3, nst=800, ngroups=64, split=(20, 10, 4, 10), n_e=3.0e6, rmin=0.1e16 if kind == "planck" else 0.0)
    S = make_surface(P, 1500, kind)
    o, g = make_oracle(P), make_gpu(P)
    for cyc in (1, 2):
        ro = oracle_step(o, P, cyc, field=cyc > 1, surf=S)
        rg = gpu_step(g, P, cyc, field=cyc > 1, surf=S)
        bad = compare(rg, ro, tol=1e-10, census_tol=1e-6)
        assert not bad, (cyc, bad)
    assert rg["counters"]["histories"] > int(P["nsv"].sum()) + 1400
    g.close()


@pytest.mark.parametrize("name,cyc", [("small", 1), ("dense", 2)])
def test_step_prologue_on_device_matches_reference_binary(emu, name, cyc):
    """c2d_step_begin_gen (volume_em + volume-source budget, csrc/c2d_gen.cu) against the binary's own tables, then a
    tracked step from those device-built tables equals the step from the host-supplied ones."""
    from gen_util import NST, check_against_binary, plasma_from_snapshot
    from golden_util import Golden
    from parity_util import gpu_step, make_gpu
    g = Golden(name)
    P, inp, ref = plasma_from_snapshot(g, cyc)
    gpu = make_gpu(P)
    out = gpu.imcgen2d_device(ncycle=cyc, nst=NST[name], f_pair=P["f_pair"], k_gg=P["k_gg"], tbbl=P["tbbl"], **inp)
    check_against_binary(out, gpu.step_tables(), ref)
    gpu.imcvol2d(); gpu.sync()
    t_dev = gpu.tallies()
    gpu.close()
    # the same step from host tables (the binary's): identical integer tallies
    P2 = dict(P, kappa_tot=ref["kappa_tot"], eps_tot=ref["eps_tot"], eps_th=ref["eps_th"], nsv=ref["nsv"], ewsv=ref["ewsv"],
              Eloss_th=ref["Eloss_th"], Eloss_tot=ref["Eloss_tot"], f_nt=inp["f_nt"], Pnt=inp["Pnt"], n_e=inp["n_e"], tea=inp["tea"],
              dt=inp["dt"], time=inp["time"])
    gpu = make_gpu(P2)
    t_host = gpu_step(gpu, P2, cyc, field=False)
    gpu.close()
    assert np.array_equal(t_dev["npcen"], t_host["npcen"]) and t_dev["counters"]["ndxout"] == t_host["counters"]["ndxout"]


def test_census_export_import_roundtrip(emu):
    """write_record / read_record contract (census2d.f): export -> import into a fresh context -> same step."""
    from parity_util import gpu_step, make_gpu
    from synth import make_problem
    P = make_problem(nz=3, nr=3, nst=2000, ngroups=32, split=(10, 5, 2, 10))
    a = make_gpu(P)
    gpu_step(a, P, 1, field=False)
    d, i, s = a.census()
    ra = gpu_step(a, P, 2)
    b = make_gpu(P)
    b.census_import(d, i, s)
    rb = gpu_step(b, P, 2)
    assert np.array_equal(ra["npcen"], rb["npcen"]) and ra["counters"]["ndxout"] == rb["counters"]["ndxout"]
    assert np.allclose(ra["edep"], rb["edep"], rtol=1e-13, atol=0)
    a.close(); b.close()


def test_error_paths(emu):
    from compton2d_b200 import C2dError
    from parity_util import gpu_step, make_gpu
    from synth import make_problem
    P = make_problem(nz=3, nr=3, nst=2000, ngroups=32, split=(10, 5, 2, 10))
    g = make_gpu(P, census_capacity=50)                 # 'too many photons' (imctrk2d.f:581-585)
    with pytest.raises(C2dError) as e:
        gpu_step(g, P, 1, field=False)
    assert e.value.code == 4
    g.close()
    g = make_gpu(P)
    with pytest.raises(C2dError) as e:                  # step not begun
        g.imcvol2d()
    assert e.value.code == 6
    g.close()
    P2 = dict(P)
    with pytest.raises(C2dError) as e:                  # Compton reflection off the outer cylinder is not built
        from compton2d_b200.host import Compton2d
        Compton2d(nz=P["nz"], nr=P["nr"], z=P["z"], r=P["r"], E_ph=P["E_ph"], E_field=P["E_field"], E_gg=P["E_gg"],
                  gnt=P["gnt"], hu=P["hu"], cr_sent=2)
    assert e.value.code == 2


WORKER = r'''
import os, sys
sys.path[:0] = [r"{root}", r"{root}/tests", r"{root}/tests/emu", r"{root}/oracle"]
import numpy as np, torch, torch.distributed as dist
import build_emu
from compton2d_b200 import lib as L
L.LIBPATH = build_emu.build()
from parity_util import make_gpu
from synth import make_problem
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
P = make_problem(nz=4, nr=3, nst=3000, ngroups=32, split=(20, 10, 3, 10), n_e=3.0e5)
g = make_gpu(P, rank=rank, world=world)
red = dict(gam_bulk=10.0, rmax=float(P["r"][-1]), t_offset=0.0, t_lo=np.arange(32) * 0.4 * P["dt"], t_hi=(np.arange(32) + 1) * 0.4 * P["dt"],
           mu0=np.array([-1.0, 0.9]), mu1=np.array([0.9, 1.0]), mu_closed=0, E0=np.array([1e-9, 1e-3]), E1=np.array([1e-3, 1e3]),
           first_band_only=0)
g.lc_begin(red)
tot = None
for cyc in (1, 2):
    g.imcgen2d(ncycle=cyc, time=0.0, dt=P["dt"], kappa_tot=P["kappa_tot"], eps_tot=P["eps_tot"], eps_th=P["eps_th"],
               f_nt=P["f_nt"], Pnt=P["Pnt"], n_e=P["n_e"], tea=P["tea"], f_pair=P["f_pair"], k_gg=P["k_gg"], nsv=P["nsv"],
               ewsv=P["ewsv"], Eloss_th=P["Eloss_th"], Eloss_tot=P["Eloss_tot"], zsurf=P["zsurf"], tbbl=P["tbbl"])
    if cyc > 1:
        g.imcfield2d()
    g.imcvol2d()
    g.sync()
    g.lc_accumulate()
    t = g.tallies()
    # the reduction the CUDA path does in place on the device, here over gloo on the host copies
    pack = np.concatenate([t["edep"], t["ecens"], t["n_field"].ravel(), t["fout"].ravel(), t["E_IC"]])
    ints = np.concatenate([t["npcen"], t["nelectron"], [t["counters"][k] for k in ("ndxout", "cmcount1", "cmcount2", "nevents")]])
    tp, ti = torch.from_numpy(pack), torch.from_numpy(ints)
    dist.all_reduce(tp); dist.all_reduce(ti)
    sid = g.census()[2]
    gathered = [None] * world
    dist.all_gather_object(gathered, sid)
lc1, lc2, lcn = g.lc_read()
if rank == 0:
    np.savez(r"{out}", pack=tp.numpy(), ints=ti.numpy(), sid=np.sort(np.concatenate(gathered)), lc1=lc1, lc2=lc2, lcn=lcn)
dist.destroy_process_group()
'''


def test_two_ranks_equal_one(emu, tmp_path):
    """world_size 2 over gloo: histories sharded h mod 2, census kept per rank, tallies all-reduced ==
    the single-rank run (integers exact, FP64 to 1e-12) after two steps."""
    from parity_util import gpu_step, make_gpu
    from synth import make_problem
    P = make_problem(nz=4, nr=3, nst=3000, ngroups=32, split=(20, 10, 3, 10), n_e=3.0e5)
    g = make_gpu(P)
    red = dict(gam_bulk=10.0, rmax=float(P["r"][-1]), t_offset=0.0, t_lo=np.arange(32) * 0.4 * P["dt"],
               t_hi=(np.arange(32) + 1) * 0.4 * P["dt"], mu0=np.array([-1.0, 0.9]), mu1=np.array([0.9, 1.0]), mu_closed=0,
               E0=np.array([1e-9, 1e-3]), E1=np.array([1e-3, 1e3]), first_band_only=0)
    g.lc_begin(red)
    gpu_step(g, P, 1, field=False); g.lc_accumulate()
    one = gpu_step(g, P, 2); g.lc_accumulate()
    lc_one = g.lc_read()
    g.close()
    out = tmp_path / "two.npz"
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT), out=str(out)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29541", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    two = np.load(out)
    ints1 = np.concatenate([one["npcen"], one["nelectron"], [one["counters"][k] for k in ("ndxout", "cmcount1", "cmcount2", "nevents")]])
    assert np.array_equal(two["ints"], ints1)
    pack1 = np.concatenate([one["edep"], one["ecens"], one["n_field"].ravel(), one["fout"].ravel(), one["E_IC"]])
    assert np.abs(two["pack"] - pack1).max() <= 1e-12 * np.abs(pack1).max()
    assert np.array_equal(two["sid"], np.sort(one["census"][2]))
    # light-curve accumulators of the two ranks, summed, == the single-rank ones (every rank bins its own escape records)
    assert np.array_equal(two["lcn"], lc_one[2]) and lc_one[2].sum() > 20
    assert np.allclose(two["lc1"], lc_one[0], rtol=1e-12, atol=0) and np.allclose(two["lc2"], lc_one[1], rtol=1e-12, atol=0)


def test_event_list_reduction_host_logic(emu):
    """SURVEY 8f N2 through the C ABI (emu build: the transform + bin functions of csrc/c2d_lc.cu run serially): records from
    host memory against the reference's own plcm / pspt outputs, then the records of a tracked step binned where they lie
    against the oracle's reduction of the drained list, accumulating over two steps."""
    import orc
    from compton2d_b200 import C2dError
    from compton2d_b200.lightcurve import lc_reduction, sed_reduction
    from lc_util import check_against_plcm, check_against_pspt, load_case, load_sed_case
    from parity_util import gpu_step, make_gpu
    from synth import make_problem
    P = make_problem(nz=4, nr=3, nst=1500, ngroups=64, split=(40, 20, 6, 10), n_e=3.0e5)
    g = make_gpu(P)
    with pytest.raises(C2dError):
        g.lc_accumulate()                                  # c2d_lc_begin first
    gold, cfg, events = load_case("small")
    g.lc_begin(lc_reduction(cfg))
    g.lc_accumulate_events(events[:1000]); g.lc_accumulate_events(events[1000:])
    check_against_plcm(gold, cfg, *g.lc_read())
    gold, cfg, events = load_sed_case("dense")
    g.lc_begin(sed_reduction(cfg))
    g.lc_accumulate_events(events)
    s1, s2, cnt = g.lc_read()
    check_against_pspt(gold, cfg, s1, cnt)
    # two reducers fed by the same records: plcm's light curve in slot 0, pspt's SED in slot 1
    gold_lc, cfg_lc, ev_small = load_case("small")
    gold_sed, cfg_sed, _ = load_sed_case("small")
    g.lc_begin(lc_reduction(cfg_lc), slot=0); g.lc_begin(sed_reduction(cfg_sed), slot=1)
    g.lc_accumulate_events(ev_small)
    check_against_plcm(gold_lc, cfg_lc, *g.lc_read(slot=0))
    s1, s2, cnt = g.lc_read(slot=1)
    check_against_pspt(gold_sed, cfg_sed, s1, cnt)
    g.lc_begin(sed_reduction(cfg), slot=1)               # re-beginning one slot leaves the other alone
    check_against_plcm(gold_lc, cfg_lc, *g.lc_read(slot=0))
    # a tracked step: the blob of make_problem, wide angle bins, arrival times of a few light-crossing times
    red = dict(gam_bulk=10.0, rmax=float(P["r"][-1]), t_offset=0.0, t_lo=np.arange(64) * 0.2 * P["dt"],
               t_hi=(np.arange(64) + 1) * 0.2 * P["dt"], mu0=np.array([-1.0, 0.9]), mu1=np.array([0.9, 1.0]), mu_closed=0,
               E0=np.array([1e-9, 1e-3, 1e-5]), E1=np.array([1e-3, 1e3, 1e1]), first_band_only=0)
    g.lc_begin(red)
    ev_all = []
    for cyc in (1, 2):
        t = gpu_step(g, P, cyc, field=cyc > 1)
        g.lc_accumulate()
        ev_all.append(t["events"])
    ev_all = np.concatenate(ev_all)
    assert len(ev_all) > 50
    o1, o2, oc = orc.events_reduce(ev_all, red)
    s1, s2, cnt = g.lc_read()
    assert np.array_equal(cnt, oc) and cnt.sum() > 50
    assert np.allclose(s1, o1, rtol=1e-13, atol=0) and np.allclose(s2, o2, rtol=1e-13, atol=0)
    bad = dict(red, t_hi=red["t_hi"][::-1].copy())
    with pytest.raises(C2dError) as e:
        g.lc_begin(bad)
    assert e.value.code == 2
    g.close()


@pytest.mark.parametrize("cr_sent,spec_switch", [(1, 0), (4, 1)])
def test_mirror_at_the_lower_surface(emu, cr_sent, spec_switch):
    """SURVEY 8f N4, the live part: specular reflection at the lower surface (imcleak2d.f:107-122, cr_sent = 1 with a cold
    boundary / cr_sent = 4).  Recombined flights and scattered children bounce (wmu -> -wmu, bins re-derived, the flight goes
    on in row 1); nothing escapes through z = zmin."""
    from compton2d_b200 import C2dError
    from parity_util import compare, gpu_step, make_gpu, make_oracle, oracle_step
    from synth import make_problem
    P = make_problem(nz=3, nr=3, nst=1500, ngroups=64, split=(40, 20, 6, 10), n_e=3.0e5)
    P["cr_sent"] = cr_sent; P["spec_switch"] = spec_switch
    o, g = make_oracle(P), make_gpu(P)
    for cyc in (1, 2):
        ro = oracle_step(o, P, cyc, field=cyc > 1)
        rg = gpu_step(g, P, cyc, field=cyc > 1)
        bad = compare(rg, ro, tol=1e-11, census_tol=1e-6)
        assert not bad, (cyc, bad)
    assert rg["counters"]["cmcount2"] > 0 and rg["counters"]["lkcount"] > 0
    ev = rg["events"]
    assert len(ev) > 10 and not np.any((ev[:, 4] <= P["zmin"]) & (ev[:, 5] < 0))       # no record leaves downwards through z = zmin
    if cr_sent == 1:                                       # a hot lower boundary would need the reflection matrices
        with pytest.raises(C2dError) as e:
            g.imcgen2d(ncycle=3, time=0.0, dt=P["dt"], kappa_tot=P["kappa_tot"], eps_tot=P["eps_tot"], eps_th=P["eps_th"],
                       f_nt=P["f_nt"], Pnt=P["Pnt"], n_e=P["n_e"], nsv=P["nsv"], ewsv=P["ewsv"], tbbl=np.full(P["nr"], 0.5))
        assert e.value.code == 2
    g.close()


@pytest.mark.parametrize("cr_sent", [3, 2])
def test_compton_reflection_tables(emu, cr_sent):
    """Compton reflection through the Green's-function tables of ref_matrix.f (inputs; here the binary's own): off a hot lower
    boundary (imcleak2d.f:124-165, the flight goes on with a new energy) and off the outer cylinder (l.216-273)."""
    from compton2d_b200 import C2dError
    from parity_util import compare, gpu_step, make_gpu, make_oracle, oracle_step, reflection_tables
    from synth import make_problem
    P = make_problem(nz=3, nr=3, nst=1500, ngroups=64, split=(40, 20, 6, 10), n_e=3.0e5)
    P["cr_sent"] = cr_sent; P["tbbl"] = np.full(P["nr"], 0.5)
    with pytest.raises(C2dError) as e:                  # without the tables
        make_gpu(P)
    assert e.value.code == 2
    reflection_tables(P)
    o, g = make_oracle(P), make_gpu(P)
    for cyc in (1, 2):
        ro = oracle_step(o, P, cyc, field=cyc > 1)
        rg = gpu_step(g, P, cyc, field=cyc > 1)
        bad = compare(rg, ro, tol=1e-11, census_tol=1e-6)
        assert not bad, (cyc, bad)
    assert rg["counters"]["lkcount"] > 0 and rg["counters"]["cmcount2"] > 0
    if cr_sent == 3:
        assert rg["Ed_ref"].sum() > 0 and (rg["census"][0][:, 5] >= 1.0).sum() > 50      # reflected photons (>= 1 keV) in the census
    g.close()


def test_degenerate_inputs(emu, tmp_path):
    """Empty and ragged inputs: a step without sources, an empty census through the file formats, the event-list reduction of
    no records, an event buffer that overflows (records beyond event_capacity are counted, not stored: C2D_EEVENTS)."""
    from compton2d_b200 import C2dError
    from parity_util import gpu_step, make_gpu
    from synth import make_problem
    P = make_problem(nz=3, nr=2, nst=400, ngroups=16, split=(10, 5, 2, 10))
    P0 = dict(P, nsv=np.zeros_like(P["nsv"]))
    g = make_gpu(P0)
    t = gpu_step(g, P0, 1, field=True)                       # empty census re-tracked, no volume sources
    assert t["counters"]["histories"] == 0 and t["counters"]["ndxout"] == 0 and not t["edep"].any() and len(t["events"]) == 0
    assert g.write_cens(tmp_path / "empty.dat") == 0 and (tmp_path / "empty.dat").read_text() == ""
    g.read_cens(tmp_path / "empty.dat", 0)
    assert g.census_count() == 0
    assert g.write_events(tmp_path / "ev.dat", append=False) == 0
    red = dict(gam_bulk=10.0, rmax=1e16, t_offset=0.0, t_lo=np.arange(4.0), t_hi=np.arange(4.0) + 1, mu0=[-1.0], mu1=[1.0], mu_closed=1,
               E0=[1e-9], E1=[1e9], first_band_only=1)
    g.lc_begin(red)
    g.lc_accumulate()
    g.lc_accumulate_events(np.zeros((0, 7)))
    assert g.lc_read()[2].sum() == 0
    g.close()
    # event buffer overflow
    g = make_gpu(P, event_capacity=5)
    with pytest.raises(C2dError) as e:                       # gpu_step drains the events
        gpu_step(g, P, 1, field=False)
    assert e.value.code == 5 and g.tallies()["counters"]["nevents"] > 5
    g.lc_begin(red)
    with pytest.raises(C2dError) as e:                       # an incomplete list must not be binned silently
        g.lc_accumulate()
    assert e.value.code == 5 and g.lc_read()[2].sum() == 0
    g.close()
    # a census record whose indices would become out-of-bounds device addresses is refused
    g = make_gpu(P)
    d = np.ones((2, 6)); i = np.array([[1, 1, 1, 1, 1, 7], [1, 1, 1, P["nz"] + 1, 1, 7]], np.int32)
    with pytest.raises(C2dError) as e:
        g.census_import(d, i)
    assert e.value.code == 2 and "record 2" in str(e.value)
    g.close()


def test_closed_step_host_logic(emu):
    """c2d_step on the resident plasma (emu build, T_const = 1: no FP kernel there) against the oracle running the same
    phases in the reference's order for cycles 0-2, cycle 0 with the MPI build's quirk (dt = 0, no groups):
    device-built tables (volume_em), budget, bank carry-over, ec_old / bingo bookkeeping, cumulative fout."""
    from golden_util import Golden
    from loop_util import compare_plasma, gpu_closed, gpu_cycle, initial_state, oracle_cycle
    from parity_util import compare
    o, st, P, scal = initial_state(Golden("small"))
    scal = dict(scal, nst=1600)
    g = gpu_closed(P, st, scal, T_const=1)
    time, fout_total = 0.0, 0.0
    for cyc in range(3):
        ro = oracle_cycle(o, cyc, time, scal, T_const=1)
        rg = gpu_cycle(g, scal["dt"])
        assert rg["report"]["n_volume"] == int(ro["nsv"].sum()) and rg["report"]["n_census_in"] == ro["n_census_in"]
        bad = compare(rg, ro, tol=1e-10, census_tol=1e-6)
        assert not bad, (cyc, bad)
        assert not compare_plasma(rg["plasma"], ro["plasma"], after_fp=False)
        assert abs(rg["report"]["bingo"] / ro["report"]["bingo"] - 1) < 1e-12
        assert abs(rg["report"]["luminosity"] - ro["report"]["luminosity"]) <= 1e-10 * max(ro["report"]["luminosity"], 1e-300)
        fout_total = fout_total + ro["fout"]
        assert np.allclose(rg["fout_total"], fout_total, rtol=1e-10, atol=0)
        if cyc == 0:                                   # unflown, un-binned, one census photon per source photon
            assert rg["counters"]["ndxout"] == int(ro["nsv"].sum()) and not rg["census"][1][:, :3].any()
        if cyc > 0:
            time += scal["dt"]
    assert rg["counters"]["nevents"] > 10 and g.ncycle == 3 and abs(g.time - 2 * scal["dt"]) < 1e-9 * g.time
    g.close()


WORKER_STEP = r'''
import os, sys
sys.path[:0] = [r"{root}", r"{root}/tests", r"{root}/tests/emu", r"{root}/oracle"]
import numpy as np, torch, torch.distributed as dist
import build_emu
from compton2d_b200 import lib as L
L.LIBPATH = build_emu.build()
from golden_util import Golden
from loop_util import gpu_closed, initial_state
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
o, st, P, scal = initial_state(Golden("small"))
scal = dict(scal, nst=1600)
g = gpu_closed(P, st, scal, T_const=1, rank=rank, world=world)
reps = []
for cyc in range(3):
    reps.append(g.step(scal["dt"]))
t = g.tallies()
sid = g.census()[2]
gathered = [None] * world
dist.all_gather_object(gathered, sid)
if rank == 0:
    np.savez(r"{out}", edep=t["edep"], ecens=t["ecens"], n_field=t["n_field"], fout_total=t["fout_total"], npcen=t["npcen"],
             ndxout=reps[-1]["ndxout"], nevents=reps[-1]["nevents"], hist=reps[-1]["histories"], sid=np.sort(np.concatenate(gathered)),
             local_ndxout=t["counters"]["ndxout"], global_ndxout=t["counters_global"]["ndxout"])
dist.destroy_process_group()
'''


def test_closed_step_two_ranks_library_collectives(emu, tmp_path):
    """world_size 2 over gloo with the collectives INSIDE the library (c2d_tally_reduce through the transport handed to
    c2d_comm_init_ext; on the GPU the same calls go to NCCL): three closed steps on two ranks == on one rank."""
    from golden_util import Golden
    from loop_util import gpu_closed, initial_state
    o, st, P, scal = initial_state(Golden("small"))
    scal = dict(scal, nst=1600)
    g = gpu_closed(P, st, scal, T_const=1)
    for cyc in range(3):
        rep = g.step(scal["dt"])
    one = g.tallies()
    sid1 = np.sort(g.census()[2])
    g.close()
    out = tmp_path / "two.npz"
    script = tmp_path / "worker_step.py"
    script.write_text(WORKER_STEP.format(root=str(ROOT), out=str(out)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29543")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29543", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    two = np.load(out)
    assert np.array_equal(two["npcen"], one["npcen"]) and int(two["ndxout"]) == rep["ndxout"] and int(two["nevents"]) == rep["nevents"]
    assert int(two["hist"]) == rep["histories"] and np.array_equal(two["sid"], sid1)
    assert int(two["global_ndxout"]) == rep["ndxout"] and 0 < int(two["local_ndxout"]) < rep["ndxout"]
    for k in ("edep", "ecens", "n_field", "fout_total"):
        assert np.abs(two[k] - one[k]).max() <= 1e-12 * np.abs(one[k]).max(), k


def test_pref_calc_matches_reference_binary(emu):
    """c2d_pref_calc (k_pref_calc, csrc/c2d_gen.cu) == Pref_calc (ref_matrix.f:7-89): E_ref and the 500 x 500 reflection Green's
    function against COMMON /reflection/ of the unmodified binary (tests/golden/refbin_reflect.npz, snapshot c0_A).  ifort's pow
    and glibc's differ in the last bits: 1e-12."""
    from compton2d_b200.host import pref_calc
    g = np.load(ROOT / "tests" / "golden" / "refbin_reflect.npz")
    E, P = pref_calc(500)
    assert np.abs(E / g["c0_A/e_ref"] - 1.0).max() < 1e-13
    ref = g["c0_A/p_ref"]
    ref = ref if np.all(ref[0, 1:] == 1.0) else ref.T          # [n_in][n_out]: above the incoming energy the CDF is 1
    assert np.array_equal(P == 1.0, ref == 1.0) and np.array_equal(P == 0.0, ref == 0.0)
    assert np.abs(P - ref).max() < 1e-12
    assert (np.diff(P, axis=1) >= 0).all()
