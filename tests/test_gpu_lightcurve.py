"""Event-list reduction on the device (csrc/c2d_lc.cu: k_lc_reduce) through the C ABI:
 * against what the reference's own post-processors wrote for the event lists of the reference-binary golden runs
   (tests/golden/plcm_*.npz, pspt_*.npz: plcm / pspt compiled unmodified from postprocessing/*.c) -- photon counts per
   (time, mu, band) cell bit-exact, luminosities to the 7 digits the programs print;
 * against the oracle's restatement (oracle/c2d_oracle.c: orc_events_reduce) on the escape records of tracked steps and on
   a large synthetic list -- counts bit-exact, FP64 sums to 1e-12 (different summation order);
 * both accumulator placements of the kernel (per-block shared-memory copy / straight to HBM)."""
import numpy as np
import pytest

import orc
from compton2d_b200.lightcurve import lc_reduction, sed_reduction
from lc_util import check_against_plcm, check_against_pspt, load_case, load_sed_case
from parity_util import gpu_step, make_gpu
from synth import make_problem

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    P = make_problem(nz=4, nr=3, nst=20000, ngroups=64, split=(40, 20, 6, 10), n_e=3.0e5)
    g = make_gpu(P)
    yield g, P
    g.close()


@pytest.mark.parametrize("name", ["small", "dense"])
@pytest.mark.parametrize("st_max", [64, 1024])          # 64 time bins: shared-memory accumulators; 1024 (plcm's): HBM
def test_light_curve_matches_reference_plcm(ctx, name, st_max):
    g, _ = ctx
    gold, cfg, events = load_case(name)
    g.lc_begin(lc_reduction(cfg, st_max=st_max))
    g.lc_accumulate_events(events[:777]); g.lc_accumulate_events(events[777:])      # odd split: unaligned tail chunk
    check_against_plcm(gold, cfg, *g.lc_read())


@pytest.mark.parametrize("name", ["small", "dense"])
def test_sed_matches_reference_pspt(ctx, name):
    g, _ = ctx
    gold, cfg, events = load_sed_case(name)
    g.lc_begin(sed_reduction(cfg))
    g.lc_accumulate_events(events)
    s1, s2, cnt = g.lc_read()
    check_against_pspt(gold, cfg, s1, cnt)
    o1, o2, oc = orc.events_reduce(events, sed_reduction(cfg))
    assert np.array_equal(cnt, oc)
    assert np.allclose(s1, o1, rtol=1e-12, atol=0) and np.allclose(s2, o2, rtol=1e-12, atol=0)


def _wide_bins(P, n_t, n_bands):
    dtb = 12.0 * P["dt"] / n_t
    edges = np.exp(np.linspace(np.log(1e-9), np.log(1e6), n_bands + 1))
    return dict(gam_bulk=10.0, rmax=float(P["r"][-1]), t_offset=0.0, t_lo=np.arange(n_t) * dtb, t_hi=(np.arange(n_t) + 1) * dtb,
                mu0=np.array([-1.0, 0.5, 0.97]), mu1=np.array([0.5, 0.97, 1.0]), mu_closed=0, E0=edges[:-1].copy(),
                E1=edges[1:] * 1.5, first_band_only=0)      # bands overlap their upper neighbour


@pytest.mark.parametrize("n_t,n_bands", [(40, 12), (900, 20)])
def test_resident_records_of_tracked_steps(ctx, n_t, n_bands):
    """The records are binned where imcleak left them (no drain); accumulators persist over steps."""
    g, P = ctx
    red = _wide_bins(P, n_t, n_bands)
    g.lc_begin(red)
    ev_all = []
    for cyc in (1, 2, 3):
        t = gpu_step(g, P, cyc, field=cyc > 1)
        g.lc_accumulate()
        ev_all.append(t["events"])
    ev_all = np.concatenate(ev_all)
    assert len(ev_all) > 2000
    o1, o2, oc = orc.events_reduce(ev_all, red)
    s1, s2, cnt = g.lc_read()
    assert np.array_equal(cnt, oc) and cnt.sum() > 2000
    assert np.allclose(s1, o1, rtol=1e-12, atol=0) and np.allclose(s2, o2, rtol=1e-12, atol=0)


@pytest.mark.parametrize("first_band_only,mu_closed", [(0, 0), (1, 1)])
def test_large_synthetic_list(ctx, first_band_only, mu_closed):
    """2e6 + 1 records (odd: the last 16-byte chunk is half a record), heavy cell contention."""
    g, P = ctx
    rng = np.random.default_rng(5)
    n = 2_000_001
    ev = np.stack([rng.uniform(0, 3e6, n), np.exp(rng.uniform(np.log(1e-8), np.log(1e3), n)), rng.uniform(1e30, 2e30, n),
                   0.75e16 * np.sqrt(rng.random(n)), 1e16 * rng.random(n), rng.uniform(-1, 1, n), np.pi * rng.random(n)], axis=1)
    edges = np.exp(np.linspace(np.log(1e-6), np.log(1e4), 31))
    red = dict(gam_bulk=33.0, rmax=0.75e16, t_offset=1.0e3, t_lo=2e3 * np.arange(50), t_hi=2e3 * np.arange(50) + 2e3,
               mu0=np.array([0.9990, 0.9995]), mu1=np.array([0.9995, 1.0]), mu_closed=mu_closed, E0=edges[:-1].copy(),
               E1=edges[1:] * 2.0, first_band_only=first_band_only)
    g.lc_begin(red)
    g.lc_accumulate_events(ev)
    o1, o2, oc = orc.events_reduce(ev, red)
    s1, s2, cnt = g.lc_read()
    assert oc.sum() > 10000
    assert np.array_equal(cnt, oc)
    assert np.allclose(s1, o1, rtol=1e-12, atol=0) and np.allclose(s2, o2, rtol=1e-12, atol=0)


def test_bench_hook(ctx):
    g, _ = ctx
    gold, cfg, _ = load_case("small")
    g.lc_begin(lc_reduction(cfg, st_max=128))
    ms, nbytes = g.lc_bench(1 << 22)
    assert ms > 0 and nbytes == 56 << 22
    assert g.lc_read()[2].sum() > 0


def test_two_reducers_fed_by_the_same_records(ctx):
    """plcm's light curve (slot 0) and pspt's SED (slot 1) from one pass over the records, as the reference runs both programs on
    the same event files."""
    g, _ = ctx
    gold_lc, cfg_lc, events = load_case("dense")
    gold_sed, cfg_sed, _ = load_sed_case("dense")
    g.lc_begin(lc_reduction(cfg_lc), slot=0); g.lc_begin(sed_reduction(cfg_sed), slot=1)
    g.lc_accumulate_events(events)
    check_against_plcm(gold_lc, cfg_lc, *g.lc_read(slot=0))
    s1, s2, cnt = g.lc_read(slot=1)
    check_against_pspt(gold_sed, cfg_sed, s1, cnt)
