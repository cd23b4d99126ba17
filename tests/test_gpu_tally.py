"""Stand-alone tally-commit kernels (csrc/c2d_tally.cu, c2d_tally_bench): the device generates the event list from a fixed
integer hash; numpy rebuilds it and the tallies must agree -- counts exactly, FP64 sums to 1e-12 (different summation
order: warp butterfly -> shared-memory bins -> one HBM atomic per block and bin)."""
import numpy as np
import pytest

from parity_util import make_gpu
from synth import make_problem

pytestmark = pytest.mark.gpu


def _events(n, nzone, nf, sorted_=False):
    i = np.arange(n, dtype=np.uint64)
    h = ((i * np.uint64(0x9E3779B97F4A7C15)) >> np.uint64(33)).astype(np.uint32)
    zone = (i * np.uint64(nzone) // np.uint64(n)).astype(np.int64) if sorted_ else (h % np.uint32(nzone)).astype(np.int64)
    de = 1.0 + 1.0e-3 * (h & np.uint32(1023)).astype(float)
    return zone, de, 0.5 * de, ((h >> np.uint32(10)) % np.uint32(nf)).astype(np.int64)


def _begin(g, P):
    g.imcgen2d(ncycle=1, time=0.0, dt=P["dt"], kappa_tot=P["kappa_tot"], eps_tot=P["eps_tot"], eps_th=P["eps_th"],
               f_nt=P["f_nt"], Pnt=P["Pnt"], n_e=P["n_e"], tea=P["tea"], f_pair=P["f_pair"], k_gg=P["k_gg"], nsv=P["nsv"],
               ewsv=P["ewsv"], Eloss_th=P["Eloss_th"], Eloss_tot=P["Eloss_tot"], zsurf=P["zsurf"], tbbl=P["tbbl"])


@pytest.mark.parametrize("kind", [0, 0x10, 0x20])
def test_deposition_commit(kind):
    P = make_problem(nz=16, nr=16, nst=100, ngroups=32)
    g = make_gpu(P, census_capacity=1000)
    try:
        _begin(g, P)
        n = 1_000_003
        ms, nb = g.tally_bench(n, kind)
        assert nb == 56 * n and ms > 0
        t = g.tallies()
        zone, de, dp, _ = _events(n, P["nzone"], P["E_field"].size, sorted_=bool(kind & 0x20))
        assert np.allclose(t["edep"], np.bincount(zone, de, P["nzone"]), rtol=1e-12, atol=0)
        assert np.allclose(t["prdep"], np.bincount(zone, dp, P["nzone"]), rtol=1e-12, atol=0)
    finally:
        g.close()


def test_census_commit():
    P = make_problem(nz=8, nr=8, nst=100, ngroups=32)
    n = 300_007
    g = make_gpu(P, census_capacity=n + 10)
    try:
        _begin(g, P)
        ms, nb = g.tally_bench(n, 1)
        assert nb == 112 * n
        t = g.tallies()
        zone, de, dp, b = _events(n, P["nzone"], P["E_field"].size)
        assert np.array_equal(t["npcen"], np.bincount(zone, minlength=P["nzone"]))
        assert np.allclose(t["ecens"], np.bincount(zone, de, P["nzone"]), rtol=1e-12, atol=0)
        nf = P["E_field"].size
        ref = np.bincount(zone * nf + b, 6.25e8 * de / dp, P["nzone"] * nf).reshape(P["nzone"], nf)
        assert np.allclose(t["n_field"], ref, rtol=1e-12, atol=0)
    finally:
        g.close()
