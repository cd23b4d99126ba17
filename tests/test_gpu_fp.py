"""GPU parity of the batched Fokker-Planck kernel (c2d_update_fp) against the CPU oracle's FP_calc, which is
itself pinned to the reference binary (tests/test_oracle_vs_refbin.py).

Tolerance (north_star: "FP electron distributions within a stated relative tolerance in FP64"): f_nt and Pnt
within 1e-9 of each zone's maximum, n_e / amxwl 1e-10, energies 1e-9; gmin, gmax, p_nth and Te_new are lattice
values chosen by comparisons and must be identical.  The device sums bins lane-strided (different order
from the serial loops) and its exp/log/pow differ from glibc's by <= 2 ulp, hence not bit-exact."""
import numpy as np
import pytest

from golden_util import Golden
from oracle_golden import cname, load_snapshot, new_oracle
from parity_util import make_gpu, problem_from_snapshot

pytestmark = pytest.mark.gpu

INJ = ("inj_switch", "inj_dis", "g2var_switch", "pick_sw", "inj_g1", "inj_g2", "inj_p", "inj_t", "inj_gg", "inj_sigma",
       "inj_L", "inj_v", "pick_rate")


def _oracle_fp(E, F, mutate=None):
    o = new_oracle(E)
    load_snapshot(o, F)
    a = o.a
    for v in ("f_nt", "pnt", "n_e", "amxwl", "gmin", "gmax", "p_nth", "te_new", "f_pair"):
        setattr(a, cname(v), E[v])
    if mutate:
        mutate(o)
    for s_ in ("E_tot_old", "E_tot_new", "hr_total", "hr_st_total", "dT_max"):
        o[s_] = 0.0
    nz, nr = o.dims["nz"], o.dims["nr"]
    nzone = nz * nr
    inp = dict(f_nt=a.f_nt[: nzone * 200].copy(), Pnt=a.Pnt[: nzone * 200].copy(), n_e=a.n_e[:nzone].copy(),
               B_field=a.B_field[:nzone].copy(), tea=a.tea[:nzone].copy(), tna=a.tna[:nzone].copy(), vol=a.vol[:nzone].copy(),
               Eloss_sy=a.Eloss_sy[:nzone].copy(), ec_old=a.ec_old[:nzone].copy(), turb_lev=a.turb_lev[:nzone].copy(),
               f_pair=a.f_pair[:nzone].copy(), n_field=a.n_field[: nzone * 400].copy(), ecens=a.ecens[:nzone].copy(),
               F_IC=a.F_IC[: 200 * 400].copy(), gmin=a.gmin[:nzone].copy(), gmax=a.gmax[:nzone].copy(),
               p_nth=a.p_nth[:nzone].copy(), amxwl=a.amxwl[:nzone].copy(), time=o["time"], dt=o["dt"], lnL=o["lnL"],
               r_esc=o["r_esc"], r_acc=o["r_acc"], inj={k: o[k] for k in INJ})
    for zone in range(1, nzone + 1):
        o.fp_calc(zone)
    ref = dict(f_nt=a.f_nt[: nzone * 200].reshape(nzone, 200).copy(), Pnt=a.Pnt[: nzone * 200].reshape(nzone, 200).copy(),
               n_e=a.n_e[:nzone].copy(), Te_new=a.Te_new[:nzone].copy(), gmin=a.gmin[:nzone].copy(), gmax=a.gmax[:nzone].copy(),
               p_nth=a.p_nth[:nzone].copy(), amxwl=a.amxwl[:nzone].copy(), fp_steps=a.fp_steps[:nzone].copy(),
               E_tot_old=o["E_tot_old"], E_tot_new=o["E_tot_new"], hr_total=o["hr_total"], hr_st_total=o["hr_st_total"],
               dT_max=o["dT_max"])
    return inp, ref


def _check(out, ref):
    assert np.array_equal(out["fp_steps"], ref["fp_steps"])
    for k in ("gmin", "gmax", "p_nth", "Te_new"):
        assert np.array_equal(out[k], ref[k]), (k, out[k][:4], ref[k][:4])
    for k, tol in (("f_nt", 1e-9), ("Pnt", 1e-9)):
        err = np.abs(out[k] - ref[k]).max(axis=1) / np.abs(ref[k]).max(axis=1)
        assert err.max() < tol, (k, err.max())
    for k, tol in (("n_e", 1e-12), ("amxwl", 1e-9)):
        assert np.abs(out[k] - ref[k]).max() <= tol * np.abs(ref[k]).max(), k
    for k in ("E_tot_old", "E_tot_new", "hr_total", "hr_st_total", "dT_max"):
        assert abs(out[k] - ref[k]) <= 1e-9 * abs(ref[k]), (k, out[k], ref[k])


@pytest.mark.parametrize("name,cyc", [("small", 1), ("small", 2), ("annulus", 2)])
def test_fp_parity_reference_state(name, cyc):
    """The reference binary's own pre-FP state (tallied n_field of a real tracking step, shipped injection
    block: pick-up Gaussian on, shock power-law injection armed)."""
    g = Golden(name)
    E, F = g.snap(f"c{cyc}_E"), g.snap(f"c{cyc}_F")
    inp, ref = _oracle_fp(E, F)
    P = problem_from_snapshot(g.snap(f"c{cyc}_B"), g.snap(f"c{cyc}_C"))
    gpu = make_gpu(P)
    try:
        out = gpu.update(**inp)
    finally:
        gpu.close()
    _check(out, ref)


def test_fp_parity_shock_injection_and_cooling():
    """Shock front inside the grid (inj_t = 0 so that row j = 1 is being injected), strong photon field so that
    inverse-Compton cooling shapes f_new, cold thermal zones so that the temperature walk goes down."""
    g = Golden("small")
    E, F = g.snap("c2_E"), g.snap("c2_F")

    def mutate(o):
        o["inj_t"] = 0.0
        o["time"] = 0.3 * o["dt"]
        o.a.n_field[:] *= 3.0e3
        o.a.tea[:3] = 8.0

    inp, ref = _oracle_fp(E, F, mutate)
    P = problem_from_snapshot(g.snap("c2_B"), g.snap("c2_C"))
    gpu = make_gpu(P)
    try:
        out = gpu.update(**inp)
    finally:
        gpu.close()
    _check(out, ref)


def test_ic_loss_device_table():
    """c2d_ic_loss (icloss2d.f) against the table the reference binary built."""
    g = Golden("small")
    F = g.snap("c1_F")
    P = problem_from_snapshot(g.snap("c1_B"), g.snap("c1_C"))
    gpu = make_gpu(P)
    try:
        fic = gpu.IC_loss()
    finally:
        gpu.close()
    assert np.abs(fic / F.f_ic - 1).max() < 1e-8


def test_gamma_bar_and_predictor_table():
    """gamma_bar on the device (lane-parallel McDonald quadratures) against the oracle's scalar loop, and the interpolation
    table that predicts the temperature walk against the quadrature: the walk trusts the table only where it clears gbar by
    1e-5 (kGtabMargin, csrc/c2d_fp.cu), so its error must stay well below that (measured: 3e-9 typical, 5e-8 worst) -- across the fit/quadrature switch at
    Theta = 0.2 too."""
    import orc
    from synth import make_problem
    P = make_problem(nz=2, nr=2, nst=100, ngroups=16)
    g = make_gpu(P, census_capacity=1000, event_capacity=1000)
    rng = np.random.default_rng(3)
    th = np.concatenate([np.exp(rng.uniform(np.log(0.01), np.log(3.0e4), 400)), 0.2 * 1.005 ** np.arange(-6, 30),
                         [0.19999, 0.2, 0.20001, 1.0, 1.957, 100.0]])
    ge, gt = g.gamma_bar(th)
    ref = np.array([orc.lib().orc_gamma_bar(float(t)) for t in th])
    assert np.abs(ge / ref - 1).max() < 1e-11
    inside = gt > 0
    err = np.abs(gt[inside] / ge[inside] - 1)
    assert inside.sum() > 400 and err.max() < 2e-7, (err.max(), th[inside][err.argmax()])
    print("predictor table: max rel err %.2e at Theta = %.4g, median %.2e" % (err.max(), th[inside][err.argmax()], np.median(err)))
    assert np.all(th[~inside] > 1.0e4)
    g.close()


def test_fp_zone_partition_equals_whole():
    """The multi-GPU partition of the FP farm (zone % world == rank, c2d_fp_in.first_zone / zone_stride): two contexts that each
    update their half, merged the way _exchange_fp merges them, equal the single-context update bit for bit."""
    from compton2d_b200.host import Compton2d
    g = Golden("small")
    E, F = g.snap("c2_E"), g.snap("c2_F")
    inp, _ = _oracle_fp(E, F)
    P = problem_from_snapshot(g.snap("c2_B"), g.snap("c2_C"))
    whole = make_gpu(P)
    try:
        ref = whole.update(**inp)
    finally:
        whole.close()
    nzone = P["nz"] * P["nr"]
    parts = []
    for rank in (0, 1):
        gpu = make_gpu(P, rank=rank, world=2)
        gpu._exchange_fp = lambda out, group=None: None            # the exchange itself is covered by the gloo test
        try:
            parts.append(gpu.update(**inp))
        finally:
            gpu.close()
    for k in ("f_nt", "Pnt", "n_e", "Te_new", "gmin", "gmax", "p_nth", "amxwl", "fp_steps", "zone_E_old", "zone_E_new"):
        merged = np.array(parts[0][k]).copy().reshape(nzone, -1)
        merged[1::2] = np.array(parts[1][k]).reshape(nzone, -1)[1::2]
        assert np.array_equal(merged, np.array(ref[k]).reshape(nzone, -1)), k
    assert parts[0]["fp_steps"][1::2].sum() == 0 and parts[1]["fp_steps"][0::2].sum() == 0      # foreign zones untouched
