"""BASELINE.json configs[3] and configs[4] as parity cases (the bench line is configs[1]).

configs[3]  64 x 64 zones x 200 photon energy groups -- beyond the reference build's nphomax = 128 (general.pa), so the
            comparison is against the CPU oracle only, at a few thousand histories (the full 1e9 histories / 8 GPUs is
            a matter of capacity, exercised by `bench.py --config c3`).
configs[4]  64 x 64 zones x 512 electron momentum bins (the reference is compiled for num_nt = 200): the FP kernel and the
            IC_loss table on the finer grid gamma - 1 = 0.2 q^(i-2), q = (1.1^198)^(1/510), which spans the reference's range
            (SURVEY.md 8d), against the oracle's FP_calc restatement (itself pinned to the binary at 200 bins).  The
            per-zone state is the reference binary's own pre-FP state (snapshot c2_E of refbin_small, zone (1,1)) resampled
            onto the finer grid, photon field jittered per zone, cold and hot zones mixed so that the temperature walk
            goes both ways."""
import numpy as np
import pytest

import orc
from golden_util import Golden
from oracle_golden import SCALARS
from parity_util import compare, gpu_step, make_gpu, make_oracle, oracle_step
from synth import make_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("sigma_mode", [0, 1])
def test_c3_shape_replay_parity(sigma_mode):
    P = make_problem(nz=64, nr=64, nst=8000, ngroups=200, jitter=0.3)
    o, g = make_oracle(P), make_gpu(P, sigma_mode=sigma_mode)
    try:
        for cyc in (1, 2):
            ro = oracle_step(o, P, cyc, field=cyc > 1)
            rg = gpu_step(g, P, cyc, field=cyc > 1)
            bad = compare(rg, ro, tol=1e-10, census_tol=1e-5)
            assert not bad, (cyc, bad)
        assert rg["counters"]["histories"] > 7000 and rg["fout"].shape[1] == 200
    finally:
        g.close()


def _fine_grid(n=512):
    q = (1.1 ** 198) ** (1.0 / (n - 2))
    g = 0.2 * q ** (np.arange(1, n + 1) - 2.0)
    return g


@pytest.mark.parametrize("N", [512, 128])       # 128 bins: nphfield > 2 N, the dg_ic loop without its shared-memory staging
def test_c4_shape_fp_parity(N):
    nz, nr = 8, 8                              # 64 of the 4096 zones: the oracle needs ~0.2 s per zone at 512 bins
    E = Golden("small").snap("c2_E")
    gnt = _fine_grid(N)
    P = make_problem(nz=nz, nr=nr, nst=100, ngroups=32)
    nzone = nz * nr
    # resample the binary's electron spectrum (zone (1,1)) onto the fine grid, renormalise, rebuild the running CDF
    f200 = np.maximum(np.asarray(E.f_nt[0, 0], float), 1e-300)
    f = np.exp(np.interp(np.log(gnt), np.log(np.asarray(E.gnt)), np.log(f200)))
    f[f < 1e-250] = 0.0; f[-1] = 0.0
    f = f / (f[:-1] * np.diff(gnt)).sum()
    Pnt = np.concatenate(([0.0], np.cumsum(f[:-1] * np.diff(gnt))))
    rng = np.random.default_rng(11)
    vr = P["vol"] / float(E.vol[0, 0]); tr = P["dt"] / float(E.dt[0])
    n_field = np.tile(np.asarray(E.n_field[:, 0, 0], float), (nzone, 1)) * vr[:, None] * (1.0 + 0.4 * (rng.random((nzone, 1)) - 0.5))
    n_field[::3] *= 3.0e3                       # strong photon field: inverse-Compton cooling shapes f_new
    tea = np.full(nzone, float(E.tea[0, 0])); tea[::4] = 8.0; tea[1::4] = 400.0
    zone = dict(f_nt=np.tile(f, (nzone, 1)), Pnt=np.tile(Pnt, (nzone, 1)), n_e=np.full(nzone, float(E.n_e[0, 0])),
                B_field=np.full(nzone, float(E.b_field[0, 0])), tea=tea, tna=np.full(nzone, float(E.tna[0, 0])), vol=P["vol"],
                Eloss_sy=float(E.eloss_sy[0, 0]) * vr * tr, ec_old=float(E.ec_old[0, 0]) * vr,
                turb_lev=np.full(nzone, float(E.turb_lev[0, 0])), gmin=np.full(nzone, float(E.gmin[0, 0])),
                gmax=np.full(nzone, float(E.gmax[0, 0])), p_nth=np.full(nzone, float(E.p_nth[0, 0])),
                amxwl=np.full(nzone, float(E.amxwl[0, 0])), ecens=float(E.ecens[0, 0]) * vr, n_field=n_field)
    # ---- oracle
    o = orc.Oracle(nz, nr, num_nt=N, nphtotal=32)
    a = o.a
    a.z = P["z"]; a.r = P["r"]; a.gnt = gnt; a.E_field = P["E_field"]
    for g_, c_ in SCALARS.items():
        if c_ not in ("rmin", "zmin", "time", "ncycle", "split1", "split2", "split3", "spl3_trg", "pair_switch", "cr_sent",
                      "spec_switch"):
            o[c_] = float(E[g_])
    o["rmin"] = P["rmin"]; o["zmin"] = P["zmin"]; o["pair_switch"] = 0; o["dt"] = P["dt"]; o["time"] = 2.0 * P["dt"]; o["ncycle"] = 2
    for k_, v in zone.items():
        getattr(a, k_)[: np.size(v)] = np.asarray(v, float).ravel()
    a.f_pair[:] = 0.0; a.q_turb[:] = float(E.q_turb[0, 0])
    for nm in ("n_pos", "dn_pp", "dne_pa", "dnp_pa"):
        getattr(a, nm)[:] = 0.0
    o.L.orc_IC_loss(o.p)
    F_IC = a.F_IC[: N * 400].reshape(N, 400).copy()
    for s_ in ("E_tot_old", "E_tot_new", "hr_total", "hr_st_total", "dT_max"):
        o[s_] = 0.0
    for z_ in range(1, nzone + 1):
        o.fp_calc(z_)
    ref = dict(f_nt=a.f_nt[: nzone * N].reshape(nzone, N).copy(), Pnt=a.Pnt[: nzone * N].reshape(nzone, N).copy(),
               n_e=a.n_e[:nzone].copy(), Te_new=a.Te_new[:nzone].copy(), gmin=a.gmin[:nzone].copy(), gmax=a.gmax[:nzone].copy(),
               p_nth=a.p_nth[:nzone].copy(), amxwl=a.amxwl[:nzone].copy(), fp_steps=a.fp_steps[:nzone].copy())
    # ---- CUDA
    P["gnt"] = gnt; P["f_nt"] = zone["f_nt"]; P["Pnt"] = zone["Pnt"]
    gpu = make_gpu(P, census_capacity=1000, event_capacity=1000)
    try:
        fic = gpu.IC_loss()
        assert np.abs(fic / F_IC - 1.0).max() < 1e-8
        inj = {k: float(E[k.lower()]) for k in ("inj_g1", "inj_g2", "inj_p", "inj_t", "inj_gg", "inj_sigma", "inj_L", "inj_v", "pick_rate")}
        inj.update({k: int(E[k]) for k in ("inj_switch", "inj_dis", "g2var_switch", "pick_sw")})
        out = gpu.update(time=2.0 * P["dt"], dt=P["dt"], F_IC=F_IC, lnL=float(E.lnl), r_esc=float(E.r_esc), r_acc=float(E.r_acc),
                         inj=inj, **zone)
    finally:
        gpu.close()
    assert np.array_equal(out["fp_steps"], ref["fp_steps"]) and ref["fp_steps"].min() >= 1
    for k_ in ("gmin", "gmax", "p_nth", "Te_new"):
        assert np.array_equal(out[k_], ref[k_]), (k_, out[k_][:6], ref[k_][:6])
    for k_ in ("f_nt", "Pnt"):
        err = np.abs(out[k_] - ref[k_]).max(axis=1) / np.abs(ref[k_]).max(axis=1)
        assert err.max() < 1e-9, (k_, err.max())
    assert np.abs(out["n_e"] - ref["n_e"]).max() <= 1e-12 * np.abs(ref["n_e"]).max()
    assert len(set(ref["Te_new"])) > 2          # the zones really differ
