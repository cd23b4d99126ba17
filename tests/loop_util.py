"""The closed IMC time step (xec2d.f:100-110 / 141-193) run twice: by the CPU oracle, phase by phase in the reference's
order, and by the library's c2d_step on its device-resident plasma.  Shared by the GPU replay test, the CPU host-logic test
(emu build, T_const = 1: the FP kernel is warp-cooperative and not part of that build) and the shipped-input test.

Order of one cycle (master + worker views merged):
    ec_old <- ecens                      imcgen2d.f:90-96
    volume_em, Eloss_*, nsv, ewsv        imcgen2d.f:203-336,405-459   (orc_imcgen_volume)
    zero tallies                         imcgen2d.f:87-104,534-551,559-601
    field_calc, vol_calc (all zones)     imcfield2d.f:57-148, imcvol2d_para.f:90-414
    FP_calc (all zones), tea clamp       update2d.f:337-1741, 262-276   [T_const = 0 and ncycle > 0]
quirk_cycle0: in cycle 0 the reference's workers track with dt = 0 and no energy groups."""
import numpy as np

from oracle_golden import cname, load_snapshot, new_oracle
from parity_util import NLC, make_gpu, problem_from_snapshot

INJ = ("inj_switch", "inj_dis", "g2var_switch", "pick_sw", "inj_g1", "inj_g2", "inj_p", "inj_t", "inj_gg", "inj_sigma",
       "inj_L", "inj_v", "pick_rate")
PLASMA = ("f_nt", "Pnt", "n_e", "tea", "tna", "B_field", "f_pair", "gmin", "gmax", "p_nth", "amxwl", "turb_lev", "vol", "zsurf")


def initial_state(g, cens_cap=400000, ev_cap=400000):
    """Oracle holding the plasma the reference binary starts cycle 1's FP farm with (== its initial plasma: no update has
    run yet) plus every run scalar, and that plasma as dense arrays for c2d_plasma_set."""
    # (snapshot F -- after the FP farm, where the per-zone job messages have delivered the master's tna, B_field, turb_lev,
    # ... -- exists only in goldens whose trace matched make_golden.py's F signature; E holds the same values except
    # turb_lev, which reaches the worker with the first job message: taken from cycle 2's E)
    E = g.snap("c1_E")
    F = g.snap("c1_F") if "c1_F" in g.labels else E
    o = new_oracle(E, cens_cap=cens_cap, ev_cap=ev_cap)
    load_snapshot(o, F)
    a = o.a
    if F is E:
        a.turb_lev = g.snap("c2_E").turb_lev
        a.F_IC[:] = 0.0
        o.L.orc_IC_loss(o.p)                     # the worker has not received F_IC yet at E: icloss2d.f restated (pinned to 1e-8)
    for v in ("f_nt", "pnt", "n_e", "amxwl", "gmin", "gmax", "p_nth", "te_new", "f_pair"):
        setattr(a, cname(v), E[v])
    nzone = o.dims["nzone"]
    N = o.dims["num_nt"]
    st = {}
    for k in PLASMA:
        n = nzone * N if k in ("f_nt", "Pnt") else nzone
        st[k] = getattr(a, k)[:n].copy()
    o["rand_switch"] = 3
    o.rseed = int(E.rseed)
    o["ndxout"] = 0
    a.ecens[:] = 0.0
    a.fout[:] = 0.0
    P = problem_from_snapshot(g.snap("c1_B"), g.snap("c1_C"))
    scal = dict(dt=float(F.dt[0]), nst=int(F.nst), lnL=float(o["lnL"]), r_esc=float(o["r_esc"]), r_acc=float(o["r_acc"]),
                inj={k: o[k] for k in INJ}, T_const=int(F.t_const), rseed=int(E.rseed))
    return o, st, P, scal


def oracle_cycle(o, ncycle, time, scal, quirk_cycle0=True, T_const=0):
    """One cycle on the oracle state `o` (which carries the plasma from cycle to cycle).  -> dict like parity_util.oracle_step
    plus the plasma after the update and the scalars of the report."""
    a = o.a
    d = o.dims
    nz, nr, nzone, N = d["nz"], d["nr"], d["nzone"], d["num_nt"]
    dt = scal["dt"]
    a.ec_old[:nzone] = a.ecens[:nzone] if ncycle > 0 else 0.0
    ecens_prev = a.ecens[:nzone].copy() if ncycle > 0 else np.zeros(nzone)
    o["dt"] = dt; o["time"] = time; o["ncycle"] = ncycle
    o.imcgen_volume(scal["nst"])
    bingo = 0.0
    for zi in range(nzone):
        bingo = bingo + 1.0 * (ecens_prev[zi] + a.Eloss_tot[zi])
    o.philox_key(o.rseed, ncycle)
    o.step_zero()
    a.fout[:] = 0; a.prdep[:] = 0; a.n_ph[:] = 0; a.nelectron[:] = 0
    o["nev"] = 0; o["n_multiple"] = 0; o["n_killed"] = 0; o["n_trials"] = 0
    saved = None
    if quirk_cycle0 and ncycle == 0:
        saved = (o["nphtotal"], o["nph_lc"], o["nmu"])
        o["dt"] = 0.0; o["nphtotal"] = 0; o["nph_lc"] = 0; o["nmu"] = 0
    n_in = int(o["ndxout"])
    if n_in > 0:
        o.field_calc()
    else:
        o["ndxout"] = 0
    o.vol_all()
    if saved:
        o["dt"] = dt; o["nphtotal"], o["nph_lc"], o["nmu"] = saved
    assert int(o["fatal"]) == 0
    npt = d["nphtotal"]
    cd, ci, cs = o.census()
    res = dict(edep=a.edep[:nzone].copy(), prdep=a.prdep[:nzone].copy(), ecens=a.ecens[:nzone].copy(),
               npcen=a.npcen[:nzone].copy(), n_field=a.n_field[: nzone * d["nphfield"]].reshape(nzone, -1).copy(),
               n_ph=a.n_ph[: nzone * d["n_gg"]].reshape(nzone, -1).copy(), E_IC=a.E_IC[:N].copy(),
               nelectron=a.nelectron[: N + 2].copy(), fout=a.fout[: 32 * npt].reshape(32, npt).copy(),
               edout=a.edout[: 32 * NLC].reshape(32, NLC).copy(), erlki=a.erlki[:nz].copy(), erlko=a.erlko[:nz].copy(),
               erlku=a.erlku[:nr].copy(), erlkl=a.erlkl[:nr].copy(), Ed_in=a.Ed_in[:nr].copy(),
               counters=dict(ndxout=int(o["ndxout"]), cmcount1=int(o["cmcount1"]), cmcount2=int(o["cmcount2"]),
                             lkcount=int(o["lkcount"]), nevents=int(o["nev"]), multiple=int(o["n_multiple"]),
                             killed=int(o["n_killed"]), trials=int(o["n_trials"])),
               census=(cd, ci, cs), events=o.events(), n_census_in=n_in,
               nsv=a.nsv[:nzone].copy(), Eloss_sy=a.Eloss_sy[:nzone].copy())
    E_out = 0.0
    for j in range(nz):
        E_out = E_out + a.erlki[j] + a.erlko[j]
    for k in range(nr):
        E_out = E_out + a.erlku[k] + a.erlkl[k]
    rep = dict(bingo=bingo, luminosity=E_out / dt, fp_ran=0, E_tot_old=0.0, E_tot_new=E_out)
    if T_const == 0 and ncycle > 0:
        for s_ in ("E_tot_old", "E_tot_new", "hr_total", "hr_st_total", "dT_max"):
            o[s_] = 0.0
        for zone in range(1, nzone + 1):
            o.fp_calc(zone)
        m = a.tna[:nzone] > 1.0
        a.tea[:nzone][m] = np.minimum(1.0e3, np.maximum(5.0, a.Te_new[:nzone][m]))
        dT = float(o["dT_max"])
        if ncycle <= 1:
            dT = max(dT, 0.25)
        Eo, En = float(o["E_tot_old"]), E_out + float(o["E_tot_new"])
        rep.update(fp_ran=1, E_tot_old=Eo, E_tot_new=En, dE_fraction=(En - Eo) / Eo, dT_max=dT, hr_total=float(o["hr_total"]),
                   hr_st_total=float(o["hr_st_total"]), fp_steps=a.fp_steps[:nzone].copy())
    res["report"] = rep
    res["plasma"] = dict(f_nt=a.f_nt[: nzone * N].reshape(nzone, N).copy(), Pnt=a.Pnt[: nzone * N].reshape(nzone, N).copy(),
                         n_e=a.n_e[:nzone].copy(), tea=a.tea[:nzone].copy(), gmin=a.gmin[:nzone].copy(), gmax=a.gmax[:nzone].copy(),
                         p_nth=a.p_nth[:nzone].copy(), amxwl=a.amxwl[:nzone].copy(), Te_new=a.Te_new[:nzone].copy())
    return res


def gpu_closed(P, st, scal, quirk_cycle0=True, T_const=0, sigma_mode=0, **kw):
    """A context with the plasma uploaded and the run configured, ready for g.step(dt)."""
    g = make_gpu(P, rseed=scal["rseed"], sigma_mode=sigma_mode, **kw)
    g.plasma_set(**st)
    g.run_config(nst=scal["nst"], T_const=T_const, quirk_cycle0=int(quirk_cycle0), inj=scal["inj"], r_esc=scal["r_esc"],
                 r_acc=scal["r_acc"])
    return g


def gpu_cycle(g, dt):
    rep = g.step(dt)
    t = g.tallies()
    t["census"] = g.census()
    t["events"] = g.events()
    t["report"] = rep
    t["plasma"] = g.plasma_get()
    return t


def compare_plasma(pg, po, after_fp):
    """FP tolerances of tests/test_gpu_fp.py: lattice-valued outputs identical, distributions 1e-9 of the zone maximum."""
    bad = {}
    if not after_fp:
        for k in ("f_nt", "Pnt", "n_e", "tea"):
            if not np.array_equal(np.asarray(pg[k]).ravel(), np.asarray(po[k]).ravel()):
                bad[k] = "changed without an update"
        return bad
    for k in ("gmin", "gmax", "p_nth", "tea", "Te_new"):
        if not np.array_equal(pg[k], po[k]):
            bad[k] = (pg[k][:3], po[k][:3])
    for k in ("f_nt", "Pnt"):
        err = np.abs(pg[k] - po[k]).max(axis=1) / np.abs(po[k]).max(axis=1)
        if err.max() > 1e-8:
            bad[k] = float(err.max())
    for k in ("n_e", "amxwl"):
        if np.abs(pg[k] - po[k]).max() > 1e-9 * np.abs(po[k]).max():
            bad[k] = float(np.abs(pg[k] - po[k]).max())
    return bad
