"""Shared machinery of the parity tests: build a problem (from a reference-binary snapshot or
synthetic), run it through the CPU oracle in Philox replay mode and through the CUDA library, compare."""
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
for p in (REPO, REPO / "oracle", REPO / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import orc  # noqa: E402

NLC = 16


def problem_from_snapshot(s_in, s_vol, split=None, nsv_scale=1.0, dt=None):
    """Dense problem description from reference-binary snapshots: `s_in` = state before the phase
    (tables the worker tracks with), `s_vol` = snapshot holding the volume-job inputs of the same cycle."""
    nz, nr = int(s_in.nz), int(s_in.nr)
    nzone = nz * nr
    npt = int(s_vol.nphtotal)
    P = dict(nz=nz, nr=nr, nzone=nzone, rmin=float(s_in.rmin), zmin=float(s_in.zmin),
             z=np.array(s_in.z), r=np.array(s_in.r), E_ph=np.array(s_vol.e_ph), E_field=np.array(s_vol.e_field),
             E_gg=np.array(s_in.e_gg), gnt=np.array(s_in.gnt), hu=np.array(s_vol.hu[: npt + 1]),
             Elcmin=np.array(s_vol.elcmin[: int(s_vol.nph_lc)]), Elcmax=np.array(s_vol.elcmax[: int(s_vol.nph_lc)]),
             mu=np.array(s_vol.mu[: int(s_vol.nmu)]), pair_switch=int(s_in.pair_switch), spec_switch=int(s_in.spec_switch),
             split=tuple(split) if split else (int(s_in.split1), int(s_in.split2), int(s_in.split3), int(s_in.spl3_trg)),
             dt=float(dt if dt is not None else s_vol.dt[0]), time=float(s_in.time))
    t3 = lambda a: np.ascontiguousarray(np.transpose(a, (1, 2, 0))).reshape(nzone, -1)
    P.update(kappa_tot=t3(s_vol.kappa_tot), eps_tot=t3(s_vol.eps_tot), eps_th=t3(s_vol.eps_th), k_gg=t3(s_in.k_gg),
             f_nt=np.array(s_in.f_nt).reshape(nzone, -1), Pnt=np.array(s_in.pnt).reshape(nzone, -1),
             n_e=np.array(s_in.n_e).ravel(), tea=np.array(s_in.tea).ravel(), f_pair=np.array(s_in.f_pair).ravel(),
             nsv=np.maximum((np.array(s_vol.nsv).ravel() * nsv_scale).astype(np.int32), 0),
             ewsv=np.array(s_vol.ewsv).ravel(), Eloss_th=np.array(s_vol.eloss_th).ravel(),
             Eloss_tot=np.array(s_vol.eloss_tot).ravel(), zsurf=np.array(s_vol.zsurf).ravel(),
             tbbl=np.array(s_vol.tbbl[:, 0]))
    return P


def make_oracle(P, rseed=9857, cens_cap=400000, ev_cap=400000):
    o = orc.Oracle(P["nz"], P["nr"], n_vol=P["E_ph"].size, num_nt=P["gnt"].size, nphfield=P["E_field"].size,
                   n_gg=P["E_gg"].size, nphtotal=P["hu"].size - 1, nph_lc=P["Elcmin"].size, nmu=P["mu"].size,
                   cens_cap=cens_cap, ev_cap=ev_cap)
    a = o.a
    for k in ("z", "r", "E_ph", "E_field", "E_gg", "gnt", "hu", "Elcmin", "Elcmax", "mu"):
        setattr(a, k, P[k])
    o["rmin"] = P["rmin"]; o["zmin"] = P["zmin"]
    o["pair_switch"] = P["pair_switch"]; o["spec_switch"] = P["spec_switch"]; o["cr_sent"] = P.get("cr_sent", 0)
    if "E_ref" in P:                                   # Compton-reflection tables, P_ref / W_abs flat in COMMON order
        a.E_ref = P["E_ref"]; a.P_ref = P["P_ref"]; a.W_abs = P["W_abs"]
    for k, v in zip(("split1", "split2", "split3", "spl3_trg"), P["split"]):
        o[k] = v
    o["rand_switch"] = 3
    o.rseed = rseed
    return o


SURF_KEYS = ("nsurfu", "nsurfl", "nsurfi", "nsurfo", "ewsurfu", "ewsurfl", "ewsurfi", "ewsurfo", "tbbu", "tbbl", "tbbi", "tbbo")


def oracle_surface(o, S):
    """Load the surface-source description S (dict, see make_surface) into the oracle and run imcsurf2d."""
    a = o.a
    for k in SURF_KEYS:
        if S.get(k) is not None:
            setattr(a, k, S[k])
    o["star_switch"] = int(S.get("star_switch", 0))
    for q, t in enumerate(S.get("spectra", ())):
        o.set_spectrum(q, t["E_file"], t["F_file"], t["P_file"], t["a1"], t["I_file"], t["nfile"])
    for k in ("spec_u", "spec_l", "spec_i", "spec_o"):
        getattr(a, k)[:] = 0 if S.get(k) is None else S[k]
    o.surf_all()


def _step_tbbl(P, surf):
    """tbbl(k, ti) is ONE COMMON array in the reference: the lower-face leak test of imcleak (imcleak2d.f:95) reads the values the
    surface source uses, in every phase of the step."""
    return surf["tbbl"] if surf is not None and surf.get("tbbl") is not None else P["tbbl"]


def oracle_step(o, P, ncycle, time=None, field=True, vol=True, surf=None):
    a = o.a
    for k in ("kappa_tot", "eps_tot", "eps_th", "k_gg", "f_nt", "Pnt", "n_e", "tea", "f_pair", "nsv", "ewsv",
              "Eloss_th", "Eloss_tot", "zsurf"):
        setattr(a, k, P[k])
    a.tbbl = _step_tbbl(P, surf)
    o["dt"] = P["dt"]; o["time"] = P["time"] if time is None else time; o["ncycle"] = ncycle
    o.philox_key(o.rseed, ncycle)
    o.step_zero()
    a.fout[:] = 0; a.prdep[:] = 0; a.n_ph[:] = 0; a.nelectron[:] = 0
    o["nev"] = 0; o["n_multiple"] = 0; o["n_killed"] = 0; o["n_trials"] = 0
    if field:
        o.field_calc()
    else:
        o["ndxout"] = 0
    if vol:
        o.vol_all()
    if surf is not None:
        oracle_surface(o, surf)
    assert int(o["fatal"]) == 0, f"oracle fatal {int(o['fatal'])}"
    nz, nr, nzone = P["nz"], P["nr"], P["nzone"]
    npt = P["hu"].size - 1
    d, i, s = o.census()
    return dict(edep=a.edep[:nzone].copy(), prdep=a.prdep[:nzone].copy(), ecens=a.ecens[:nzone].copy(),
                npcen=a.npcen[:nzone].copy(), n_field=a.n_field[: nzone * P["E_field"].size].reshape(nzone, -1).copy(),
                n_ph=a.n_ph[: nzone * P["E_gg"].size].reshape(nzone, -1).copy(), E_IC=a.E_IC[: P["gnt"].size].copy(),
                nelectron=a.nelectron[: P["gnt"].size + 2].copy(), fout=a.fout[: 32 * npt].reshape(32, npt).copy(),
                edout=a.edout[: 32 * NLC].reshape(32, NLC).copy(), erlki=a.erlki[:nz].copy(), erlko=a.erlko[:nz].copy(),
                erlku=a.erlku[:nr].copy(), erlkl=a.erlkl[:nr].copy(), Ed_in=a.Ed_in[:nr].copy(),
                counters=dict(ndxout=int(o["ndxout"]), cmcount1=int(o["cmcount1"]), cmcount2=int(o["cmcount2"]),
                              lkcount=int(o["lkcount"]), nevents=int(o["nev"]), multiple=int(o["n_multiple"]),
                              killed=int(o["n_killed"]), trials=int(o["n_trials"])),
                census=(d, i, s), events=o.events())


def make_gpu(P, **kw):
    from compton2d_b200.synth import context_for
    return context_for(P, **kw)


def gpu_step(g, P, ncycle, time=None, field=True, vol=True, surf=None):
    g.imcgen2d(ncycle=ncycle, time=P["time"] if time is None else time, dt=P["dt"], kappa_tot=P["kappa_tot"],
               eps_tot=P["eps_tot"], eps_th=P["eps_th"], f_nt=P["f_nt"], Pnt=P["Pnt"], n_e=P["n_e"], tea=P["tea"],
               f_pair=P["f_pair"], k_gg=P["k_gg"], nsv=P["nsv"], ewsv=P["ewsv"], Eloss_th=P["Eloss_th"],
               Eloss_tot=P["Eloss_tot"], zsurf=P["zsurf"], tbbl=_step_tbbl(P, surf))
    if field:
        g.imcfield2d()
    if vol:
        g.imcvol2d()
    if surf is not None:
        g.imcsurf2d(**surf)
    g.sync()
    t = g.tallies()
    t["census"] = g.census()
    t["events"] = g.events()
    return t


def rel(a, b):
    a = np.asarray(a, float); b = np.asarray(b, float)
    sc = max(np.abs(b).max() if b.size else 0.0, 1e-300)
    return float(np.abs(a - b).max() / sc) if a.size else 0.0


INT_COUNTERS = ("ndxout", "cmcount1", "cmcount2", "lkcount", "nevents", "multiple", "killed", "trials")


def compare(gpu, ref, tol=1e-11, census_tol=1e-9):
    """-> dict of mismatches.  Integers exact; FP64 tallies to `tol` of the array's max; census records
    matched by replay stream id."""
    bad = {}
    for k in INT_COUNTERS:
        if gpu["counters"][k] != ref["counters"][k]:
            bad[k] = (gpu["counters"][k], ref["counters"][k])
    for k in ("npcen", "nelectron"):
        if not np.array_equal(gpu[k], ref[k]):
            bad[k] = int(np.abs(gpu[k] - ref[k]).sum())
    for k in ("edep", "prdep", "ecens", "n_field", "n_ph", "E_IC", "fout", "edout", "erlki", "erlko", "erlku", "erlkl", "Ed_in"):
        e = rel(gpu[k], ref[k])
        if e > tol:
            bad[k] = e
    (gd, gi, gs), (rd, ri, rs) = gpu["census"], ref["census"]
    if len(gs) != len(rs):
        bad["census_n"] = (len(gs), len(rs))
    else:
        go, ro = np.argsort(gs, kind="stable"), np.argsort(rs, kind="stable")
        if not np.array_equal(gs[go], rs[ro]):
            bad["census_sid"] = int((gs[go] != rs[ro]).sum())
        else:
            if not np.array_equal(gi[go], ri[ro]):
                bad["census_int"] = int((gi[go] != ri[ro]).any(axis=1).sum())
            r = np.abs(gd[go] - rd[ro]) / np.maximum(np.abs(rd[ro]), 1e-300)
            if r.size and r.max() > census_tol:
                bad["census_dbl"] = float(r.max())
    ge, re_ = gpu["events"], ref["events"]
    if ge.shape != re_.shape:
        bad["events_n"] = (ge.shape, re_.shape)
    elif ge.size:
        key = lambda e: np.lexsort((e[:, 6], e[:, 5], e[:, 4], e[:, 3], e[:, 1], e[:, 2]))
        a, b = ge[key(ge)], re_[key(re_)]
        r = np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
        if r.max() > census_tol:
            bad["events"] = float(r.max())
    return bad


def reflection_tables(P):
    """E_ref, P_ref, W_abs of the reference binary's setup (Pref_calc / Wref_calc, ref_matrix.f) from the `reflect` golden, flat in
    COMMON order, into problem P."""
    from golden_util import Golden
    s = Golden("reflect").snap("c1_B")
    P["E_ref"] = np.array(s.e_ref)
    P["P_ref"] = np.ascontiguousarray(np.asarray(s.p_ref).T).ravel()
    P["W_abs"] = np.ascontiguousarray(np.asarray(s.w_abs).T).ravel()
    return P
