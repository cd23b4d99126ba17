"""Shared by the step-prologue tests (imcgen2d / volume_em on the device): the plasma state of a
reference-binary snapshot as dense arrays, and the binary's own tables of the same cycle to compare with."""
import numpy as np

from parity_util import problem_from_snapshot

NST = dict(small=6000, annulus=5000, dense=1200, surface=1200)      # nst of oracle/refbin/make_golden.py CONFIGS


def plasma_from_snapshot(g, cyc):
    """-> (P, plasma inputs, reference outputs).  Inputs: what the master holds at the top of cycle `cyc` (snapshot A);
    outputs: eps_tot / eps_th / E_ph / Eloss / nsv / ewsv as delivered with the volume jobs (snapshot C) and kappa_tot as
    delivered by z_surf_bcast (snapshot D; it lags one step on the worker)."""
    A, B, C, D = (g.snap(f"c{cyc}_{x}") for x in "ABCD")
    P = problem_from_snapshot(B, C)
    nz, nr = P["nz"], P["nr"]
    nzone = nz * nr
    t3 = lambda a: np.ascontiguousarray(np.transpose(a, (1, 2, 0))).reshape(nzone, -1)
    inp = dict(f_nt=np.array(A.f_nt).reshape(nzone, -1), Pnt=np.array(A.pnt).reshape(nzone, -1), n_e=np.array(A.n_e).ravel(),
               tea=np.array(A.tea).ravel(), B_field=np.array(A.b_field).ravel(), vol=np.array(C.vol).ravel(),
               zsurf=np.array(C.zsurf).ravel(), dt=float(C.dt[0]), time=float(A.time))
    ref = dict(kappa_tot=t3(D.kappa_tot), eps_tot=t3(C.eps_tot), eps_th=t3(C.eps_th), nsv=np.array(C.nsv).ravel(),
               ewsv=np.array(C.ewsv).ravel(), Eloss_tot=np.array(C.eloss_tot).ravel(), Eloss_th=np.array(C.eloss_th).ravel(),
               Eloss_sy=np.array(D.eloss_sy).ravel() if False else None)
    return P, inp, ref


def rel(a, b):
    a = np.asarray(a, float).ravel(); b = np.asarray(b, float).ravel()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def check_against_binary(out, tables, ref, tol=1e-11):
    assert np.array_equal(out["nsv"], ref["nsv"]), (out["nsv"], ref["nsv"])
    for k in ("ewsv", "Eloss_tot", "Eloss_th"):
        assert rel(out[k], ref[k]) < tol, (k, rel(out[k], ref[k]))
    for k in ("kappa_tot", "eps_tot", "eps_th"):
        assert rel(tables[k], ref[k]) < tol, (k, rel(tables[k], ref[k]))
