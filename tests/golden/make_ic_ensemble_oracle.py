#!/usr/bin/env python
"""The oracle's side of the inverse-Compton ensemble test: the closed loop of tests/loop_util.py (cycles 0-2, FP on, Philox
streams) on the `dense` configuration for N seeds, reduced exactly like oracle/refbin/make_ic_ensemble.py reduces the
reference binary's runs.  -> tests/golden/ic_ensemble_oracle.npz.  TEST INFRASTRUCTURE.

usage: python tests/golden/make_ic_ensemble_oracle.py [n_seeds=32]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests"), str(ROOT / "oracle")]
from golden_util import Golden  # noqa: E402
from loop_util import PLASMA, initial_state, oracle_cycle  # noqa: E402

NST = 1200          # input.dat nst of the `dense` configuration (the worker's copy of nst is 0 in the snapshots)


def new_stats(ens):
    return dict(cm1=np.zeros(2, np.int64), cm2=np.zeros(2, np.int64), n_sc=np.zeros(2, np.int64), w_sc=np.zeros(2),
                n_un=np.zeros(2, np.int64), w_un=np.zeros(2), spec_sc=np.zeros(len(ens["edges"]) - 1))


def add_cycle(out, c, counters, ev, ens):
    """Fold cycle c + 1 (c = 0, 1) into the ensemble statistics: the step's collision counters and its escape records."""
    w_med, edges = float(ens["w_median"]), ens["edges"]
    heavy = (ev[:, 2] > 0.3 * w_med) & (ev[:, 2] < 3.0 * w_med) & (ev[:, 1] < 10.0)
    out["cm1"][c] = counters["cmcount1"]; out["cm2"][c] = counters["cmcount2"]
    out["n_sc"][c] = int((~heavy).sum()); out["w_sc"][c] = float(ev[~heavy, 2].sum())
    out["n_un"][c] = int(heavy.sum()); out["w_un"][c] = float(ev[heavy, 2].sum())
    out["spec_sc"] += np.histogram(np.log10(ev[~heavy, 1]), bins=edges, weights=ev[~heavy, 2])[0]


def member(o, st, scal, seed, ens):
    """Cycles 0-2 from the initial plasma with Philox seed `seed` -> the per-cycle statistics of the ensemble fixture."""
    a = o.a
    for k in PLASMA:
        getattr(a, k)[: st[k].size] = st[k]
    a.ecens[:] = 0.0; a.fout[:] = 0.0; a.ec_old[:] = 0.0
    o["ndxout"] = 0
    o.rseed = int(seed)
    out = new_stats(ens)
    time = 0.0
    for cyc in range(3):
        res = oracle_cycle(o, cyc, time, scal, T_const=scal["T_const"])
        if cyc > 0:
            time += scal["dt"]
            add_cycle(out, cyc - 1, res["counters"], res["events"], ens)
    return out


def setup():
    ens = np.load(ROOT / "tests" / "golden" / "ic_ensemble_dense.npz")
    o, st, P, scal = initial_state(Golden("dense"))
    scal = dict(scal, nst=NST)
    assert abs(scal["dt"] / float(ens["dt"]) - 1.0) < 1e-12
    return o, st, scal, ens


def main(n_seeds=32):
    o, st, scal, ens = setup()
    seeds = [500 + 13 * i for i in range(n_seeds)]
    rows = [member(o, st, scal, s, ens) for s in seeds]
    out = {k: np.stack([r[k] for r in rows]) for k in rows[0]}
    np.savez_compressed(ROOT / "tests" / "golden" / "ic_ensemble_oracle.npz", seeds=np.array(seeds), **out)
    print("collisions per cycle:", out["cm1"].mean(axis=0), "scattered escapes per cycle:", out["n_sc"].mean(axis=0))


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:]))
