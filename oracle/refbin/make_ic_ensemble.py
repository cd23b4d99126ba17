#!/usr/bin/env python
"""Ensemble of the UNMODIFIED reference binary for the statistical test of the scattered (inverse-Compton) component.

A single run of the `dense` configuration (oracle/refbin/make_golden.py: 3x3 zones, n_e = 3e5, split 60/40/12, nst = 1200)
has O(20) primary collisions per cycle, each of which fans out into ~300 correlated children: one run cannot test the
scattered component bin by bin (SURVEY.md 8 notes).  This script runs the same inputs with N different `rseed` values and
rand_switch = 1 (cycles 0-2) -- with the shipped rand_switch = 2 the worker's ran1 stream does not depend on rseed at all:
fp_mpi.f:68-72 calls ran1(rseed) with a POSITIVE argument while ran1's state is still empty, and rand.f:344 then seeds with
max(-idum, 1) = 1 whatever rseed is (every such run is bit-identical); the lagged-Fibonacci generator is seeded through
seed_zone(rseed) and does follow it -- and keeps, per seed and cycle, what the worker prints (cmcount1, cmcount2: xec2d.f:191) and what
it writes to its event file (p001_evb.dat: every escape with t_bound, xnu, ew), reduced to counts, weights and a coarse
spectrum of the scattered escapes.  -> tests/golden/ic_ensemble_dense.npz, read by tests/test_statistical_parity.py.

usage: python oracle/refbin/make_ic_ensemble.py [n_seeds=32] [n_workers=8]"""
from __future__ import annotations

import re
import shutil
import sys
from concurrent.futures import ProcessPoolExecutor
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent))
import refrun  # noqa: E402
from make_golden import CONFIGS  # noqa: E402

REPO = Path(__file__).resolve().parent.parent.parent
GOLD = REPO / "tests" / "golden"
EDGES = np.array([-9.0, -5.0, -3.0, -1.0, 1.0, 3.0, 5.0, 7.0, 12.0])      # log10(xnu / keV) of the scattered escapes


def one(seed: int):
    cfg = CONFIGS["dense"]
    rd = Path("/tmp") / f"c2d_ens_{seed}"
    shutil.rmtree(rd, ignore_errors=True)
    rd.mkdir()
    refrun.write_inputs(rd, dict(cfg["over"], rseed=seed, rand_switch=1), cfg["zone"])
    r = refrun.run(rd, trace=False, cpu=None)
    assert r.returncode == 0, (rd / "stdout.txt").read_text()[-2000:]
    so = (rd / "stdout.txt").read_text()
    cm = [(float(t), int(a), int(b)) for t, a, b in
          re.findall(r"time=\s*([0-9.Ee+-]+)\s+cmcount1=\s*(\d+)\s*\n?\s*cmcount2=\s*(\d+)", so)]
    txt = (rd / "p001_evb.dat").read_text()
    ev = np.loadtxt(rd / "p001_evb.dat", ndmin=2) if txt.strip() else np.zeros((0, 7))
    shutil.rmtree(rd, ignore_errors=True)
    return seed, cm, ev


def main(n_seeds=32, n_workers=8):
    seeds = [1000 + 37 * i for i in range(n_seeds)]
    with ProcessPoolExecutor(n_workers) as ex:
        res = list(ex.map(one, seeds))
    # the cycle length: time printed with the second cmcount line of any run (xec2d.f:191 prints `time` before it advances)
    dt = min(t for _, cm, _ in res for t, _, _ in cm if t > 0)
    w_med = float(np.median(np.concatenate([ev[:, 2] for _, _, ev in res])))
    out = dict(seeds=np.array(seeds), dt=np.array(dt), w_median=np.array(w_med), edges=EDGES,
               cm1=np.zeros((n_seeds, 2), np.int64), cm2=np.zeros((n_seeds, 2), np.int64),
               n_sc=np.zeros((n_seeds, 2), np.int64), w_sc=np.zeros((n_seeds, 2)), n_un=np.zeros((n_seeds, 2), np.int64),
               w_un=np.zeros((n_seeds, 2)), spec_sc=np.zeros((n_seeds, len(EDGES) - 1)))
    for i, (seed, cm, ev) in enumerate(res):
        for t, a, b in cm:
            c = int(round(t / dt))                      # 0 -> cycle 1, 1 -> cycle 2
            if c < 2:
                out["cm1"][i, c] = a; out["cm2"][i, c] = b
        heavy = (ev[:, 2] > 0.3 * w_med) & (ev[:, 2] < 3.0 * w_med) & (ev[:, 1] < 10.0)     # unscattered source photons
        for c in (0, 1):
            m = (ev[:, 0] >= c * dt) & (ev[:, 0] < (c + 1) * dt)
            out["n_sc"][i, c] = int((m & ~heavy).sum()); out["w_sc"][i, c] = float(ev[m & ~heavy, 2].sum())
            out["n_un"][i, c] = int((m & heavy).sum()); out["w_un"][i, c] = float(ev[m & heavy, 2].sum())
        sc = ~heavy & (ev[:, 0] < 2 * dt)
        out["spec_sc"][i] = np.histogram(np.log10(ev[sc, 1]), bins=EDGES, weights=ev[sc, 2])[0]
    np.savez_compressed(GOLD / "ic_ensemble_dense.npz", **out)
    print("wrote", GOLD / "ic_ensemble_dense.npz", "collisions per cycle:", out["cm1"].mean(axis=0),
          "scattered escapes per cycle:", out["n_sc"].mean(axis=0))


if __name__ == "__main__":
    main(*(int(a) for a in sys.argv[1:]))
