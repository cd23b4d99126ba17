"""Fixture generator (TEST INFRASTRUCTURE): the CPU oracle on BASELINE configs[0] as shipped, cycles 0-1, Philox replay.
Prints the counters kept in tests/golden/oracle_c0_replay.json.  ~10 min on one core."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[2]
sys.path[:0] = [str(ROOT), str(ROOT / "tests"), str(ROOT / "oracle")]
import numpy as np
from golden_util import Golden
from loop_util import initial_state, oracle_cycle
gd=Golden("c0")
o, st, P, scal = initial_state(gd, cens_cap=2_000_000, ev_cap=400_000)
scal["nst"]=800000
w_ref=float(gd.d["events_w_median"])
t=0.0
for cyc in range(2):
    t0=time.time()
    r=oracle_cycle(o,cyc,t,scal)
    ev=r["events"]
    u=(ev[:,2]>0.3*w_ref)&(ev[:,2]<3*w_ref)&(ev[:,1]<10) if len(ev) else np.zeros(0,bool)
    print(cyc, r["counters"], "heavy escapes", int(u.sum()), "of", len(ev), "%.0fs"%(time.time()-t0), flush=True)
    if cyc>0: t+=scal["dt"]
