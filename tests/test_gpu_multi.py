"""Two B200s, the collectives INSIDE the library over NCCL: tally all-reduce, the scattering cascade dealt
out across the ranks (children and 3rd-split sub-children, escalation ladder included), FP zones partitioned and
all-gathered.  Three closed steps (c2d_step) on two ranks must equal the same steps on one rank: integers exactly, FP64
tallies and the plasma to summation-order rounding.  Skipped on a box with one GPU (tests/test_host_logic_cpu.py runs the
same comparison on the g++ build over gloo)."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys
sys.path[:0] = [r"{root}", r"{root}/tests", r"{root}/oracle"]
import numpy as np, torch, torch.distributed as dist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
from golden_util import Golden
from loop_util import gpu_closed, initial_state
o, st, P, scal = initial_state(Golden("{golden}"))
scal = dict(scal, nst={nst})
g = gpu_closed(P, st, scal, T_const=0, sigma_mode=1, device=int(os.environ["LOCAL_RANK"]), rank=rank, world=world)
reps = [g.step(scal["dt"]) for _ in range(3)]
t = g.tallies()
pl = g.plasma_get()
sid = g.census()[2]
gathered = [None] * world
dist.all_gather_object(gathered, sid)
if rank == 0:
    np.savez(r"{out}", edep=t["edep"], ecens=t["ecens"], n_field=t["n_field"], fout_total=t["fout_total"], npcen=t["npcen"],
             nelectron=t["nelectron"], E_IC=t["E_IC"], ndxout=reps[-1]["ndxout"], nevents=reps[-1]["nevents"],
             hist=reps[-1]["histories"], sid=np.sort(np.concatenate(gathered)), f_nt=pl["f_nt"], tea=pl["tea"], n_e=pl["n_e"],
             cm1=t["counters_global"]["cmcount1"], cm2=t["counters_global"]["cmcount2"], local_ndxout=t["counters"]["ndxout"])
g.close()
dist.destroy_process_group()
'''


@pytest.mark.parametrize("golden,nst", [("dense", 40000)])
def test_closed_steps_on_two_gpus_equal_one(tmp_path, golden, nst):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    sys.path[:0] = [str(ROOT / "tests"), str(ROOT / "oracle")]
    from golden_util import Golden
    from loop_util import gpu_closed, initial_state
    o, st, P, scal = initial_state(Golden(golden))
    scal = dict(scal, nst=nst)
    g = gpu_closed(P, st, scal, T_const=0, sigma_mode=1)
    for _ in range(3):
        rep = g.step(scal["dt"])
    one, pl1, sid1 = g.tallies(), g.plasma_get(), np.sort(g.census()[2])
    g.close()
    assert one["counters_global"]["cmcount2"] > one["counters_global"]["cmcount1"] > 0        # children were dealt out
    out, script = tmp_path / "two.npz", tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT), out=str(out), golden=golden, nst=nst))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29547", str(script)], capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, MASTER_ADDR="127.0.0.1"))
    assert r.returncode == 0, r.stderr[-3000:]
    two = np.load(out)
    for k in ("npcen", "nelectron"):
        assert np.array_equal(two[k], one[k]), k
    assert int(two["ndxout"]) == rep["ndxout"] and int(two["nevents"]) == rep["nevents"] and int(two["hist"]) == rep["histories"]
    assert int(two["cm1"]) == one["counters_global"]["cmcount1"] and int(two["cm2"]) == one["counters_global"]["cmcount2"]
    assert 0 < int(two["local_ndxout"]) < rep["ndxout"] and np.array_equal(two["sid"], sid1)
    for k in ("edep", "ecens", "n_field", "fout_total", "E_IC"):
        assert np.abs(two[k] - one[k]).max() <= 1e-11 * np.abs(one[k]).max(), k
    assert np.array_equal(two["tea"], pl1["tea"])                                             # lattice-valued
    for k in ("f_nt", "n_e"):
        assert np.abs(two[k] - pl1[k]).max() <= 1e-9 * np.abs(pl1[k]).max(), k
