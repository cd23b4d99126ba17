"""The C++ host mirror (include/compton2d_b200.hpp) on the device: tests/cpp/xec_mirror.cpp linked against the CUDA library
itself must reproduce the Python mirror's run -- integers exactly, FP64 tallies to 1e-12 (atomic order differs run to run)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from parity_util import gpu_step, make_gpu
from synth import make_problem
from test_cpp_host import _dump

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_cpp_mirror_on_the_device(tmp_path):
    from compton2d_b200 import lib as L
    libdir = Path(L.LIBPATH).parent
    exe = tmp_path / "xec_mirror"
    subprocess.check_call(["g++", "-std=c++17", "-O1", f"-I{ROOT / 'include'}", str(ROOT / "tests/cpp/xec_mirror.cpp"), "-o", str(exe),
                           f"-L{libdir}", "-lcompton2d_b200", f"-Wl,-rpath,{libdir}", "-Wl,--allow-shlib-undefined"])
    P = make_problem(nz=4, nr=3, nst=20000, ngroups=64, split=(40, 20, 6, 10), n_e=3.0e5)
    g = make_gpu(P)
    gpu_step(g, P, 1, field=False)
    t = gpu_step(g, P, 2)
    g.close()
    _dump(P, tmp_path)
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    rd = lambda n, dt: np.fromfile(tmp_path / f"out_{n}", dtype=dt)
    assert np.array_equal(rd("npcen", np.int64), t["npcen"]) and np.array_equal(rd("nelectron", np.int64), t["nelectron"])
    c = rd("counters", np.int64)
    from compton2d_b200.lib import COUNTERS
    for k in ("ndxout", "cmcount1", "cmcount2", "lkcount", "nevents", "histories", "trials"):
        assert c[COUNTERS.index(k)] == t["counters"][k], k
    for k in ("edep", "ecens", "E_IC"):
        a, b = rd(k, np.float64), np.asarray(t[k]).ravel()
        assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max(), k
    assert np.array_equal(np.sort(rd("sid", np.uint64)), np.sort(t["census"][2])) and t["counters"]["cmcount2"] > 0
