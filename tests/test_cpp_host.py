"""The C++ host mirror (include/compton2d_b200.hpp: c2d::Compton2d, the reference's xec call sequence name by name) drives the
same library as the Python mirror: tests/cpp/xec_mirror.cpp, compiled with g++ and linked against the g++ debug build of the
kernel sources (tests/emu, TEST INFRASTRUCTURE), must produce bit-identical tallies, census and event records, and must turn
the library's error codes into the reference's print-and-stop behaviour."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "emu"))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    import build_emu
    so = build_emu.build()
    out = tmp_path_factory.mktemp("cpp") / "xec_mirror"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", f"-I{ROOT / 'include'}", str(ROOT / "tests/cpp/xec_mirror.cpp"),
                           "-o", str(out), f"-L{so.parent}", "-lc2d_emu", f"-Wl,-rpath,{so.parent}"])
    return out


def _dump(P, d):
    d.mkdir(parents=True, exist_ok=True)
    for k in ("z", "r", "E_ph", "E_field", "E_gg", "gnt", "hu", "Elcmin", "Elcmax", "mu", "kappa_tot", "eps_tot", "eps_th", "f_nt", "Pnt",
              "n_e", "tea", "f_pair", "k_gg", "ewsv", "Eloss_th", "Eloss_tot", "zsurf", "tbbl"):
        np.ascontiguousarray(P[k], dtype=np.float64).tofile(d / f"{k}.f64")
    np.ascontiguousarray(P["nsv"], dtype=np.int32).tofile(d / "nsv.i32")
    sc = dict(nz=P["nz"], nr=P["nr"], rmin=P["rmin"], zmin=P["zmin"], split1=P["split"][0], split2=P["split"][1], split3=P["split"][2],
              spl3_trg=P["split"][3], pair_switch=P["pair_switch"], spec_switch=P["spec_switch"], rseed=9857, time=P["time"], dt=P["dt"])
    (d / "scalars.txt").write_text("".join(f"{k} {float(v)!r}\n" for k, v in sc.items()))


def test_cpp_mirror_equals_python_mirror(exe, tmp_path, monkeypatch):
    import build_emu
    from compton2d_b200 import lib as L
    monkeypatch.setattr(L, "LIBPATH", build_emu.build())
    monkeypatch.setattr(L, "_lib", None)
    from parity_util import gpu_step, make_gpu
    from synth import make_problem
    P = make_problem(nz=4, nr=3, nst=1500, ngroups=64, split=(40, 20, 6, 10), n_e=3.0e5)
    g = make_gpu(P)
    gpu_step(g, P, 1, field=False)
    t = gpu_step(g, P, 2)
    g.close()
    monkeypatch.setattr(L, "_lib", None)
    _dump(P, tmp_path)
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    rd = lambda n, dt: np.fromfile(tmp_path / f"out_{n}", dtype=dt)
    for k in ("edep", "prdep", "ecens", "E_IC", "erlku"):
        assert np.array_equal(rd(k, np.float64), np.asarray(t[k]).ravel()), k
    assert np.array_equal(rd("n_field", np.float64), t["n_field"].ravel()) and np.array_equal(rd("fout", np.float64), t["fout"].ravel())
    assert np.array_equal(rd("npcen", np.int64), t["npcen"]) and np.array_equal(rd("nelectron", np.int64), t["nelectron"])
    from compton2d_b200.lib import COUNTERS
    assert np.array_equal(rd("counters", np.int64), [t["counters"][k] for k in COUNTERS])
    d, iv, sid = t["census"]
    assert np.array_equal(rd("dbuf", np.float64), d.ravel()) and np.array_equal(rd("ibuf", np.int32), iv.ravel())
    assert np.array_equal(rd("sid", np.uint64), sid)
    assert np.array_equal(rd("events", np.float64), t["events"].ravel()) and len(t["events"]) > 10
    assert len((tmp_path / "p001_census.dat").read_text().splitlines()) == 2 * len(d)
    assert "2 cycles" in r.stdout


def test_cpp_mirror_error_behaviour(exe, tmp_path):
    """cr_sent = 2 (Compton reflection off the outer cylinder, not built): the library returns C2D_EINVAL, the mirror throws, the host prints and stops."""
    from synth import make_problem
    P = make_problem(nz=2, nr=2, nst=100, ngroups=16)
    _dump(P, tmp_path)
    r = subprocess.run([str(exe), str(tmp_path), "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 12 and "cr_sent" in r.stderr


def test_cpp_closed_step_equals_python(exe, tmp_path, monkeypatch):
    """c2d::Compton2d::step (the closed time step, c2d_step) against compton2d_b200.host.Compton2d.step on the same library:
    three cycles on the resident plasma, T_const = 1 (the emu build has no FP kernel)."""
    import build_emu
    from compton2d_b200 import lib as L
    monkeypatch.setattr(L, "LIBPATH", build_emu.build())
    monkeypatch.setattr(L, "_lib", None)
    from golden_util import Golden
    from loop_util import gpu_closed, initial_state
    o, st, P, scal = initial_state(Golden("small"))
    del o
    scal = dict(scal, nst=1200, inj=dict(scal["inj"], inj_switch=0, pick_sw=0))
    g = gpu_closed(P, st, scal, T_const=1)
    reps = [g.step(scal["dt"]) for _ in range(3)]
    t = g.tallies()
    g.close()
    monkeypatch.setattr(L, "_lib", None)
    _dump(dict(P, **{k: st[k] for k in ("f_nt", "Pnt", "n_e", "tea", "zsurf")}), tmp_path)
    for k in ("tna", "B_field", "gmin", "gmax", "p_nth", "amxwl", "turb_lev", "vol"):
        np.ascontiguousarray(st[k], dtype=np.float64).tofile(tmp_path / f"{k}.f64")
    with open(tmp_path / "scalars.txt", "a") as f:
        f.write(f"nst {float(scal['nst'])!r}\nT_const 1.0\n")
    lines = [l for l in (tmp_path / "scalars.txt").read_text().splitlines() if not l.startswith(("dt ", "rseed "))]
    (tmp_path / "scalars.txt").write_text("\n".join(lines) + f"\ndt {scal['dt']!r}\nrseed {float(scal['rseed'])!r}\n")
    r = subprocess.run([str(exe), str(tmp_path), "closed"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    cnt = np.fromfile(tmp_path / "out_closed_counts", dtype=np.int64).reshape(3, 4)
    assert np.array_equal(cnt, [[x["ndxout"], x["nevents"], x["n_volume"], x["histories"]] for x in reps])
    bl = np.fromfile(tmp_path / "out_closed_bingo", dtype=np.float64).reshape(3, 2)
    assert np.array_equal(bl, [[x["bingo"], x["luminosity"]] for x in reps])
    assert np.array_equal(np.fromfile(tmp_path / "out_closed_ecens", dtype=np.float64), t["ecens"])
    assert np.array_equal(np.fromfile(tmp_path / "out_closed_fout_total", dtype=np.float64), t["fout_total"].ravel())
    assert "ncycle 3" in r.stdout
