"""Reference file formats of the device-owned state (csrc/c2d_io.cu): census file (write_cens / read_cens,
census2d.f) and event file (format 105, imcleak2d.f:181).

Pin: tests/golden/evb_dense.txt holds the first 600 lines of the event file the reference BINARY wrote for the `dense`
golden run (oracle/refbin/make_golden.py).  The library's E14.7 formatter must reproduce that text byte for byte from the
parsed values, and -- rounding included -- from the full-precision values of the oracle's replay of the same run."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / "tests" / "golden"
sys.path.insert(0, str(ROOT / "tests" / "emu"))


@pytest.fixture(scope="module")
def fmt():
    from compton2d_b200 import build, lib as L
    build.build()
    lib = L.load()
    buf = C.create_string_buffer(16)

    def f(v):
        lib.c2d_format_e14_7(float(v), buf)
        return buf.raw[:14].decode()
    return f


def _line(fmt, rec):
    return " ".join(fmt(v) for v in rec)


def test_e14_7_reproduces_the_reference_binarys_event_file(fmt):
    lines = (GOLD / "evb_dense.txt").read_text().splitlines()
    assert len(lines) == 600 and all(len(l) == 104 for l in lines)
    for l in lines:
        rec = [float(l[15 * k: 15 * k + 14]) for k in range(7)]
        assert _line(fmt, rec) == l


def test_e14_7_edge_cases(fmt):
    assert fmt(0.0) == " 0.0000000E+00"
    assert fmt(1.0) == " 0.1000000E+01"
    assert fmt(-0.7126318) == "-0.7126318E+00"
    assert fmt(9.9999996) == " 0.1000000E+02"            # carry into the exponent
    assert fmt(0.99999994) == " 0.9999999E+00" and fmt(0.99999996) == " 0.1000000E+01"
    assert fmt(1.2345678e99) == " 0.1234568+100" and fmt(-1.5e-120) == "-0.1500000-119"   # three-digit exponent: no 'E'
    assert fmt(1.2345678e98) == " 0.1234568E+99"
    assert fmt(float("nan")).strip() == "NaN"


def test_rounding_against_the_binary_from_full_precision_values(fmt):
    """The oracle replays the dense run bit-exactly in its integers and to 1e-11 in FP64: formatting ITS full-precision escape
    records must give the binary's text, except where a value sits within ~1e-11 of a 7-digit rounding boundary."""
    from golden_util import Golden
    from test_oracle_vs_refbin import _run_phase
    g = Golden("dense")
    got = []
    for cyc in (1, 2):
        for phase in ("field", "vol"):
            o, _ = _run_phase(g, cyc, phase)
            got.append(o.events())
    got = np.concatenate(got)[:600]
    lines = (GOLD / "evb_dense.txt").read_text().splitlines()
    same = total = 0
    for rec, l in zip(got, lines):
        for k in range(1, 7):                             # t_bound: stale COMMON value at the lower face
            ref = l[15 * k: 15 * k + 14]
            mine = fmt(rec[k])
            total += 1
            if mine == ref:
                same += 1
            else:                                         # scattered photons differ from the ifort binary at ~1e-7 at worst
                assert abs(float(mine) - float(ref)) <= 2e-6 * abs(float(ref)), (mine, ref)
    assert same >= 0.97 * total, (same, total)


@pytest.fixture()
def emu(monkeypatch):
    import build_emu
    from compton2d_b200 import lib as L
    so = build_emu.build()
    monkeypatch.setattr(L, "LIBPATH", so)
    monkeypatch.setattr(L, "_lib", None)
    yield
    monkeypatch.setattr(L, "_lib", None)


def check_roundtrip(g, P, tmp_path, gpu_step):
    """write_cens -> file in the reference's layout -> read_cens into a fresh bank: 7 digits, integers exact; event file
    parses back to the drained records at 7 digits and is what plcm's fscanf of 7 %lf reads."""
    gpu_step(g, P, 1, field=False)
    t = gpu_step(g, P, 2)
    d, iv, _ = t["census"]
    n = g.write_cens(tmp_path / "p001_census.dat")
    assert n == len(d) > 100
    lines = (tmp_path / "p001_census.dat").read_text().splitlines()
    assert len(lines) == 2 * n and all(len(l) == 84 for l in lines[0::2]) and all(len(l) == 30 for l in lines[1::2])
    dd = np.array([[float(l[14 * k: 14 * k + 14]) for k in range(6)] for l in lines[0::2]])
    ii = np.array([[int(l[5 * k: 5 * k + 5]) for k in range(6)] for l in lines[1::2]])
    assert np.array_equal(ii, iv)
    assert np.allclose(dd, d, rtol=5.1e-7, atol=0)
    g.read_cens(tmp_path / "p001_census.dat", n)
    d2, i2, _ = g.census()
    assert np.array_equal(i2, iv) and np.array_equal(d2, dd)
    from compton2d_b200 import C2dError
    with pytest.raises(C2dError):
        g.read_cens(tmp_path / "p001_census.dat", n + 1)      # file shorter than ndxout says
    # event file: appended step by step like unit 1000+myid
    ev = t["events"]
    m = g.write_events(tmp_path / "p001_evb.dat", append=False)
    m2 = g.write_events(tmp_path / "p001_evb.dat", append=True)
    assert m == m2 == len(ev) > 10
    back = np.loadtxt(tmp_path / "p001_evb.dat", ndmin=2)
    assert back.shape == (2 * m, 7) and np.allclose(back[:m], ev, rtol=5.1e-7, atol=0) and np.array_equal(back[:m], back[m:])
    assert all(len(l) == 104 for l in (tmp_path / "p001_evb.dat").read_text().splitlines())


def test_census_and_event_files_roundtrip_host_logic(emu, tmp_path):
    from parity_util import gpu_step, make_gpu
    from synth import make_problem
    P = make_problem(nz=4, nr=3, nst=1500, ngroups=64, split=(40, 20, 6, 10), n_e=3.0e5)
    g = make_gpu(P)
    check_roundtrip(g, P, tmp_path, gpu_step)
    g.close()
