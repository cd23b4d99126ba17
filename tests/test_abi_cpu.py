"""CPU-only checks of the drop-in boundary: the CUDA library builds (nvcc cross-compiles without a GPU),
loads, exports every symbol include/compton2d_b200.h declares, and refuses to run without a device
(no CPU fallback).  No compute call is made here."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from compton2d_b200 import build, lib as L
    build.build()
    return L.load()


def test_header_symbols_are_exported(lib):
    hdr = (ROOT / "include" / "compton2d_b200.h").read_text()
    declared = sorted(set(re.findall(r"\b(c2d_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    from compton2d_b200 import lib as L
    assert sorted(L.EXPORTS) == declared


def test_no_cpu_fallback(lib):
    """Without a CUDA device c2d_create must fail with C2D_ENODEV (this box has no GPU)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.c2d_create(C.byref(h), 0)
    assert rc == 1 and b"no CUDA device" in lib.c2d_last_error()
    from compton2d_b200 import C2dError
    from parity_util import make_gpu
    from synth import make_problem
    with pytest.raises(C2dError):
        make_gpu(make_problem(nz=2, nr=2, nst=10))


def test_struct_layouts_match_header(lib):
    """ctypes mirrors of the header structs: field order and count (guards against silent ABI drift)."""
    from compton2d_b200 import lib as L
    hdr = (ROOT / "include" / "compton2d_b200.h").read_text()

    def fields(struct_name):
        end = hdr.index("} " + struct_name + ";")
        body = hdr[hdr.rindex("typedef struct {", 0, end) + len("typedef struct {"): end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            decl = re.sub(r"^(const\s+)?(unsigned\s+)?\w+\s+", "", stmt)
            for part in decl.split(","):
                names.append(re.sub(r"[\*\s]|\[.*\]", "", part))
        return names

    for cls, name in ((L.Dims, "c2d_dims"), (L.StepIn, "c2d_step_in"), (L.Tallies, "c2d_tallies"), (L.FpIn, "c2d_fp_in"),
                      (L.FpOut, "c2d_fp_out"), (L.LcConfig, "c2d_lc_config"), (L.GenIn, "c2d_gen_in"), (L.SurfIn, "c2d_surf_in"),
                      (L.Plasma, "c2d_plasma"), (L.PlasmaOut, "c2d_plasma_out"), (L.SurfRun, "c2d_surf_run"), (L.Run, "c2d_run"),
                      (L.StepReport, "c2d_step_report"), (L.Common, "c2d_common")):
        assert [f[0] for f in cls._fields_] == fields(name), name


def test_pack_helpers_roundtrip(lib):
    """COMMON (padded, column-major) <-> dense [zone][bin] for both index orders (commonblock.f)."""
    jmax, kmax, nz, nr = 7, 6, 4, 3
    rng = np.random.default_rng(0)
    a = rng.random((kmax, jmax, 10))            # Fortran a(10, jmax, kmax): memory order [k][j][i]
    dense = np.zeros((nz * nr, 8))
    lib.c2d_pack_bin_first(a.ctypes.data, 10, jmax, kmax, 8, nz, nr, dense.ctypes.data)
    for j in range(nz):
        for k in range(nr):
            assert np.array_equal(dense[j * nr + k], a[k, j, :8])
    back = np.zeros_like(a)
    lib.c2d_unpack_bin_first(dense.ctypes.data, 10, jmax, kmax, 8, nz, nr, back.ctypes.data)
    assert np.array_equal(back[:nr, :nz, :8], a[:nr, :nz, :8])
    b = rng.random((8, kmax, jmax))             # Fortran b(jmax, kmax, 8): memory order [i][k][j]
    dense2 = np.zeros((nz * nr, 8))
    lib.c2d_pack_bin_last(b.ctypes.data, jmax, kmax, 8, nz, nr, dense2.ctypes.data)
    for j in range(nz):
        for k in range(nr):
            assert np.array_equal(dense2[j * nr + k], b[:, k, j])
    back2 = np.zeros_like(b)
    lib.c2d_unpack_bin_last(dense2.ctypes.data, jmax, kmax, 8, nz, nr, back2.ctypes.data)
    assert np.array_equal(back2[:, :nr, :nz], b[:, :nr, :nz])
