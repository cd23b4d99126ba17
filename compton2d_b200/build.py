"""Build libcompton2d_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libcompton2d_b200.so"
SOURCES = ["c2d_track.cu", "c2d_abi.cu", "c2d_fp.cu", "c2d_sigma.cu", "c2d_gen.cu", "c2d_lc.cu", "c2d_io.cu", "c2d_step.cu", "c2d_comm.cu", "c2d_bank.cu", "c2d_common.cu"]
# -fmad=false: the tracker replays against the CPU oracle decision by decision; contraction into FMAs
# is enabled only where a kernel asks for it explicitly (fma()).
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
         "-Xcompiler", "-fPIC", "-shared", "-ldl"]


def build(force: bool = False, verbose: bool = False) -> Path:
    srcs = [CSRC / s for s in SOURCES]
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + \
        [HERE.parent / "include" / "compton2d_b200.h"]
    if not force and LIB.exists() and LIB.stat().st_mtime >= max(p.stat().st_mtime for p in deps):
        return LIB
    extra = os.environ.get("C2D_NVCC_EXTRA", "").split()       # experiments, e.g. -DC2D_HIST_MINB=5
    out = Path(os.environ.get("C2D_LIB_OUT", LIB))              # ... built next to the product library (C2D_LIBPATH selects it)
    cmd = ["nvcc", *FLAGS, *extra, *(["-Xptxas", "-v"] if verbose else []), "-o", str(out), *map(str, srcs)]
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
