"""TEST INFRASTRUCTURE ONLY: compile the CUDA sources of compton2d_b200/csrc with g++ against
tests/emu/fake/cuda_runtime.h into tests/emu/libc2d_emu.so, so host logic and kernel control flow can
be debugged on a box without a GPU.  Never loaded by the product (compton2d_b200/lib.py only ever
opens its own in-tree CUDA library)."""
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
LIB = HERE / "libc2d_emu.so"


def build(force=False):
    srcs = [HERE / "emu_launch.cpp", ROOT / "compton2d_b200/csrc/c2d_abi.cu", ROOT / "compton2d_b200/csrc/c2d_fp.cu",
            ROOT / "compton2d_b200/csrc/c2d_lc.cu", ROOT / "compton2d_b200/csrc/c2d_io.cu", ROOT / "compton2d_b200/csrc/c2d_step.cu",
            ROOT / "compton2d_b200/csrc/c2d_comm.cu", ROOT / "compton2d_b200/csrc/c2d_bank.cu",
            ROOT / "compton2d_b200/csrc/c2d_common.cu"]
    deps = srcs + list((ROOT / "compton2d_b200/csrc").glob("*")) + [HERE / "fake/cuda_runtime.h"]
    if not force and LIB.exists() and LIB.stat().st_mtime >= max(p.stat().st_mtime for p in deps):
        return LIB
    cmd = ["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-w", f"-I{HERE / 'fake'}",
           "-x", "c++", *map(str, srcs), "-o", str(LIB)]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
