"""SURVEY.md 8f N3 on the device: the census bank and the escape records through the reference's own file formats
(csrc/c2d_io.cu; formatter pinned to the reference binary in tests/test_record_files_cpu.py)."""
import pytest

from parity_util import gpu_step, make_gpu
from synth import make_problem
from test_record_files_cpu import check_roundtrip

pytestmark = pytest.mark.gpu


def test_census_and_event_files_roundtrip(tmp_path):
    P = make_problem(nz=4, nr=3, nst=20000, ngroups=64, split=(40, 20, 6, 10), n_e=3.0e5)
    g = make_gpu(P)
    check_roundtrip(g, P, tmp_path, gpu_step)
    g.close()
