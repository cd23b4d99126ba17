"""compton2d_b200 -- B200-native (sm_100a CUDA) hot path of Compton2d behind a C ABI.

`lib` loads the in-tree shared library (no CPU fallback); `host.Compton2d` mirrors the
reference's per-step subroutine interface (xec2d.f:141-193)."""
from .host import Compton2d  # noqa: F401
from .lib import C2dError, load  # noqa: F401

__all__ = ["Compton2d", "C2dError", "load"]
