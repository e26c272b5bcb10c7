"""ntm_tracker_b200 -- B200-native NTM-cell hot path behind the reference's API.

    from ntm_tracker_b200 import NTMCell, LoopNTMTracker

mirrors ``from ntm_cell import NTMCell`` / ``from ntm_tracker_new import
LoopNTMTracker`` of JeffOwOSun/ntm-tracker.  Compute lives in libntm_b200.so
(include/ntm_b200.h); importing this package loads it and fails loudly if it has
not been built.
"""
from . import _cabi

_cabi.load()   # no library -> RuntimeError here, never a silent fallback

from .ntm_cell import NTMCell, random_uniform_initializer  # noqa: E402
from .ntm_tracker_new import LoopNTMTracker  # noqa: E402
from .training import NTMTrainer  # noqa: E402
from .session import ResidentTracker  # noqa: E402

__all__ = ["NTMCell", "LoopNTMTracker", "NTMTrainer", "ResidentTracker", "random_uniform_initializer"]
