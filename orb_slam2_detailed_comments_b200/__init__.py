"""B200-native ORB feature front-end: drop-in for ORB-SLAM2's ORBextractor / ORBmatcher hot path.

The compute path is hand-written sm_100a CUDA behind the C ABI of include/orb_b200.h
(lib/liborb_b200.so). This package is the thin host-side mirror of the reference's interface.
"""
from ._lib import KP_DTYPE, LIB_PATH, OrbError, lib  # noqa: F401
from .extractor import ORBextractor  # noqa: F401
from .matcher import FrameView, ORBmatcher, int_pipe_peak  # noqa: F401
from . import frame  # noqa: F401,E402
from . import search  # noqa: F401,E402
from . import input  # noqa: F401,E402
