"""asr_b200 — B200-native (sm_100a) implementation of the octree-conv -> SDF hot
path of adaptive-surface-reconstruction.  Host side only: ctypes binding of
libasr_b200.so plus the mirror of the reference's model/pipeline interfaces."""
from . import _lib, ops  # noqa: F401

__version__ = "0.1.0"
