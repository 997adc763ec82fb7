"""rtsds_b200 — B200-native (sm_100a) kernels and execution plans for the RTSDS
segmentation hot path.  Importing the package does not touch the GPU; the
shared library is loaded (and built in-tree if absent) on first use."""
from ._lib import RtsdsError, build, lib  # noqa: F401

__version__ = "0.1.0"
