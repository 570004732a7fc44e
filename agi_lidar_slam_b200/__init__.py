"""B200-native S-FAST_LIO hot path (undistort -> voxel downsample -> IESKF update) behind a C-ABI.

The compute lives in liblio_b200.so (hand-written sm_100a kernels, include/lio_b200.h; the reference's C++ call surface on
top of it is include/lio_facade.hpp).  This package is the thin Python host side: ctypes binding (`_cabi`), the reference
main loop on the C-ABI (`replay`), the sharded-map driver (`sharded`), PCD / odometry formats (`formats`) and synthetic data
(`synth`).  There is no CPU fallback: contexts can only be created on a B200-class GPU.
"""
from . import synth  # noqa: F401

__all__ = ["synth"]
