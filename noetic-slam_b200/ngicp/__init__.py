"""ngicp — host-side mirror of the reference's nano_gicp interface over libngicp_b200.so (CUDA, sm_100a).

The repo directory is `noetic-slam_b200/` (not an importable name), so callers put it on sys.path:
    sys.path.insert(0, "<repo>/noetic-slam_b200"); import ngicp
"""
from .binding import (NgicpError, REG_FROBENIUS, REG_MIN_EIG, REG_NONE, REG_NORMALIZED_MIN_EIG, REG_PLANE, SOURCE, TARGET,
                      LIB_PATH, build, lib)
from .gicp import CropBox, KdTreeFLANN, NanoGICP, VoxelGrid

__all__ = ["NanoGICP", "KdTreeFLANN", "CropBox", "VoxelGrid", "NgicpError", "build", "lib", "LIB_PATH", "REG_NONE", "REG_MIN_EIG",
           "REG_NORMALIZED_MIN_EIG", "REG_PLANE", "REG_FROBENIUS", "SOURCE", "TARGET"]
