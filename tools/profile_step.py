"""Small driver for ncu: resident 1M-pt submap, then a few scan registrations (the bench step)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import bench, ngicp

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
tgt, bounds, scans = bench.make_workload(0)
g = bench.configure(ngicp.NanoGICP(0))
_, m4, _ = g.batchCovariances(tgt, bounds, want_mat4=True)
g.setInputTarget(tgt); g.setTargetCovariances(m4)
import ctypes
try:
    _rt = ctypes.CDLL("libcudart.so")
except OSError:
    import torch  # noqa: F401  (loads the CUDA runtime it ships)
    _rt = ctypes.CDLL([m for m in open("/proc/self/maps").read().split() if "libcudart" in m][0])
for i in range(2):          # warm-up outside the profiled range
    g.setInputSource(scans[i % len(scans)].copy()); g.calculateSourceCovariances(); T = g.align()
g.synchronize()
_rt.cudaProfilerStart()
for i in range(steps):
    g.setInputSource(scans[i % len(scans)].copy()); g.calculateSourceCovariances(); T = g.align()
g.synchronize()
_rt.cudaProfilerStop()
print("iters", g.nr_iterations_, "launches", g.timings()["kernel_launches"])
