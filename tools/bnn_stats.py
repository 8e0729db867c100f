"""Development: statistics of the bounded nearest-neighbour shortcut of K4a (needs a -DNGICP_STATS build)."""
import sys, ctypes
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import bench, ngicp
from ngicp import synth
L = ngicp.lib()
NAMES = ["queries", "?", "refused", "L=base", "L=base+1", "L=base+2", "probes", "points", "point batches", "seed@base", "seed@base+1", "-", "capped", "cycles bounded", "cycles fallback", "items"]
def bnn():
    out = (ctypes.c_ulonglong * 16)()
    L.ngicp_debug_bnn(out, 1)
    return dict(zip(NAMES, list(out)))
def wk():
    out = (ctypes.c_ulonglong * 8)()
    L.ngicp_debug_stats_lin(out, 1)
    return list(out)
tgt, bounds, scans = bench.make_workload(0)
g = bench.configure(ngicp.NanoGICP(0))
g.setInputTarget(tgt); g.calculateTargetCovariances()
keys, lo, h0 = g.target_kdtree_.voxel_keys()
print("target h0", h0)
g.setInputSource(scans[1]); g.calculateSourceCovariances()
bnn(); wk()
g.linearize(np.eye(4))
print("first call :", bnn(), "wknn passes/M/members/calls/refused", wk()[:5])
g.linearize(synth.se3((0, 0, 0.002), (0.01, 0.0, 0.0)))
print("second call:", bnn(), "wknn", wk()[:5])
