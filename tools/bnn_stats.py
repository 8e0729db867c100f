"""Development: statistics of the bounded nearest-neighbour search of K4a (needs a -DNGICP_STATS build)."""
import sys, ctypes
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import bench, ngicp
L = ngicp.lib()
NAMES = ["fast queries", "heavy: r>2h", "heavy: >16 cells", "-", "hinted", "no bound", "warp queries", "L=base", "L=base+1", "L=base+2",
         "declined", "sum M", "sum R", "sum cells", "-", "-"]
def bnn():
    out = (ctypes.c_ulonglong * 16)()
    L.ngicp_debug_bnn(out, 1)
    return {k: v for k, v in zip(NAMES, list(out)) if k != "-"}
tgt, bounds, scans = bench.make_workload(0)
g = bench.configure(ngicp.NanoGICP(0))
g.setInputTarget(tgt); g.calculateTargetCovariances()
for i in range(3):
    g.setInputSource(scans[i]); g.calculateSourceCovariances()
    bnn()
    g.linearize(np.eye(4))
    print("first linearize:", bnn())
    g.align()
    t = g.timings()
    print("align          :", bnn())
