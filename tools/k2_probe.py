"""Development: K2 alone on the bench's eight scans (source side of a step): mean / max knn_ms over the scans."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import bench, ngicp
scans = np.load("/tmp/k2_scans.npy") if Path("/tmp/k2_scans.npy").exists() else None
if scans is None:
    _, _, sc = bench.make_workload(0)
    scans = np.stack(sc)
    np.save("/tmp/k2_scans.npy", scans)
g = bench.configure(ngicp.NanoGICP(0))
g.enableTiming(True)
ts = []
for rep in range(3):
    for s in scans:
        g.timings(reset=True)
        g.setInputSource(s.copy()); g.calculateSourceCovariances(); g.synchronize()
        t = g.timings(reset=True)
        if rep:
            ts.append(t["knn_ms"])
print("knn_ms mean %.4f min %.4f max %.4f" % (np.mean(ts), np.min(ts), np.max(ts)))
