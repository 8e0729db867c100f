"""Development: the slowest K2 work items and what they did (needs a -DNGICP_STATS build)."""
import sys, ctypes
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import bench, ngicp
L = ngicp.lib()
tgt, bounds, scans = bench.make_workload(0)
g = bench.configure(ngicp.NanoGICP(0))
g.setInputSource(scans[0]); g.calculateSourceCovariances()
g.setInputSource(scans[1]); g.calculateSourceCovariances()
buf = (ctypes.c_uint * (8 * 8192))()
L.ngicp_debug_items_knn(buf, 8192)
a = np.array(buf[:], dtype=np.int64).reshape(8192, 8)
names = ["cycles", "passes", "sumM", "maxM", "heavy_rounds", "refused", "M_after_filter", "chunks"]
print("mean", dict(zip(names, a.mean(0).round(1))))
order = np.argsort(-a[:, 0])
for i in order[:12]:
    print(int(i), dict(zip(names, a[i].tolist())))
print("cycles percentiles 50/90/99/99.9/max", np.percentile(a[:, 0], [50, 90, 99, 99.9, 100]).astype(int).tolist())
big = a[:, 0] > 100000
print("items > 100k cycles:", int(big.sum()), "share of all cycles", round(a[big, 0].sum() / a[:, 0].sum(), 3), "mean passes", a[big, 1].mean().round(2), "mean sumM", a[big, 2].mean().round(0), "mean heavy rounds", a[big, 4].mean().round(2))
