"""Development: start / end time of every K2 work item of one 65,536-point scan (needs the -DNGICP_STATS build:
make -C noetic-slam_b200/csrc BUILD=build_stats OUT=../libngicp_b200_stats.so EXTRA=-DNGICP_STATS; run with
NGICP_LIB=noetic-slam_b200/libngicp_b200_stats.so)."""
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
for rep in range(3):
    g.setInputSource(scans[rep]); g.calculateSourceCovariances()
n_items = 8192
buf = (ctypes.c_ulonglong * (4 * n_items))()
L.ngicp_debug_items_knn(buf, n_items)
a = np.array(buf[:], dtype=np.int64).reshape(n_items, 4)
passes, maxM, sumM = a[:, 2] >> 32, a[:, 2] & 0xffffffff, a[:, 3]
t0 = a[:, 0].min()
start, end = (a[:, 0] - t0) / 1e3, (a[:, 1] - t0) / 1e3     # us
dur = end - start
print("kernel span %.1f us; item duration us: mean %.2f median %.2f p90 %.2f p99 %.2f max %.2f" % (end.max(), dur.mean(), np.median(dur), *np.percentile(dur, [90, 99]), dur.max()))
print("sum of durations %.0f us -> %.1f us at 148*16 concurrent" % (dur.sum(), dur.sum() / (148 * 16)))
order = np.argsort(-end)
print("last items to finish: (index, start, duration)", [(int(i), round(float(start[i]), 1), round(float(dur[i]), 1)) for i in order[:10]])
slow = np.argsort(-dur)[:20]
print("slowest items: (index, start, duration, passes, sumM, maxM)", [(int(i), round(float(start[i]), 1), round(float(dur[i]), 1), int(passes[i]), int(sumM[i]), int(maxM[i])) for i in slow])
print("all items: passes mean %.2f, sumM mean %.0f, maxM mean %.0f; corr(dur, sumM) %.3f corr(dur, passes) %.3f corr(dur, maxM) %.3f" % (passes.mean(), sumM.mean(), maxM.mean(), np.corrcoef(dur, sumM)[0, 1], np.corrcoef(dur, passes)[0, 1], np.corrcoef(dur, maxM)[0, 1]))
for lo, hi in [(0, 15), (15, 25), (25, 40), (40, 60), (60, 200)]:
    m = (dur >= lo) & (dur < hi)
    if m.any(): print("  dur [%d,%d) us: %d items, passes %.1f, sumM %.0f, maxM %.0f" % (lo, hi, m.sum(), passes[m].mean(), sumM[m].mean(), maxM[m].mean()))
# how many items are running over time
ts = np.linspace(0, end.max(), 12)
print("running items at t:", [(round(float(t), 0), int(((start <= t) & (end > t)).sum())) for t in ts])
# heavy items by index decile
dec = np.array_split(np.arange(n_items), 16)
print("mean duration by index 16-quantile:", [round(float(dur[d].mean()), 1) for d in dec])
