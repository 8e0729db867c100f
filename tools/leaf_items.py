"""Development (stats build): timeline of the K2 work items of one 65,536-point scan."""
import sys, ctypes
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import bench, ngicp
L = ngicp.lib()
from ngicp import synth
import scenarios as S
which = sys.argv[1] if len(sys.argv) > 1 else "bench"
if which == "bench":
    tgt, bounds, scans = bench.make_workload(0)
    cloud = scans[0]
elif which == "diag":
    sc = synth.Scene(0); rng = np.random.default_rng(2)
    tgt, bounds, poses = synth.make_submap(sc, 100_000, 0, n_keyframes=40)
    T_ws = poses[20] @ synth.se3((0, 0, 0.02), (0.3, 0.1, 0.0))
    cloud = synth.transform_points(T_ws, synth.scan(sc, T_ws, rng, keep_all=True))
else:
    a, _, _ = S.scan_pair(6, w=128)
    cloud = a.copy(); cloud[::7, 2] = 0.0; cloud[::11, 0] = 0.0
print("cloud", which, cloud.shape)
g = bench.configure(ngicp.NanoGICP(0))
for rep in range(3):
    g.setInputSource(cloud.copy()); g.calculateSourceCovariances()
st = (ctypes.c_ulonglong * 16)(); L.ngicp_debug_stats_leaf(st, 1)
g.setInputSource(cloud.copy()); g.calculateSourceCovariances()
L.ngicp_debug_stats_leaf(st, 1)
n = min(int(st[0]), 8192)
buf = np.zeros(4 * 8192, np.uint64)
L.ngicp_debug_leaf_items(buf.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong)), 8192)
it = buf.reshape(-1, 4)[:n]
t0 = it[:, 0].min()
start = (it[:, 0] - t0).astype(np.float64) / 1e3; end = (it[:, 1] - t0).astype(np.float64) / 1e3
dur = end - start
level = (it[:, 2] >> np.uint64(56)).astype(int); members = ((it[:, 2] >> np.uint64(48)) & np.uint64(0xff)).astype(int)
passes = ((it[:, 2] >> np.uint64(40)) & np.uint64(0xff)).astype(int); exact = ((it[:, 2] >> np.uint64(32)) & np.uint64(0xff)).astype(int)
scanned = (it[:, 2] & np.uint64(0xffffffff)).astype(int); maxM = (it[:, 3] >> np.uint64(32)).astype(int)
print(f"items {n}; kernel span {end.max():.1f} us; last start {start.max():.1f} us; duration mean {dur.mean():.1f} median {np.median(dur):.1f} p90 {np.percentile(dur, 90):.1f} p99 {np.percentile(dur, 99):.1f} max {dur.max():.1f} us")
print("start-time percentiles (us):", np.percentile(start, [10, 50, 90, 99]).round(1))
ph = np.zeros(4 * 4096, np.uint32)
L.ngicp_debug_leaf_phase(ph.ctypes.data_as(ctypes.POINTER(ctypes.c_uint)), 4096)
ph = ph.reshape(-1, 4)
order = np.argsort(-end)[:15]
for i in order:
    print(f"  item {i}: start {start[i]:.1f} end {end[i]:.1f} dur {dur[i]:.1f} us level {level[i]} members {members[i]} passes {passes[i]} exact {exact[i]} scanned {scanned[i]} maxM {maxM[i]} pos {int(it[i, 3] & np.uint64(0xffffffff))}"
          + (f" | kcyc probe {ph[i,0]/1e3:.1f} select {ph[i,1]/1e3:.1f} (flush {ph[i,2]/1e3:.1f}, tma wait {ph[i,3]/1e3:.1f}) total {dur[i]*1.9:.0f}" if i < 4096 else ""))
for lv in sorted(set(level.tolist())):
    m = level == lv
    print(f"  level {lv}: items {m.sum()} mean dur {dur[m].mean():.1f} max {dur[m].max():.1f} mean members {members[m].mean():.1f}")

print("passes histogram:", np.bincount(passes)[:12].tolist(), " scanned per item: mean", scanned.mean().round(1), "p50", np.median(scanned), "p99", np.percentile(scanned, 99), "max", scanned.max())
for lo, hi in ((0, 200), (200, 500), (500, 1000), (1000, 2000), (2000, 5000), (5000, 10**9)):
    m = (scanned >= lo) & (scanned < hi)
    if m.any(): print(f"  scanned {lo}-{hi}: items {m.sum()} mean dur {dur[m].mean():.1f} us mean passes {passes[m].mean():.2f}")
