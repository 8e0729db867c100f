"""Development: pass / candidate statistics of the warp-cooperative search (needs a -DNGICP_STATS build)."""
import sys, ctypes
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import bench, ngicp
L = ngicp.lib()
def stats(which, reset=True):
    out = (ctypes.c_ulonglong * 8)()
    getattr(L, f"ngicp_debug_stats_{which}")(out, int(reset))
    return list(out)
def show(tag, s, nq):
    passes, M, members, calls, refused = s[:5]
    print(f"{tag}: warps {calls} passes/warp {passes/max(calls,1):.2f} cand/pass {M/max(passes,1):.1f} members/pass {members/max(passes,1):.1f} "
          f"refused {refused} ({100*refused/max(nq,1):.2f}% of queries) cand/warp {M*1.0/max(calls,1):.0f}; maxM {s[5]} passes>2048: {s[6]} holding {100*s[7]/max(M,1):.1f}% of all candidates")
tgt, bounds, scans = bench.make_workload(0)
g = bench.configure(ngicp.NanoGICP(0))
stats("knn"); stats("lin")
g.setInputTarget(tgt); g.calculateTargetCovariances()
show("K2 target 1M", stats("knn"), len(tgt))
g.setInputSource(scans[0]); g.calculateSourceCovariances()
show("K2 source 65k raw", stats("knn"), 65536)
e, H, b = g.linearize(np.eye(4))
show("K4 65k vs 1M", stats("lin"), 65536)
from ngicp import synth
v = synth.voxel_filter(scans[0])
g.setInputSource(v); g.calculateSourceCovariances()
show(f"K2 source voxel-filtered {len(v)}", stats("knn"), len(v))
# per-item / per-warp time distribution of the correspondence kernel (cycles, log2 buckets)
def hist():
    out = (ctypes.c_ulonglong * 32)()
    L.ngicp_debug_lin_hist(out, 1)
    return list(out)
hist()
T = np.eye(4)
g.setInputSource(scans[1]); g.calculateSourceCovariances()
stats("lin")
e, H, b = g.linearize(T)
h1 = hist()
show("K4a first call", stats("lin"), 65536)
print("item cycles  :", {f"<2^{10+i}": h1[i] for i in range(16) if h1[i]})
print("warp cycles  :", {f"<2^{10+i}": h1[16+i] for i in range(16) if h1[16+i]})
e, H, b = g.linearize(synth.se3((0,0,0.002),(0.01,0.0,0.0)))
h2 = hist()
show("K4a second call (prev hints)", stats("lin"), 65536)
print("item cycles  :", {f"<2^{10+i}": h2[i] for i in range(16) if h2[i]})
print("warp cycles  :", {f"<2^{10+i}": h2[16+i] for i in range(16) if h2[16+i]})
# which work items are the slow ones?
g.setInputSource(scans[2]); g.calculateSourceCovariances()
e, H, b = g.linearize(np.eye(4))
ic = (ctypes.c_uint * 8192)()
L.ngicp_debug_item_cycles(ic, 8192)
ic = np.array(ic[:], dtype=np.int64)
order = np.argsort(-ic)[:12]
print("slowest items (cycles):", ic[order].tolist(), "median", int(np.median(ic)), "sum/2368warps", int(ic.sum() / 3552))
tree = g.source_kdtree_
# sorted order of the source = Morton order; recover the points of the slow items through the public kNN on the source tree
keys, lo, h0 = tree.voxel_keys()
perm = np.argsort(keys, kind="stable")          # approximately the device order (sorted on 27 bits, stable)
for it in order[:6]:
    pts_it = scans[2][perm[it * 8:(it + 1) * 8]]
    d_t, _ = g.target_kdtree_.nearestKSearch(pts_it, 1)[1], None
    print(f" item {it}: cycles {ic[it]} centroid {pts_it.mean(0).round(2).tolist()} spread {np.ptp(pts_it, 0).round(2).tolist()} nn-dist {np.sqrt(d_t[:,0]).round(3).tolist()}")
print("target bbox", tgt.min(0).round(1).tolist(), tgt.max(0).round(1).tolist())
