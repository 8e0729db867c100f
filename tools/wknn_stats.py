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
