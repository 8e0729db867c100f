"""Development: pass / candidate statistics of the leaf-scheduled K2 (needs a -DNGICP_STATS build:
make -C noetic-slam_b200/csrc BUILD=build_stats OUT=../libngicp_b200_stats.so EXTRA=-DNGICP_STATS; NGICP_LIB=.../libngicp_b200_stats.so)."""
import sys, ctypes
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import bench, ngicp
import scenarios as S
from ngicp import synth
L = ngicp.lib()
NAMES = ["items", "passes", "cand_staged", "member_lanes", "lanes_fast", "lanes_exact", "passes_exact_resel", "refused_lanes", "prune_calls",
         "cand_before_prune", "cand_after_prune", "flushes", "max_M", "passes_streaming", "tie_redos"]
def stats(reset=True):
    out = (ctypes.c_ulonglong * 16)()
    L.ngicp_debug_stats_leaf(out, int(reset))
    return dict(zip(NAMES, list(out)))
def show(tag, n):
    s = stats()
    it, ps = max(s["items"], 1), max(s["passes"], 1)
    print(f"{tag}: n={n} items {s['items']} ({n/it:.1f} pts/item) passes/item {ps/it:.3f} cand/pass {s['cand_staged']/ps:.1f} lanes/pass {s['member_lanes']/ps:.1f} "
          f"flushes/pass {s['flushes']/ps:.2f} fast lanes {s['lanes_fast']} exact lanes {s['lanes_exact']} ({100*s['lanes_exact']/max(n,1):.3f}%) exact passes {s['passes_exact_resel']} "
          f"({100*s['passes_exact_resel']/ps:.2f}%) refused {s['refused_lanes']} prune {s['prune_calls']} ({s['cand_before_prune']}->{s['cand_after_prune']}) "
          f"maxM {s['max_M']} streaming passes {s['passes_streaming']} tie redos {s['tie_redos']}", flush=True)
tgt, bounds, scans = bench.make_workload(0)
g = bench.configure(ngicp.NanoGICP(0))
stats()
g.setInputSource(scans[0]); g.calculateSourceCovariances(); show("raw scan k=16", 65536)
v = synth.voxel_filter(scans[0]); g.setInputSource(v); g.calculateSourceCovariances(); show("voxel-filtered k=16", len(v))
kf = tgt[bounds[3]:bounds[4]]; g.setInputSource(kf); g.calculateSourceCovariances(); show("keyframe k=16", len(kf))
a, _, _ = S.scan_pair(3, w=128)
for k in (5, 20):
    g.setCorrespondenceRandomness(k); g.setInputSource(a.copy()); g.calculateSourceCovariances(); show(f"scan_pair k={k}", len(a))
