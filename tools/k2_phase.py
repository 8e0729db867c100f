"""Development: where K2 warps spend their cycles (needs a -DNGICP_STATS build)."""
import sys, ctypes
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import bench, ngicp
L = ngicp.lib()
def phase():
    out = (ctypes.c_ulonglong * 40)()
    L.ngicp_debug_phase_knn(out, 1)
    return list(out)
tgt, bounds, scans = bench.make_workload(0)
g = bench.configure(ngicp.NanoGICP(0))
g.setInputSource(scans[0]); g.calculateSourceCovariances()
for i in (1, 2):
    phase()
    g.setInputSource(scans[i]); g.calculateSourceCovariances()
    p = phase()
    tot = p[2]
    print(f"scan {i}: cycles/warp {tot // 8192}  heavy rounds {100 * p[0] / tot:.1f}%  shared rounds {100 * p[1] / tot:.1f}%")
    print("   per-warp cycles histogram:", {f"<2^{b + 1}": p[8 + b] for b in range(32) if p[8 + b]})
