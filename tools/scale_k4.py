"""Development: K4a/K4b/K5 device time vs number of source points (same 1M-point target)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import bench, ngicp
tgt, bounds, scans = bench.make_workload(0)
g = bench.configure(ngicp.NanoGICP(0))
g.setInputTarget(tgt); g.calculateTargetCovariances()
g.enableTiming(True)
rng = np.random.default_rng(0)
for n in (512, 4096, 16384, 65536):
    src = scans[0][np.sort(rng.choice(65536, n, replace=False))]
    g.setInputSource(src); g.calculateSourceCovariances()
    for rep in range(3):
        g.timings(reset=True)
        T = np.eye(4); T[0, 3] = 0.001 * rep
        g.linearize(T); g.compute_error(T)
        t = g.timings(reset=True)
        print(f"n={n:6d} rep{rep}: correspond {1e3*t['correspond_ms']:.1f} us  linearise {1e3*t['linearize_ms']:.1f} us  error {1e3*t['error_ms']:.1f} us")
