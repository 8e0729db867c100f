"""Small driver for ncu: the judged bulk launches — K2/K3 over KEYFRAMES keyframes (BASELINE config 3: 256 x 65,536 points) and K4b over 64 batched scans."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import bench, ngicp
tgt, bounds, scans = bench.make_workload(0)
g = bench.configure(ngicp.NanoGICP(0))
_, m4, _ = g.batchCovariances(tgt, bounds, want_mat4=True)
g.setInputTarget(tgt); g.setTargetCovariances(m4)
hbm = bench.peak_hbm()
out = bench.bulk_covariance_cfg3(g, scans, hbm, list(range(int(os.environ.get("KEYFRAMES", 256)))))
out.update(bench.bulk_linearize(g, scans, hbm))
print({k: v for k, v in out.items() if "ms" in k})
