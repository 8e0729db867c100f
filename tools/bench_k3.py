"""Development: bulk covariance build timing (K1/K2/K3 device times) + a parity spot check."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import bench, ngicp, oracle
import scenarios as S
from ngicp import synth
nk = int(sys.argv[1]) if len(sys.argv) > 1 else 64
sc = synth.Scene(0); rng = np.random.default_rng(2)
poses = synth.trajectory(sc, 8, 0, step=1.0)
scans = [synth.transform_points(P, synth.scan(sc, P, rng, keep_all=True)) for P in poses]
g = bench.configure(ngicp.NanoGICP(0))
hbm = bench.peak_hbm()
out = bench.bulk_covariance(g, scans, hbm, n_keyframes=nk)
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in out.items() if k != "roofline_K3"})
print("K3 roofline", out["roofline_K3"])
# parity spot check incl. exact zeros in the coordinates
a, _, _ = S.scan_pair(6, w=128)
a = a.copy(); a[::7, 2] = 0.0; a[::11, 0] = 0.0
gg = S.configure(ngicp.NanoGICP(0)); o = S.configure(oracle.OracleGICP("port"))
gg.setInputSource(a); o.setInputSource(a); gg.calculateSourceCovariances(); o.calculateSourceCovariances()
idx, _ = oracle.KdTree(a, "port").knn(a, 16); ok = S.spectral_gap_ok(a, idx)
print("cov max abs err (gap-ok rows)", float(np.abs(gg.getSourceCovariances() - o.getSourceCovariances())[ok].max()))
