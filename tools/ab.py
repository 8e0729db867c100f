"""Development: A/B of the judged bulk kernels under the NGICP_* development switches (one process per variant).
usage: python tools/ab.py k3|k4b   -> one line with the kernel time and, for k3, the parity spot check."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import bench, ngicp, oracle
import scenarios as S
from ngicp import synth
env = {k: v for k, v in os.environ.items() if k.startswith("NGICP_")}
hbm = bench.peak_hbm()
g = bench.configure(ngicp.NanoGICP(0))
if sys.argv[1] == "k3":
    sc = synth.Scene(0); rng = np.random.default_rng(2)
    poses = synth.trajectory(sc, 8, 0, step=1.0)
    scans = [synth.transform_points(P, synth.scan(sc, P, rng, keep_all=True)) for P in poses]
    out = bench.bulk_covariance(g, scans, hbm, n_keyframes=64)
    a, _, _ = S.scan_pair(6, w=128)
    a = a.copy(); a[::7, 2] = 0.0; a[::11, 0] = 0.0
    gg = S.configure(ngicp.NanoGICP(0)); o = S.configure(oracle.OracleGICP("port"))
    gg.setInputSource(a); o.setInputSource(a); gg.calculateSourceCovariances(); o.calculateSourceCovariances()
    idx, _ = oracle.KdTree(a, "port").knn(a, 16); ok = S.spectral_gap_ok(a, idx)
    err = np.abs(gg.getSourceCovariances() - o.getSourceCovariances())
    print("K3", env, "index_ms", round(out["index_ms"], 3), "knn_ms", round(out["knn_ms"], 3), "cov_ms", round(out["covariance_ms"], 4), "frac", round(out["roofline_K3"]["frac"], 3),
          "err(gap-ok)", float(err[ok].max()), "err(all)", float(err.max()), "rows", int(ok.sum()), "/", len(ok))
else:
    tgt, bounds, scans = bench.make_workload(0)
    g.setInputTarget(tgt); g.calculateTargetCovariances()
    out = bench.bulk_linearize(g, scans, hbm)
    print("K4b", env, "lin_ms", round(out["batch_linearize_ms"], 4), "frac", round(out["roofline_K4b"]["frac"], 3), "corr_ms", round(out["batch_correspond_ms"], 3),
          "matched", round(out["batch_matched_fraction"], 3))
