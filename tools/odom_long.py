"""BASELINE config 4 at reduced length: a seeded OS1-64 sequence (65,536-point scans, sensor moving during the scan) through
the odom loop over the CUDA path and over the CPU oracle; prints agreement and timing.
usage: python tools/odom_long.py [n_scans=120] [step_m=0.5]"""
import sys, time, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import ngicp, oracle
import scenarios as S
from ngicp import odom, synth
from odom_backends import OracleBackend

n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
step = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
groups = 2
scene = synth.Scene(4)
t0 = time.time()
poses = odom.synthetic_poses(scene, n, 4, step)


def make(i):
    return odom.synthetic_scan(scene, poses, i, 4, 1024, groups)


import multiprocessing as mp, os
with mp.get_context("fork").Pool(min(16, os.cpu_count() or 1)) as pool:      # before any CUDA context exists
    seq = pool.map(make, range(n), chunksize=4)
print(f"generated {n} scans in {time.time() - t0:.1f}s", flush=True)
rng = np.random.default_rng(8)
drift = [synth.random_se3(rng, 0.03, 0.3) for _ in range(n)]


def run(backend):
    loop = odom.OdomLoop(backend, odom.OdomParams())
    res, ts = [], []
    for i, (rec, Ts, block, col_t) in enumerate(seq):
        t = time.perf_counter()
        if i == 0:
            loop.T = Ts[groups // 2].astype(np.float32)
            loop.propagateGICP()
            res.append(loop.callbackPointCloud(rec, None))
        else:
            def prior(stamps, Ts=Ts, i=i):
                k = np.minimum((stamps.astype(np.int64) * groups) // 100_000_000, groups - 1)
                return (drift[i] @ Ts)[k].astype(np.float32)
            res.append(loop.callbackPointCloud(rec, prior))
        ts.append(time.perf_counter() - t)
    return loop, res, ts


import scipy.spatial  # noqa: F401  (the hull code imports it lazily; keep the one-off import out of the per-scan times)
g = S.configure(ngicp.NanoGICP(0), max_corr=0.5, max_iter=32, rot_eps=0.01, trans_eps=0.01)
lg, rg, tg = run(odom.DeviceBackend(g))
o = S.configure(oracle.OracleGICP("ref" if oracle.available("ref") else "port"), max_corr=0.5, max_iter=32, rot_eps=0.01, trans_eps=0.01)
lo, ro, to = run(OracleBackend(o))
same_kf = [a.new_keyframe == b.new_keyframe for a, b in zip(rg, ro)]
same_sub = [a.submap == b.submap for a, b in zip(rg, ro)]
same_it = [a.iterations == b.iterations and a.converged == b.converged for a, b in zip(rg, ro)]
dt = max(float(np.abs(a.T[:3, 3] - b.T[:3, 3]).max()) for a, b in zip(rg, ro))
dr = max(float(np.abs(a.T[:3, :3] - b.T[:3, :3]).max()) for a, b in zip(rg, ro))
err = max(float(np.abs(r.T[:3, 3] - s[1][groups // 2][:3, 3]).max()) for r, s in zip(rg[1:], seq[1:]))
out = {"scans": n, "points_per_scan": 65536, "path_m": n * step, "keyframes": len(lg.keyframes), "keyframes_oracle": len(lo.keyframes),
       "max_submap_keyframes": max(len(r.submap) for r in rg), "same_keyframe_decisions": all(same_kf), "same_submap_sets": all(same_sub),
       "same_iterations_and_convergence": sum(same_it), "max_pose_diff_m": dt, "max_rotation_entry_diff": dr,
       "max_position_error_vs_truth_m": err, "device_ms_per_scan_median": 1e3 * float(np.median(tg[3:])), "device_ms_per_scan_mean": 1e3 * float(np.mean(tg[3:])),
       "oracle_ms_per_scan_median": 1e3 * float(np.median(to[3:])), "oracle_threads": o.p.get("num_threads", 0) if hasattr(o, "p") else None}
print(json.dumps(out))
