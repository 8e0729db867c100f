"""Development: where the C++ odom loop's host time goes (cfg-4 shaped sequence, native loop only).
PRE=bulk runs a 4M-point batched covariance build first (what bench.py does before its cfg-4 leg); PRE=py runs the Python loop first."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")]
import numpy as np
import bench

n = int(os.environ.get("SCANS", 200))
pre = os.environ.get("PRE", "")
seq = bench.generate_sequences([(4, n, 0.4, 2, False)], 1)[0]
import ngicp
from ngicp import odom, synth
rng = np.random.default_rng(8)
drift = [synth.random_se3(rng, 0.03, 0.3) for _ in range(n)]
if "torch" in pre:
    import torch
    torch.zeros(1 << 20, device="cuda:0").sum().item()
if "bulk" in pre:
    g0 = bench.configure(ngicp.NanoGICP(0))
    pts = np.concatenate([np.stack([s[0]["x"], s[0]["y"], s[0]["z"]], 1) for s in seq[:64]])
    pts = pts[np.isfinite(pts).all(1)]
    m = len(pts) // 64
    g0.batchCovariances(pts[:64 * m], np.arange(65, dtype=np.int64) * m)
    print("bulk done", len(pts))
if "py" in pre:
    lp = odom.OdomLoop(odom.DeviceBackend(bench.configure(ngicp.NanoGICP(0))), odom.OdomParams())
    bench.drive_loop(lp, seq, drift, 2, 0, min(n, 100))
for rep in range(2):
    loop = odom.NativeOdomLoop(bench.configure(ngicp.NanoGICP(0)), odom.OdomParams())
    ts, res, st = [], [], []
    for i in range(n):
        bench.drive_loop(loop, seq, drift, 2, i, i + 1, ts, res)
        st.append(loop.profile(reset=True))
    t = 1e3 * np.array(ts[3:])
    print("rep", rep, "median %.3f mean %.3f p90 %.3f p99 %.3f max %.3f" % (np.median(t), t.mean(), np.percentile(t, 90), np.percentile(t, 99), t.max()))
    print({k: round(float(np.median([s[k] for s in st[3:]])), 4) for k in loop.STAGES})
    slow = np.argsort(-t)[:4]
    print("slowest:", [(int(i + 3), round(float(t[i]), 2), res[i + 3].new_keyframe, {k: round(v, 2) for k, v in st[i + 3].items() if k != "scans" and v > 0.3}) for i in slow])
