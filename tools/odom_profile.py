"""Development: where the C++ odom loop's host time goes (cfg-4 shaped sequence, native loop only)."""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")]
import numpy as np
import bench

n = int(os.environ.get("SCANS", 200))
seq = bench.generate_sequences([(4, n, 0.4, 2, False)], 1)[0]
import ngicp
from ngicp import odom, synth
rng = np.random.default_rng(8)
drift = [synth.random_se3(rng, 0.03, 0.3) for _ in range(n)]
for rep in range(2):
    loop = odom.NativeOdomLoop(bench.configure(ngicp.NanoGICP(0)), odom.OdomParams())
    ts, res = [], []
    bench.drive_loop(loop, seq, drift, 2, 0, n, ts, res)
    t = 1e3 * np.array(ts[3:])
    print("rep", rep, "median %.3f mean %.3f p90 %.3f p99 %.3f max %.3f" % (np.median(t), t.mean(), np.percentile(t, 90), np.percentile(t, 99), t.max()))
    print({k: round(v, 4) if isinstance(v, float) else v for k, v in loop.profile().items()})
    slow = np.argsort(-t)[:8]
    print("slowest:", [(int(i + 3), round(float(t[i]), 2), res[i + 3].new_keyframe, res[i + 3].submap_changed, len(res[i + 3].submap)) for i in slow])
    quiet = [t[i] for i in range(len(t)) if not res[i + 3].new_keyframe and (i + 4 >= len(res) or not res[i + 4].submap_changed)]
    print("scans without keyframe / rebuild: median %.3f mean %.3f (%d)" % (np.median(quiet), np.mean(quiet), len(quiet)))
