"""Development: where the wall time of one callbackPointCloud goes on the device path (per backend call + host policy)."""
import sys, time, collections
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import ngicp
import scenarios as S
from ngicp import odom, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
groups = 2
scene = synth.Scene(4)
poses = odom.synthetic_poses(scene, n, 4, 0.4)
def make(i): return odom.synthetic_scan(scene, poses, i, 4, 1024, groups)
import multiprocessing as mp, os
with mp.get_context("fork").Pool(min(16, os.cpu_count() or 1)) as pool:
    seq = pool.map(make, range(n), chunksize=4)
import scipy.spatial  # noqa
rng = np.random.default_rng(8)
drift = [synth.random_se3(rng, 0.03, 0.3) for _ in range(n)]
acc = collections.defaultdict(float); cnt = collections.defaultdict(int); mx = collections.defaultdict(float)
class Timed(odom.DeviceBackend):
    pass
for name in ("set_max_correspondence_distance", "ingest", "deskew_filter_set_source", "calculate_source_covariances", "align", "capture_keyframe", "transform_keyframe", "set_submap"):
    def wrap(f, name=name):
        def w(self, *a, **k):
            t = time.perf_counter(); r = f(self, *a, **k); d = time.perf_counter() - t; acc[name] += d; cnt[name] += 1; mx[name] = max(mx[name], d); return r
        return w
    setattr(Timed, name, wrap(getattr(odom.DeviceBackend, name)))
g = S.configure(ngicp.NanoGICP(0), max_corr=0.5, max_iter=32, rot_eps=0.01, trans_eps=0.01)
loop = odom.OdomLoop(Timed(g), odom.OdomParams())
tot = 0.0
for i, (rec, Ts, block, col_t) in enumerate(seq):
    if i == 3: acc.clear(); cnt.clear(); mx.clear(); tot = 0.0
    t = time.perf_counter()
    if i == 0:
        loop.T = Ts[groups // 2].astype(np.float32); loop.propagateGICP(); loop.callbackPointCloud(rec, None)
    else:
        def prior(stamps, Ts=Ts, i=i):
            k = np.minimum((stamps.astype(np.int64) * groups) // 100_000_000, groups - 1)
            return (drift[i] @ Ts)[k].astype(np.float32)
        loop.callbackPointCloud(rec, prior)
    tot += time.perf_counter() - t
m = n - 3
print("ms per scan: total %.3f" % (1e3 * tot / m), {k: round(1e3 * v / m, 3) for k, v in acc.items()}, "host policy %.3f" % (1e3 * (tot - sum(acc.values())) / m))
print("calls", dict(cnt), "max ms", {k: round(1e3 * v, 2) for k, v in mx.items()}, "keyframes", len(loop.keyframes))
