"""Development: where the host-buffer path spends its time (pageable vs page-locked scan buffers)."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import bench, ngicp
from ngicp import synth
tgt, bounds, scans = bench.make_workload(0)
g = bench.configure(ngicp.NanoGICP(0))
_, m4, _ = g.batchCovariances(tgt, bounds, want_mat4=True)
g.setInputTarget(tgt); g.setTargetCovariances(m4)
h_scans = [synth.to_aos32(s) for s in scans]
pins = [torch.empty(h_scans[0].shape, dtype=torch.float32).pin_memory() for _ in range(2)]
def run(kind, reps=30):
    ts = {"set": 0.0, "cov": 0.0, "align": 0.0}
    for i in range(reps + 3):
        src = h_scans[i % len(h_scans)]
        if kind == "pinned":
            buf = pins[i % 2].numpy(); buf[:] = src; hs = buf
        elif kind == "packed12":
            hs = np.ascontiguousarray(src[:, :3])
        else:
            hs = src.copy()
        t0 = time.perf_counter(); g.setInputSource(hs)
        t1 = time.perf_counter(); g.calculateSourceCovariances()
        t2 = time.perf_counter(); g.align()
        t3 = time.perf_counter()
        if i >= 3:
            ts["set"] += t1 - t0; ts["cov"] += t2 - t1; ts["align"] += t3 - t2
    print(kind, {k: round(v / reps * 1e3, 4) for k, v in ts.items()}, "total ms", round(sum(ts.values()) / reps * 1e3, 4))
for kind in ("pageable32", "packed12", "pinned", "pageable32", "pinned"):
    run(kind)
