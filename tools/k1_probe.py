"""Development: K1 alone — index build of one 65,536-point scan, repeated; prints index_ms."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import ngicp
from ngicp import synth
sc = synth.Scene(0)
rng = np.random.default_rng(1)
scan = synth.scan(sc, synth.se3((0, 0, 0.1), (1.0, 2.0, 0.0)), rng, keep_all=True)
g = ngicp.NanoGICP(0)
g.enableTiming(True)
ts = []
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    g.timings(reset=True)
    g.setInputSource(scan.copy())
    g.synchronize()
    ts.append(g.timings(reset=True)["index_ms"])
print("n", len(scan), "index_ms", [round(t, 4) for t in ts])
