#!/bin/bash
# Development (GPU box): compute-sanitizer memcheck + racecheck of the leaf search on small clouds
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys
sys.path[:0] = ['.', 'noetic-slam_b200', 'tests']
import numpy as np, ngicp, scenarios as S
a, b, _ = S.scan_pair(3, w=64)
g = S.configure(ngicp.NanoGICP(0))
for k in (5, 16, 20):
    g.setCorrespondenceRandomness(k); g.setInputSource(a.copy()); g.calculateSourceCovariances()
dup = np.concatenate([a, a[::3], a[:40]])
g.setCorrespondenceRandomness(16); g.setInputSource(dup); g.calculateSourceCovariances()
g.setInputTarget(b); g.calculateTargetCovariances(); g.setInputSource(a.copy()); g.calculateSourceCovariances(); T = g.align()
print("ok", g.nr_iterations_)
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san.py > gpurun_out/memcheck.txt 2>&1; tail -6 gpurun_out/memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python /tmp/san.py > gpurun_out/racecheck.txt 2>&1; tail -6 gpurun_out/racecheck.txt
