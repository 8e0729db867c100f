"""Development: rows where the production K2 differs from the oracle, with distances (run on the GPU box)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import ngicp, oracle
import scenarios as S
from test_gpu_k2 import ref_sqdist
for seed, w, k in ((3, 128, 5), (3, 128, 8), (3, 128, 16), (3, 128, 20), (3, 128, 32)):
    a, _, _ = S.scan_pair(seed, w=w)
    g = S.configure(ngicp.NanoGICP(0), k=k)
    g.setInputSource(a)
    gi, dens = g.selfNeighbours(0, k)
    gd = ref_sqdist(a, a[gi])
    oi, od = oracle.KdTree(a, "port").knn(a, k)
    gi_c, gd_c = S.canonical_rows(gi, gd)
    oi_c, od_c = S.canonical_rows(oi.astype(np.int32), od)
    ex, tie, bad = S.knn_rows_equivalent(gi_c, gd_c, oi_c, od_c)
    print(f"k={k} n={len(a)} exact {ex} tie {tie} bad {bad}", flush=True)
    rows = np.nonzero(~((gi_c == oi_c).all(1) & (gd_c == od_c).all(1)))[0][:6]
    for r in rows:
        print(" row", r, "self first:", gi[r, 0] == r, "dups in row:", k - len(set(gi[r].tolist())))
        print("   gpu idx", gi_c[r].tolist()); print("   gpu d  ", [f"{x:.9g}" for x in gd_c[r]])
        print("   orc idx", oi_c[r].tolist()); print("   orc d  ", [f"{x:.9g}" for x in od_c[r]])
        miss = sorted(set(oi_c[r].tolist()) - set(gi_c[r].tolist())); extra = sorted(set(gi_c[r].tolist()) - set(oi_c[r].tolist()))
        print("   missing", miss, [f"{float(ref_sqdist(a[r:r+1], a[None, [m]])[0,0]):.9g} bits {ref_sqdist(a[r:r+1], a[None, [m]]).view(np.uint32)[0,0]:#x}" for m in miss],
              "extra", extra, [f"{float(ref_sqdist(a[r:r+1], a[None, [m]])[0,0]):.9g} bits {ref_sqdist(a[r:r+1], a[None, [m]]).view(np.uint32)[0,0]:#x}" for m in extra])
