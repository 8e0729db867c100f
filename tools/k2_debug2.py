"""Development (stats build): records of lanes that finish with an incomplete or repeated list."""
import sys, ctypes
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import ngicp
import scenarios as S
L = ngicp.lib()
a, _, _ = S.scan_pair(3, w=128)
for k in (5, 8):
    g = S.configure(ngicp.NanoGICP(0), k=k)
    g.setInputSource(a)
    gi, dens = g.selfNeighbours(0, k)
    rec = np.zeros((64, 24), np.uint32); n = ctypes.c_uint(0)
    L.ngicp_debug_leaf_records(rec.ctypes.data_as(ctypes.POINTER(ctypes.c_uint)), ctypes.byref(n))
    print(f"k={k}: {n.value} records")
    for r in rec[:min(n.value, 64)]:
        print(f" j(pos) {r[0]} Lg {r[1]} M {r[2]} sb {r[3]} bound {r[4]:#x} ({np.uint32(min(r[4], 0x7f800000)).view(np.float32)}) thr {r[5]:#x} beff {r[6]:#x} next {r[7]:#x} R {r[8]} first {r[9]} "
              f"item.start {r[10]} count {r[11] >> 8} level {r[11] & 255} tau {r[12]:#x} members {r[13]:#x}")
        print("    dl", [f"{x:#x}" for x in r[14:22]], [float(np.uint32(x & ~np.uint32((1 << int(r[3])) - 1)).view(np.float32)) if x < 0x7f800000 else None for x in r[14:22]])
