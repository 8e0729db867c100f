"""Development: the cfg-5 leg alone (8 MulRan-shaped sequences through the C++ loop on one GPU), repeated."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import numpy as np
import bench
specs = [(100 + k, bench.CFG5_SCANS, 0.4, 1, True) for k in range(8)]
seqs = bench.generate_sequences(specs, 1)
import torch
torch.cuda.set_device(0)
out = []
for rep in range(int(os.environ.get("REPS", 4))):
    r = bench.multi_sequence_cfg5(0, 0, 1, seqs, list(range(8)), lambda: torch.cuda.synchronize(), lambda t, u: (t, u))
    out.append(round(r["scans_per_s"]))
print("cfg5 scans/s", out, "K1_CLUSTER", os.environ.get("NGICP_K1_CLUSTER", "1"), "SORT_CLUSTER", os.environ.get("NGICP_SORT_CLUSTER", "1"))
