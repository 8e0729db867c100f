"""Generates tests/golden/knn_ref_full.npz: FULL-SIZE k-NN tables from the REFERENCE's own nanoflann.h (oracle variant "ref",
compiled in place from /root/reference/src/dlio/include/nano_gicp/nanoflann.h, see oracle/Makefile). Run HERE (the container
that has /root/reference):      python tests/golden/make_golden_full.py
Contents (the clouds are regenerated from their seeds by tests/test_gpu_pins.py:full_size_clouds; their SHA-256 is stored so
that generator drift is detected rather than mis-reported as a k-NN mismatch):
  self16   (65536, 16) int32  canonical (distance, index) rows of every point of a full OS1-64 scan, stored as index - row
  corr1_a / corr1_b  (65536,) int32  1-NN of the scan (fp32-transformed by pose A / B) in the 1,000,000-point submap
  q16      (16384, 16) int32  canonical rows of 16,384 mixed queries into the submap
Distances are not stored: they follow from the indices with the reference's fp32 metric (nanoflann.h:509-520)."""
import hashlib
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import oracle  # noqa: E402
import scenarios as S  # noqa: E402
from test_gpu_pins import full_size_clouds, transform_f32, POSES, mixed_queries  # noqa: E402

OUT = Path(__file__).resolve().parent


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    oracle.build(ref=True)
    assert oracle.lib("ref").orc_tree_kind() == b"reference-nanoflann"
    scan, tgt = full_size_clouds()
    i16, d16 = oracle.KdTree(scan, "ref").knn(scan, 16)
    i16, d16 = S.canonical_rows(i16.astype(np.int32), d16)
    tree = oracle.KdTree(tgt, "ref")
    out = {"scan_sha": sha(scan), "tgt_sha": sha(tgt), "self16": (i16 - np.arange(len(scan), dtype=np.int32)[:, None]).astype(np.int32)}
    for name, T in POSES.items():
        i1, d1 = tree.knn(transform_f32(T, scan), 1)
        out["corr1_" + name] = i1[:, 0].astype(np.int32)
    q = mixed_queries(scan, tgt, 16384)
    iq, dq = tree.knn(q, 16)
    iq, dq = S.canonical_rows(iq.astype(np.int32), dq)
    out["q16"] = iq
    np.savez_compressed(OUT / "knn_ref_full.npz", **out)
    print("knn_ref_full.npz", (OUT / "knn_ref_full.npz").stat().st_size // 1024, "KiB")


if __name__ == "__main__":
    main()
