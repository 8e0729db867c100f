"""Generates tests/golden/*.npz. Run HERE (the container that has /root/reference):
    python tests/golden/make_golden.py
k-NN tables come from the REFERENCE's own nanoflann.h (oracle variant "ref", compiled in place from
/root/reference/src/dlio/include/nano_gicp/nanoflann.h, see oracle/Makefile) in the reference's raw
result order (KD-visit order on ties). Covariance / linearise / LM vectors come from the oracle
restatement running over that same reference tree. The reference itself ships no fixtures for this
path, so these files are the pin.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import oracle  # noqa: E402
import scenarios as S  # noqa: E402
from ngicp import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def main():
    oracle.build(ref=True)
    assert oracle.lib("ref").orc_tree_kind() == b"reference-nanoflann"
    a, b, T_true = S.scan_pair(0, w=64)          # ~3.3k points each
    print("clouds", a.shape, b.shape)
    tree = oracle.KdTree(a, "ref")
    idx16, sqd16 = tree.knn(a, 16, canonical=False)          # self k-NN, reference order
    idx1, sqd1 = tree.knn(b, 1, canonical=False)             # correspondence-style 1-NN
    idx5, sqd5 = tree.knn(b[:512] + np.float32([0.3, -0.2, 0.1]), 5, canonical=False)
    np.savez_compressed(OUT / "knn_ref.npz", cloud=a, queries=b, idx16=idx16, sqd16=sqd16, idx1=idx1, sqd1=sqd1, idx5=idx5, sqd5=sqd5)

    g = S.configure(oracle.OracleGICP("ref", num_threads=1))
    g.setInputSource(a)
    g.setInputTarget(b)
    out = {}
    for name, reg in (("plane", oracle.REG_PLANE), ("none", oracle.REG_NONE), ("min_eig", oracle.REG_MIN_EIG),
                      ("norm_min_eig", oracle.REG_NORMALIZED_MIN_EIG), ("frobenius", oracle.REG_FROBENIUS)):
        g.setRegularizationMethod(reg)
        g.calculateSourceCovariances()
        out["cov_" + name] = g.getSourceCovariances()[:, :3, :3]
        out["density_" + name] = np.float32(g.source_density_)
    g.setRegularizationMethod(oracle.REG_PLANE)
    g.calculateSourceCovariances()
    g.calculateTargetCovariances()
    out["cov_target_plane"] = g.getTargetCovariances()[:, :3, :3]
    T0 = synth.se3((0.01, -0.02, 0.015), (0.2, -0.1, 0.05))
    T1 = synth.se3((0.001, 0.002, -0.001), (0.01, 0.02, -0.01)) @ T0
    err, H, bb = g.linearize(T0)
    corr, sqd, mah = g.update_correspondences(T0)
    out.update(T0=T0, T1=T1, lin_err=err, lin_H=H, lin_b=bb, lin_ncorr=g.num_correspondences, corr=corr, corr_sqd=sqd,
               mahal=mah[:, :3, :3], err_T1=g.compute_error(T1))
    T = g.align()
    out.update(align_T=T, align_iters=g.nr_iterations_, align_converged=g.converged_, align_final_err=g.getFinalError(),
               align_H=g.getFinalHessian(), T_true=T_true)
    np.savez_compressed(OUT / "gicp_oracle.npz", source=a, target=b, **out)
    for f in ("knn_ref.npz", "gicp_oracle.npz"):
        print(f, (OUT / f).stat().st_size // 1024, "KiB")


if __name__ == "__main__":
    main()
