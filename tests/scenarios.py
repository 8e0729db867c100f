"""Seeded inputs shared by the CPU and GPU tests (synthetic OS1-64 scans, see ngicp/synth.py)."""
from functools import lru_cache

import numpy as np

from ngicp import synth

DLIO = dict(k=16, max_corr=0.5, max_iter=32, rot_eps=0.01, trans_eps=0.01)  # reference cfg/params.yaml:57-63


def configure(g, k=16, max_corr=0.5, max_iter=32, rot_eps=0.01, trans_eps=0.01, reg=None):
    """Apply DLIO's setters (reference src/dlio/odom.cc:89-101) to a NanoGICP-like object."""
    g.setCorrespondenceRandomness(k)
    g.setMaxCorrespondenceDistance(max_corr)
    g.setMaximumIterations(max_iter)
    g.setRotationEpsilon(rot_eps)
    g.setTransformationEpsilon(trans_eps)
    if reg is not None:
        g.setRegularizationMethod(reg)
    return g


@lru_cache(maxsize=None)
def scan_pair(seed=0, w=128, leaf=0.25):
    """Two consecutive voxel-filtered scans (sensor frames) and the true relative pose (BASELINE cfg 1 shape,
    w columns instead of 1024 to keep CPU tests fast)."""
    sc = synth.Scene(seed)
    rng = np.random.default_rng(seed + 1)
    poses = synth.trajectory(sc, 2, seed)
    a = synth.scan(sc, poses[0], rng, w=w)
    b = synth.scan(sc, poses[1], rng, w=w)
    if leaf > 0:
        a, b = synth.voxel_filter(a, leaf), synth.voxel_filter(b, leaf)
    T_true = np.linalg.inv(poses[0]) @ poses[1]   # maps scan-1 points into scan-0's frame
    return a, b, T_true


@lru_cache(maxsize=None)
def scan_to_submap(seed=0, n_target=60000, n_keyframes=6, w=256, max_t=0.2, max_deg=2.0):
    """Small-scale BASELINE cfg 2: a raw scan (world frame, perturbed) against a keyframe submap."""
    sc = synth.Scene(seed)
    rng = np.random.default_rng(seed + 11)
    tgt, bounds, poses = synth.make_submap(sc, n_target, seed, n_keyframes=n_keyframes, w=w)
    T_ws = poses[len(poses) // 2] @ synth.se3((0, 0, 0.02), (0.3, 0.1, 0.0))
    src_sensor = synth.scan(sc, T_ws, rng, w=w)
    src_world = synth.transform_points(T_ws, src_sensor)
    T_off = synth.random_se3(rng, max_t, max_deg)
    src = synth.transform_points(np.linalg.inv(T_off), src_world)   # aligning src must recover ~T_off
    return src, tgt, bounds, T_off


def canonical_rows(idx, sqd):
    """Sort every k-NN row by (distance, index)."""
    order = np.lexsort((idx, sqd), axis=1)
    return np.take_along_axis(idx, order, 1), np.take_along_axis(sqd, order, 1)


def knn_rows_equivalent(idx_a, sqd_a, idx_b, sqd_b):
    """Compare two canonical k-NN tables. Distances must be bit-identical everywhere. Indices must be
    identical except in rows where the k-th distance is tied with a neighbour that did not make the
    cut (either side may keep either of the tied points: the reference keeps KD-visit order,
    nanoflann.h:207-240). Returns (n_exact_rows, n_tie_rows, n_bad_rows)."""
    same_d = (sqd_a == sqd_b).all(1)
    same_i = (idx_a == idx_b).all(1)
    exact = same_d & same_i
    # tie rows: distances equal, index sets differ only among entries whose distance equals the row's last distance
    tie = np.zeros(len(idx_a), bool)
    for r in np.nonzero(same_d & ~same_i)[0]:
        last = sqd_a[r, -1]
        keep = sqd_a[r] < last
        tie[r] = np.array_equal(idx_a[r][keep], idx_b[r][keep])
    bad = ~(exact | tie)
    return int(exact.sum()), int(tie.sum()), int(bad.sum())


def spectral_gap_ok(points, idx, rel=1e-2):
    """Rows whose neighbourhood covariance has a well separated smallest eigenvalue: the PLANE
    regularisation U diag(1,1,eps) V^T is only well defined there (SURVEY.md hard part 2)."""
    nb = points[idx].astype(np.float64)
    c = nb - nb.mean(1, keepdims=True)
    cov = np.einsum("nki,nkj->nij", c, c) / idx.shape[1]
    w = np.linalg.eigvalsh(cov)
    # an exactly singular neighbourhood (smallest singular value == 0, e.g. all z identical) leaves the sign of the
    # last singular vectors of U and V independent of each other, so U diag(1,1,eps) V^T is arbitrary there too
    return ((w[:, 1] - w[:, 0]) > rel * w[:, 2]) & (w[:, 0] > 1e-12 * w[:, 2])
