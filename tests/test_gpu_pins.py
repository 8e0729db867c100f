"""Full-size pins of the CUDA k-NN on the reference's own KD-tree (reference src/dlio/include/nano_gicp/nanoflann.h:1436-1460,
called from nano_gicp.cc:224 and :343), on the GPU box:
  (i)   all 65,536 self k-NN rows (k = 16) of a full OS1-64 scan            — production K2 (leaf search)
  (ii)  all 65,536 1-NN correspondences into the 1,000,000-point submap at two poses — production K4a (bounded search)
  (iii) k = 16 queries into the 1,000,000-point submap                       — public nearestKSearch
against oracle.KdTree(..., "ref") when the oracle build that contains the reference's nanoflann.h travelled to this box
(oracle/_ref), and ALWAYS against tests/golden/knn_ref_full.npz, which tests/golden/make_golden_full.py made from that
same reference tree. Plus the 1-thread / N-thread envelope of the oracle's LM iteration counts (SURVEY.md hard part 3: the
reference's OpenMP reductions make nr_iterations_ depend on the thread count)."""
import hashlib
import os
from functools import lru_cache
from pathlib import Path

import numpy as np
import pytest

import oracle
import scenarios as S
from ngicp import synth

G = Path(__file__).resolve().parent / "golden"

POSES = {"a": synth.se3((0.0, 0.0, 0.0), (0.0, 0.0, 0.0)), "b": synth.se3((0.01, -0.015, 0.03), (0.15, -0.1, 0.05))}


@lru_cache(maxsize=None)
def full_size_clouds():
    """(scan (65536,3), submap (1000000,3)): BASELINE config 2 shaped, seeded."""
    sc = synth.Scene(0)
    rng = np.random.default_rng(2)
    tgt, bounds, poses = synth.make_submap(sc, 1_000_000, 0, n_keyframes=40)
    T_ws = poses[20] @ synth.se3((0, 0, 0.02), (0.3, 0.1, 0.0))
    scan = synth.transform_points(T_ws, synth.scan(sc, T_ws, rng, keep_all=True))
    assert scan.shape == (65536, 3) and tgt.shape == (1_000_000, 3)
    return scan, tgt


def transform_f32(T, pts):
    """((r0*x + r1*y) + r2*z) + t in fp32, no FMA — the reference's `trans_f * pt` (nano_gicp.cc:210,222)."""
    R = np.asarray(T, np.float64)[:3, :3].astype(np.float32)
    t = np.asarray(T, np.float64)[:3, 3].astype(np.float32)
    p = np.asarray(pts, np.float32)
    out = np.empty_like(p)
    for r in range(3):
        out[:, r] = ((R[r, 0] * p[:, 0] + R[r, 1] * p[:, 1]).astype(np.float32) + R[r, 2] * p[:, 2]).astype(np.float32) + t[r]
    return out.astype(np.float32)


def mixed_queries(scan, tgt, n):
    rng = np.random.default_rng(5)
    a = tgt[rng.choice(len(tgt), n // 2, replace=False)] + rng.normal(0, 0.05, (n // 2, 3)).astype(np.float32)
    b = scan[rng.choice(len(scan), n - n // 2, replace=False)] + rng.normal(0, 0.3, (n - n // 2, 3)).astype(np.float32)
    return np.concatenate([a, b]).astype(np.float32)


def ref_sqdist(q, p):
    d = (q[:, None, :] - p).astype(np.float32)
    return ((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32) + d[..., 2] * d[..., 2]).astype(np.float32)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def golden():
    z = np.load(G / "knn_ref_full.npz")
    scan, tgt = full_size_clouds()
    assert str(z["scan_sha"]) == sha(scan) and str(z["tgt_sha"]) == sha(tgt), "the seeded generators changed: regenerate knn_ref_full.npz"
    return z


@pytest.fixture(scope="module")
def big_gpu():
    import ngicp
    scan, tgt = full_size_clouds()
    g = S.configure(ngicp.NanoGICP(0))
    g.setInputTarget(tgt)
    g.calculateTargetCovariances()
    g.setInputSource(scan)
    g.calculateSourceCovariances()
    return g


def test_port_tree_matches_the_reference_fixture_at_full_size(golden):
    """CPU: the oracle's own k-d tree (what the GPU box falls back to when oracle/_ref did not travel) against the tables the
    reference's nanoflann.h produced, on samples of the full-size clouds."""
    scan, tgt = full_size_clouds()
    rows = np.random.default_rng(3).choice(len(scan), 4096, replace=False)
    pi, pd = oracle.KdTree(scan, "port").knn(scan[rows], 16)
    pi, pd = S.canonical_rows(pi.astype(np.int32), pd)
    oi = (golden["self16"] + np.arange(len(scan), dtype=np.int32)[:, None])[rows]
    assert S.knn_rows_equivalent(pi, pd, oi, ref_sqdist(scan[rows], scan[oi]))[2] == 0
    q = mixed_queries(scan, tgt, 16384)[:2048]
    ti, td = oracle.KdTree(tgt, "port").knn(q, 16)
    ti, td = S.canonical_rows(ti.astype(np.int32), td)
    assert S.knn_rows_equivalent(ti, td, golden["q16"][:2048], ref_sqdist(q, tgt[golden["q16"][:2048]]))[2] == 0


@pytest.mark.gpu
def test_self_knn_all_rows_of_a_full_scan_match_reference_nanoflann(big_gpu, golden):
    scan, _ = full_size_clouds()
    gi, _ = big_gpu.selfNeighbours(0, 16)
    gd = ref_sqdist(scan, scan[gi])
    gi_c, gd_c = S.canonical_rows(gi, gd)
    oi = golden["self16"] + np.arange(len(scan), dtype=np.int32)[:, None]
    od = ref_sqdist(scan, scan[oi])
    exact, tie, bad = S.knn_rows_equivalent(gi_c, gd_c, oi, od)
    assert bad == 0 and exact + tie == 65536
    if oracle.available("ref"):
        ri, rd = oracle.KdTree(scan, "ref").knn(scan, 16)
        ri, rd = S.canonical_rows(ri.astype(np.int32), rd)
        assert (rd == od).all(), "fixture distances recomputed from indices == the reference tree's own distances"
        assert S.knn_rows_equivalent(gi_c, gd_c, ri, rd)[2] == 0


@pytest.mark.gpu
def test_all_correspondences_into_the_1m_submap_match_reference_nanoflann(big_gpu, golden):
    scan, tgt = full_size_clouds()
    thr2 = np.float64(0.5) ** 2
    tree = oracle.KdTree(tgt, "ref") if oracle.available("ref") else None
    for name, T in POSES.items():
        corr, sqd, _ = big_gpu.update_correspondences(T)
        q = transform_f32(T, scan)
        oi = golden["corr1_" + name]
        od = ref_sqdist(q, tgt[oi][:, None, :])[:, 0]
        want = np.where(od.astype(np.float64) < thr2, oi, -1)              # strict gate, float promoted to double (nano_gicp.cc:227)
        differ = np.nonzero(corr != want)[0]
        # the only admissible difference: another target point at exactly the same distance (tie), both inside the gate
        for r in differ:
            assert corr[r] >= 0 and want[r] >= 0 and ref_sqdist(q[r:r + 1], tgt[None, [corr[r]]])[0, 0] == od[r]
        assert len(differ) <= 8
        ok = want >= 0
        assert (sqd[ok] == od[ok]).all(), "distances of the correspondences are bit-identical"
        if tree is not None:
            ri, rd = tree.knn(q, 1)
            assert (rd[:, 0] == od).all()


@pytest.mark.gpu
def test_queries_into_the_1m_submap_match_reference_nanoflann(big_gpu, golden):
    scan, tgt = full_size_clouds()
    q = mixed_queries(scan, tgt, 16384)
    gi, gd = big_gpu.target_kdtree_.nearestKSearch(q, 16)
    oi = golden["q16"]
    od = ref_sqdist(q, tgt[oi])
    assert S.knn_rows_equivalent(gi, gd, oi, od)[2] == 0
    if oracle.available("ref"):
        qq = mixed_queries(scan, tgt, 100_000)
        gi, gd = big_gpu.target_kdtree_.nearestKSearch(qq, 16)
        ri, rd = oracle.KdTree(tgt, "ref").knn(qq, 16)
        ri, rd = S.canonical_rows(ri.astype(np.int32), rd)
        assert S.knn_rows_equivalent(gi, gd, ri, rd)[2] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_lm_iterations_inside_the_oracle_thread_envelope(seed):
    """The reference accumulates H, b and the error in per-thread OpenMP partials (nano_gicp.cc:252-299), so its iteration
    count may differ between thread counts; the CUDA path (fixed-order compensated reduction) must land inside the envelope
    of the oracle run with 1 thread and with all host threads, and agree on the pose with both."""
    import ngicp
    src, tgt, bounds, T_off = S.scan_to_submap(seed)
    runs = []
    for threads in (1, os.cpu_count() or 2):
        o = S.configure(oracle.OracleGICP("port", num_threads=threads))
        o.setInputSource(src); o.setInputTarget(tgt)
        To = o.align()
        runs.append((o.nr_iterations_, bool(o.converged_), To))
    g = S.configure(ngicp.NanoGICP(0))
    g.setInputSource(src); g.setInputTarget(tgt)
    T = g.align()
    its = [r[0] for r in runs]
    assert min(its) <= g.nr_iterations_ <= max(its)
    assert bool(g.hasConverged()) in {r[1] for r in runs}
    for _, _, To in runs:
        assert np.abs(T[:3, 3] - To[:3, 3]).max() < 1e-4                      # 1e-4 m (north star)
        D = T[:3, :3].astype(np.float64).T @ To[:3, :3].astype(np.float64) - np.eye(3)
        assert np.linalg.norm(D) / np.sqrt(2.0) < 1e-5                         # 1e-5 rad


@pytest.mark.gpu
def test_degenerate_neighbourhoods_keep_the_plane_normal_in_the_oracle_null_space(big_gpu):
    """Rows the spectral-gap mask drops from the 1e-4 covariance comparison (two smallest eigenvalues nearly equal: the far,
    single-ring neighbourhoods of a LiDAR scan): U diag(1,1,eps) V^T is then only defined up to a rotation inside the
    near-null space, but the direction the device squeezes must lie INSIDE that space (nano_gicp.cc:365-385)."""
    scan, _ = full_size_clouds()
    C = big_gpu.getSourceCovariances()[:, :3, :3]
    idx, _ = oracle.KdTree(scan, "port").knn(scan, 16)
    ok = S.spectral_gap_ok(scan, idx)
    drop = np.nonzero(~ok)[0]
    assert 0 < len(drop) < 0.35 * len(scan)
    nb = scan[idx[drop]].astype(np.float64)
    c = nb - nb.mean(1, keepdims=True)
    A = np.einsum("nki,nkj->nij", c, c) / 16.0
    w, V = np.linalg.eigh(A)                                                  # ascending
    # device normal: eigenvector of the eigenvalue 1e-3 of I - (1 - 1e-3) n n^T
    wg, Vg = np.linalg.eigh(C[drop])
    assert np.abs(wg - np.array([1e-3, 1.0, 1.0])).max() < 1e-5
    n = Vg[:, :, 0]
    # oracle near-null space: every eigenvector whose eigenvalue is within the mask's gap of the smallest one
    inside = (w - w[:, :1]) <= 1e-2 * w[:, 2:3] + 1e-12 * w[:, 2:3]
    inside[:, 0] = True
    comp = np.einsum("nij,ni->nj", V, n)                                      # components of n along the oracle eigenvectors
    outside = np.sqrt(((comp ** 2) * (~inside)).sum(1))
    assert outside.max() < 2e-6, f"largest component outside the oracle's near-null space: {outside.max():.3e}"
