"""CPU tests of the oracle (oracle/): pinned against the reference's own nanoflann (golden k-NN
tables made by tests/golden/make_golden.py) and cross-checked against numpy/scipy."""
from pathlib import Path

import numpy as np
import pytest

import oracle
import scenarios as S
from ngicp import synth
from oracle import voxel_keys as vk

G = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def knn_gold():
    return np.load(G / "knn_ref.npz")


@pytest.fixture(scope="module")
def gicp_gold():
    return np.load(G / "gicp_oracle.npz")


def test_port_knn_matches_reference_nanoflann(knn_gold):
    """The self-contained port search returns what the reference's nanoflann returned (golden),
    bit-identical distances, identical indices up to ties at the k-th boundary."""
    tree = oracle.KdTree(knn_gold["cloud"], "port")
    for k, q, gi, gd in ((16, knn_gold["cloud"], knn_gold["idx16"], knn_gold["sqd16"]),
                         (1, knn_gold["queries"], knn_gold["idx1"], knn_gold["sqd1"]),
                         (5, knn_gold["queries"][:512] + np.float32([0.3, -0.2, 0.1]), knn_gold["idx5"], knn_gold["sqd5"])):
        idx, sqd = tree.knn(q, k)
        ri, rd = S.canonical_rows(gi, gd)
        exact, tie, bad = S.knn_rows_equivalent(idx, sqd, ri, rd)
        assert bad == 0, (k, exact, tie, bad)
        assert (sqd == rd).all()


@pytest.mark.skipif(not oracle.available("ref"), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_build_reproduces_golden(knn_gold):
    assert oracle.lib("ref").orc_tree_kind() == b"reference-nanoflann"
    idx, sqd = oracle.KdTree(knn_gold["cloud"], "ref").knn(knn_gold["cloud"], 16, canonical=False)
    assert (idx == knn_gold["idx16"]).all() and (sqd == knn_gold["sqd16"]).all()


def test_knn_against_bruteforce():
    rng = np.random.default_rng(0)
    p = rng.normal(size=(700, 3)).astype(np.float32)
    p[100:110] = p[0]                     # exact duplicates: ties broken by index
    q = rng.normal(size=(50, 3)).astype(np.float32)
    idx, sqd = oracle.KdTree(p, "port").knn(q, 7)
    d = (q[:, None, :] - p[None, :, :]).astype(np.float32)
    d2 = ((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32) + d[..., 2] * d[..., 2]).astype(np.float32)
    order = np.lexsort((np.broadcast_to(np.arange(len(p)), d2.shape), d2), axis=1)[:, :7]
    assert (idx == order).all()
    assert (sqd == np.take_along_axis(d2, order, 1)).all()


def test_knn_fewer_points_than_k_pads():
    p = np.eye(3, dtype=np.float32)
    idx, sqd = oracle.KdTree(p, "port").knn(p[:1], 5)
    assert (idx[0, 3:] == -1).all() and np.isinf(sqd[0, 3:]).all() and idx[0, 0] == 0


def test_covariance_matches_numpy(gicp_gold):
    a = gicp_gold["source"]
    g = S.configure(oracle.OracleGICP("port", num_threads=2))
    g.setInputSource(a)
    idx, sqd = oracle.KdTree(a, "port").knn(a, 16)
    nb = a[idx].astype(np.float64)
    c = nb - nb.mean(1, keepdims=True)
    cov = np.einsum("nki,nkj->nij", c, c) / 16.0      # divide by k, self included (nano_gicp.cc:353-354)
    ok = S.spectral_gap_ok(a, idx)
    assert ok.mean() > 0.9
    U, s, Vt = np.linalg.svd(cov)
    expect = {
        oracle.REG_NONE: cov,
        oracle.REG_PLANE: np.einsum("nij,j,njk->nik", U, np.array([1, 1, 1e-3]), Vt),
        oracle.REG_MIN_EIG: np.einsum("nij,nj,njk->nik", U, np.maximum(s, 1e-3), Vt),
        oracle.REG_NORMALIZED_MIN_EIG: np.einsum("nij,nj,njk->nik", U, np.maximum(s / s[:, :1], 1e-3), Vt),
    }
    lam = cov + 1e-3 * np.eye(3)
    ci = np.linalg.inv(lam)
    expect[oracle.REG_FROBENIUS] = np.linalg.inv(ci / np.linalg.norm(ci, axis=(1, 2), keepdims=True))
    for reg, want in expect.items():
        g.setRegularizationMethod(reg)
        g.calculateSourceCovariances()
        got = g.getSourceCovariances()
        assert np.abs(got[:, 3, :]).max() == 0 and np.abs(got[:, :, 3]).max() == 0
        err = np.abs(got[ok, :3, :3] - want[ok]).max()
        assert err < 1e-9 * max(1.0, np.abs(want[ok]).max()), (reg, err)
    # density: sum_{j>=1} d2 / ((k-1)(k+2)/2 as int), averaged (nano_gicp.cc:345-346,389)
    dens = (sqd[:, 1:].astype(np.float64).sum(1) / 135).mean()
    assert abs(g.source_density_ - dens) < 1e-4 * dens


def test_oracle_reproduces_golden(gicp_gold):
    d = gicp_gold
    g = S.configure(oracle.OracleGICP("port", num_threads=3))
    g.setInputSource(d["source"])
    g.setInputTarget(d["target"])
    g.calculateSourceCovariances()
    g.calculateTargetCovariances()
    assert np.abs(g.getSourceCovariances()[:, :3, :3] - d["cov_plane"]).max() < 1e-12
    err, H, b = g.linearize(d["T0"])
    assert abs(err - d["lin_err"]) < 1e-9 * abs(d["lin_err"])
    assert np.abs(H - d["lin_H"]).max() < 1e-9 * np.abs(d["lin_H"]).max()
    assert np.abs(b - d["lin_b"]).max() < 1e-9 * np.abs(d["lin_b"]).max()
    assert g.num_correspondences == int(d["lin_ncorr"])
    assert abs(g.compute_error(d["T1"]) - d["err_T1"]) < 1e-9 * abs(d["err_T1"])
    T = g.align()
    assert g.nr_iterations_ == int(d["align_iters"]) and g.converged_ == bool(d["align_converged"])
    assert np.abs(T - d["align_T"]).max() < 1e-6


def test_linearize_matches_numpy(gicp_gold):
    """Independent vectorised restatement of nano_gicp.cc:206-302 from the oracle's own correspondences."""
    d = gicp_gold
    a, b, T = d["source"].astype(np.float64), d["target"].astype(np.float64), d["T0"]
    corr = d["corr"]
    v = corr >= 0
    R, t = T[:3, :3], T[:3, 3]
    CA, CB = d["cov_plane"][v], d["cov_target_plane"][corr[v]]
    M = np.linalg.inv(CB + R @ CA @ R.T)
    assert np.abs(M - d["mahal"][v]).max() < 1e-8 * np.abs(M).max()
    q = a[v] @ R.T + t
    e = b[corr[v]] - q
    J = np.zeros((v.sum(), 3, 6))
    J[:, 0, 1], J[:, 0, 2], J[:, 1, 0], J[:, 1, 2], J[:, 2, 0], J[:, 2, 1] = -q[:, 2], q[:, 1], q[:, 2], -q[:, 0], -q[:, 1], q[:, 0]
    J[:, :, 3:] = -np.eye(3)
    H = np.einsum("nki,nkl,nlj->ij", J, M, J)
    bb = np.einsum("nki,nkl,nl->i", J, M, e)
    err = np.einsum("ni,nij,nj->", e, M, e)
    assert abs(err - d["lin_err"]) < 1e-9 * err
    assert np.abs(H - d["lin_H"]).max() < 1e-9 * np.abs(H).max()
    assert np.abs(bb - d["lin_b"]).max() < 1e-9 * np.abs(bb).max()
    assert int((corr > 0).sum()) == int(d["lin_ncorr"])      # index 0 is not counted (nano_gicp.cc:244)
    # 1-NN gate: strict d2 < thr^2
    assert (d["corr_sqd"][v] < 0.25).all() and (d["corr_sqd"][~v] >= 0.25).all()


def test_align_recovers_known_transform():
    a, b, T_true = S.scan_pair(1, w=96)
    g = S.configure(oracle.OracleGICP("port"))
    g.setInputSource(b)
    g.setInputTarget(a)
    T = g.align()
    assert g.hasConverged()
    assert np.abs(T[:3, 3] - T_true[:3, 3]).max() < 0.02 and np.abs(T[:3, :3] - T_true[:3, :3]).max() < 5e-3


def test_thread_count_only_changes_last_bits(gicp_gold):
    """The reference's OpenMP reductions are order dependent (SURVEY.md hard part 3): the oracle at 1
    and N threads brackets that noise, far inside the parity tolerances."""
    d = gicp_gold
    res = []
    for nt in (1, 4):
        g = S.configure(oracle.OracleGICP("port", num_threads=nt))
        g.setInputSource(d["source"]); g.setInputTarget(d["target"])
        g.calculateSourceCovariances(); g.calculateTargetCovariances()
        res.append(g.linearize(d["T0"]))
    assert abs(res[0][0] - res[1][0]) < 1e-10 * abs(res[0][0])
    assert np.abs(res[0][1] - res[1][1]).max() < 1e-10 * np.abs(res[0][1]).max()


def test_voxel_key_spec():
    p = np.array([[0, 0, 0], [1, 0, 0], [0, 2, 0], [0, 0, 3.9]], np.float32)
    lo, h0 = vk.grid_params(p)
    assert (lo == 0).all() and h0 == np.float32(2.0 ** -10)      # extent 3.9*1.01 < 4 = 2^2 -> h0 = 2^(2-12)
    k = vk.voxel_keys(p, lo, h0)
    assert k[0] == 0 and k[1] == sum(1 << (3 * i) for i in range(11) if (1024 >> i) & 1)
    assert k[2] == 2 * sum(1 << (3 * i) for i in range(12) if (2048 >> i) & 1)
    assert vk.voxel_keys(p, lo, h0, seg=3)[0] == 3 << 36
