"""The thread-block-cluster kernels (csrc/cluster_sort.cuh, csrc/index_cluster.cuh) at the sizes where their shape changes:
one CTA, a ragged last CTA, 4096- vs 8192-record tiles (65,536 / 65,537 points), the largest cluster (131,072) and the first
size that falls back to the multi-kernel path — index keys, k-NN after the build, VoxelGrid and the time sort, each against
the oracle, and the cluster path against the multi-kernel path (NGICP_K1_CLUSTER / NGICP_SORT_CLUSTER are read once per
process, so that comparison runs in a child process)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

import oracle
from oracle import voxel_keys as vk

pytestmark = pytest.mark.gpu

SIZES = [2, 33, 4095, 4096, 4097, 40000, 65535, 65536, 65537, 100003, 131072, 131073]


def _cloud(n, seed=0, dup=True):
    """A LiDAR-like mixture: a dense blob next to the sensor, a ground plane, far walls, and (dup) repeated points."""
    rng = np.random.default_rng([seed, n])
    k = max(n // 3, 1)
    near = rng.normal(0, 0.4, (k, 3))
    ground = np.stack([rng.uniform(-60, 60, k), rng.uniform(-60, 60, k), rng.normal(-1.8, 0.02, k)], 1)
    far = rng.uniform(-90, 90, (n - 2 * k if n > 2 * k else 0, 3))
    c = np.concatenate([near, ground, far])[:n].astype(np.float32)
    if len(c) < n:
        c = np.concatenate([c, rng.uniform(-1, 1, (n - len(c), 3)).astype(np.float32)])
    if dup and n >= 64:
        c[n // 2:n // 2 + 16] = c[7]                     # exact duplicates: equal keys, stability decides their order
    return c


@pytest.mark.parametrize("n", SIZES)
def test_index_keys_after_a_cluster_build_are_the_specified_keys(n):
    import ngicp
    a = _cloud(n)
    t = ngicp.KdTreeFLANN()
    t.setInputCloud(a)
    keys, lo, h0 = t.voxel_keys()
    lo_o, h0_o = vk.grid_params(a)
    assert (lo == lo_o).all() and h0 == h0_o
    assert (keys == vk.voxel_keys(a, lo_o, h0_o)).all()


@pytest.mark.parametrize("n", [33, 4097, 65536, 65537, 131072])
def test_knn_and_self_neighbours_on_a_cluster_built_index(n):
    """Everything downstream of the build (sorted points, inverse permutation, every level of the hash) through exact k-NN:
    queries against the public search, and the production self-search's neighbour sets."""
    import ngicp
    import scenarios as S
    a = _cloud(n, seed=1)
    k = min(8, n)
    t = ngicp.KdTreeFLANN()
    t.setInputCloud(a)
    rng = np.random.default_rng(n)
    q = np.concatenate([a[rng.integers(0, n, 300)], rng.uniform(-80, 80, (200, 3)).astype(np.float32)])
    idx, sqd = t.nearestKSearch(q, k)
    oi, od = oracle.KdTree(a, "port").knn(q, k)
    ri, rd = S.canonical_rows(oi.astype(np.int32), od)
    assert (sqd == rd).all()
    exact, tie, bad = S.knn_rows_equivalent(idx, sqd, ri, rd)
    assert bad == 0
    if n >= 4097:
        g = ngicp.NanoGICP(0)
        g.setCorrespondenceRandomness(16)
        g.setInputSource(a)
        g.calculateSourceCovariances()
        gi = g.selfNeighbours(k=16)[0]
        rows = rng.integers(0, n, 2000)
        oi, od = oracle.KdTree(a, "port").knn(a[rows], 16)
        gd = ((a[gi[rows]].astype(np.float32) - a[rows][:, None]) ** 2)
        gd = ((gd[..., 0] + gd[..., 1]) + gd[..., 2]).astype(np.float32)
        gi_c, gd_c = S.canonical_rows(gi[rows], gd)
        oi_c, od_c = S.canonical_rows(oi.astype(np.int32), od)
        exact, tie, bad = S.knn_rows_equivalent(gi_c, gd_c, oi_c, od_c)
        assert bad == 0


@pytest.mark.parametrize("n", [33, 4096, 4097, 65536, 65537, 131072, 131073])
def test_voxel_grid_through_the_cluster_sort(n):
    import ngicp
    a = _cloud(n, seed=2)
    a[::97] = np.nan
    g = ngicp.NanoGICP(0)
    v = ngicp.VoxelGrid(g); o = oracle.VoxelGrid()
    for f in (v, o):
        f.setLeafSize(0.25); f.setInputCloud(a)
    out, ref = v.filter(), o.filter()
    assert out.shape == ref.shape and (out == ref).all()


@pytest.mark.parametrize("n,dtype", [(4097, np.uint32), (65536, np.uint32), (65537, np.float32), (131072, np.uint32), (65536, np.float64), (70000, np.uint32)])
def test_time_sort_through_the_cluster_sort(n, dtype):
    """ngicp_scan_ingest: stable sort by time stamp (ties keep arrival order), dropped points last; float64 stamps need all
    64 key bits and stay on the multi-kernel sort."""
    import ngicp
    rng = np.random.default_rng(n)
    name = {np.uint32: "t", np.float32: "time", np.float64: "timestamp"}[dtype]
    fields = [("x", np.float32), ("y", np.float32), ("z", np.float32), ("w", np.float32), ("intensity", np.float32), ("pad0", np.uint32), (name, dtype)]
    if dtype != np.float64:
        fields.append(("pad", np.uint32, 1))
    rec = np.zeros(n, np.dtype(fields, align=False))
    xyz = _cloud(n, seed=3, dup=False)
    rec["x"], rec["y"], rec["z"] = xyz.T
    col = rng.integers(0, 1024, n)
    rec[name] = (col * 97656).astype(dtype) if dtype == np.uint32 else (col * 9.7656e-5).astype(dtype)
    rec["x"][::131] = np.nan
    crop = ([-1.0] * 3, [1.0] * 3, True)
    g = ngicp.NanoGICP(0)
    stamps, kept = g.ingestScan(rec, name, crop=crop)
    xyz_o, grp_o, stamps_o = oracle.scan_ingest(rec, name, crop=crop)
    assert kept == len(xyz_o) and (stamps == stamps_o).all()
    out = g.deskewScan(np.eye(4, dtype=np.float32), leaf=None, set_source=False)
    assert out.shape == xyz_o.shape and (out == xyz_o).all()           # identity frames: the time-ordered cloud itself


def test_cluster_and_multi_kernel_paths_agree_bit_for_bit():
    """Same clouds through NGICP_K1_CLUSTER=1 / NGICP_SORT_CLUSTER=1 and through the multi-kernel path in a child process:
    identical sorted keys, covariances and voxel-grid output."""
    code = textwrap.dedent("""
        import sys, hashlib
        import numpy as np
        sys.path[:0] = %r
        import ngicp
        from test_gpu_cluster import _cloud
        h = hashlib.sha256()
        for n in (4097, 65536, 100003):
            a = _cloud(n, seed=5)
            g = ngicp.NanoGICP(0)
            g.setCorrespondenceRandomness(16)
            g.setInputSource(a)
            g.calculateSourceCovariances()
            h.update(np.ascontiguousarray(g.getSourceCovariances()).tobytes())
            h.update(np.ascontiguousarray(g.source_kdtree_.voxel_keys()[0]).tobytes())
            v = ngicp.VoxelGrid(g); v.setLeafSize(0.3); v.setInputCloud(a)
            h.update(np.ascontiguousarray(v.filter()).tobytes())
        print(h.hexdigest())
    """) % ([p for p in sys.path if p],)
    out = {}
    for flag in ("1", "0"):
        env = dict(os.environ, NGICP_K1_CLUSTER=flag, NGICP_SORT_CLUSTER=flag)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        out[flag] = r.stdout.strip().splitlines()[-1]
    assert out["1"] == out["0"]
