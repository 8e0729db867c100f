"""GPU parity of the PRODUCTION self k-NN (K2, leaf-scheduled search, csrc/lknn.cuh) against the oracle's KD-tree.

The public nearestKSearch goes through a different (per-query) kernel, so these tests look at the neighbour sets K2
actually hands to the covariance kernel (ngicp_self_neighbours): the k nearest under the (distance, index) tie-break
of every point of the cloud, reference src/dlio/src/nano_gicp/nano_gicp.cc:343 -> nanoflann.h:1436-1460.
Bar: distances bit-identical, index sets identical except for ties at the k-th distance (either side may keep either
tied point; the reference keeps KD-visit order, nanoflann.h:207-240)."""
import numpy as np
import pytest

import ngicp
import oracle
import scenarios as S
from ngicp import synth

pytestmark = pytest.mark.gpu


def ref_sqdist(q, p):
    """((dx*dx)+(dy*dy))+(dz*dz) in fp32, query minus data (nanoflann.h:509-520)."""
    d = (q[:, None, :] - p).astype(np.float32)
    return ((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32) + d[..., 2] * d[..., 2]).astype(np.float32)


def check_self_sets(cloud, k, variant="port", g=None, which=0):
    cloud = np.ascontiguousarray(cloud, np.float32)
    if g is None:
        g = S.configure(ngicp.NanoGICP(0), k=k)
        g.setInputSource(cloud)
    gi, dens = g.selfNeighbours(which, k)
    n = len(cloud)
    assert gi.shape == (n, k)
    assert (gi[:, 0] == np.arange(n)).all(), "entry 0 of a row is the point itself"
    assert ((gi >= 0) & (gi < n)).all()
    for r in (0, n // 2, n - 1):
        assert len(set(gi[r].tolist())) == k, "a row must not repeat a neighbour"
    gd = ref_sqdist(cloud, cloud[gi])
    oi, od = oracle.KdTree(cloud, variant).knn(cloud, k)
    gi_c, gd_c = S.canonical_rows(gi, gd)
    oi_c, od_c = S.canonical_rows(oi.astype(np.int32), od)
    exact, tie, bad = S.knn_rows_equivalent(gi_c, gd_c, oi_c, od_c)
    assert bad == 0, f"{bad} rows differ from the oracle beyond ties ({exact} exact, {tie} tie rows)"
    norm = ((k - 1) * (2 + k)) // 2
    if norm > 0:
        want = od_c[:, 1:].astype(np.float64).sum(1) / norm
        assert np.allclose(dens, want, rtol=1e-13, atol=0), "density terms (nano_gicp.cc:345-347)"
    return exact, tie


@pytest.mark.parametrize("k", [1, 2, 5, 8, 16, 20, 32])
def test_self_neighbour_sets_match_the_oracle(k):
    a, _, _ = S.scan_pair(3, w=128)
    exact, tie = check_self_sets(a, k)
    assert exact > 0.99 * len(a)


def test_self_neighbour_sets_raw_scan_dense_and_sparse_regions():
    """An unfiltered OS1-64 scan: leaves with hundreds of points next to the sensor (bucket pruning, streamed chunks) and
    isolated far returns (refused members climbing levels)."""
    sc = synth.Scene(5)
    rng = np.random.default_rng(6)
    a = synth.scan(sc, synth.trajectory(sc, 1, 5)[0], rng, w=512, keep_all=True)
    check_self_sets(a, 16)
    check_self_sets(a, 20)


def test_self_neighbour_sets_with_duplicates_and_exact_ties():
    """Duplicate points and a regular lattice: many exact distance ties at the k-th place, resolved by original index."""
    a, _, _ = S.scan_pair(4, w=64)
    dup = np.concatenate([a, a[::3], a[::3], a[:40]])
    rng = np.random.default_rng(1)
    dup = dup[rng.permutation(len(dup))]
    gx = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(6), indexing="ij"), -1).reshape(-1, 3).astype(np.float32) * np.float32(0.25)
    for cloud in (dup, gx):
        g = S.configure(ngicp.NanoGICP(0))
        g.setInputSource(cloud)
        gi, _ = g.selfNeighbours(0, 16)
        gd = ref_sqdist(cloud, cloud[gi])
        oi, od = oracle.KdTree(cloud, "port").knn(cloud, 16)
        gi_c, gd_c = S.canonical_rows(gi, gd)
        # exact statement of the documented tie-break: brute force, k smallest (distance, index)
        sel = np.random.default_rng(2).choice(len(cloud), 300, replace=False)
        d2 = ref_sqdist(cloud[sel], cloud[None, :, :])
        order = np.lexsort((np.broadcast_to(np.arange(len(cloud)), d2.shape), d2), axis=1)[:, :16]
        bd = np.take_along_axis(d2, order, 1)
        assert (gd_c[sel] == bd).all()
        assert (np.sort(od, 1) == gd_c).all(), "distances bit-identical to the oracle KD-tree"
        # the set: everything strictly closer than the k-th distance is forced; ties at the k-th distance go to the smallest
        # original indices, except that the row always keeps the query itself (a duplicate of it is the same point)
        for r, s in enumerate(sel):
            want = set(order[r].tolist())
            got = set(gi[s].tolist())
            if want != got:
                last = bd[r, -1]
                forced = {int(i) for i, d in zip(order[r], bd[r]) if d < last}
                assert forced <= got or last == 0.0
                extra = got - forced
                assert all(ref_sqdist(cloud[s:s + 1], cloud[None, [e]])[0, 0] == last for e in extra)


def test_self_neighbour_sets_tiny_and_degenerate_clouds():
    rng = np.random.default_rng(0)
    for n in (16, 17, 33, 100):
        check_self_sets(rng.normal(size=(n, 3)).astype(np.float32), 16)
    line = np.zeros((200, 3), np.float32)
    line[:, 0] = np.arange(200) * 0.01
    check_self_sets(line, 16)
    flat = rng.uniform(-5, 5, (3000, 3)).astype(np.float32)
    flat[:, 2] = 0
    check_self_sets(flat, 16)
    far = np.concatenate([rng.normal(size=(500, 3)), rng.normal(size=(40, 3)) * 0.01 + 300.0]).astype(np.float32)
    check_self_sets(far, 16)


def test_self_neighbour_sets_per_keyframe_in_a_batched_cloud():
    """Batched index (BASELINE config 3): neighbours never cross a keyframe boundary."""
    sc = synth.Scene(3)
    rng = np.random.default_rng(5)
    clouds = [synth.voxel_filter(synth.scan(sc, P, rng, w=96)) for P in synth.trajectory(sc, 3, 3)]
    pts = np.concatenate(clouds)
    off = np.cumsum([0] + [len(c) for c in clouds])
    g = S.configure(ngicp.NanoGICP(0))
    g.setInputSourceBatch(pts, off)
    gi, _ = g.selfNeighbours(0, 16)
    for s, c in enumerate(clouds):
        rows = gi[off[s]:off[s + 1]] - off[s]
        assert ((rows >= 0) & (rows < len(c))).all()
        gd = ref_sqdist(c, c[rows])
        oi, od = oracle.KdTree(c, "port").knn(c, 16)
        assert S.knn_rows_equivalent(*S.canonical_rows(rows, gd), *S.canonical_rows(oi.astype(np.int32), od))[2] == 0


def test_self_neighbours_full_os1_scan_against_reference_nanoflann():
    """All 65,536 rows of a full OS1-64 scan (k = 16) against the reference's own nanoflann.h KD-tree when the oracle
    build that contains it travelled to this box (oracle/_ref), else against the port."""
    sc = synth.Scene(0)
    rng = np.random.default_rng(2)
    a = synth.scan(sc, synth.trajectory(sc, 1, 0)[0], rng, keep_all=True)
    assert a.shape == (65536, 3)
    variant = "ref" if oracle.available("ref") else "port"
    exact, tie = check_self_sets(a, 16, variant)
    assert exact + tie == 65536
