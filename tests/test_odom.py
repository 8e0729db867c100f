"""The odom-loop harness: host policy on CPU; the loop over the CUDA path against the same loop over
the oracle on the GPU box."""
import numpy as np
import pytest

from ngicp import odom, synth
from odom_backends import FakeBackend


def test_push_submap_indices_keeps_ties_and_k_smallest():
    # pushSubmapIndices (reference src/dlio/src/dlio/odom.cc:1628-1652): <= k-th smallest, ties included
    assert odom.push_submap_indices([3.0, 1.0, 2.0, 2.0, 5.0], 2, [10, 11, 12, 13, 14]) == [11, 12, 13]
    assert odom.push_submap_indices([], 3, []) == []
    assert odom.push_submap_indices([1.0, 2.0], 10, [7, 8]) == [7, 8]


def test_convex_hull_planar_and_spatial():
    sq = np.array([[0, 0, 0], [4, 0, 0], [4, 4, 0], [0, 4, 0], [2, 2, 0], [1, 3, 0]], float)
    assert odom.convex_hull_indices(sq) == [0, 1, 2, 3]                       # coplanar keyframes: the 2-D hull
    cube = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)] + [[0.5, 0.5, 0.5]], float)
    assert odom.convex_hull_indices(cube) == list(range(8))
    line = np.array([[i, 2 * i, 0.0] for i in range(6)])
    assert odom.convex_hull_indices(line) == [0, 5]


def test_concave_hull_follows_a_c_shaped_path():
    ang = np.deg2rad(np.arange(0, 271, 10.0))
    outer = np.stack([5 * np.cos(ang), 5 * np.sin(ang), 0 * ang], 1)
    inner = np.stack([4 * np.cos(ang), 4 * np.sin(ang), 0 * ang], 1)
    P = np.concatenate([outer, inner])
    hull = odom.concave_hull_indices(P, alpha=1.0)
    convex = odom.convex_hull_indices(P)
    n = len(ang)
    assert set(range(n, 2 * n)) <= set(hull)                                  # the inner arc is on the alpha shape ...
    assert not (set(range(n + 2, 2 * n - 2)) & set(convex))                   # ... and not on the convex hull
    assert set(range(n)) <= set(hull)
    assert odom.concave_hull_indices(P, alpha=0.1) == []                      # alpha below every circumradius: nothing kept


def test_quaternion_helpers():
    R = synth.rot_from_rotvec([0.1, -0.2, 0.7])
    q = odom.quat_from_rot(R)
    assert abs(np.linalg.norm(q) - 1) < 1e-12
    assert abs(odom.quat_angle_deg(q, q)) < 1e-6
    r = odom.quat_from_rot(synth.rot_from_rotvec([0, 0, np.deg2rad(50.0)]))
    assert abs(odom.quat_angle_deg(r, [1, 0, 0, 0]) - 50.0) < 1e-9
    assert abs(odom.quat_angle_deg(r, [-1, 0, 0, 0]) - 50.0) < 1e-9           # sign fix (odom.cc:1560-1568)


def _records(n=2000, seed=0):
    rng = np.random.default_rng(seed)
    rec = np.zeros(n, odom.OS1_RECORD)
    xyz = rng.uniform(-20, 20, (n, 3)).astype(np.float32)
    rec["x"], rec["y"], rec["z"] = xyz.T
    rec["t"] = rng.integers(0, 4, n)
    return rec


def test_keyframe_policy_and_submap_changes_with_a_scripted_backend():
    """Straight drive, 0.4 m per scan, identity corrections: a keyframe whenever the closest one is farther than the
    adaptive threshold (spaciousness clamps it into [0.5, 5]), every new keyframe transformed exactly once, the submap
    re-assembled only when the keyframe set changes (odom.cc:1517-1598, 1700-1741)."""
    n = 24
    be = FakeBackend([np.eye(4, dtype=np.float32)] * n)
    loop = odom.OdomLoop(be, odom.OdomParams(adaptive=True))
    rec = _records()
    pose = lambda i: synth.se3((0, 0, 0), (0.4 * i, 0, 0)).astype(np.float32)
    out = [loop.callbackPointCloud(rec, None)]
    for i in range(1, n):
        out.append(loop.callbackPointCloud(rec, lambda st, i=i: np.repeat(pose(i)[None], len(st), 0)))
    thr = loop.keyframe_thresh_dist_
    assert 0.5 <= thr <= 5.0
    kf_x = np.array([k[0][0] for k in loop.keyframes])
    assert len(kf_x) == be.captured and len(kf_x) >= 2
    assert (np.diff(kf_x) > thr - 1e-6).all()                                 # never closer than the threshold
    assert (np.diff(kf_x) <= thr + 0.4 + 1e-6).all()                          # and added as soon as it is exceeded
    assert sorted(be.transformed) == list(range(be.captured))                 # each keyframe transformed once
    assert be.max_corr and all(abs(d - 0.25) < 1e-12 for d in be.max_corr)    # adaptive: 0.5 * 0.5 (odom.cc:1616)
    changes = sum(r.new_keyframe for r in out)
    assert len(be.submaps) == changes                                          # one assembly per change of the set
    assert be.submaps[-1] == sorted(set(be.submaps[-1]))
    assert out[5].T[0, 3] == pytest.approx(0.4 * 5)


def test_rotation_only_keyframe_needs_no_neighbours():
    """abs(dd) <= threshD and theta > threshR only makes a keyframe when at most one keyframe is nearby (odom.cc:1586-1588)."""
    be = FakeBackend([np.eye(4, dtype=np.float32)] * 4)
    loop = odom.OdomLoop(be, odom.OdomParams(adaptive=False))
    rec = _records()
    loop.callbackPointCloud(rec, None)
    turn = synth.se3((0, 0, np.deg2rad(60.0)), (0, 0, 0)).astype(np.float32)
    r = loop.callbackPointCloud(rec, lambda st: np.repeat(turn[None], len(st), 0))
    assert r.new_keyframe and len(loop.keyframes) == 2
    turn2 = synth.se3((0, 0, np.deg2rad(120.0)), (0, 0, 0)).astype(np.float32)
    r = loop.callbackPointCloud(rec, lambda st: np.repeat(turn2[None], len(st), 0))
    assert not r.new_keyframe                                                 # two keyframes nearby now


@pytest.mark.gpu
@pytest.mark.parametrize("params,mulran", [(odom.OdomParams(), False), (odom.OdomParams(adaptive=False, keyframe_thresh_dist=0.5), False),
                                           (odom.OdomParams(adaptive=False, keyframe_thresh_dist=0.5), True)],
                         ids=["adaptive", "dense-keyframes", "mulran-shaped"])
def test_odom_loop_on_the_device_matches_the_loop_over_the_oracle(params, mulran):
    """BASELINE config 4 in miniature: the same seeded OS1-64 sequence (sensor moving during every scan, deskewed with
    per-stamp priors, voxel filtered, keyframes and submap rebuilt) through the loop over the CUDA path and through the
    loop over the CPU oracle: same keyframes, same submap sets, same iteration counts, poses within 1e-4 m / 1e-5 rad."""
    import ngicp, oracle
    import scenarios as S
    from odom_backends import OracleBackend
    scene = synth.Scene(3)
    # MulRan-shaped input (BASELINE config 5): every time stamp is zero, so a scan is one deskew group
    # (file_player_mulran/src/ROSThread.cpp:509-518) and the sensor is taken to be at rest during a scan
    n, w, groups = 14, 256, (1 if mulran else 8)
    seq = list(odom.synthetic_sequence(scene, n, seed=3, step=0.3, w=w, groups=groups, mulran=mulran))
    rng = np.random.default_rng(5)
    drift = [synth.random_se3(rng, 0.03, 0.3) for _ in range(n)]

    def run(backend):
        loop = odom.OdomLoop(backend, params)
        res = []
        for i, (rec, Ts, block, col_t) in enumerate(seq):
            if i == 0:
                loop.T = Ts[groups // 2].astype(np.float32)
                loop.propagateGICP()
                res.append(loop.callbackPointCloud(rec, None))
                continue
            # the prior the IMU integration would deliver: true motion inside the scan, a small seeded drift on top
            def prior(stamps, Ts=Ts, i=i):
                g = np.minimum((stamps.astype(np.int64) * groups) // 100_000_000, groups - 1)
                return np.stack([(drift[i] @ Ts[k]).astype(np.float32) for k in g])
            res.append(loop.callbackPointCloud(rec, prior))
        return loop, res

    g = S.configure(ngicp.NanoGICP(0), max_corr=0.5, max_iter=32, rot_eps=0.01, trans_eps=0.01)
    o = S.configure(oracle.OracleGICP("port"), max_corr=0.5, max_iter=32, rot_eps=0.01, trans_eps=0.01)
    lg, rg = run(odom.DeviceBackend(g))
    lo, ro = run(OracleBackend(o))
    assert len(lg.keyframes) == len(lo.keyframes) >= 2
    for a, b in zip(rg, ro):
        assert a.n_points == b.n_points and a.new_keyframe == b.new_keyframe and a.submap == b.submap
        assert a.iterations == b.iterations and a.converged == b.converged
        assert np.abs(a.T[:3, 3] - b.T[:3, 3]).max() < 1e-4
        assert np.abs(a.T[:3, :3] - b.T[:3, :3]).max() < 1e-5
    # and the loop tracks the truth: the drift of the prior is removed by the registration
    err = [np.abs(r.T[:3, 3] - s[1][groups // 2][:3, 3]).max() for r, s in zip(rg[1:], seq[1:])]
    assert max(err) < 0.02


def _native_hull(P, concave, alpha=0.0):
    import ctypes as C
    from ngicp import binding as B
    P = np.ascontiguousarray(P, np.float64)
    out = np.empty(len(P), np.int32)
    m = B.lib().ngicp_hull_planar(P.ctypes.data_as(C.POINTER(C.c_double)), len(P), int(concave), float(alpha), out.ctypes.data_as(C.POINTER(C.c_int)))
    return m if m < 0 else out[:m].tolist()


def test_native_planar_hulls_match_the_qhull_ones():
    """csrc/odom_loop.cu's own 2-D convex hull and alpha shape (what the C++ loop uses for planar keyframe sets) against
    the scipy/qhull restatement of PCL's ConvexHull / ConcaveHull used by the Python loop; host code, no GPU needed."""
    rng = np.random.default_rng(0)
    for trial in range(40):
        n = int(rng.integers(4, 120))
        P = np.zeros((n, 3))
        if trial % 2:       # a drive: a noisy path
            P[:, :2] = np.cumsum(rng.normal(0, 1.0, (n, 2)) + [1.0, 0.2], 0)
        else:
            P[:, :2] = rng.uniform(-20, 20, (n, 2))
        P[:, 2] = rng.normal(0, 1e-3, n)
        R = synth.rot_from_rotvec(rng.normal(0, 0.3, 3)) if trial % 3 == 0 else np.eye(3)
        P = (P @ R.T).astype(np.float32).astype(np.float64)
        assert _native_hull(P, False) == odom.convex_hull_indices(P)
        for alpha in (0.7, 2.0, 10.0):
            assert _native_hull(P, True, alpha) == odom.concave_hull_indices(P, alpha), (trial, alpha)
    line = np.array([[i, 2 * i, 0.0] for i in range(6)])
    assert _native_hull(line, False) == [0, 5]
    cube = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)] + [[0.5, 0.5, 0.5]], float)
    assert _native_hull(cube, False) == -3                                     # spatial: the callback case


def _native_hull_3d(P, concave, alpha=0.0):
    import ctypes as C
    from ngicp import binding as B
    P = np.ascontiguousarray(P, np.float64)
    out = np.empty(len(P), np.int32)
    m = B.lib().ngicp_hull_spatial(P.ctypes.data_as(C.POINTER(C.c_double)), len(P), int(concave), float(alpha), out.ctypes.data_as(C.POINTER(C.c_int)))
    return m if m < 0 else out[:m].tolist()


def test_native_spatial_hulls_match_the_qhull_ones():
    """The 3-D convex hull (incremental insertion) and 3-D alpha shape (Bowyer-Watson Delaunay) the C++ loop uses for spatial
    keyframe sets when no callbacks are installed, against the scipy/qhull restatement of PCL's hulls; host code, no GPU."""
    rng = np.random.default_rng(1)
    for trial in range(45):
        n = int(rng.integers(5, 160))
        if trial % 3 == 0:
            P = rng.uniform(-20, 20, (n, 3))
        elif trial % 3 == 1:
            P = np.cumsum(rng.normal(0, 1.0, (n, 3)) + [1.0, 0.2, 0.3], 0)          # a flight path
        else:
            P = rng.normal(0, 5, (n, 3)) * [1, 1, 0.3]
        P = P.astype(np.float32).astype(np.float64)
        assert odom._hull_dimension(P)[0] == 3
        assert _native_hull_3d(P, False) == odom.convex_hull_indices(P), trial
        for alpha in (1.0, 3.0, 8.0, 50.0):
            assert _native_hull_3d(P, True, alpha) == odom.concave_hull_indices(P, alpha), (trial, alpha)
    cube = np.array([[x, y, z] for x in (0, 1) for y in (0, 1) for z in (0, 1)] + [[0.5, 0.5, 0.5]], float)
    assert _native_hull_3d(cube, False) == list(range(8)) == _native_hull_3d(cube, True, 2.0)      # eight cospherical corners
    flat = np.concatenate([rng.uniform(-5, 5, (30, 2)), np.zeros((30, 1))], 1)
    assert _native_hull_3d(flat, False) == -2                                  # coplanar: the planar route's business


@pytest.mark.gpu
@pytest.mark.parametrize("params,mulran", [(odom.OdomParams(), False), (odom.OdomParams(adaptive=False, keyframe_thresh_dist=0.5), False),
                                           (odom.OdomParams(adaptive=False, keyframe_thresh_dist=0.5), True)],
                         ids=["adaptive", "dense-keyframes", "mulran-shaped"])
def test_the_cpp_loop_makes_the_python_loops_decisions(params, mulran):
    _cpp_vs_python_loop(params, mulran, climb=0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("callbacks", [False, True], ids=["native-3d-hulls", "qhull-callbacks"])
def test_the_cpp_loop_on_a_climbing_trajectory_uses_spatial_hulls(callbacks):
    """The sensor also moves up and down (0.9 m amplitude): the keyframe set becomes spatial, PCL's dimension switch picks
    3-D hulls, and the C++ loop — with its native 3-D hull / alpha shape, and with the qhull restatement installed as
    callbacks — still makes the Python loop's decisions."""
    _cpp_vs_python_loop(odom.OdomParams(adaptive=False, keyframe_thresh_dist=0.5), False, climb=0.9, callbacks=callbacks)


def _cpp_vs_python_loop(params, mulran, climb, callbacks=False):
    """ngicp_odom_* (csrc/odom_loop.cu) against OdomLoop over DeviceBackend, both on the CUDA path, same sequence: same
    points, keyframes, submap sets, iteration counts; poses to fp32 rounding of one 4x4 product."""
    import ngicp
    import scenarios as S
    scene = synth.Scene(3)
    n, w, groups = 30, 256, (1 if mulran else 8)
    if climb:
        poses = odom.synthetic_poses(scene, n, seed=3, step=0.3)
        poses = [P.copy() for P in poses]
        for i, P in enumerate(poses):
            P[2, 3] += climb * np.sin(0.45 * max(i - 1, 0))                     # at rest during scan 0, then up and down
            # (a steady climb along a nearly straight path is still a PLANAR keyframe set for PCL's dimension switch)
        seq = [odom.synthetic_scan(scene, poses, i, 3, w, groups, mulran) for i in range(n)]
    else:
        seq = list(odom.synthetic_sequence(scene, n, seed=3, step=0.3, w=w, groups=groups, mulran=mulran))
    rng = np.random.default_rng(5)
    drift = [synth.random_se3(rng, 0.03, 0.3) for _ in range(n)]

    def run(make):
        g = S.configure(ngicp.NanoGICP(0), max_corr=0.5, max_iter=32, rot_eps=0.01, trans_eps=0.01)
        loop = make(g)
        res = []
        for i, (rec, Ts, block, col_t) in enumerate(seq):
            if i == 0:
                if isinstance(loop, odom.NativeOdomLoop):
                    loop.set_pose(Ts[groups // 2])
                else:
                    loop.T = Ts[groups // 2].astype(np.float32)
                    loop.propagateGICP()
                res.append(loop.callbackPointCloud(rec, None))
                continue
            def prior(stamps, Ts=Ts, i=i):
                g_ = np.minimum((stamps.astype(np.int64) * groups) // 100_000_000, groups - 1)
                return np.stack([(drift[i] @ Ts[k]).astype(np.float32) for k in g_])
            res.append(loop.callbackPointCloud(rec, prior))
        return res

    rp = run(lambda g: odom.OdomLoop(odom.DeviceBackend(g), params))
    rn = run(lambda g: odom.NativeOdomLoop(g, params, hull_callbacks=callbacks))
    assert sum(r.new_keyframe for r in rn) >= 3
    if climb:
        kf = np.array([r.T[:3, 3] for r in rp if r is not None and r.new_keyframe])
        assert len(kf) >= 5 and odom._hull_dimension(kf.astype(np.float64))[0] == 3       # the spatial route was taken
    for a, b in zip(rn, rp):
        assert a.n_points == b.n_points and a.new_keyframe == b.new_keyframe and a.submap == b.submap and a.submap_changed == b.submap_changed
        assert a.iterations == b.iterations and a.converged == b.converged
        assert np.abs(a.T - b.T).max() < 2e-6 and np.abs(a.T_corr - b.T_corr).max() < 2e-6


def test_synthetic_sequence_is_seeded_per_scan_and_shaped_like_the_reference_point():
    """Records are the reference's 32-byte dlio::Point (include/dlio/dlio.h:85-108); every scan has its own seeded generator,
    so scan i is the same whether it is made alone, in order, or in another process; MulRan-shaped scans carry zero stamps."""
    assert odom.OS1_RECORD.itemsize == 32 and odom.OS1_RECORD.fields["t"][1] == 20
    scene = synth.Scene(1)
    poses = odom.synthetic_poses(scene, 4, seed=7, step=0.3)
    assert np.array_equal(poses[0], poses[1])                                  # at rest during scan 0 (not deskewed, odom.cc:656-664)
    seq = list(odom.synthetic_sequence(scene, 4, seed=7, step=0.3, w=64, groups=4))
    rec2, Ts2, block, col_t = odom.synthetic_scan(scene, poses, 2, seed=7, w=64, groups=4)
    assert rec2.tobytes() == seq[2][0].tobytes() and np.array_equal(Ts2, seq[2][1])
    assert len(rec2) == 64 * 64 and (rec2["w"] == 1.0).all()
    assert rec2["t"].max() == col_t.max() and len(np.unique(rec2["t"])) == 64   # one stamp per column (os_ros.cpp:135-151)
    assert len(Ts2) == 4 and sorted(set(block.tolist())) == [0, 1, 2, 3]
    mul, _, _, _ = odom.synthetic_scan(scene, poses, 2, seed=7, w=64, groups=1, mulran=True)
    assert (mul["t"] == 0).all()                                               # file_player_mulran/src/ROSThread.cpp:509-518
