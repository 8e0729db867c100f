import numpy as np

from ngicp import synth


def test_os1_64_scan_shape_and_determinism():
    sc = synth.Scene(0)
    T = synth.trajectory(sc, 1, 0)[0]
    s1 = synth.scan(sc, T, np.random.default_rng(3), keep_all=True)
    s2 = synth.scan(sc, T, np.random.default_rng(3), keep_all=True)
    assert s1.shape == (65536, 3) and s1.dtype == np.float32 and (s1 == s2).all()
    r = np.linalg.norm(s1, axis=1)
    assert r.min() >= 0.99 and r.max() <= 120.01
    d = synth.ray_dirs()
    assert d.shape == (65536, 3) and np.allclose(np.linalg.norm(d, axis=1), 1.0)


def test_voxel_filter_and_submap():
    sc = synth.Scene(1)
    tgt, bounds, poses = synth.make_submap(sc, 20000, 1, n_keyframes=4, w=128)
    assert tgt.shape == (20000, 3) and bounds[-1] == 20000 and len(poses) == 4
    p = np.random.default_rng(0).uniform(-1, 1, (5000, 3)).astype(np.float32)
    v = synth.voxel_filter(p, 0.25)
    assert len(v) <= 512 and len(v) > 400
    aos = synth.to_aos32(v)
    assert aos.shape[1] == 8 and (aos[:, 3] == 1).all() and aos.strides[0] == 32
