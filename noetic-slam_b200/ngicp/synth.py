"""Seeded synthetic Ouster OS1-64-shaped scans, keyframe submaps and trajectories.

Ray directions follow the formula the reference's vendored Ouster SDK uses to build its XYZ lookup
table (reference src/ouster/ouster-sdk/ouster_client/src/lidar_scan.cpp:326-356):
    enc = 2*pi - v*2*pi/W ; az = -azimuth_deg[u]*pi/180 ; alt = altitude_deg[u]*pi/180
    dir = (cos(enc+az)cos(alt), sin(enc+az)cos(alt), sin(alt))
with an OS1-64-shaped beam table (64 near-uniform altitude angles +16.729..-16.679 deg, azimuth
offsets cycling {+3.07,+0.90,-1.26,-3.38} deg; shape of the table in
src/ouster/ouster-sdk/tests/metadata/2_1_2_os1-991913000010-64.json, not a copy of it).
1024 columns x 64 beams = 65,536 returns per scan, 10 Hz.

The scene is analytic (ground + 80x50x12 m hall + boxes + pillars); ranges get N(0, 2 cm) noise so
that no neighbourhood is exactly planar (SURVEY.md hard part 2). Everything is numpy, fp32 out.
"""
from __future__ import annotations

import numpy as np

H_BEAMS = 64
W_COLS = 1024


def os1_64_beams():
    alt = np.linspace(16.729, -16.679, H_BEAMS)
    az = np.tile(np.array([3.07, 0.90, -1.26, -3.38]), H_BEAMS // 4)
    return alt, az


def ray_dirs(w: int = W_COLS) -> np.ndarray:
    """(64*w, 3) float64 unit directions in the sensor frame, row-major over (beam u, column v)."""
    alt, az = os1_64_beams()
    v = np.arange(w)
    enc = 2 * np.pi - v * 2 * np.pi / w
    ang = enc[None, :] + (-az * np.pi / 180.0)[:, None]
    ca = np.cos(alt * np.pi / 180.0)[:, None]
    sa = np.sin(alt * np.pi / 180.0)[:, None]
    d = np.stack([np.cos(ang) * ca, np.sin(ang) * ca, np.broadcast_to(sa, ang.shape)], axis=-1)
    return d.reshape(-1, 3)


# ----------------------------------------------------------------------------- SE(3) helpers
def rot_from_rotvec(w) -> np.ndarray:
    w = np.asarray(w, np.float64)
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def se3(rotvec=(0, 0, 0), t=(0, 0, 0)) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = rot_from_rotvec(rotvec)
    T[:3, 3] = t
    return T


def transform_points(T, pts) -> np.ndarray:
    p = np.asarray(pts, np.float64)[:, :3]
    return (p @ T[:3, :3].T + T[:3, 3]).astype(np.float32)


def random_se3(rng, max_t: float, max_deg: float) -> np.ndarray:
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = np.deg2rad(rng.uniform(0.3, 1.0) * max_deg)
    t = rng.normal(size=3)
    t = t / np.linalg.norm(t) * rng.uniform(0.3, 1.0) * max_t
    return se3(ax * ang, t)


# ----------------------------------------------------------------------------- scene
class Scene:
    """Ground z=0 inside an 80 x 50 x 12 m hall, plus seeded boxes and vertical pillars."""

    def __init__(self, seed: int = 0, n_boxes: int = 14, n_pillars: int = 10):
        rng = np.random.default_rng(seed + 7919)
        self.lo = np.array([-40.0, -25.0, 0.0])
        self.hi = np.array([40.0, 25.0, 12.0])
        c = np.stack([rng.uniform(-36, 36, n_boxes), rng.uniform(-22, 22, n_boxes)], 1)
        s = rng.uniform(0.8, 4.0, (n_boxes, 2))
        h = rng.uniform(0.8, 5.0, n_boxes)
        self.box_lo = np.concatenate([c - s / 2, np.zeros((n_boxes, 1))], 1)
        self.box_hi = np.concatenate([c + s / 2, h[:, None]], 1)
        self.pil_c = np.stack([rng.uniform(-36, 36, n_pillars), rng.uniform(-22, 22, n_pillars)], 1)
        self.pil_r = rng.uniform(0.25, 0.9, n_pillars)
        self.pil_h = rng.uniform(3.0, 12.0, n_pillars)

    def clear_of_obstacles(self, xy, margin=1.5) -> bool:
        xy = np.asarray(xy)
        inside = np.all((xy > self.box_lo[:, :2] - margin) & (xy < self.box_hi[:, :2] + margin), axis=1)
        near = np.linalg.norm(self.pil_c - xy, axis=1) < self.pil_r + margin
        return not (inside.any() or near.any())

    def cast(self, o, d) -> np.ndarray:
        """Nearest hit range along rays o + s*d (o: (3,), d: (M,3) unit). inf where nothing is hit."""
        with np.errstate(divide="ignore", invalid="ignore"):
            # hall interior: exit distance through the enclosing box
            s_ax = np.where(d > 0, (self.hi - o) / d, np.where(d < 0, (self.lo - o) / d, np.inf))
            best = s_ax.min(axis=1)
            # obstacle boxes: slab test
            for lo, hi in zip(self.box_lo, self.box_hi):
                t1 = (lo - o) / d
                t2 = (hi - o) / d
                tn = np.minimum(t1, t2).max(axis=1)
                tf = np.maximum(t1, t2).min(axis=1)
                hit = (tf >= tn) & (tn > 0)
                best = np.where(hit & (tn < best), tn, best)
            # pillars: vertical cylinders
            for c, r, h in zip(self.pil_c, self.pil_r, self.pil_h):
                oc = o[:2] - c
                a = d[:, 0] ** 2 + d[:, 1] ** 2
                b = 2 * (oc[0] * d[:, 0] + oc[1] * d[:, 1])
                cc = oc @ oc - r * r
                disc = b * b - 4 * a * cc
                s = (-b - np.sqrt(np.where(disc > 0, disc, np.nan))) / (2 * a)
                z = o[2] + s * d[:, 2]
                hit = (disc > 0) & (s > 0) & (z >= 0) & (z <= h)
                best = np.where(hit & (s < best), s, best)
        return best


def scan(scene: Scene, T_ws, rng, w: int = W_COLS, noise: float = 0.02, rmin: float = 1.0, rmax: float = 120.0,
         keep_all: bool = False) -> np.ndarray:
    """One OS1-64 scan taken at world pose T_ws (sensor->world). Returns SENSOR-frame points, fp32 (N,3).

    Returns dropped by the range gates are removed (DLIO's crop box, params.yaml:43) unless
    keep_all, in which case they are clamped into range so the scan has exactly 64*w points.
    """
    dirs = ray_dirs(w)
    o = T_ws[:3, 3]
    dw = dirs @ T_ws[:3, :3].T
    r = scene.cast(o, dw)
    r = r + rng.normal(0.0, noise, r.shape)
    ok = np.isfinite(r) & (r >= rmin) & (r <= rmax)
    if keep_all:
        r = np.clip(np.where(np.isfinite(r), r, rmax), rmin, rmax)
        return (dirs * r[:, None]).astype(np.float32)
    return (dirs[ok] * r[ok, None]).astype(np.float32)


def voxel_filter(pts, leaf: float = 0.25) -> np.ndarray:
    """pcl::VoxelGrid centroid semantics: one centroid per occupied voxel,
    emitted in ascending linear voxel index (x fastest)."""
    p = np.asarray(pts, np.float32)[:, :3]
    if len(p) == 0:
        return p.copy()
    inv = np.float32(1.0 / leaf)
    mn = np.floor(p.min(0) * inv).astype(np.int64)
    mx = np.floor(p.max(0) * inv).astype(np.int64)
    ijk = np.floor(p * inv).astype(np.int64) - mn
    div = mx - mn + 1
    lin = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    uniq, inverse, counts = np.unique(lin, return_inverse=True, return_counts=True)
    out = np.zeros((len(uniq), 3), np.float64)
    for a in range(3):
        out[:, a] = np.bincount(inverse, weights=p[:, a].astype(np.float64), minlength=len(uniq))
    return (out / counts[:, None]).astype(np.float32)


def trajectory(scene: Scene, n: int, seed: int = 0, step: float = 0.5, height: float = 1.8) -> list[np.ndarray]:
    """n sensor->world poses on a smooth seeded path that stays clear of the obstacles."""
    rng = np.random.default_rng(seed + 104729)
    poses = []
    xy = np.array([-20.0, -8.0])
    yaw = rng.uniform(-0.3, 0.3)
    for _ in range(200):
        if scene.clear_of_obstacles(xy):
            break
        xy = np.array([rng.uniform(-30, -10), rng.uniform(-15, 0)])
    while len(poses) < n:
        roll, pitch = rng.normal(0, 0.01, 2)
        T = se3((0, 0, yaw), (xy[0], xy[1], height)) @ se3((roll, pitch, 0))
        poses.append(T)
        for _ in range(50):
            dyaw = rng.normal(0, np.deg2rad(2.0))
            nxt = xy + step * np.array([np.cos(yaw + dyaw), np.sin(yaw + dyaw)])
            if abs(nxt[0]) < 34 and abs(nxt[1]) < 20 and scene.clear_of_obstacles(nxt):
                xy, yaw = nxt, yaw + dyaw
                break
            yaw += np.deg2rad(25.0)
    return poses


def pad_or_trim(pts: np.ndarray, n: int, rng) -> np.ndarray:
    """Exactly n points: random subset, or top up with jittered copies (never exact duplicates)."""
    m = len(pts)
    if m >= n:
        sel = np.sort(rng.choice(m, n, replace=False))
        return pts[sel]
    extra = pts[rng.integers(0, m, n - m)] + rng.normal(0, 0.01, (n - m, 3)).astype(np.float32)
    return np.concatenate([pts, extra.astype(np.float32)], 0)


def make_submap(scene: Scene, n_target: int, seed: int = 0, n_keyframes: int = 24, leaf: float = 0.25, w: int = W_COLS):
    """Concatenation of voxel-filtered keyframe scans in the WORLD frame (the reference builds its
    submap this way, src/dlio/odom.cc:1719-1729). Returns (points (n_target,3) fp32, keyframe
    boundaries, poses). Each keyframe's points stay contiguous so per-keyframe covariances can be
    computed once and concatenated (covariance reuse)."""
    rng = np.random.default_rng(seed)
    poses = trajectory(scene, n_keyframes, seed, step=1.0)
    per = n_target // n_keyframes
    clouds, bounds = [], [0]
    for i, T in enumerate(poses):
        s = scan(scene, T, rng, w=w)
        s = voxel_filter(s, leaf) if leaf > 0 else s
        want = per if i < n_keyframes - 1 else n_target - per * (n_keyframes - 1)
        s = pad_or_trim(s, want, rng)
        clouds.append(transform_points(T, s))
        bounds.append(bounds[-1] + len(s))
    return np.concatenate(clouds, 0), np.array(bounds, np.int64), poses


def to_aos32(pts: np.ndarray) -> np.ndarray:
    """(N,3) -> the reference's 32-byte dlio::Point AoS (x,y,z,1,intensity,t,pad,pad) as (N,8) fp32
    (reference src/dlio/include/dlio/dlio.h:85-108)."""
    a = np.zeros((len(pts), 8), np.float32)
    a[:, :3] = pts[:, :3]
    a[:, 3] = 1.0
    return a
