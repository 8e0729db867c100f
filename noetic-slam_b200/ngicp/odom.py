"""The GICP-relevant part of dlio::OdomNode's per-scan loop, host side, over the device path.

What is restated here is the control flow that decides WHAT the registration kernels see, with the reference's names:
callbackPointCloud (src/dlio/src/dlio/odom.cc:737-837), preprocessPoints / deskewPointcloud (:528-706), computeSpaciousness /
computeDensity / setAdaptiveParams (:1398-1436, :1600-1626), setInputSource (:720-723), initializeInputTarget (:708-718),
getNextPose (:984-1018), propagateGICP (:1230-1246), updateKeyframes (:1517-1598), pushSubmapIndices / buildSubmap /
buildKeyframesAndSubmap (:1628-1780), computeConvexHull / computeConcaveHull (:1438-1515).

Out of scope and therefore inputs of the loop, not parts of it: IMU integration (`integrateImu`, :1056-1228 — the caller
supplies the per-time-stamp prior poses the integration would have produced), the geometric observer (`updateState`,
:1286-1344 — the state position is the GICP pose), ROS publishing, debug output. The reference builds the submap on a
std::async thread and adopts it at the next scan's getNextPose; here the build runs at the end of the scan and is adopted at
the next getNextPose, which is the order of effects the reference has whenever its background build finishes within one
scan period.

The arithmetic (deskew, voxel filter, index build, covariances, alignment, keyframe transform, submap concatenation) runs
behind a small backend interface: `DeviceBackend` below is the product (everything in HBM through the C ABI); the tests
drive the same loop with a backend made of the CPU oracle to check poses, keyframe sets and submap sets scan by scan.

Hulls: PCL's ConvexHull / ConcaveHull wrap qhull, and so does scipy.spatial; PCL itself is third-party and absent from the
reference tree, so the dimension switch and the alpha test are restated from PCL 1.10 (surface/convex_hull.hpp,
concave_hull.hpp) — parity unpinned.
"""
from __future__ import annotations

import heapq
from dataclasses import dataclass, field

import numpy as np


@dataclass
class OdomParams:
    """src/dlio/cfg/params.yaml + cfg/dlio.yaml values that reach this loop."""
    crop_size: float = 1.0                # preprocessing/cropBoxFilter/size (params.yaml:43)
    voxel_res: float = 0.25               # preprocessing/voxelFilter/res (:45)
    keyframe_thresh_dist: float = 1.0     # keyframe/threshD (:48)
    keyframe_thresh_rot: float = 45.0     # keyframe/threshR (:49)
    submap_knn: int = 10                  # submap/keyframe/knn (:53)
    submap_kcv: int = 10                  # kcv (:54)
    submap_kcc: int = 10                  # kcc (:55)
    gicp_min_num_points: int = 64         # gicp/minNumPoints (:57)
    gicp_max_corr_dist: float = 0.5       # gicp/maxCorrespondenceDistance (:59)
    adaptive: bool = True                 # dlio.yaml:17
    time_field: str = "t"                 # Ouster: uint32 ns offsets (odom.cc:603-611)
    time_scale: float = 1e-9


# ------------------------------------------------------------------------------------------------------------- hulls
def _hull_dimension(P: np.ndarray) -> tuple[int, int]:
    """PCL's calculateInputDimension: 2 when the smallest covariance eigenvalue is < 1e-3 of the largest, else 3; for the
    planar case the coordinate dropped is the one the plane normal is most aligned with."""
    c = P - P.mean(0)
    w, V = np.linalg.eigh(c.T @ c / max(len(P) - 1, 1))
    if w[2] <= 0 or w[0] / w[2] < 1e-3:
        return 2, int(np.argmax(np.abs(V[:, 0])))
    return 3, -1


def convex_hull_indices(P: np.ndarray) -> list[int]:
    """Indices of the points on the convex hull (pcl::ConvexHull::getHullPointIndices, odom.cc:1459-1472)."""
    from scipy.spatial import ConvexHull, QhullError
    P = np.asarray(P, np.float64)
    dim, drop = _hull_dimension(P)
    Q = P if dim == 3 else np.delete(P, drop, axis=1)
    try:
        return sorted(int(i) for i in ConvexHull(Q).vertices)
    except QhullError:   # collinear keyframes: the two extremes along the line
        c = Q - Q.mean(0)
        s = c @ np.linalg.svd(c, full_matrices=False)[2][0]
        return sorted({int(np.argmin(s)), int(np.argmax(s))})


def concave_hull_indices(P: np.ndarray, alpha: float) -> list[int]:
    """Indices of the points on the alpha shape (pcl::ConcaveHull with setAlpha, odom.cc:1496-1512): Delaunay simplices
    whose circumradius is <= alpha are kept; the hull is made of the faces that belong to exactly one kept simplex."""
    from scipy.spatial import Delaunay, QhullError
    P = np.asarray(P, np.float64)
    dim, drop = _hull_dimension(P)
    Q = P if dim == 3 else np.delete(P, drop, axis=1)
    try:
        tri = Delaunay(Q)
    except QhullError:
        return []
    # circumradius of every simplex at once: 2 (V_i - V_0) . c = |V_i|^2 - |V_0|^2
    S = tri.simplices
    V = Q[S]                                                      # (m, d+1, d)
    A = 2.0 * (V[:, 1:] - V[:, :1])
    rhs = (V[:, 1:] ** 2).sum(2) - (V[:, :1] ** 2).sum(2)
    det = np.linalg.det(A)
    good = np.abs(det) > 1e-300
    centre = np.zeros((len(S), Q.shape[1]))
    if good.any():
        centre[good] = np.linalg.solve(A[good], rhs[good][..., None])[..., 0]
    keep = good & (np.linalg.norm(centre - V[:, 0], axis=1) <= alpha)
    faces: dict[tuple, int] = {}
    for simplex in S[keep]:
        for skip in range(len(simplex)):
            f = tuple(sorted(int(v) for i, v in enumerate(simplex) if i != skip))
            faces[f] = faces.get(f, 0) + 1
    return sorted({v for f, c in faces.items() if c == 1 for v in f})


def push_submap_indices(dists, k: int, frames) -> list[int]:
    """pushSubmapIndices (odom.cc:1628-1652): every frame whose distance is <= the k-th smallest (ties all kept)."""
    if len(dists) == 0:
        return []
    heap: list[float] = []          # max-heap of the k smallest, as negatives
    for d in dists:
        d = float(np.float32(d))
        if len(heap) >= k and -heap[0] > d:
            heapq.heapreplace(heap, -d)
        elif len(heap) < k:
            heapq.heappush(heap, -d)
    kth = -heap[0]
    return [int(frames[i]) for i in range(len(dists)) if float(np.float32(dists[i])) <= kth]


def quat_from_rot(R) -> np.ndarray:
    """(w, x, y, z) of a rotation matrix, normalised (propagateGICP, odom.cc:1234-1245)."""
    R = np.asarray(R, np.float64)
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[1 + i] = 0.25 * s
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    return q / np.linalg.norm(q)


def quat_angle_deg(q, r) -> float:
    """Angle of q * r^-1 with the sign fix of updateKeyframes (odom.cc:1560-1572)."""
    q = np.asarray(q, np.float64); r = np.asarray(r, np.float64)
    if q @ r < 0:
        r = -r
    ri = np.array([r[0], -r[1], -r[2], -r[3]]) / (r @ r)
    w = q[0] * ri[0] - q[1:] @ ri[1:]
    v = q[0] * ri[1:] + ri[0] * q[1:] + np.cross(q[1:], ri[1:])
    return float(np.degrees(2.0 * np.arctan2(np.linalg.norm(v), w)))


# ----------------------------------------------------------------------------------------------------------- backend
class DeviceBackend:
    """Everything the loop computes, on one B200 through the C ABI (no host copy of a cloud or covariance list except
    the deskewed scan DLIO publishes)."""

    def __init__(self, gicp):
        self.gicp = gicp

    def set_max_correspondence_distance(self, d): self.gicp.setMaxCorrespondenceDistance(d)

    def ingest(self, records, time_field, crop):
        return self.gicp.ingestScan(records, time_field, crop=crop)

    def deskew_filter_set_source(self, frames, leaf):
        return len(self.gicp.deskewScan(frames, leaf=(leaf,) * 3 if leaf else None, set_source=True))

    def calculate_source_covariances(self):
        self.gicp.calculateSourceCovariances()
        return self.gicp.source_density_

    def align(self):
        T = self.gicp.align()
        return T, self.gicp.hasConverged(), self.gicp.nr_iterations_

    def capture_keyframe(self): return self.gicp.captureKeyframe()

    def transform_keyframe(self, kf, T): kf.transform(T)

    def set_submap(self, keyframes): self.gicp.assembleSubmap(keyframes)


# -------------------------------------------------------------------------------------------------------------- loop
@dataclass
class ScanResult:
    T: np.ndarray                  # lidar pose after the scan (4x4 float32)
    T_corr: np.ndarray
    converged: bool
    iterations: int
    n_points: int
    new_keyframe: bool
    submap: list[int] = field(default_factory=list)
    submap_changed: bool = False


class OdomLoop:
    def __init__(self, backend, params: OdomParams | None = None):
        self.b = backend
        self.p = params or OdomParams()
        self.T = np.eye(4, dtype=np.float32)
        self.T_prior = np.eye(4, dtype=np.float32)
        self.T_corr = np.eye(4, dtype=np.float32)
        self.lidar_p = np.zeros(3, np.float32)
        self.lidar_q = np.array([1.0, 0, 0, 0])
        self.keyframes: list[tuple[np.ndarray, np.ndarray, object]] = []     # (p, q, backend keyframe)
        self.keyframe_transformations: list[np.ndarray] = []
        self.num_processed_keyframes = 0
        self.keyframe_convex: list[int] = []
        self.keyframe_concave: list[int] = []
        self.submap_kf_idx_curr: list[int] = []
        self.submap_kf_idx_prev: list[int] = []
        self.submap_hasChanged = True
        self.keyframe_thresh_dist_ = self.p.keyframe_thresh_dist
        self.spaciousness: list[float] = []
        self.density: list[float] = []
        self._median_prev = None
        self._density_prev = None
        self.first_opt_done = False
        self.source_density_ = 0.0
        self.trajectory: list[np.ndarray] = []
        self.concave_alpha = self.p.keyframe_thresh_dist

    # ---- metrics and adaptive parameters (odom.cc:1398-1436, 1600-1626) ----
    def computeSpaciousness(self, x, y):
        d = np.sqrt(x * x + y * y)
        median_curr = float(np.partition(d, len(d) // 2)[len(d) // 2])
        if self._median_prev is None:
            self._median_prev = median_curr
        lpf = np.float32(0.95) * np.float32(self._median_prev) + np.float32(0.05) * np.float32(median_curr)
        self._median_prev = float(lpf)
        self.spaciousness.append(float(lpf))

    def computeDensity(self):
        density = self.source_density_ if self.first_opt_done else 0.0
        if self._density_prev is None:
            self._density_prev = density
        lpf = np.float32(0.95) * np.float32(self._density_prev) + np.float32(0.05) * np.float32(density)
        self._density_prev = float(lpf)
        self.density.append(float(lpf))

    def setAdaptiveParams(self):
        sp = min(max(self.spaciousness[-1], 0.5), 5.0)
        self.keyframe_thresh_dist_ = sp
        den = min(max(self.density[-1], 0.5 * self.p.gicp_max_corr_dist), 2.0 * self.p.gicp_max_corr_dist)
        if sp < 5.0: den = 0.5 * self.p.gicp_max_corr_dist
        if sp > 5.0: den = 2.0 * self.p.gicp_max_corr_dist
        self.b.set_max_correspondence_distance(den)
        self.concave_alpha = self.keyframe_thresh_dist_

    # ---- one scan ----
    def callbackPointCloud(self, records: np.ndarray, prior_frames) -> ScanResult | None:
        """records: structured scan (x, y, z float32 + the time field), sensor frame. prior_frames(stamps) -> (n, 4, 4)
        float32 world poses of the sensor at the unique time stamps (what integrateImu x extrinsics returns); None for
        the first scan (the reference assumes no motion, odom.cc:656-664)."""
        p = self.p
        crop = (np.full(3, -p.crop_size, np.float32), np.full(3, p.crop_size, np.float32), True)
        stamps, kept = self.b.ingest(records, p.time_field, crop)
        if kept == 0:
            return None
        median = len(stamps) // 2
        first = not self.keyframes
        if first or prior_frames is None:
            frames = self.T[None].astype(np.float32)
            self.T_prior = self.T.copy()
        else:
            frames = np.asarray(prior_frames(stamps), np.float32)
            self.T_prior = frames[median].copy()
        n_src = self.b.deskew_filter_set_source(frames, p.voxel_res)
        if n_src <= p.gicp_min_num_points:
            return None                                              # "Low number of points in the cloud!" (odom.cc:764-767)
        # the reference's order (odom.cc:769-779): metrics (the density is still the previous scan's), adaptive parameters,
        # then the source covariances — which is also when the device path starts the first correspondence search
        # original_scan = the cloud after removeNaN + CropBox (odom.cc:490-526); only its planar ranges are needed
        x, y, z = (np.ascontiguousarray(records[f]) for f in ("x", "y", "z"))
        m = np.maximum(np.maximum(np.abs(x), np.abs(y)), np.abs(z))       # NaN propagates and fails both comparisons
        ok = (m >= np.float32(p.crop_size)) & (m < np.float32(np.inf))
        self.computeSpaciousness(x[ok], y[ok])
        self.computeDensity()
        if p.adaptive:
            self.setAdaptiveParams()
        self.source_density_ = self.b.calculate_source_covariances()
        if first:
            self.initializeInputTarget()
            self.buildKeyframesAndSubmap(self.lidar_p)
            return ScanResult(self.T.copy(), self.T_corr.copy(), True, 0, n_src, True, list(self.submap_kf_idx_curr), True)
        changed = self.submap_hasChanged
        res = self.getNextPose()
        new_kf = self.updateKeyframes()
        self.buildKeyframesAndSubmap(self.lidar_p)
        self.trajectory.append(self.T.copy())
        self.first_opt_done = True
        return ScanResult(self.T.copy(), self.T_corr.copy(), res[1], res[2], n_src, new_kf, list(self.submap_kf_idx_curr), changed)

    def initializeInputTarget(self):
        self.keyframes.append((self.lidar_p.copy(), self.lidar_q.copy(), self.b.capture_keyframe()))
        self.keyframe_transformations.append(self.T_corr.copy())

    def getNextPose(self):
        # the submap built after the previous scan is already the target of the backend (set_submap); nothing to adopt
        self.submap_hasChanged = False
        T_corr, conv, it = self.b.align()
        self.T_corr = np.asarray(T_corr, np.float32)
        self.T = (self.T_corr @ self.T_prior).astype(np.float32)
        self.propagateGICP()
        return self.T, conv, it

    def propagateGICP(self):
        self.lidar_p = self.T[:3, 3].copy()
        self.lidar_q = quat_from_rot(self.T[:3, :3])

    def updateKeyframes(self) -> bool:
        P = np.array([k[0] for k in self.keyframes], np.float32)
        d = np.sqrt(((self.lidar_p[None].astype(np.float32) - P) ** 2).sum(1, dtype=np.float32))
        num_nearby = int((d <= np.float32(self.keyframe_thresh_dist_ * 1.5)).sum())
        closest = int(np.argmin(d))
        dd = float(d[closest])
        theta_deg = quat_angle_deg(self.lidar_q, self.keyframes[closest][1])
        new = dd > self.keyframe_thresh_dist_ or abs(theta_deg) > self.p.keyframe_thresh_rot
        if dd <= self.keyframe_thresh_dist_:
            new = False
        if dd <= self.keyframe_thresh_dist_ and abs(theta_deg) > self.p.keyframe_thresh_rot and num_nearby <= 1:
            new = True
        if new:
            self.keyframes.append((self.lidar_p.copy(), self.lidar_q.copy(), self.b.capture_keyframe()))
            self.keyframe_transformations.append(self.T_corr.copy())
        return new

    def buildKeyframesAndSubmap(self, state_p):
        for i in range(self.num_processed_keyframes, len(self.keyframes)):
            self.b.transform_keyframe(self.keyframes[i][2], self.keyframe_transformations[i])      # odom.cc:1757-1762
            self.num_processed_keyframes += 1
        self.buildSubmap(state_p)

    def buildSubmap(self, state_p):
        n = self.num_processed_keyframes
        P = np.array([k[0] for k in self.keyframes[:n]], np.float32)
        ds = np.sqrt(((np.asarray(state_p, np.float32)[None] - P) ** 2).sum(1, dtype=np.float32))
        cur = push_submap_indices(ds, self.p.submap_knn, list(range(n)))
        if n >= 4:
            self.keyframe_convex = convex_hull_indices(P)
        cur += push_submap_indices([ds[c] for c in self.keyframe_convex], self.p.submap_kcv, self.keyframe_convex)
        if n >= 5:
            self.keyframe_concave = concave_hull_indices(P, self.concave_alpha)
        cur += push_submap_indices([ds[c] for c in self.keyframe_concave], self.p.submap_kcc, self.keyframe_concave)
        self.submap_kf_idx_curr = sorted(set(cur))
        if self.submap_kf_idx_curr != self.submap_kf_idx_prev:
            self.submap_hasChanged = True
            self.b.set_submap([self.keyframes[k][2] for k in self.submap_kf_idx_curr])                # odom.cc:1719-1738
            self.submap_kf_idx_prev = list(self.submap_kf_idx_curr)


# -------------------------------------------------------------------------------------------------- the C++ loop
class NativeOdomLoop:
    """The same loop in C++ behind the C ABI (csrc/odom_loop.cu: ngicp_odom_*): one ctypes call before and one after the
    caller's IMU integration per scan, instead of ~12 Python-level calls. Same decisions as OdomLoop scan by scan
    (tests/test_odom.py). Hulls are native (2-D for planar keyframe sets, 3-D otherwise)."""

    def __init__(self, gicp, params: OdomParams | None = None, record_dtype=None, hull_callbacks: bool = False):
        import ctypes as C
        from . import binding as B
        self._C, self._B, self._L = C, B, B.lib()
        self.gicp = gicp
        self.p = params or OdomParams()
        rd = np.dtype(record_dtype if record_dtype is not None else OS1_RECORD)
        tdt, toff = rd.fields[self.p.time_field][:2]
        cp = B.OdomParamsC()
        self._L.ngicp_odom_default_params(C.byref(cp))
        for f in ("crop_size", "voxel_res", "keyframe_thresh_dist", "keyframe_thresh_rot", "submap_knn", "submap_kcv", "submap_kcc",
                  "gicp_min_num_points", "gicp_max_corr_dist"):
            setattr(cp, f, getattr(self.p, f))
        cp.adaptive = int(self.p.adaptive)
        cp.time_offset_bytes = int(toff)
        cp.time_type = {np.dtype(np.uint32): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}[tdt]
        self._o = C.c_void_p(None)
        B.check(gicp._h, self._L.ngicp_odom_create(gicp._h, C.byref(cp), C.byref(self._o)))

        def _cb(fn):
            def call(xyz, n, alpha, out, _user):
                try:
                    P = np.ctypeslib.as_array(xyz, shape=(n, 3)).copy()
                    idx = fn(P, alpha)
                    for j, v in enumerate(idx):
                        out[j] = int(v)
                    return len(idx)
                except Exception:       # noqa: BLE001 - reported by the C side as a failed callback
                    return -1
            return B.HULL_FN(call)
        # planar AND spatial keyframe sets are handled natively; hull_callbacks=True routes spatial sets through the qhull
        # restatement above instead (what a C++ caller would do with pcl::ConvexHull / pcl::ConcaveHull)
        self._cbs = None
        if hull_callbacks:
            self._cbs = (_cb(lambda P, a: convex_hull_indices(P)), _cb(lambda P, a: concave_hull_indices(P, a)))
            self._L.ngicp_odom_set_hull_callbacks(self._o, self._cbs[0], self._cbs[1], None)
        self._stamps = np.empty(0, np.float64)
        self._ids = np.empty(4096, np.int32)
        self.n_keyframes = 0

    def _check(self, rc):
        if rc != self._B.OK:
            msg = self._L.ngicp_odom_last_error(self._o)
            raise self._B.NgicpError(rc, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "_o", None) is not None and self._o.value:
            self._L.ngicp_odom_destroy(self._o)
            self._o = self._C.c_void_p(None)

    def __del__(self):
        try:
            if self.gicp._h.value:
                self.close()
        except Exception:       # noqa: BLE001
            pass

    STAGES = ("ingest", "ranges", "deskew_voxel_index", "median_params", "covariances", "align", "keyframe", "submap")

    def profile(self, reset=False) -> dict:
        """Mean host wall clock per stage of the C++ loop, milliseconds per scan."""
        C = self._C
        sec = (C.c_double * len(self.STAGES))()
        n = C.c_long(0)
        self._check(self._L.ngicp_odom_get_profile(self._o, sec, C.byref(n), int(reset)))
        return {"scans": n.value, **{k: 1e3 * v / max(n.value, 1) for k, v in zip(self.STAGES, sec)}}

    def set_pose(self, T):
        """Initial pose of the lidar (the reference starts at the configured origin; the tests start on the truth)."""
        t = np.ascontiguousarray(np.asarray(T, np.float32).T).reshape(16)
        self._check(self._L.ngicp_odom_set_pose(self._o, t.ctypes.data_as(self._C.POINTER(self._C.c_float))))

    def callbackPointCloud(self, records: np.ndarray, prior_frames) -> ScanResult | None:
        C = self._C
        rec = np.ascontiguousarray(records)
        if len(self._stamps) < len(rec):
            self._stamps = np.empty(len(rec), np.float64)
        nu, nk = C.c_size_t(0), C.c_size_t(0)
        self._check(self._L.ngicp_odom_scan_begin(self._o, rec.ctypes.data, len(rec), rec.dtype.itemsize,
                                                 self._stamps.ctypes.data_as(C.POINTER(C.c_double)), C.byref(nu), C.byref(nk)))
        fp = C.POINTER(C.c_float)
        res = self._B.OdomResultC()
        ids = self._ids.ctypes.data_as(C.POINTER(C.c_int))
        if prior_frames is None or nk.value == 0:
            rc = self._L.ngicp_odom_scan_finish(self._o, None, 0, C.byref(res), ids, len(self._ids))
        else:
            F = np.asarray(prior_frames(self._stamps[:nu.value]), np.float32)
            Fc = np.ascontiguousarray(F.transpose(0, 2, 1)).reshape(-1, 16)
            rc = self._L.ngicp_odom_scan_finish(self._o, Fc.ctypes.data_as(fp), len(Fc), C.byref(res), ids, len(self._ids))
        self._check(rc)
        if not res.valid:
            return None
        self.n_keyframes = res.n_keyframes
        T = np.array(res.T, np.float32).reshape(4, 4).T.copy()
        Tc = np.array(res.T_corr, np.float32).reshape(4, 4).T.copy()
        return ScanResult(T, Tc, bool(res.converged), res.iterations, res.n_points, bool(res.new_keyframe),
                          self._ids[:res.n_submap].tolist(), bool(res.submap_changed))


# ---------------------------------------------------------------------------------------------- synthetic sequences
OS1_RECORD = np.dtype([("x", np.float32), ("y", np.float32), ("z", np.float32), ("w", np.float32), ("intensity", np.float32),
                       ("t", np.uint32), ("pad", np.uint32, 2)])      # 32 B, dlio::Point (include/dlio/dlio.h:85-108)


def synthetic_poses(scene, n_scans: int, seed: int = 0, step: float = 0.25):
    """Sensor poses at the scan boundaries; the sensor is at rest during scan 0, which the reference does not deskew
    (odom.cc:656-664)."""
    from . import synth
    poses = synth.trajectory(scene, n_scans + 1, seed, step=step)
    return [poses[0]] + poses


def synthetic_scan(scene, poses, i: int, seed: int = 0, w: int = 1024, groups: int = 8, mulran: bool = False):
    """Scan i of a seeded OS1-64 sequence at 10 Hz (BASELINE configs 4/5): (records, sensor poses at the `groups` column
    blocks, column block of every column, column time stamps). The sensor moves DURING the scan: block g of columns is
    cast from the pose interpolated at its time, so the deskew frames matter. mulran=True zeroes the time field (one deskew
    stamp, file_player_mulran/src/ROSThread.cpp:509-518). Every scan has its own seeded generator, so scans can be made in
    any order (or in parallel)."""
    from . import synth
    rng = np.random.default_rng([seed, i])
    col_t = (np.arange(w) * (100e6 / w)).astype(np.uint32)              # ns since the scan start (os_ros.cpp:135-151)
    block = np.minimum((np.arange(w) * groups) // w, groups - 1)
    A, Bp = poses[i], poses[i + 1]
    rel = np.linalg.inv(A) @ Bp
    rv = _rotvec(rel[:3, :3])
    Ts = [A @ synth.se3(rv * (g + 0.5) / groups, rel[:3, 3] * (g + 0.5) / groups) for g in range(groups)]
    pts = np.empty((64, w, 3), np.float32)
    for g in range(groups):
        full = synth.scan(scene, Ts[g], rng, w=w, keep_all=True).reshape(64, w, 3)
        pts[:, block == g] = full[:, block == g]
    rec = np.zeros(64 * w, OS1_RECORD)
    flat = pts.reshape(-1, 3)
    rec["x"], rec["y"], rec["z"], rec["w"] = flat[:, 0], flat[:, 1], flat[:, 2], 1.0
    rec["t"] = 0 if mulran else np.broadcast_to(col_t[None], (64, w)).reshape(-1)
    return rec, np.asarray(Ts, np.float64), block, col_t


def synthetic_sequence(scene, n_scans: int, seed: int = 0, step: float = 0.25, w: int = 1024, groups: int = 8, mulran: bool = False):
    poses = synthetic_poses(scene, n_scans, seed, step)
    for i in range(n_scans):
        yield synthetic_scan(scene, poses, i, seed, w, groups, mulran)


def _rotvec(R) -> np.ndarray:
    c = np.clip((np.trace(R) - 1) / 2, -1, 1)
    th = np.arccos(c)
    if th < 1e-12:
        return np.zeros(3)
    return th / (2 * np.sin(th)) * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
