"""TEST INFRASTRUCTURE: the odom loop's backend interface (ngicp/odom.py) implemented with the CPU oracle — the host
pipeline the reference runs: deskew + VoxelGrid restated in oracle.cc, nano_gicp through OracleGICP, keyframes as host
arrays transformed and concatenated the way dlio::OdomNode does (reference src/dlio/src/dlio/odom.cc:1719-1762)."""
import numpy as np
import oracle


class OracleBackend:
    def __init__(self, gicp: "oracle.OracleGICP"):
        self.gicp = gicp
        self._scan = None
        self._src = None

    def set_max_correspondence_distance(self, d): self.gicp.setMaxCorrespondenceDistance(d)

    def ingest(self, records, time_field, crop):
        xyz, grp, stamps = oracle.scan_ingest(records, time_field, crop=crop)
        self._scan = (xyz, grp)
        return stamps, len(xyz)

    def deskew_filter_set_source(self, frames, leaf):
        cloud = oracle.scan_deskew(self._scan[0], self._scan[1], frames)
        if leaf:
            v = oracle.VoxelGrid(); v.setLeafSize(leaf); v.setInputCloud(cloud)
            cloud = v.filter()
        self._src = cloud
        self.gicp.setInputSource(cloud)
        return len(cloud)

    def calculate_source_covariances(self):
        self.gicp.calculateSourceCovariances()
        return self.gicp.source_density_

    def align(self):
        T = self.gicp.align()
        return T, self.gicp.hasConverged(), self.gicp.nr_iterations_

    def capture_keyframe(self):
        return {"xyz": self._src.copy(), "cov": self.gicp.getSourceCovariances().copy()}

    def transform_keyframe(self, kf, T):
        T = np.asarray(T, np.float32)
        p = kf["xyz"]
        # pcl::transformPointCloud, fp32: ((m0 x + m1 y) + m2 z) + m3
        kf["xyz"] = np.stack([((T[r, 0] * p[:, 0] + T[r, 1] * p[:, 1]) + T[r, 2] * p[:, 2]) + T[r, 3] for r in range(3)], 1).astype(np.float32)
        Td = T.astype(np.float64)
        kf["cov"] = Td[None] @ kf["cov"] @ Td.T[None]

    def set_submap(self, keyframes):
        self.gicp.setInputTarget(np.concatenate([k["xyz"] for k in keyframes]))
        self.gicp.setTargetCovariances(np.concatenate([k["cov"] for k in keyframes]))


class FakeBackend:
    """No arithmetic at all: scripted poses, for the host-logic tests of the keyframe / submap policy."""

    def __init__(self, corrections):
        self.corrections = list(corrections)
        self.submaps, self.captured, self.transformed, self.max_corr = [], 0, [], []

    def set_max_correspondence_distance(self, d): self.max_corr.append(d)
    def ingest(self, records, time_field, crop): return np.arange(4, dtype=np.float64), len(records)
    def deskew_filter_set_source(self, frames, leaf): return 1000
    def calculate_source_covariances(self): return 0.1
    def align(self): return self.corrections.pop(0), True, 3
    def capture_keyframe(self): self.captured += 1; return self.captured - 1
    def transform_keyframe(self, kf, T): self.transformed.append(kf)
    def set_submap(self, keyframes): self.submaps.append(list(keyframes))
