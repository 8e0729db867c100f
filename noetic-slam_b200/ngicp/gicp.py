"""Host-side mirror of the reference interface for the hot path, over the C ABI.

Names, argument meaning and error behaviour follow nano_gicp::NanoGICP<PointSource,PointTarget>
(reference src/dlio/include/nano_gicp/nano_gicp.h:63-150, src/nano_gicp/nano_gicp.cc) and
nanoflann::KdTreeFLANN<PointT> (reference include/nano_gicp/nanoflann_adaptor.h:57-152) so that the
parity tests read like tests of the reference. Clouds are numpy arrays (N, >=3) float32 — the
reference's 32-byte dlio::Point AoS is `synth.to_aos32(...)`, any stride works. 4x4 matrices are
ordinary row-major numpy arrays here; the binding transposes to the ABI's column-major.

All compute happens in libngicp_b200.so (CUDA, sm_100a). Nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import binding as B


def _pts(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] < 3:
        raise ValueError(f"cloud must be (N, >=3) float32, got {a.shape}")
    return a


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _colmajor64(T) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(T, np.float64).T).reshape(16)


class KdTreeFLANN:
    """nanoflann::KdTreeFLANN<PointT>: setInputCloud builds the Morton/voxel-hash index on the GPU,
    nearestKSearch is an exact k-NN (rows ordered by distance, then index)."""

    def __init__(self, owner: "NanoGICP | None" = None, device: int = 0):
        self._L = B.lib()
        self._owner = owner or NanoGICP(device)
        self._idx = C.c_void_p(None)
        self._cloud = None

    @classmethod
    def _adopt(cls, owner, idx_ptr, cloud):
        t = cls(owner)
        t._idx = C.c_void_p(idx_ptr)
        t._L.ngicp_index_retain(t._idx)
        t._cloud = cloud
        return t

    def __del__(self):
        if getattr(self, "_idx", None) and self._idx.value:
            self._L.ngicp_index_release(self._idx)
            self._idx = C.c_void_p(None)

    def setInputCloud(self, cloud):
        cloud = _pts(cloud)
        if self._idx.value:
            self._L.ngicp_index_release(self._idx)
            self._idx = C.c_void_p(None)
        h = self._owner._h
        B.check(h, self._L.ngicp_index_build(h, cloud.ctypes.data, cloud.shape[0], cloud.strides[0], C.byref(self._idx)))
        self._cloud = cloud

    def getInputCloud(self):
        return self._cloud

    def size(self) -> int:
        return int(self._L.ngicp_index_size(self._idx)) if self._idx.value else 0

    def nearestKSearch(self, queries, k: int):
        """Batched: queries (Q, >=3) -> (k_indices (Q,k) int32, k_sqr_distances (Q,k) float32)."""
        if not self._idx.value:
            raise RuntimeError("[nanoflann] findNeighbors() called before building the index.")  # nanoflann.h:1442-1445
        q = _pts(np.atleast_2d(queries))
        idx = np.empty((q.shape[0], k), np.int32)
        sqd = np.empty((q.shape[0], k), np.float32)
        h = self._owner._h
        B.check(h, self._L.ngicp_knn(h, self._idx, q.ctypes.data, q.shape[0], q.strides[0], k, _ptr(idx, C.c_int), _ptr(sqd, C.c_float)))
        return idx, sqd

    def voxel_keys(self):
        """(keys uint64 (N,) in original point order, origin xyz, h0) — the documented key spec."""
        n = self.size()
        keys = np.empty(n, np.uint64)
        oh = np.zeros(4, np.float32)
        h = self._owner._h
        B.check(h, self._L.ngicp_index_keys(h, self._idx, _ptr(keys, C.c_uint64), _ptr(oh, C.c_float)))
        return keys, oh[:3].copy(), float(oh[3])


class Keyframe:
    """A keyframe held in HBM: the cloud and covariance list DLIO stores per keyframe (reference src/dlio/odom.cc:1592),
    captured from the scan that was just registered. Additive API."""

    def __init__(self, owner: "NanoGICP", ptr):
        self._owner, self._L, self._kf = owner, owner._L, C.c_void_p(ptr)

    def __del__(self):
        if getattr(self, "_kf", None) and self._kf.value and self._owner._h.value:
            self._L.ngicp_keyframe_release(self._owner._h, self._kf)
            self._kf = C.c_void_p(None)

    def size(self) -> int:
        return int(self._L.ngicp_keyframe_size(self._kf))

    def transform(self, T):
        """pcl::transformPointCloud + cov <- Td cov Td^T (odom.cc:1757-1762), in place, on the device."""
        t = np.ascontiguousarray(np.asarray(T, np.float32).T).reshape(16)
        B.check(self._owner._h, self._L.ngicp_keyframe_transform(self._owner._h, self._kf, _ptr(t, C.c_float)))

    def download(self):
        n = self.size()
        xyz = np.empty((n, 3), np.float32)
        cov = np.empty((n, 4, 4), np.float64)
        B.check(self._owner._h, self._L.ngicp_keyframe_download(self._owner._h, self._kf, _ptr(xyz, C.c_float), _ptr(cov, C.c_double)))
        return xyz, cov.transpose(0, 2, 1).copy()


class NanoGICP:
    """nano_gicp::NanoGICP<PointT,PointT> on one B200 (one CUDA stream per object)."""

    def __init__(self, device: int = 0):
        self._L = B.lib()
        self._h = C.c_void_p(None)
        B.check(None, self._L.ngicp_create(device, C.byref(self._h)))
        self._p = B.Params()
        self._L.ngicp_default_params(C.byref(self._p))
        self.device = device
        self._input = None
        self._target = None
        self.source_kdtree_ = None
        self.target_kdtree_ = None
        self.source_density_ = 0.0
        self.target_density_ = 0.0
        self.num_correspondences = 0
        self.converged_ = False
        self.nr_iterations_ = 0
        self.final_transformation_ = np.eye(4, dtype=np.float32)

    def close(self):
        """Release the device object now. The trees this object handed out (source_kdtree_ / target_kdtree_) keep a
        reference to it, which is a cycle: without close() the device memory waits for Python's cyclic collector."""
        if getattr(self, "_h", None) and self._h.value:
            for t in (self.source_kdtree_, self.target_kdtree_):
                if t is not None and getattr(t, "_owner", None) is self:
                    t.__del__()            # drop the index reference while the handle is alive
            self.source_kdtree_ = None
            self.target_kdtree_ = None
            self._L.ngicp_destroy(self._h)
            self._h = C.c_void_p(None)

    def __enter__(self): return self
    def __exit__(self, *exc): self.close()

    def __del__(self):
        self.close()

    def _push(self):
        B.check(self._h, self._L.ngicp_set_params(self._h, C.byref(self._p)))

    # ---- setters (nano_gicp.cc:71-94, lsq_registration.cc:73-95) ----
    def setNumThreads(self, n): pass  # the reference's OpenMP team size has no meaning on the GPU
    def setCorrespondenceRandomness(self, k): self._p.k_correspondences = int(k); self._push()
    def setMaxCorrespondenceDistance(self, d): self._p.max_corr_dist = float(d); self._push()
    def setRegularizationMethod(self, m): self._p.regularization = int(m); self._push()
    def setMaximumIterations(self, n): self._p.max_iterations = int(n); self._push()
    def setRotationEpsilon(self, e): self._p.rotation_epsilon = float(e); self._push()
    def setTransformationEpsilon(self, e): self._p.transformation_epsilon = float(e); self._push()
    def setInitialLambdaFactor(self, f): self._p.lm_init_lambda_factor = float(f); self._push()
    def setGaussNewton(self, on=True): self._p.use_gauss_newton = int(bool(on)); self._push()

    # ---- clouds / trees (nano_gicp.cc:97-161) ----
    def _set_input(self, which, cloud):
        cloud = _pts(cloud)
        B.check(self._h, self._L.ngicp_set_input(self._h, which, cloud.ctypes.data, cloud.shape[0], cloud.strides[0]))
        return KdTreeFLANN._adopt(self, self._L.ngicp_get_index(self._h, which), cloud)

    def setInputSource(self, cloud):
        if cloud is self._input:      # pointer-identity early-out, nano_gicp.cc:136
            return
        self.source_kdtree_ = self._set_input(B.SOURCE, cloud)
        self._input = cloud

    def setInputTarget(self, cloud):
        if cloud is self._target:     # nano_gicp.cc:151
            return
        self.target_kdtree_ = self._set_input(B.TARGET, cloud)
        self._target = cloud

    def setInputSourceDevice(self, dev_ptr: int, n: int, token=None):
        """setInputSource for a scan that is already in HBM as packed float4 (bench's device-resident arm)."""
        B.check(self._h, self._L.ngicp_set_input_device(self._h, B.SOURCE, C.c_void_p(dev_ptr), n))
        self.source_kdtree_ = KdTreeFLANN._adopt(self, self._L.ngicp_get_index(self._h, B.SOURCE), token)
        self._input = token

    def setInputSourceFiltered(self, cloud, crop=None, leaf=None):
        """CropBox -> VoxelGrid -> setInputSource without leaving the device (DLIO's getScanFromROS / preprocessPoints
        order, odom.cc:500-502 and :575-584). crop = (min_xyz, max_xyz, negative) or None; leaf = (lx, ly, lz) or None.
        Returns the filtered cloud (host copy, the cloud DLIO would have handed to setInputSource)."""
        out = _filter_scan(self, cloud, crop, leaf, B.SOURCE)
        self.source_kdtree_ = KdTreeFLANN._adopt(self, self._L.ngicp_get_index(self._h, B.SOURCE), out)
        self._input = out
        return out

    def ingestScan(self, records: np.ndarray, time_field: str, crop=None):
        """First half of dlio::OdomNode::deskewPointcloud (odom.cc:588-650) on the device: `records` is a structured
        array with float32 x, y, z at offsets 0/4/8 and a per-point time stamp field (uint32 `t`, float32 `time` or
        float64 `timestamp`). Drops non-finite points, applies the optional CropBox, sorts by time stamp and returns
        (unique stamps ascending as float64 raw field values, points kept)."""
        rec = np.ascontiguousarray(records)
        dt = rec.dtype.fields[time_field][0]
        ttype = {np.dtype(np.uint32): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}[dt]
        fp = C.POINTER(C.c_float)
        if crop is not None:
            mn = np.ascontiguousarray(crop[0], np.float32); mx = np.ascontiguousarray(crop[1], np.float32)
            a_mn, a_mx, neg = mn.ctypes.data_as(fp), mx.ctypes.data_as(fp), int(bool(crop[2]))
        else:
            a_mn = a_mx = None; neg = 0
        stamps = np.empty(len(rec), np.float64)
        nu, nk = C.c_size_t(0), C.c_size_t(0)
        B.check(self._h, self._L.ngicp_scan_ingest(self._h, rec.ctypes.data, len(rec), rec.dtype.itemsize, rec.dtype.fields[time_field][1], ttype,
                                                   a_mn, a_mx, neg, stamps.ctypes.data_as(C.POINTER(C.c_double)), C.byref(nu), C.byref(nk)))
        self._ingested = nk.value
        return stamps[:nu.value].copy(), nk.value

    def deskewScan(self, frames, leaf=None, set_source=True):
        """Second half (odom.cc:690-701, then :575-584 and :721): frames = (n_unique, 4, 4) float32 poses (already times the
        extrinsic), or one (4, 4) matrix for the no-IMU paths. Returns the deskewed (and voxel-filtered) cloud."""
        F = np.asarray(frames, np.float32)
        if F.ndim == 2:
            F = F[None]
        Fc = np.ascontiguousarray(F.transpose(0, 2, 1)).reshape(-1, 16)        # column-major
        lf = None if leaf is None else np.ascontiguousarray(leaf, np.float32)
        fp = C.POINTER(C.c_float)
        out = np.empty((self._ingested, 3), np.float32)
        n_out = C.c_size_t(0)
        B.check(self._h, self._L.ngicp_scan_deskew(self._h, Fc.ctypes.data_as(fp), len(Fc), None if lf is None else lf.ctypes.data_as(fp),
                                                   B.SOURCE if set_source else -1, out.ctypes.data_as(fp), C.byref(n_out)))
        out = out[:n_out.value].copy()
        if set_source:
            self.source_kdtree_ = KdTreeFLANN._adopt(self, self._L.ngicp_get_index(self._h, B.SOURCE), out)
            self._input = out
        return out

    def registerInputSource(self, cloud):  # nano_gicp.cc:119-124: stores the cloud only
        self._input = cloud

    def registerInputTarget(self, cloud):  # nano_gicp.cc:127-132
        self._target = cloud

    def setSourceTree(self, tree: KdTreeFLANN):
        """`source_kdtree_ = tree` (public member assignment in the reference)."""
        B.check(self._h, self._L.ngicp_attach_index(self._h, B.SOURCE, tree._idx))
        self.source_kdtree_ = tree

    def setTargetTree(self, tree: KdTreeFLANN):
        """`gicp.target_kdtree_ = submap_kdtree` (reference src/dlio/odom.cc:995)."""
        B.check(self._h, self._L.ngicp_attach_index(self._h, B.TARGET, tree._idx))
        self.target_kdtree_ = tree

    def swapSourceAndTarget(self):
        B.check(self._h, self._L.ngicp_swap_source_and_target(self._h))
        self._input, self._target = self._target, self._input
        self.source_kdtree_, self.target_kdtree_ = self.target_kdtree_, self.source_kdtree_

    def clearSource(self):
        B.check(self._h, self._L.ngicp_clear(self._h, B.SOURCE)); self._input = None; self.source_kdtree_ = None

    def clearTarget(self):
        B.check(self._h, self._L.ngicp_clear(self._h, B.TARGET)); self._target = None; self.target_kdtree_ = None

    # ---- covariances (nano_gicp.cc:164-191,330-392) ----
    def _calc(self, which):
        d = C.c_float(0)
        B.check(self._h, self._L.ngicp_compute_covariances(self._h, which, C.byref(d)))
        return float(d.value)

    def calculateSourceCovariances(self):
        self.source_density_ = self._calc(B.SOURCE)
        return True

    def calculateTargetCovariances(self):
        self.target_density_ = self._calc(B.TARGET)
        return True

    def selfNeighbours(self, which=B.SOURCE, k=None):
        """Inspection: the neighbour sets calculate_covariances works on (nano_gicp.cc:343), from the production K2 search.
        -> (idx (N,k) int32 original indices, row i = [i, its other k-1 nearest in unspecified order], density terms (N,))."""
        tree = self.source_kdtree_ if which == B.SOURCE else self.target_kdtree_
        n = tree.size()
        k = int(k or self.k_correspondences_)
        idx = np.empty((n, k), np.int32)
        dens = np.empty(n, np.float64)
        B.check(self._h, self._L.ngicp_self_neighbours(self._h, which, k, _ptr(idx, C.c_int), _ptr(dens, C.c_double)))
        return idx, dens

    def _get_covs(self, which):
        n = C.c_size_t(0)
        if not self._L.ngicp_has_covariances(self._h, which, C.byref(n)):
            return None
        out = np.empty((n.value, 4, 4), np.float64)
        B.check(self._h, self._L.ngicp_get_covariances(self._h, which, _ptr(out, C.c_double), n.value))
        return out.transpose(0, 2, 1).copy()

    def getSourceCovariances(self): return self._get_covs(B.SOURCE)
    def getTargetCovariances(self): return self._get_covs(B.TARGET)

    def _set_covs(self, which, covs):
        c = np.ascontiguousarray(np.asarray(covs, np.float64).transpose(0, 2, 1))
        B.check(self._h, self._L.ngicp_set_covariances(self._h, which, _ptr(c, C.c_double), c.shape[0]))

    def setSourceCovariances(self, covs): self._set_covs(B.SOURCE, covs)
    def setTargetCovariances(self, covs): self._set_covs(B.TARGET, covs)

    # ---- registration (nano_gicp.cc:194-326, lsq_registration.cc:108-229) ----
    def update_correspondences(self, T):
        n = self.source_kdtree_.size()
        corr = np.empty(n, np.int32)
        sqd = np.empty(n, np.float32)
        mah = np.empty((n, 4, 4), np.float64)
        nc = C.c_int(0)
        t = _colmajor64(T)
        B.check(self._h, self._L.ngicp_update_correspondences(self._h, _ptr(t, C.c_double), _ptr(corr, C.c_int), _ptr(sqd, C.c_float),
                                                              _ptr(mah, C.c_double), C.byref(nc)))
        self.num_correspondences = nc.value
        return corr, sqd, mah.transpose(0, 2, 1).copy()

    def linearize(self, T):
        H = np.zeros((6, 6), np.float64)
        b = np.zeros(6, np.float64)
        e = C.c_double(0)
        nc = C.c_int(0)
        t = _colmajor64(T)
        B.check(self._h, self._L.ngicp_linearize(self._h, _ptr(t, C.c_double), _ptr(H, C.c_double), _ptr(b, C.c_double), C.byref(e), C.byref(nc)))
        self.num_correspondences = nc.value
        return e.value, H, b

    def compute_error(self, T):
        e = C.c_double(0)
        t = _colmajor64(T)
        B.check(self._h, self._L.ngicp_compute_error(self._h, _ptr(t, C.c_double), C.byref(e)))
        return e.value

    def align(self, guess=None):
        """pcl::Registration::align(output, guess): returns final_transformation_ (4x4 float32)."""
        g = None
        if guess is not None:
            g = np.ascontiguousarray(np.asarray(guess, np.float32).T).reshape(16)
        out = np.zeros(16, np.float32)
        it, conv = C.c_int(0), C.c_int(0)
        H = np.zeros((6, 6), np.float64)
        fe = C.c_double(0)
        rc = self._L.ngicp_align(self._h, _ptr(g, C.c_float) if g is not None else None, _ptr(out, C.c_float), C.byref(it), C.byref(conv),
                                 _ptr(H, C.c_double), C.byref(fe))
        B.check(self._h, rc, allow=(B.ERR_LM_NOT_CONVERGED,))  # the reference prints "lm not converged!!" and carries on
        self.lm_failed_ = rc == B.ERR_LM_NOT_CONVERGED
        self.final_transformation_ = out.reshape(4, 4).T.copy()
        self.nr_iterations_ = it.value
        self.converged_ = bool(conv.value)
        self.final_hessian_ = H
        self.final_error_ = fe.value
        return self.final_transformation_

    def hasConverged(self): return self.converged_
    def getFinalTransformation(self): return self.final_transformation_
    def getFinalHessian(self): return self.final_hessian_
    def getFinalError(self): return self.final_error_

    def transformSource(self, T):
        """pcl::transformPointCloud(*input_, output, final_transformation_) (lsq_registration.cc:133)."""
        n = self.source_kdtree_.size()
        out = np.zeros((n, 3), np.float32)
        t = np.ascontiguousarray(np.asarray(T, np.float32).T).reshape(16)
        B.check(self._h, self._L.ngicp_transform_source(self._h, _ptr(t, C.c_float), out.ctypes.data, n, out.strides[0]))
        return out

    # ---- batched units / instrumentation ----
    def batchCovariances(self, points, seg_offsets, want_mat4=False):
        """Covariances of many keyframes in one pass (BASELINE config 3). Returns (cov6 (N,6) float32
        in original order [, mat4 (N,4,4)], per-keyframe density)."""
        p = _pts(points)
        so = np.ascontiguousarray(seg_offsets, np.int64)
        ns = len(so) - 1
        cov6 = np.empty((p.shape[0], 6), np.float32)
        dens = np.empty(ns, np.float32)
        m4 = np.empty((p.shape[0], 4, 4), np.float64) if want_mat4 else None
        B.check(self._h, self._L.ngicp_batch_covariances(self._h, p.ctypes.data, p.shape[0], p.strides[0], _ptr(so, C.c_int64), ns,
                                                         _ptr(m4, C.c_double) if want_mat4 else None, _ptr(cov6, C.c_float), _ptr(dens, C.c_float)))
        if want_mat4:
            return cov6, m4.transpose(0, 2, 1).copy(), dens
        return cov6, dens

    def setInputSourceBatch(self, points, seg_offsets):
        """Many scans stored back to back become the source (one segment per scan) — batched registration units."""
        p = _pts(points)
        so = np.ascontiguousarray(seg_offsets, np.int64)
        B.check(self._h, self._L.ngicp_set_input_batch(self._h, B.SOURCE, p.ctypes.data, p.shape[0], p.strides[0], _ptr(so, C.c_int64), len(so) - 1))
        self.source_kdtree_ = KdTreeFLANN._adopt(self, self._L.ngicp_get_index(self._h, B.SOURCE), p)
        self._input = points
        self._n_scans = len(so) - 1

    def batchLinearize(self, Ts):
        """linearize for every scan of the batched source at its own pose: (errors (S,), H (S,6,6), b (S,6), ncorr (S,))."""
        Ts = np.asarray(Ts, np.float64)
        S_ = Ts.shape[0]
        t = np.ascontiguousarray(Ts.transpose(0, 2, 1)).reshape(S_, 16)
        H = np.zeros((S_, 6, 6)); b = np.zeros((S_, 6)); e = np.zeros(S_); nc = np.zeros(S_, np.int32)
        B.check(self._h, self._L.ngicp_batch_linearize(self._h, S_, _ptr(t, C.c_double), _ptr(H, C.c_double), _ptr(b, C.c_double), _ptr(e, C.c_double),
                                                       _ptr(nc, C.c_int)))
        return e, H, b, nc

    def captureKeyframe(self) -> Keyframe:
        """Snapshot the current source cloud + covariances as a device-resident keyframe."""
        kf = C.c_void_p(None)
        B.check(self._h, self._L.ngicp_keyframe_capture(self._h, C.byref(kf)))
        return Keyframe(self, kf.value)

    def assembleSubmap(self, keyframes):
        """buildSubmap (odom.cc:1719-1738) on the device: concatenated keyframes become the target cloud, its index and
        its covariance list."""
        arr = (C.c_void_p * len(keyframes))(*[k._kf.value for k in keyframes])
        B.check(self._h, self._L.ngicp_submap_assemble(self._h, arr, len(keyframes)))
        self.target_kdtree_ = KdTreeFLANN._adopt(self, self._L.ngicp_get_index(self._h, B.TARGET), None)
        self._target = object()

    def enableTiming(self, on=True):
        B.check(self._h, self._L.ngicp_enable_timing(self._h, int(on)))

    def timings(self, reset=True) -> dict:
        t = B.Timings()
        B.check(self._h, self._L.ngicp_get_timings(self._h, C.byref(t), int(reset)))
        return {f: getattr(t, f) for f, _ in B.Timings._fields_}

    def stream_ptr(self) -> int:
        return int(self._L.ngicp_stream(self._h) or 0)

    def synchronize(self):
        B.check(self._h, self._L.ngicp_synchronize(self._h))


# ---- scan pre-filters on the device: PCL's class names as DLIO uses them (odom.cc:114-118, :500-502, :575-584) -------------
def _filter_scan(owner: "NanoGICP", cloud, crop, leaf, set_as: int):
    p = _pts(cloud)
    fp = C.POINTER(C.c_float)
    if crop is not None:
        mn = np.ascontiguousarray(crop[0], np.float32); mx = np.ascontiguousarray(crop[1], np.float32)
        a_mn, a_mx, neg = mn.ctypes.data_as(fp), mx.ctypes.data_as(fp), int(bool(crop[2]))
    else:
        a_mn = a_mx = None; neg = 0
    lf = None if leaf is None else np.ascontiguousarray(leaf, np.float32)
    out = np.empty((p.shape[0], 3), np.float32)
    n_out = C.c_size_t(0)
    B.check(owner._h, owner._L.ngicp_filter_scan(owner._h, p.ctypes.data, p.shape[0], p.strides[0], a_mn, a_mx, neg,
                                                 None if lf is None else lf.ctypes.data_as(fp), set_as, out.ctypes.data_as(fp), C.byref(n_out)))
    return out[:n_out.value].copy()


class CropBox:
    """pcl::CropBox on the device (ngicp_filter_scan): setMin / setMax / setNegative / setInputCloud / filter."""

    def __init__(self, owner: "NanoGICP"):
        self._o = owner
        self._min = np.full(3, -1.0, np.float32); self._max = np.full(3, 1.0, np.float32); self._neg = False; self._cloud = None

    def setMin(self, v): self._min = np.asarray(v, np.float32)[:3].copy()
    def setMax(self, v): self._max = np.asarray(v, np.float32)[:3].copy()
    def setNegative(self, b): self._neg = bool(b)
    def setInputCloud(self, cloud): self._cloud = cloud

    def filter(self):
        return _filter_scan(self._o, self._cloud, (self._min, self._max, self._neg), None, -1)


class VoxelGrid:
    """pcl::VoxelGrid on the device (ngicp_filter_scan): setLeafSize / setInputCloud / filter."""

    def __init__(self, owner: "NanoGICP"):
        self._o = owner
        self._leaf = np.full(3, 0.05, np.float32); self._cloud = None

    def setLeafSize(self, lx, ly=None, lz=None):
        self._leaf = np.array([lx, lx if ly is None else ly, lx if lz is None else lz], np.float32)

    def setInputCloud(self, cloud): self._cloud = cloud

    def filter(self):
        return _filter_scan(self._o, self._cloud, None, self._leaf, -1)
