"""TEST INFRASTRUCTURE — ctypes front-end of the CPU oracle (oracle/oracle.cc).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package, and only as the checker / the reported CPU baseline. The product
(noetic-slam_b200/) never imports it.

Two shared libraries come out of oracle/oracle.cc (see its header and oracle/Makefile):
  variant "port" -> oracle/_build/liboracle_port.so  self-contained restatement (own exact k-NN)
  variant "ref"  -> oracle/_ref/liboracle_ref.so     same restatement over the REFERENCE's own
                    nanoflann.h, compiled in place from /root/reference (k-NN is reference code)
Method names mirror nano_gicp::NanoGICP (reference include/nano_gicp/nano_gicp.h:84-137).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIBS = {"port": _HERE / "_build" / "liboracle_port.so", "ref": _HERE / "_ref" / "liboracle_ref.so"}
_loaded: dict[str, C.CDLL] = {}

REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS = range(5)


def build(ref: bool | None = None) -> None:
    """Compile the oracle. `ref=None` builds the reference variant only when the tree is present."""
    subprocess.run(["make", "-C", str(_HERE), "-s"], check=True)
    have_ref = Path("/root/reference/src/dlio/include/nano_gicp/nanoflann.h").exists()
    if ref or (ref is None and have_ref):
        subprocess.run(["make", "-C", str(_HERE), "-s", "ref"], check=True)


def available(variant: str) -> bool:
    return _LIBS[variant].exists()


def _f32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _f64p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i32p(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def lib(variant: str = "port") -> C.CDLL:
    if variant in _loaded:
        return _loaded[variant]
    path = _LIBS[variant]
    if not path.exists():
        if variant == "port":
            build(ref=False)
        else:
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
    L = C.CDLL(str(path))
    vp, sz, i, d = C.c_void_p, C.c_size_t, C.c_int, C.c_double
    fp, dp, ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.orc_tree_kind.restype = C.c_char_p
    L.orc_max_threads.restype = i
    L.orc_tree_build.restype = vp
    L.orc_tree_build.argtypes = [fp, sz, sz]
    L.orc_tree_free.argtypes = [vp]
    L.orc_knn.restype = i
    L.orc_knn.argtypes = [vp, fp, sz, sz, i, ip, fp, i, i]
    L.orc_gicp_create.restype = vp
    L.orc_gicp_destroy.argtypes = [vp]
    L.orc_gicp_set_params.argtypes = [vp, i, i, d, i, i, d, d, d, i, i]
    L.orc_gicp_set_cloud.argtypes = [vp, i, fp, sz, sz]
    L.orc_gicp_calc_covs.restype = i
    L.orc_gicp_calc_covs.argtypes = [vp, i, fp]
    L.orc_gicp_get_covs.restype = sz
    L.orc_gicp_get_covs.argtypes = [vp, i, dp]
    L.orc_gicp_set_covs.argtypes = [vp, i, dp, sz]
    L.orc_gicp_update_correspondences.argtypes = [vp, dp, ip, fp, dp]
    L.orc_crop_box.restype = sz
    L.orc_crop_box.argtypes = [fp, sz, sz, fp, fp, i, fp]
    L.orc_scan_ingest.restype = sz
    L.orc_scan_ingest.argtypes = [vp, sz, sz, sz, i, fp, fp, i, fp, ip, dp, C.POINTER(sz)]
    L.orc_scan_deskew.argtypes = [fp, ip, sz, fp, sz, fp]
    L.orc_voxel_grid.restype = sz
    L.orc_voxel_grid.argtypes = [fp, sz, sz, fp, fp, ip, ip]
    L.orc_gicp_num_correspondences.restype = i
    L.orc_gicp_num_correspondences.argtypes = [vp]
    L.orc_gicp_linearize.restype = d
    L.orc_gicp_linearize.argtypes = [vp, dp, dp, dp]
    L.orc_gicp_compute_error.restype = d
    L.orc_gicp_compute_error.argtypes = [vp, dp]
    L.orc_gicp_align.restype = i
    L.orc_gicp_align.argtypes = [vp, fp, fp, ip, ip, dp, dp]
    _loaded[variant] = L
    return L


def _as_points(p) -> np.ndarray:
    p = np.ascontiguousarray(p, dtype=np.float32)
    assert p.ndim == 2 and p.shape[1] >= 3, p.shape
    return p


def _colmajor(T) -> np.ndarray:
    """4x4 (row-major numpy) -> 16 doubles column-major, what the C ABI takes."""
    return np.ascontiguousarray(np.asarray(T, dtype=np.float64).T).reshape(16)


class KdTree:
    """nanoflann::KdTreeFLANN<PointT> stand-in (reference include/nano_gicp/nanoflann_adaptor.h:57-152)."""

    def __init__(self, points, variant: str = "port"):
        self._L = lib(variant)
        self._pts = _as_points(points)
        self._h = self._L.orc_tree_build(_f32p(self._pts), self._pts.shape[0], self._pts.shape[1])

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_tree_free(self._h)
            self._h = None

    def knn(self, queries, k: int, canonical: bool = True, num_threads: int = 0):
        q = _as_points(queries)
        idx = np.empty((q.shape[0], k), np.int32)
        sqd = np.empty((q.shape[0], k), np.float32)
        self._L.orc_knn(self._h, _f32p(q), q.shape[0], q.shape[1], k, _i32p(idx), _f32p(sqd), int(canonical), num_threads)
        return idx, sqd


class OracleGICP:
    """CPU restatement of nano_gicp::NanoGICP<PointT,PointT> (reference src/nano_gicp/nano_gicp.cc)."""

    def __init__(self, variant: str = "port", num_threads: int = 0):
        self._L = lib(variant)
        self._h = self._L.orc_gicp_create()
        self.variant = variant
        self.p = dict(num_threads=num_threads, k=20, max_corr_dist=float(np.finfo(np.float32).max), reg_method=REG_PLANE,
                      max_iterations=64, rot_eps=2e-3, trans_eps=5e-4, lm_init_lambda_factor=1e-9,
                      lm_max_iterations=10, gauss_newton=0)
        self._n = [0, 0]
        self._push()

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_gicp_destroy(self._h)
            self._h = None

    def _push(self):
        p = self.p
        self._L.orc_gicp_set_params(self._h, p["num_threads"], p["k"], p["max_corr_dist"], p["reg_method"], p["max_iterations"],
                                    p["rot_eps"], p["trans_eps"], p["lm_init_lambda_factor"], p["lm_max_iterations"], p["gauss_newton"])

    # --- setters (names follow the reference) ---
    def setNumThreads(self, n): self.p["num_threads"] = n; self._push()
    def setCorrespondenceRandomness(self, k): self.p["k"] = k; self._push()
    def setMaxCorrespondenceDistance(self, d): self.p["max_corr_dist"] = float(d); self._push()
    def setRegularizationMethod(self, m): self.p["reg_method"] = int(m); self._push()
    def setMaximumIterations(self, n): self.p["max_iterations"] = n; self._push()
    def setRotationEpsilon(self, e): self.p["rot_eps"] = e; self._push()
    def setTransformationEpsilon(self, e): self.p["trans_eps"] = e; self._push()
    def setInitialLambdaFactor(self, f): self.p["lm_init_lambda_factor"] = f; self._push()

    def _set_cloud(self, which, pts):
        pts = _as_points(pts)
        self._n[which] = pts.shape[0]
        self._L.orc_gicp_set_cloud(self._h, which, _f32p(pts), pts.shape[0], pts.shape[1])

    def setInputSource(self, pts): self._set_cloud(0, pts)
    def setInputTarget(self, pts): self._set_cloud(1, pts)

    def _calc(self, which):
        dens = C.c_float(0)
        self._L.orc_gicp_calc_covs(self._h, which, C.byref(dens))
        return float(dens.value)

    def calculateSourceCovariances(self):
        self.source_density_ = self._calc(0)
        return True

    def calculateTargetCovariances(self):
        self.target_density_ = self._calc(1)
        return True

    def _get_covs(self, which):
        n = self._L.orc_gicp_get_covs(self._h, which, None)
        out = np.zeros((n, 4, 4), np.float64)
        if n:
            self._L.orc_gicp_get_covs(self._h, which, _f64p(out))
        return out.transpose(0, 2, 1).copy()  # column-major -> numpy row-major (symmetric anyway)

    def getSourceCovariances(self): return self._get_covs(0)
    def getTargetCovariances(self): return self._get_covs(1)

    def _set_covs(self, which, covs):
        c = np.ascontiguousarray(np.asarray(covs, np.float64).transpose(0, 2, 1))
        self._L.orc_gicp_set_covs(self._h, which, _f64p(c), c.shape[0])

    def setSourceCovariances(self, covs): self._set_covs(0, covs)
    def setTargetCovariances(self, covs): self._set_covs(1, covs)

    def update_correspondences(self, T):
        n = self._n[0]
        corr = np.empty(n, np.int32)
        sqd = np.empty(n, np.float32)
        mah = np.empty((n, 4, 4), np.float64)
        t = _colmajor(T)
        self._L.orc_gicp_update_correspondences(self._h, _f64p(t), _i32p(corr), _f32p(sqd), _f64p(mah))
        self.num_correspondences = self._L.orc_gicp_num_correspondences(self._h)
        return corr, sqd, mah.transpose(0, 2, 1).copy()

    def linearize(self, T):
        H = np.zeros((6, 6), np.float64)
        b = np.zeros(6, np.float64)
        t = _colmajor(T)
        err = self._L.orc_gicp_linearize(self._h, _f64p(t), _f64p(H), _f64p(b))
        self.num_correspondences = self._L.orc_gicp_num_correspondences(self._h)
        return err, H, b

    def compute_error(self, T):
        t = _colmajor(T)
        return self._L.orc_gicp_compute_error(self._h, _f64p(t))

    def align(self, guess=None):
        g = np.eye(4, dtype=np.float32) if guess is None else np.asarray(guess, np.float32)
        gc = np.ascontiguousarray(g.T).reshape(16)
        out = np.zeros(16, np.float32)
        it, conv = C.c_int(0), C.c_int(0)
        H = np.zeros((6, 6), np.float64)
        ferr = C.c_double(0)
        rc = self._L.orc_gicp_align(self._h, _f32p(gc), _f32p(out), C.byref(it), C.byref(conv), _f64p(H), C.byref(ferr))
        self.final_transformation_ = out.reshape(4, 4).T.copy()
        self.nr_iterations_ = it.value
        self.converged_ = bool(conv.value)
        self.final_hessian_ = H
        self.final_error_ = ferr.value
        self.lm_failed_ = bool(rc)
        return self.final_transformation_

    def hasConverged(self): return self.converged_
    def getFinalTransformation(self): return self.final_transformation_
    def getFinalHessian(self): return self.final_hessian_
    def getFinalError(self): return self.final_error_


# ---- scan pre-filters: PCL's names (pcl::CropBox / pcl::VoxelGrid as DLIO configures them, odom.cc:114-118) --------------
def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != 2 or a.shape[1] < 3:
        raise ValueError(f"cloud must be (N, >=3) float32, got {a.shape}")
    return a


class CropBox:
    """pcl::CropBox restated (oracle.cc:orc_crop_box): setMin / setMax / setNegative / setInputCloud / filter."""

    def __init__(self, variant: str = "port"):
        self._L = lib(variant)
        self._min = np.full(3, -1.0, np.float32); self._max = np.full(3, 1.0, np.float32); self._neg = False; self._cloud = None

    def setMin(self, v): self._min = np.asarray(v, np.float32)[:3].copy()
    def setMax(self, v): self._max = np.asarray(v, np.float32)[:3].copy()
    def setNegative(self, b): self._neg = bool(b)
    def setInputCloud(self, cloud): self._cloud = _f32(cloud)

    def filter(self):
        c = self._cloud
        out = np.empty((len(c), 3), np.float32)
        fp = C.POINTER(C.c_float)
        m = self._L.orc_crop_box(c.ctypes.data_as(fp), len(c), c.strides[0] // 4, self._min.ctypes.data_as(fp), self._max.ctypes.data_as(fp),
                                 int(self._neg), out.ctypes.data_as(fp))
        return out[:m].copy()


class VoxelGrid:
    """pcl::VoxelGrid restated (oracle.cc:orc_voxel_grid): setLeafSize / setInputCloud / filter."""

    def __init__(self, variant: str = "port"):
        self._L = lib(variant)
        self._leaf = np.full(3, 0.05, np.float32); self._cloud = None
        self.voxel_index = None; self.voxel_count = None

    def setLeafSize(self, lx, ly=None, lz=None):
        self._leaf = np.array([lx, lx if ly is None else ly, lx if lz is None else lz], np.float32)

    def setInputCloud(self, cloud): self._cloud = _f32(cloud)

    def filter(self):
        c = self._cloud
        out = np.empty((len(c), 3), np.float32)
        vox = np.empty(len(c), np.int32); cnt = np.empty(len(c), np.int32)
        fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
        m = self._L.orc_voxel_grid(c.ctypes.data_as(fp), len(c), c.strides[0] // 4, self._leaf.ctypes.data_as(fp), out.ctypes.data_as(fp),
                                   vox.ctypes.data_as(ip), cnt.ctypes.data_as(ip))
        self.voxel_index, self.voxel_count = vox[:m].copy(), cnt[:m].copy()
        return out[:m].copy()


def scan_ingest(records: np.ndarray, time_field: str, crop=None, variant: str = "port"):
    """dlio::OdomNode::deskewPointcloud, first half (oracle.cc:orc_scan_ingest): (xyz in time order, group per point,
    unique stamps as raw field values)."""
    L = lib(variant)
    rec = np.ascontiguousarray(records)
    dt = rec.dtype.fields[time_field][0]
    ttype = {np.dtype(np.uint32): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}[dt]
    fp, ip, dp = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double)
    if crop is not None:
        mn = np.ascontiguousarray(crop[0], np.float32); mx = np.ascontiguousarray(crop[1], np.float32)
        a_mn, a_mx, neg = mn.ctypes.data_as(fp), mx.ctypes.data_as(fp), int(bool(crop[2]))
    else:
        a_mn = a_mx = None; neg = 0
    xyz = np.empty((len(rec), 3), np.float32); grp = np.empty(len(rec), np.int32); st = np.empty(len(rec), np.float64)
    nu = C.c_size_t(0)
    m = L.orc_scan_ingest(rec.ctypes.data, len(rec), rec.dtype.itemsize, rec.dtype.fields[time_field][1], ttype, a_mn, a_mx, neg,
                          xyz.ctypes.data_as(fp), grp.ctypes.data_as(ip), st.ctypes.data_as(dp), C.byref(nu))
    return xyz[:m].copy(), grp[:m].copy(), st[:nu.value].copy()


def scan_deskew(xyz, group, frames, variant: str = "port"):
    """Second half (oracle.cc:orc_scan_deskew): frames = (n_unique, 4, 4) or one (4, 4) float32 matrix."""
    L = lib(variant)
    F = np.asarray(frames, np.float32)
    if F.ndim == 2:
        F = F[None]
    Fc = np.ascontiguousarray(F.transpose(0, 2, 1)).reshape(-1, 16)
    xyz = np.ascontiguousarray(xyz, np.float32); group = np.ascontiguousarray(group, np.int32)
    out = np.empty_like(xyz)
    fp, ip = C.POINTER(C.c_float), C.POINTER(C.c_int)
    L.orc_scan_deskew(xyz.ctypes.data_as(fp), group.ctypes.data_as(ip), len(xyz), Fc.ctypes.data_as(fp), len(Fc), out.ctypes.data_as(fp))
    return out
