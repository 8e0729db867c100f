"""ctypes binding of libngicp_b200.so — the C ABI declared in include/ngicp_b200.h.

The library is the product: hand-written CUDA for sm_100a behind a plain C interface. There is no
CPU fallback here or in the library; a missing .so or a missing B200 is an error, loudly.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

PKG_ROOT = Path(__file__).resolve().parent.parent          # noetic-slam_b200/
LIB_PATH = Path(os.environ["NGICP_LIB"]) if os.environ.get("NGICP_LIB") else PKG_ROOT / "libngicp_b200.so"   # NGICP_LIB: development builds (tools/)
CSRC = PKG_ROOT / "csrc"

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_UNSUPPORTED, ERR_LM_NOT_CONVERGED = range(6)
REG_NONE, REG_MIN_EIG, REG_NORMALIZED_MIN_EIG, REG_PLANE, REG_FROBENIUS = range(5)
SOURCE, TARGET = 0, 1

# every symbol include/ngicp_b200.h declares (tests check the .so exports exactly these)
SYMBOLS = [
    "ngicp_default_params", "ngicp_version", "ngicp_last_error", "ngicp_create", "ngicp_destroy", "ngicp_set_params",
    "ngicp_get_params", "ngicp_set_async_input", "ngicp_stream", "ngicp_synchronize", "ngicp_index_build", "ngicp_index_build_device",
    "ngicp_index_retain", "ngicp_index_release", "ngicp_index_size", "ngicp_knn", "ngicp_self_neighbours", "ngicp_index_keys", "ngicp_set_input", "ngicp_set_input_device",
    "ngicp_attach_index", "ngicp_get_index", "ngicp_swap_source_and_target", "ngicp_clear", "ngicp_compute_covariances",
    "ngicp_get_covariances", "ngicp_set_covariances", "ngicp_has_covariances", "ngicp_update_correspondences",
    "ngicp_linearize", "ngicp_compute_error", "ngicp_align", "ngicp_transform_source", "ngicp_batch_covariances", "ngicp_set_input_batch", "ngicp_batch_linearize",
    "ngicp_keyframe_capture", "ngicp_keyframe_transform", "ngicp_keyframe_size", "ngicp_keyframe_release", "ngicp_keyframe_download",
    "ngicp_submap_assemble", "ngicp_filter_scan", "ngicp_scan_ingest", "ngicp_scan_deskew",
    "ngicp_odom_default_params", "ngicp_odom_create", "ngicp_odom_destroy", "ngicp_odom_last_error", "ngicp_odom_set_hull_callbacks",
    "ngicp_odom_set_pose", "ngicp_odom_get_profile", "ngicp_odom_scan_begin", "ngicp_odom_scan_finish", "ngicp_hull_planar", "ngicp_hull_spatial",
    "ngicp_enable_timing", "ngicp_get_timings",
]


class Params(C.Structure):
    _fields_ = [("k_correspondences", C.c_int), ("max_corr_dist", C.c_double), ("regularization", C.c_int),
                ("max_iterations", C.c_int), ("rotation_epsilon", C.c_double), ("transformation_epsilon", C.c_double),
                ("lm_init_lambda_factor", C.c_double), ("lm_max_iterations", C.c_int), ("use_gauss_newton", C.c_int)]


class Timings(C.Structure):
    _fields_ = [("index_ms", C.c_float), ("knn_ms", C.c_float), ("covariance_ms", C.c_float), ("linearize_ms", C.c_float),
                ("error_ms", C.c_float), ("linearize_calls", C.c_int), ("error_calls", C.c_int), ("kernel_launches", C.c_int),
                ("correspond_ms", C.c_float)]


class OdomParamsC(C.Structure):
    _fields_ = [("crop_size", C.c_float), ("voxel_res", C.c_float), ("keyframe_thresh_dist", C.c_float), ("keyframe_thresh_rot", C.c_float),
                ("submap_knn", C.c_int), ("submap_kcv", C.c_int), ("submap_kcc", C.c_int), ("gicp_min_num_points", C.c_int),
                ("gicp_max_corr_dist", C.c_float), ("adaptive", C.c_int), ("time_offset_bytes", C.c_int), ("time_type", C.c_int)]


class OdomResultC(C.Structure):
    _fields_ = [("valid", C.c_int), ("T", C.c_float * 16), ("T_corr", C.c_float * 16), ("converged", C.c_int), ("iterations", C.c_int),
                ("n_points", C.c_int), ("new_keyframe", C.c_int), ("submap_changed", C.c_int), ("n_keyframes", C.c_int), ("n_submap", C.c_int)]


HULL_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_int, C.c_double, C.POINTER(C.c_int), C.c_void_p)


class NgicpError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ngicp error {code}: {msg}")
        self.code = code


def build(verbose: bool = False) -> Path:
    """Compile libngicp_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", str(CSRC), "-j8"], capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libngicp_b200.so failed:\n" + (r.stdout or "") + (r.stderr or ""))
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise FileNotFoundError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                "(there is no CPU fallback for the CUDA path)")
    L = C.CDLL(str(LIB_PATH))
    vp, sz, i = C.c_void_p, C.c_size_t, C.c_int
    fp, dp, ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int)
    L.ngicp_default_params.argtypes = [C.POINTER(Params)]
    L.ngicp_default_params.restype = None
    L.ngicp_version.restype = C.c_char_p
    L.ngicp_last_error.restype = C.c_char_p
    L.ngicp_last_error.argtypes = [vp]
    L.ngicp_create.argtypes = [i, C.POINTER(vp)]
    L.ngicp_destroy.argtypes = [vp]
    L.ngicp_set_params.argtypes = [vp, C.POINTER(Params)]
    L.ngicp_get_params.argtypes = [vp, C.POINTER(Params)]
    L.ngicp_set_async_input.argtypes = [vp, i]
    L.ngicp_stream.restype = vp
    L.ngicp_stream.argtypes = [vp]
    L.ngicp_synchronize.argtypes = [vp]
    L.ngicp_index_build.argtypes = [vp, vp, sz, sz, C.POINTER(vp)]
    L.ngicp_index_build_device.argtypes = [vp, vp, sz, C.POINTER(vp)]
    L.ngicp_index_retain.argtypes = [vp]
    L.ngicp_index_release.argtypes = [vp]
    L.ngicp_index_size.restype = sz
    L.ngicp_index_size.argtypes = [vp]
    L.ngicp_knn.argtypes = [vp, vp, vp, sz, sz, i, ip, fp]
    L.ngicp_self_neighbours.argtypes = [vp, i, i, ip, dp]
    L.ngicp_index_keys.argtypes = [vp, vp, C.POINTER(C.c_uint64), fp]
    L.ngicp_set_input.argtypes = [vp, i, vp, sz, sz]
    L.ngicp_set_input_device.argtypes = [vp, i, vp, sz]
    L.ngicp_attach_index.argtypes = [vp, i, vp]
    L.ngicp_get_index.restype = vp
    L.ngicp_get_index.argtypes = [vp, i]
    L.ngicp_swap_source_and_target.argtypes = [vp]
    L.ngicp_clear.argtypes = [vp, i]
    L.ngicp_compute_covariances.argtypes = [vp, i, fp]
    L.ngicp_get_covariances.argtypes = [vp, i, dp, sz]
    L.ngicp_set_covariances.argtypes = [vp, i, dp, sz]
    L.ngicp_has_covariances.argtypes = [vp, i, C.POINTER(sz)]
    L.ngicp_update_correspondences.argtypes = [vp, dp, ip, fp, dp, ip]
    L.ngicp_linearize.argtypes = [vp, dp, dp, dp, dp, ip]
    L.ngicp_compute_error.argtypes = [vp, dp, dp]
    L.ngicp_align.argtypes = [vp, fp, fp, ip, ip, dp, dp]
    L.ngicp_transform_source.argtypes = [vp, fp, vp, sz, sz]
    L.ngicp_batch_covariances.argtypes = [vp, vp, sz, sz, C.POINTER(C.c_int64), i, dp, fp, fp]
    L.ngicp_set_input_batch.argtypes = [vp, i, vp, sz, sz, C.POINTER(C.c_int64), i]
    L.ngicp_batch_linearize.argtypes = [vp, i, dp, dp, dp, dp, ip]
    L.ngicp_keyframe_capture.argtypes = [vp, C.POINTER(vp)]
    L.ngicp_keyframe_transform.argtypes = [vp, vp, fp]
    L.ngicp_keyframe_size.restype = sz
    L.ngicp_keyframe_size.argtypes = [vp]
    L.ngicp_keyframe_release.argtypes = [vp, vp]
    L.ngicp_keyframe_download.argtypes = [vp, vp, fp, dp]
    L.ngicp_submap_assemble.argtypes = [vp, C.POINTER(vp), i]
    L.ngicp_filter_scan.argtypes = [vp, vp, sz, sz, fp, fp, i, fp, i, fp, C.POINTER(sz)]
    L.ngicp_scan_ingest.argtypes = [vp, vp, sz, sz, sz, i, fp, fp, i, dp, C.POINTER(sz), C.POINTER(sz)]
    L.ngicp_scan_deskew.argtypes = [vp, fp, sz, fp, i, fp, C.POINTER(sz)]
    L.ngicp_odom_default_params.argtypes = [C.POINTER(OdomParamsC)]
    L.ngicp_odom_default_params.restype = None
    L.ngicp_odom_create.argtypes = [vp, C.POINTER(OdomParamsC), C.POINTER(vp)]
    L.ngicp_odom_destroy.argtypes = [vp]
    L.ngicp_odom_last_error.restype = C.c_char_p
    L.ngicp_odom_last_error.argtypes = [vp]
    L.ngicp_odom_set_hull_callbacks.argtypes = [vp, HULL_FN, HULL_FN, vp]
    L.ngicp_odom_set_pose.argtypes = [vp, fp]
    L.ngicp_odom_scan_begin.argtypes = [vp, vp, sz, sz, dp, C.POINTER(sz), C.POINTER(sz)]
    L.ngicp_odom_scan_finish.argtypes = [vp, fp, sz, C.POINTER(OdomResultC), ip, i]
    L.ngicp_odom_get_profile.argtypes = [vp, dp, C.POINTER(C.c_long), i]
    L.ngicp_hull_planar.argtypes = [dp, i, i, C.c_double, ip]
    L.ngicp_hull_spatial.argtypes = [dp, i, i, C.c_double, ip]
    L.ngicp_enable_timing.argtypes = [vp, i]
    L.ngicp_get_timings.argtypes = [vp, C.POINTER(Timings), i]
    _lib = L
    return L


def check(handle, rc: int, allow=()):
    if rc != OK and rc not in allow:
        msg = lib().ngicp_last_error(handle)
        raise NgicpError(rc, msg.decode() if msg else "")
    return rc
