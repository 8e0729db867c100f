"""TEST INFRASTRUCTURE — numpy statement of the voxel-key spec (DESIGN.md "K1 keys").

The reference has no voxel keys (it builds a KD-tree, nanoflann.h:1025-1185); the keys are this
build's own integer index, so the parity anchor is this spec, which the CUDA path must reproduce
bit for bit:
    origin_a = min over the cloud of coordinate a                      (fp32)
    extent   = max_a fl32(max_a - origin_a), at least 2^-10
    h0       = 2^(e-12) where frexp(fl32(extent * 1.01f)) = (m, e)     (so 4096*h0 > extent)
    u_a      = fl32(p_a - origin_a);  c_a = clamp(floor(fl32(u_a * (1/h0))), 0, 4095)
    key      = (segment << 36) | morton3(c_x, c_y, c_z)                (x = lowest interleaved bit)
"""
import numpy as np


def grid_params(pts):
    p = np.asarray(pts, np.float32)[:, :3]
    lo = p.min(0)
    ext = np.float32(np.max(p.max(0) - lo))
    e = ext if ext > np.float32(9.765625e-4) else np.float32(9.765625e-4)
    _, ex = np.frexp(np.float32(e * np.float32(1.01)))
    h0 = np.float32(2.0) ** np.float32(int(ex) - 12)
    return lo, np.float32(h0)


def _expand3(v):
    x = v.astype(np.uint64) & np.uint64(0x1FFFFF)
    x = (x | (x << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    x = (x | (x << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    x = (x | (x << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    return x


def voxel_keys(pts, lo=None, h0=None, seg=0):
    p = np.asarray(pts, np.float32)[:, :3]
    if lo is None or h0 is None:
        lo, h0 = grid_params(p)
    inv = np.float32(1.0) / np.float32(h0)
    u = (p - lo.astype(np.float32)).astype(np.float32)
    c = np.floor((u * inv).astype(np.float32))
    c = np.clip(c, 0, 4095).astype(np.uint32)
    m = _expand3(c[:, 0]) | (_expand3(c[:, 1]) << np.uint64(1)) | (_expand3(c[:, 2]) << np.uint64(2))
    return (np.uint64(seg) << np.uint64(36)) | m
