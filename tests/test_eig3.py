"""The PLANE regularisation without an eigen-decomposition (csrc/eig3.cuh) on the host: the header compiles as plain C++, so
its accuracy, its refusal rule and its fallback rate are checked here against numpy.linalg.eigh on synthetic k = 16
neighbourhoods — planes, single scan rings (nearly collinear), two rings, blobs, exact planes, identical points."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("eig3") / "libeig3_host.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", str(ROOT / "noetic-slam_b200" / "csrc"),
                    str(ROOT / "tests" / "shims" / "eig3_host.cc"), "-o", str(out)], check=True)
    L = ctypes.CDLL(str(out))
    L.plane_fast.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    L.plane_fast.restype = ctypes.c_int
    return L


def covs(P):
    """(S - s s^T / k) / k about the first point, from fp32 coordinates (reference src/dlio/src/nano_gicp/nano_gicp.cc:348-354)."""
    P = P.astype(np.float32).astype(np.float64)
    d = P - P[:, :1]
    s = d.sum(1)
    S = np.einsum("nki,nkj->nij", d, d)
    return (S - np.einsum("ni,nj->nij", s, s) / P.shape[1]) / P.shape[1]


def run(lib, C):
    a = np.ascontiguousarray(np.stack([C[:, 0, 0], C[:, 0, 1], C[:, 0, 2], C[:, 1, 1], C[:, 1, 2], C[:, 2, 2]], 1))
    o = np.empty_like(a)
    refused = lib.plane_fast(a.ctypes.data, o.ctypes.data, len(a))
    w, V = np.linalg.eigh(C)
    n = V[:, :, 0]
    ref = np.eye(3)[None] - (1 - 1e-3) * np.einsum("ni,nj->nij", n, n)
    r = np.stack([ref[:, 0, 0], ref[:, 0, 1], ref[:, 0, 2], ref[:, 1, 1], ref[:, 1, 2], ref[:, 2, 2]], 1)
    ok = ~np.isnan(o[:, 0])
    gap = (w[:, 1] - w[:, 0]) / np.maximum(w.sum(1), 1e-300)
    return refused, ok, np.abs(o - r).max(1), gap


def scenes(n, rng):
    c = rng.uniform(-60, 60, (n, 1, 3))
    R = np.linalg.qr(rng.normal(size=(n, 3, 3)))[0]
    rot = lambda X: c + np.einsum("nij,nkj->nki", R, X)
    uv = rng.uniform(-1, 1, (n, 16, 2)) * rng.uniform(0.05, 2, (n, 1, 1))
    z = rng.normal(0, 0.02, (n, 16, 1))
    u = np.sort(rng.uniform(-1, 1, (n, 16, 1)), 1) * rng.uniform(0.1, 5, (n, 1, 1))
    v = rng.integers(0, 2, (n, 16, 1)) * rng.uniform(0.05, 3, (n, 1, 1))
    return {"planes": rot(np.concatenate([uv, z], 2)),
            "single ring": rot(np.concatenate([u, rng.normal(0, 0.02, (n, 16, 2))], 2)),
            "two rings": rot(np.concatenate([u, v, rng.normal(0, 0.02, (n, 16, 1))], 2)),
            "blobs": c + rng.normal(0, 1, (n, 16, 3)) * rng.uniform(0.01, 1, (n, 1, 1)),
            "exact planes": rot(np.concatenate([uv, 0 * z], 2))}


def test_fast_path_stays_below_the_conditioning_floor(lib):
    rng = np.random.default_rng(0)
    for name, P in scenes(40000, rng).items():
        refused, ok, err, gap = run(lib, covs(P))
        assert refused <= 0.002 * len(P), (name, refused)                 # the exact (Jacobi) path is the rare one
        # eigh itself resolves the normal only to ~eps / gap^2 of the trace-scaled matrix: compare above that floor
        floor = 1e-15 / np.maximum(gap, 1e-12) ** 2
        assert (err[ok] <= np.maximum(1e-9, 100 * floor[ok])).all(), (name, float(err[ok].max()))
        wide = ok & (gap > 1e-3)
        assert wide.any() and err[wide].max() < 1e-9, (name, float(err[wide].max()))
        assert not (~ok & (gap > 1e-4)).any(), name                       # never refused with a clear spectral gap


def test_degenerate_neighbourhoods_are_refused_not_guessed(lib):
    rng = np.random.default_rng(1)
    c = rng.uniform(-60, 60, (1000, 1, 3))
    refused, ok, _, _ = run(lib, covs(np.repeat(c, 16, 1)))              # sixteen identical points: the zero matrix
    assert refused == 1000 and not ok.any()
    line = c + np.linspace(-1, 1, 16)[None, :, None] * np.array([1.0, 2.0, -0.5])[None, None]
    refused, ok, _, gap = run(lib, covs(line))                            # exactly collinear: the two small eigenvalues coincide
    assert refused == 1000
    bad = np.full((4, 3, 3), np.nan)
    assert run_raw(lib, bad) == 4                                          # non-finite input never takes the fast path


def run_raw(lib, C):
    a = np.ascontiguousarray(np.stack([C[:, 0, 0], C[:, 0, 1], C[:, 0, 2], C[:, 1, 1], C[:, 1, 2], C[:, 2, 2]], 1))
    o = np.empty_like(a)
    return lib.plane_fast(a.ctypes.data, o.ctypes.data, len(a))
