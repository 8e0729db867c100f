"""The drop-in C++ surface (include/nano_gicp/*.h) compiles against mini PCL/Eigen shims and, on the GPU box,
gives the same pose as the Python mirror when driven the way dlio::OdomNode drives the reference."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import ngicp
import scenarios as S

ROOT = Path(__file__).resolve().parent.parent
CXX = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"


def build_dropin(tmp_path, pcl_minor=10, source="dropin_main.cc"):
    """pcl_minor 10: the PCL <= 1.10 flavour of the shims (cloud pointers are a distinct boost::shared_ptr template, as in
    DLIO's Ubuntu 20.04 image); 12: PCL >= 1.11 (std::shared_ptr). A std/boost mix-up in the headers fails to compile."""
    exe = tmp_path / f"{Path(source).stem}_{pcl_minor}"
    lib = ROOT / "noetic-slam_b200" / "libngicp_b200.so"
    cmd = [CXX, "-std=c++17", "-O1", "-Wall", "-Werror", f"-DSHIM_PCL_MINOR={pcl_minor}", f"-I{ROOT / 'tests' / 'shims'}", f"-I{ROOT / 'include'}",
           str(ROOT / "tests" / "shims" / source), "-o", str(exe), str(lib), f"-Wl,-rpath,{lib.parent}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


@pytest.mark.parametrize("pcl_minor", [10, 12])
def test_dropin_headers_compile_and_link(tmp_path, pcl_minor):
    assert build_dropin(tmp_path, pcl_minor).exists()
    assert build_dropin(tmp_path, pcl_minor, "dropin_bench.cc").exists()


@pytest.mark.gpu
def test_dropin_matches_python_mirror(tmp_path):
    exe = build_dropin(tmp_path)
    a, b, _ = S.scan_pair(2, w=128)
    (tmp_path / "src.bin").write_bytes(np.ascontiguousarray(b, np.float32).tobytes())
    (tmp_path / "tgt.bin").write_bytes(np.ascontiguousarray(a, np.float32).tobytes())
    r = subprocess.run([str(exe), str(tmp_path / "src.bin"), str(tmp_path / "tgt.bin")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    lines = r.stdout.strip().splitlines()
    tok = lines[0].split()
    T_cpp = np.array([float(x) for x in tok[1:17]]).reshape(4, 4)
    g = S.configure(ngicp.NanoGICP(0))
    g.setInputSource(b); g.setInputTarget(a)
    g.calculateSourceCovariances(); g.calculateTargetCovariances()
    T = g.align()
    assert np.abs(T_cpp - T).max() < 1e-6
    assert int(tok[tok.index("converged") + 1]) == int(g.hasConverged())
    assert abs(float(tok[tok.index("density") + 1]) - g.source_density_) < 1e-6 * g.source_density_
    assert int(tok[tok.index("ncov") + 1]) == len(b)
    al = [float(x) for x in tok[tok.index("aligned0") + 1: tok.index("aligned0") + 4]]
    want = (T[:3, :3].astype(np.float64) @ b[0].astype(np.float64) + T[:3, 3]).astype(np.float32)
    assert np.abs(np.array(al) - want).max() < 1e-4
    k = lines[1].split()
    idx, sqd = g.target_kdtree_.nearestKSearch(b[:1], 3)
    assert [int(x) for x in k[2:5]] == idx[0].tolist() and int(k[1]) == 3
    f = lines[2].split()
    filt = g.setInputSourceFiltered(b, crop=([-1.0] * 3, [1.0] * 3, True), leaf=(0.25, 0.25, 0.25))
    g.calculateSourceCovariances()
    T2 = g.align()
    assert int(f[1]) == len(filt) and (np.array([float(x) for x in f[3:6]], np.float32) == filt[0]).all()
    T2_cpp = np.array([float(x) for x in f[7:23]]).reshape(4, 4)
    assert np.abs(T2_cpp - T2).max() < 1e-6
    # the tree handed from one NanoGICP object to another right after its build was launched (stream ordering)
    hv = lines[3].split()
    assert hv[0] == "handover"
    T3_cpp = np.array([float(x) for x in hv[2:18]]).reshape(4, 4)
    assert np.abs(T3_cpp - T).max() < 1e-6
