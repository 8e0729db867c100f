"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/ngicp_b200.h
declares, and refuses loudly to run without a B200 (no CPU fallback)."""
import ctypes
import re
from pathlib import Path

import pytest

import ngicp
from ngicp import binding as B

ROOT = Path(__file__).resolve().parent.parent


def header_functions():
    text = (ROOT / "include" / "ngicp_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ngicp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = ngicp.lib()
    declared = header_functions()
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(B.SYMBOLS) == declared


def test_default_params_are_the_reference_constructor_defaults():
    p = B.Params()
    ngicp.lib().ngicp_default_params(ctypes.byref(p))
    assert p.k_correspondences == 20 and p.regularization == ngicp.REG_PLANE          # nano_gicp.cc:60,64
    assert p.max_corr_dist == pytest.approx(3.4028234663852886e38)                      # FLT_MAX, nano_gicp.cc:62
    assert (p.max_iterations, p.lm_max_iterations) == (64, 10)                           # lsq_registration.cc:55,62
    assert (p.rotation_epsilon, p.transformation_epsilon, p.lm_init_lambda_factor) == (2e-3, 5e-4, 1e-9)


def test_no_cpu_fallback(have_gpu):
    if have_gpu:
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p(None)
    rc = ngicp.lib().ngicp_create(0, ctypes.byref(h))
    assert rc == B.ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in ngicp.lib().ngicp_last_error(None)
    with pytest.raises(ngicp.NgicpError):
        ngicp.NanoGICP(0)


def test_product_never_imports_the_oracle():
    pat = re.compile(r"import\s+oracle|from\s+oracle|liboracle|oracle/|orc_[a-z]+\(|#include\s*[<\"].*oracle")
    for f in (ROOT / "noetic-slam_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h", ".cc") and f.is_file():
            assert not pat.search(f.read_text()), f


def test_header_is_plain_c_and_struct_layouts_match_the_python_mirror(tmp_path):
    """include/ngicp_b200.h must compile as C (the boundary is a C ABI, not a C++ one), and the ctypes mirrors of its
    structs (ngicp/binding.py) must have the C compiler's sizes and field offsets."""
    import ctypes as C
    import shutil
    import subprocess
    from ngicp import binding as B
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    structs = {"ngicp_params": B.Params, "ngicp_timings": B.Timings, "ngicp_odom_params": B.OdomParamsC, "ngicp_odom_result": B.OdomResultC}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ngicp_b200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf(" %zu", offsetof({cname}, {fname}));')
        lines.append('  printf("\\n");')
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True).stdout.strip().splitlines()
    for line in out:
        name, size, *offs = line.split()
        cls = structs[name]
        assert int(size) == C.sizeof(cls), name
        assert [int(o) for o in offs] == [getattr(cls, f).offset for f, _ in cls._fields_], name


def test_odom_loop_defaults_are_the_reference_yaml_values_and_null_arguments_are_refused():
    """ngicp_odom_default_params = reference src/dlio/cfg/params.yaml:43-59 + cfg/dlio.yaml:17; the entry points refuse NULL
    handles with NGICP_ERR_INVALID instead of touching the device (no GPU needed)."""
    import ctypes as C
    L = B.lib()
    p = B.OdomParamsC()
    L.ngicp_odom_default_params(C.byref(p))
    assert (p.crop_size, p.voxel_res, p.keyframe_thresh_dist, p.keyframe_thresh_rot) == (1.0, 0.25, 1.0, 45.0)
    assert (p.submap_knn, p.submap_kcv, p.submap_kcc, p.gicp_min_num_points, p.adaptive) == (10, 10, 10, 64, 1)
    assert abs(p.gicp_max_corr_dist - 0.5) < 1e-7 and (p.time_offset_bytes, p.time_type) == (20, 0)     # dlio::Point `t`
    out = C.c_void_p(None)
    assert L.ngicp_odom_create(None, C.byref(p), C.byref(out)) == B.ERR_INVALID and not out.value
    n = C.c_size_t(0)
    assert L.ngicp_odom_scan_begin(None, None, 0, 32, None, C.byref(n), None) == B.ERR_INVALID
    res = B.OdomResultC()
    assert L.ngicp_odom_scan_finish(None, None, 0, C.byref(res), None, 0) == B.ERR_INVALID
    assert L.ngicp_odom_destroy(None) == B.OK


def test_a_plain_c_program_links_and_calls_the_library(tmp_path):
    """The boundary from the consumer's side: a C99 program that includes the header, links libngicp_b200.so and calls the
    entry points that need no device (version, defaults, the host-side hull) — then asks for a handle and, without a B200,
    must get NGICP_ERR_NO_DEVICE and an error text instead of a CPU fallback."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "client.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "ngicp_b200.h"
int main(void) {
  ngicp_params p;
  ngicp_odom_params op;
  ngicp_default_params(&p);
  ngicp_odom_default_params(&op);
  const double sq[] = {0, 0, 0, 4, 0, 0, 4, 4, 0, 0, 4, 0, 2, 2, 0, 1, 3, 0};
  int idx[6];
  const int m = ngicp_hull_planar(sq, 6, 0, 0.0, idx);
  ngicp_handle* h = NULL;
  const int rc = ngicp_create(0, &h);
  printf("%s k=%d iters=%d knn=%d hull=%d:%d%d%d%d create=%d err=%d\n", ngicp_version(), p.k_correspondences, p.max_iterations, op.submap_knn, m,
         idx[0], idx[1], idx[2], idx[3], rc, (int)(strlen(ngicp_last_error(NULL)) > 0));
  if (rc == NGICP_OK) ngicp_destroy(h);
  return 0;
}
''')
    exe = tmp_path / "client"
    libdir = ROOT / "noetic-slam_b200"
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(src), "-o", str(exe), f"-L{libdir}", "-lngicp_b200",
                        f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    line = out.stdout.strip()
    assert " k=20 iters=64 knn=10 hull=4:0123 " in line, line
    import torch
    if not torch.cuda.is_available():
        assert f"create={B.ERR_NO_DEVICE} err=1" in line, line
