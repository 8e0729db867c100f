import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "noetic-slam_b200"), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def have_gpu():
    try:
        import ctypes
        import ngicp
        h = ctypes.c_void_p(None)
        rc = ngicp.lib().ngicp_create(0, ctypes.byref(h))
        if rc == 0:
            ngicp.lib().ngicp_destroy(h)
        return rc == 0
    except Exception:
        return False
