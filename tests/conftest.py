import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
for p in (REPO / "tests", REPO / "oracle", REPO / "mf-lbm-cuda_b200", REPO):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gpu_lib():
    """The CUDA library, loaded - fails loudly (no fallback) when missing or when no GPU is visible."""
    import ctypes
    import mflbm
    lib = mflbm.load_library()
    cudart = ctypes.CDLL("libcudart.so")
    n = ctypes.c_int(0)
    rc = cudart.cudaGetDeviceCount(ctypes.byref(n))
    assert rc == 0 and n.value > 0, "gpu-marked test run without a visible CUDA device"
    return lib
