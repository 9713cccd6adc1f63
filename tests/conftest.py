import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pkg():
    import importlib
    importlib.import_module(entry.PKG_NAME + ".build").build()
    return entry.load_package()


@pytest.fixture(scope="session")
def cpu_oracle():
    import oracle
    return oracle.cpu()


def has_gpu() -> bool:
    try:
        import ctypes
        lib = ctypes.CDLL("libcuda.so.1")
        n = ctypes.c_int(0)
        return lib.cuInit(0) == 0 and lib.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


@pytest.fixture(scope="session")
def gpu(pkg):
    """The GPU tests must run on the CUDA path: fail (not skip) if it is unavailable."""
    assert has_gpu(), "no CUDA device visible: -m gpu tests need a B200"
    return pkg


# ---- shared scene helpers -------------------------------------------------------------------
HALL_SMALL = (32, 24, 12)   # 8 x 6 x 3 m "room" (SURVEY.md §8 d)
HALL_LARGE = (48, 40, 12)   # 12 x 10 x 3 m "hall"

CAMERAS = {
    "c1": dict(W=640, H=480, f=525.0, cx=319.5, cy=239.5),
    "c2": dict(W=1280, H=720, f=900.0, cx=639.5, cy=359.5),
    "c3": dict(W=1920, H=1080, f=1400.0, cx=959.5, cy=539.5),
}


def make_calib(pkg, W, H, f, cx, cy, dist=None):
    c = pkg.CameraCalibration()
    c.loadCalibration(f, f, cx, cy, dist or [0.0] * 5, W, H)
    return c


def scaled_camera(W, H):
    """A ~70 degree camera for an arbitrary resolution."""
    return dict(W=W, H=H, f=0.73 * W, cx=(W - 1) / 2.0, cy=(H - 1) / 2.0)


def bgra_of(records: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(records[:, 3]).view(np.uint32)
