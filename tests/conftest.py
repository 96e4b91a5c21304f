import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "active-orchard-slam_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def gpu_ctx():
    from aos_gpu import lib
    ctx = lib.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session", autouse=True)
def _match_this_images_opencv():
    """The checker's Voronoi arithmetic is the cv2 of this image (4.13: Subdiv2D outer triangle at 6 x the rectangle),
    the library's default is the reference platform's OpenCV 4.5.x (3 x).  Probe the real cv2 and tell the library, as a
    node does for the OpenCV it links (include/aos_gpu.h, aos_set_subdiv_outer_factor)."""
    import ctypes
    try:
        from aos_gpu import lib
        from oracle import subdiv
        L = lib.load()
        L.aos_set_subdiv_outer_factor.argtypes = [ctypes.c_float]
        L.aos_set_subdiv_outer_factor(ctypes.c_float(subdiv.outer_factor()))
    except Exception:
        pass   # library not built: the tests that need it fail on their own
    yield
