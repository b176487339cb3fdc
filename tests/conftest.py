import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "slam-visual-odometry_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _has_cv2():
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


HAS_CV2 = _has_cv2()
needs_cv2 = pytest.mark.skipif(not HAS_CV2, reason="cv2 not importable")


@pytest.fixture(scope="session")
def built():
    """The in-tree libraries; built on demand so a fresh checkout can run the CPU suite."""
    import b200vo
    if not (os.path.exists(b200vo.LIB_PATH) and os.path.exists(b200vo.HOST_LIB_PATH)
            and os.path.exists(os.path.join(ROOT, "tests", "hostsim", "libgeom_host.so"))):
        import __graft_entry__ as g
        g.build()
    return b200vo


@pytest.fixture(scope="session")
def seq_small():
    """Three 640x240 synthetic frames with ground truth (rendered once per session)."""
    from b200vo import synth
    frames, R_wc, C_w, K = synth.make_sequence(640, 240, 3)
    return frames.numpy(), R_wc, C_w, K


@pytest.fixture(scope="session")
def seq_kitti():
    """Two 1241x376 frames (the headline configuration's frame shape)."""
    from b200vo import synth
    frames, R_wc, C_w, K = synth.make_sequence(1241, 376, 2)
    return frames.numpy(), R_wc, C_w, K


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "pair_480x160.npz")
    return np.load(path)
