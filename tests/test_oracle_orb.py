"""Pins the numpy oracle (oracle/orb.py) to the reference's arithmetic: the committed cv2 golden
vectors (tests/golden/make_golden.py) and, when cv2 is importable, live cv2 at the headline size."""
import numpy as np
import pytest

from conftest import needs_cv2
from oracle import orb as O


def test_bgr2gray_golden(golden):
    assert np.array_equal(O.bgr2gray(golden["bgr"]), golden["gray_from_bgr"])


def test_level_sizes_and_budgets():
    assert O.level_sizes(1241, 376) == [(1241, 376), (1034, 313), (862, 261), (718, 218), (598, 181), (499, 151), (416, 126), (346, 105)]
    assert O.features_per_level(1200) == [261, 217, 181, 151, 126, 105, 87, 72]
    assert O.features_per_level(2000) == [434, 362, 302, 251, 209, 175, 145, 122]
    assert O.features_per_level(3000) == [652, 543, 452, 377, 314, 262, 218, 182]
    assert O.features_per_level(8000) == [1737, 1448, 1207, 1005, 838, 698, 582, 485]
    assert O.features_per_level(20000) == [4343, 3620, 3016, 2514, 2095, 1746, 1455, 1211]


def test_pyramid_golden(golden):
    pyr = O.build_pyramid(golden["frames"][0])
    flat = golden["pyr_flat"]
    off = 0
    for l, (w, h) in enumerate(golden["pyr_sizes"]):
        assert pyr[l].shape == (h, w)
        assert np.array_equal(pyr[l].ravel(), flat[off:off + w * h]), "level %d" % l
        off += w * h


def test_fast_golden(golden):
    pyr = O.build_pyramid(golden["frames"][0])
    g = golden["fast"]
    for l in range(8):
        xs, ys, sc = O.fast_detect(pyr[l])
        ref = g[g[:, 0] == l][:, 1:]
        assert np.array_equal(np.stack([xs, ys, sc], 1), ref), "level %d" % l     # same raster order too


def test_gauss_kernel_and_blur_golden(golden):
    assert np.array_equal(O.GAUSS7.view(np.uint32), golden["gauss7"].astype(np.float32).view(np.uint32))
    assert np.array_equal(O.gaussian_blur_7x7(golden["frames"][0]), golden["blur_l0"])


def test_fast_atan2_golden(golden):
    xy = golden["atan_xy"]
    assert np.array_equal(O.fast_atan2(xy[:, 1], xy[:, 0]), golden["atan_deg"])
    assert O.fast_atan2(np.float32(0), np.float32(0)) == 0


@pytest.mark.parametrize("i", [0, 1])
def test_detect_and_compute_golden(golden, i):
    r = O.detect_and_compute(golden["frames"][i], 500)
    for key in ("level", "x", "y", "response", "angle", "pt", "size", "desc"):
        assert np.array_equal(r[key], golden["f%d_%s" % (i, key)]), key


def test_retain_best_keeps_ties():
    resp = np.array([5, 3, 3, 3, 1, 9], np.float32)
    assert O.retain_best_mask(resp, 3).tolist() == [True, True, True, True, False, True]
    assert O.retain_best_mask(resp, 10).all() and not O.retain_best_mask(resp, 0).any()


def test_brief_pattern_identity():
    import hashlib
    p = O.brief_pattern()
    assert p.shape == (256, 4) and p.sum() == -406 and np.abs(p).sum() == 6854
    assert hashlib.sha256(p.astype("<i4").tobytes()).hexdigest() == "7e645581387b82784797e8adddb9b6f0c12611859fda09ca8a9bec96d767a05f"


@needs_cv2
@pytest.mark.parametrize("nf", [1200, 3000])
def test_detect_and_compute_live_cv2_kitti(seq_kitti, nf):
    import cv2
    img = seq_kitti[0][0]
    orb = cv2.ORB_create(nf, 1.2, 8, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 10)
    k, d = orb.detectAndCompute(img, None)
    c = O.canonical_order_cv2(k, d)
    r = O.detect_and_compute(img, nf)
    assert len(k) == len(r["x"])
    for key in ("level", "x", "y", "response", "angle", "pt", "size", "desc"):
        assert np.array_equal(r[key], c[key]), key


@needs_cv2
def test_resize_live_cv2_odd_sizes():
    import cv2
    rng = np.random.default_rng(0)
    for (w, h) in ((1920, 1080), (640, 480), (346, 105), (97, 71)):
        img = rng.integers(0, 256, (h, w), dtype=np.uint8)
        dw, dh = O.level_sizes(w, h)[1]
        assert np.array_equal(O.resize_linear_exact(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT))
