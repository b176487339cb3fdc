"""Pins oracle/match.py to cv2.BFMatcher (golden vectors + live cv2) and to the reference's
adaptive threshold (visual_odometry.cpp:145-180)."""
import numpy as np

from conftest import needs_cv2
from oracle import match as M


def test_cross_check_golden(golden):
    qi, ti, d = M.cross_check_match(golden["f0_desc"], golden["f1_desc"])
    g = golden["match"]
    assert np.array_equal(qi, g[:, 0]) and np.array_equal(ti, g[:, 1]) and np.array_equal(d, g[:, 2])


def test_knn2_golden(golden):
    idx, dist = M.knn2_match(golden["f0_desc"], golden["f1_desc"])
    assert np.array_equal(idx, golden["knn2"][:, :, 0]) and np.array_equal(dist, golden["knn2"][:, :, 1])


def test_good_matches_golden(golden):
    qi, ti, d, thr, nc = M.good_matches(golden["f0_desc"], golden["f1_desc"])
    g = golden["good"]
    assert thr == float(golden["good_thr"]) and nc == len(golden["match"])
    assert np.array_equal(qi, g[:, 0]) and np.array_equal(ti, g[:, 1]) and np.array_equal(d, g[:, 2])


def test_adaptive_threshold_rule():
    assert M.adaptive_threshold(np.array([0, 10, 20, 30, 40])) == 14.0         # 0.7 * median(=20)
    assert M.adaptive_threshold(np.array([10, 11, 12])) == 30.0                # 3 * min
    assert M.adaptive_threshold(np.array([20, 60, 90, 100])) == 35.0           # capped
    assert M.adaptive_threshold(np.array([7])) == 21.0


def test_empty_inputs():
    z = np.zeros((0, 32), np.uint8)
    d = np.zeros((4, 32), np.uint8)
    assert len(M.cross_check_match(z, d)[0]) == 0 and len(M.cross_check_match(d, z)[0]) == 0
    assert len(M.good_matches(z, d)[0]) == 0


@needs_cv2
def test_cross_check_live_cv2_tie_heavy():
    import cv2
    rng = np.random.default_rng(4)
    bf = cv2.BFMatcher_create(cv2.NORM_HAMMING, True)
    for trial in range(20):
        n1, n2 = rng.integers(1, 200, 2)
        base = rng.integers(0, 256, 32, dtype=np.uint8)
        q = np.tile(base, (n1, 1)); t = np.tile(base, (n2, 1))
        q[:, 0] ^= rng.integers(0, 8, n1, dtype=np.uint8); t[:, 0] ^= rng.integers(0, 8, n2, dtype=np.uint8)
        m = bf.match(q, t)
        qi, ti, d = M.cross_check_match(q, t)
        assert np.array_equal(qi, [x.queryIdx for x in m]) and np.array_equal(ti, [x.trainIdx for x in m])
        assert np.array_equal(d, [int(x.distance) for x in m])


@needs_cv2
def test_ratio_test_live_cv2():
    """The ratio-test mode's oracle against cv2.knnMatch + Lowe's test written the usual way."""
    import cv2
    rng = np.random.default_rng(8)
    bf = cv2.BFMatcher_create(cv2.NORM_HAMMING, False)
    for trial in range(6):
        n1, n2 = rng.integers(2, 400, 2)
        t = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
        q = t[rng.integers(0, n2, n1)].copy()
        for i in range(n1):
            for b in rng.integers(0, 256, rng.integers(0, 80)):
                q[i, b // 8] ^= np.uint8(1 << (b % 8))
        ratio = [0.6, 0.75, 0.8][trial % 3]
        keep = [(m.queryIdx, m.trainIdx, int(m.distance)) for m, n in bf.knnMatch(q, t, k=2)
                if m.distance < np.float32(ratio) * n.distance]
        qi, ti, d = M.ratio_test_match(q, t, ratio)
        assert [tuple(x) for x in zip(qi.tolist(), ti.tolist(), d.tolist())] == keep


def test_ratio_test_needs_two_neighbours():
    d = np.random.default_rng(0).integers(0, 256, (5, 32), dtype=np.uint8)
    assert len(M.ratio_test_match(d, d[:1])[0]) == 0
    assert len(M.good_matches(d, d[:1], ratio=0.75)[0]) == 0
