"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never import this from the product path.

CPU (numpy) restatement of the matching stage of the reference:
  cv::BFMatcher(NORM_HAMMING, crossCheck=true).match   /root/reference/src/visual_odometry.cpp:87,153
  sort by distance, min/median, adaptive threshold       /root/reference/src/visual_odometry.cpp:155-177
following SURVEY.md App. A.8.  Pinned against cv2 4.13.0 by tests/test_oracle_match.py and the
fixtures in tests/golden/.
"""
import numpy as np

MAX_HAMMING_THRESHOLD = 35.0      # visual_odometry.cpp:147


def hamming_matrix(q, t, chunk=512):
    """(n1,n2) int32 Hamming distances between rows of u8 (n,32) descriptor matrices."""
    q = np.ascontiguousarray(q, np.uint8)
    t = np.ascontiguousarray(t, np.uint8)
    q64 = q.view(np.uint64)
    t64 = t.view(np.uint64)
    out = np.empty((len(q), len(t)), np.int32)
    for s in range(0, len(q), chunk):
        x = q64[s:s + chunk, None, :] ^ t64[None, :, :]
        out[s:s + chunk] = np.bitwise_count(x).sum(-1)
    return out


def nearest(q, t):
    """NN of each query over trains and of each train over queries, lowest index wins ties.
    Returns (nn_q idx, nn_q dist, nn_t idx, nn_t dist)."""
    n1, n2 = len(q), len(t)
    nq = np.zeros(n1, np.int32); dq = np.full(n1, 1 << 30, np.int32)
    nt = np.zeros(n2, np.int32); dt = np.full(n2, 1 << 30, np.int32)
    chunk = 1024
    for s in range(0, n1, chunk):
        D = hamming_matrix(q[s:s + chunk], t)
        a = D.argmin(1)                       # numpy argmin returns the first minimum
        nq[s:s + chunk] = a
        dq[s:s + chunk] = D[np.arange(len(a)), a]
        ca = D.argmin(0)
        cd = D[ca, np.arange(n2)]
        better = cd < dt                      # strict: earlier chunk (lower index) wins ties
        nt[better] = ca[better] + s
        dt[better] = cd[better]
    return nq, dq, nt, dt


def cross_check_match(q, t):
    """BFMatcher(NORM_HAMMING, True).match(q, t): (queryIdx, trainIdx, distance) ascending queryIdx."""
    if len(q) == 0 or len(t) == 0:
        z = np.zeros(0, np.int32)
        return z, z.copy(), z.copy()
    nq, dq, nt, _ = nearest(q, t)
    i = np.arange(len(q), dtype=np.int32)
    keep = nt[nq] == i
    return i[keep], nq[keep], dq[keep]


def knn2_match(q, t):
    """BFMatcher(NORM_HAMMING, False).knnMatch(k=2): stable (distance, index) ordering.
    Returns idx (n1,2) int32 and dist (n1,2) int32."""
    D = hamming_matrix(q, t)
    order = np.argsort(D, axis=1, kind="stable")[:, :2]
    return order.astype(np.int32), np.take_along_axis(D, order, 1)


def ratio_test_match(q, t, ratio=0.75):
    """Lowe ratio test on knnMatch(k=2) (the north star's second matching mode; the reference itself
    only cross-checks, visual_odometry.cpp:87): keep m iff m.distance < ratio * n.distance, evaluated
    on DMatch's float32 distances; a query with fewer than two neighbours is dropped.
    Returns (queryIdx, trainIdx, distance) ascending queryIdx."""
    if len(q) == 0 or len(t) < 2:
        z = np.zeros(0, np.int32)
        return z, z.copy(), z.copy()
    idx, dist = knn2_match(q, t)
    keep = dist[:, 0].astype(np.float32) < np.float32(ratio) * dist[:, 1].astype(np.float32)
    i = np.arange(len(q), dtype=np.int32)
    return i[keep], idx[keep, 0], dist[keep, 0]


def adaptive_threshold(dist_sorted):
    """visual_odometry.cpp:160-166: thr = min(max(3*min, 0.7*median), 35); median = sorted[size/2]."""
    mn = float(dist_sorted[0])
    med = float(dist_sorted[len(dist_sorted) // 2])
    return min(max(3.0 * mn, 0.7 * med), MAX_HAMMING_THRESHOLD)


def match_statistics(dist):
    """The [MatchDebug] values of the unfiltered match list (visual_odometry.cpp:119-143,160-168):
    min = sorted.front(), max = sorted.back(), mean = sum/size in double, median = element size/2.
    Returns dict(min, max, mean, median); all 0 for an empty list (the reference returns before printing)."""
    d = np.sort(np.asarray(dist, np.float64))
    if len(d) == 0:
        return {"min": 0.0, "max": 0.0, "mean": 0.0, "median": 0.0}
    return {"min": float(d[0]), "max": float(d[-1]), "mean": float(d.sum() / len(d)), "median": float(d[len(d) // 2])}


def good_matches(q, t, ratio=None):
    """get_good_matches_of_features (visual_odometry.cpp:145-180) in the canonical order
    (distance, queryIdx) -- std::sort there is unstable, so ties have no defined order.
    ratio=None: the reference's cross-check matcher; a float: the ratio-test matcher in its place.
    Returns (queryIdx, trainIdx, distance, threshold, n_cross)."""
    qi, ti, d = cross_check_match(q, t) if ratio is None else ratio_test_match(q, t, ratio)
    if len(qi) == 0:
        return qi, ti, d, 0.0, 0
    order = np.lexsort((qi, d))
    qi, ti, d = qi[order], ti[order], d[order]
    thr = adaptive_threshold(d)
    keep = d.astype(np.float64) < thr
    return qi[keep], ti[keep], d[keep], thr, len(order)
