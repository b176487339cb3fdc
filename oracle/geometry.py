"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never import this from the product path.

CPU (numpy, float64) restatement of the two-view geometry the reference obtains from OpenCV:
  cv::findEssentialMat(p1, p2, K, RANSAC, 0.999, 2.0, mask)   /root/reference/src/visual_odometry.cpp:213-216
  cv::recoverPose(E, p1, p2, K, R, t, mask)                   /root/reference/src/visual_odometry.cpp:268
  F = K^-T E K^-1                                             /root/reference/src/visual_odometry.cpp:225-246
  pose inversion / chaining                                   /root/reference/src/visual_odometry.cpp:292-314
following SURVEY.md App. A.9, A.9b, A.10.  OpenCV's RANSAC draws its own samples and exits
adaptively, so parity is defined under a SHARED, SEEDED hypothesis set:
for the same list of 5-index samples the solution sets, inlier counts, selected model and the
recovered (R, t) must agree.  Pinned against cv2 4.13.0 (five-point solution sets from
cv2.findEssentialMat on exactly 5 points, masks, recoverPose) by tests/test_oracle_geometry.py.
"""
import itertools

import numpy as np

MASK64 = (1 << 64) - 1
RANSAC_THRESHOLD_PX = 2.0          # visual_odometry.cpp:215
RECOVER_POSE_DIST = 50.0           # cv::recoverPose default distanceThresh


# ----------------------------------------------------------------------------- sampling
def _splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & MASK64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return x, z ^ (z >> 31)


def draw_samples(seed, pair_id, nhyp, m):
    """The product's hypothesis sampler (DESIGN.md 'shared hypothesis set'): 5 distinct indices
    in [0,m) per hypothesis; a pure function of (seed, pair_id, hypothesis index), never of the
    rank or the batch the pair is processed in."""
    out = np.zeros((nhyp, 5), np.int32)
    if m < 5:
        return out
    _, base = _splitmix64((seed ^ ((pair_id + 1) * 0xD1342543DE82EF95)) & MASK64)
    for h in range(nhyp):
        state = (base + h * 0xA24BAED4963EE407) & MASK64
        k = 0
        while k < 5:
            state, z = _splitmix64(state)
            cand = ((z >> 32) * m) >> 32
            if cand in out[h, :k]:
                continue
            out[h, k] = cand
            k += 1
    return out


# ----------------------------------------------------------------------------- normalisation
def normalize_points(pts, K):
    """(u,v) float32 pixels -> double normalised coords ((u-cx)/fx, (v-cy)/fy) (App. A.9)."""
    p = np.asarray(pts, np.float32).astype(np.float64).reshape(-1, 2)
    return np.stack([(p[:, 0] - K[0, 2]) / K[0, 0], (p[:, 1] - K[1, 2]) / K[1, 1]], 1)


def normalized_threshold(K, thr_px=RANSAC_THRESHOLD_PX):
    return thr_px / ((K[0, 0] + K[1, 1]) / 2.0)


# ----------------------------------------------------------------------------- five-point
# monomial order (Nister): columns of the 10x20 constraint matrix
_MONO = [(3, 0, 0), (0, 3, 0), (2, 1, 0), (1, 2, 0), (2, 0, 1), (2, 0, 0), (0, 2, 1), (0, 2, 0), (1, 1, 1), (1, 1, 0),
         (1, 0, 2), (1, 0, 1), (1, 0, 0), (0, 1, 2), (0, 1, 1), (0, 1, 0), (0, 0, 3), (0, 0, 2), (0, 0, 1), (0, 0, 0)]


def _pmul(a, b):
    out = {}
    for (ea, ca), (eb, cb) in itertools.product(a.items(), b.items()):
        e = (ea[0] + eb[0], ea[1] + eb[1], ea[2] + eb[2])
        out[e] = out.get(e, 0.0) + ca * cb
    return out


def _padd(a, b, sb=1.0):
    out = dict(a)
    for e, c in b.items():
        out[e] = out.get(e, 0.0) + sb * c
    return out


def five_point(x1, x2):
    """All real essential matrices for 5 normalised correspondences (x2^T E x1 = 0).
    Returns (k,3,3), each with Frobenius norm 1 (sign arbitrary)."""
    x1 = np.asarray(x1, np.float64); x2 = np.asarray(x2, np.float64)
    A = np.stack([x2[:, 0] * x1[:, 0], x2[:, 0] * x1[:, 1], x2[:, 0],
                  x2[:, 1] * x1[:, 0], x2[:, 1] * x1[:, 1], x2[:, 1],
                  x1[:, 0], x1[:, 1], np.ones(5)], 1)
    _, _, vt = np.linalg.svd(A)
    X, Y, Z, W = (vt[5 + i].reshape(3, 3) for i in range(4))
    E = [[{(1, 0, 0): X[r, c], (0, 1, 0): Y[r, c], (0, 0, 1): Z[r, c], (0, 0, 0): W[r, c]} for c in range(3)]
         for r in range(3)]
    # det(E)
    det = {}
    for (a, b, c, s) in ((0, 1, 2, 1), (0, 2, 1, -1), (1, 2, 0, 1), (1, 0, 2, -1), (2, 0, 1, 1), (2, 1, 0, -1)):
        det = _padd(det, _pmul(_pmul(E[0][a], E[1][b]), E[2][c]), s)
    # 2 E E^T E - tr(E E^T) E
    EEt = [[None] * 3 for _ in range(3)]
    for r in range(3):
        for c in range(3):
            acc = {}
            for k in range(3):
                acc = _padd(acc, _pmul(E[r][k], E[c][k]))
            EEt[r][c] = acc
    tr = _padd(_padd(EEt[0][0], EEt[1][1]), EEt[2][2])
    rows = [det]
    for r in range(3):
        for c in range(3):
            acc = {}
            for k in range(3):
                acc = _padd(acc, _pmul(EEt[r][k], E[k][c]), 2.0)
            acc = _padd(acc, _pmul(tr, E[r][c]), -1.0)
            rows.append(acc)
    M = np.array([[row.get(e, 0.0) for e in _MONO] for row in rows])
    try:
        B = np.linalg.solve(M[:, :10], M[:, 10:])
    except np.linalg.LinAlgError:
        return np.zeros((0, 3, 3))

    def klm(e, f):
        be, bf = B[e], B[f]
        px = np.array([-bf[0], be[0] - bf[1], be[1] - bf[2], be[2]])
        py = np.array([-bf[3], be[3] - bf[4], be[4] - bf[5], be[5]])
        p1 = np.array([-bf[6], be[6] - bf[7], be[7] - bf[8], be[8] - bf[9], be[9]])
        return px, py, p1

    Bm = [klm(4, 5), klm(6, 7), klm(8, 9)]
    pm, ps = np.polymul, np.polysub
    detB = np.polyadd(np.polyadd(
        pm(ps(pm(Bm[0][0], Bm[1][1]), pm(Bm[0][1], Bm[1][0])), Bm[2][2]),
        pm(ps(pm(Bm[0][1], Bm[1][2]), pm(Bm[0][2], Bm[1][1])), Bm[2][0])),
        pm(ps(pm(Bm[0][2], Bm[1][0]), pm(Bm[0][0], Bm[1][2])), Bm[2][1]))
    roots = np.roots(detB)
    sols = []
    for r in roots:
        if abs(r.imag) > 1e-10:
            continue
        z = r.real
        bz = np.array([[np.polyval(Bm[i][j], z) for j in range(3)] for i in range(3)])
        _, _, v = np.linalg.svd(bz)
        if abs(v[2, 2]) < 1e-10:
            continue
        x, y = v[2, 0] / v[2, 2], v[2, 1] / v[2, 2]
        Em = x * X + y * Y + z * Z + W
        sols.append(Em / np.linalg.norm(Em))
    return np.array(sols).reshape(-1, 3, 3)


# ----------------------------------------------------------------------------- scoring
def sampson_err_f32(E, x1, x2):
    """OpenCV's EMEstimatorCallback::computeError, result cast to float (App. A.9)."""
    X1 = np.concatenate([x1, np.ones((len(x1), 1))], 1)
    X2 = np.concatenate([x2, np.ones((len(x2), 1))], 1)
    Ex1 = X1 @ E.T
    Etx2 = X2 @ E
    x2tEx1 = (X2 * Ex1).sum(1)
    den = Ex1[:, 0] ** 2 + Ex1[:, 1] ** 2 + Etx2[:, 0] ** 2 + Etx2[:, 1] ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        return (x2tEx1 * x2tEx1 / den).astype(np.float32)


def inlier_mask(E, x1, x2, thr):
    t = np.float32(thr * thr)
    return sampson_err_f32(E, x1, x2) <= t


def ransac_shared(x1, x2, thr, samples, solver=five_point):
    """Score every solution of every sample over all points; first arg-max wins.
    Returns dict(E, mask, count, hyp, sol, counts=list per hyp)."""
    best = dict(E=None, mask=None, count=-1, hyp=-1, sol=-1)
    counts = []
    for h, s in enumerate(samples):
        sols = solver(x1[s], x2[s])
        cs = []
        for k, E in enumerate(sols):
            m = inlier_mask(E, x1, x2, thr)
            c = int(m.sum())
            cs.append(c)
            if c > best["count"]:
                best = dict(E=E, mask=m, count=c, hyp=h, sol=k)
        counts.append(cs)
    best["counts"] = counts
    return best


def ransac_update_num_iters(p, ep, model_points, max_iters):
    """cv::RANSACUpdateNumIters (OpenCV calib3d/ptsetreg.cpp): iterations needed for confidence p at
    outlier ratio ep.  findEssentialMat calls it after every improvement of the best model."""
    import math
    import sys
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, sys.float_info.min)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < sys.float_info.min:
        return 0
    num = math.log(num)
    denom = math.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))


def adaptive_blocks(best_count_after_block, m, max_iters, block=128, conf=0.999):
    """The product's adaptive mode: hypotheses are evaluated in blocks of `block`; after each block
    the stopping rule above is applied to the best inlier count so far.  Returns the number of blocks
    evaluated given the best count after each block."""
    nb = -(-max_iters // block)
    for k in range(nb):
        cnt = best_count_after_block[k]
        niters = max_iters
        if m > 0 and cnt >= 5:
            niters = ransac_update_num_iters(conf, (m - cnt) / m, 5, max_iters)
        if min((k + 1) * block, max_iters) >= niters or m < 5:
            return k + 1
    return nb


# ----------------------------------------------------------------------------- recoverPose
def decompose_essential(E):
    U, _, Vt = np.linalg.svd(E)
    if np.linalg.det(U) < 0:
        U = -U
    if np.linalg.det(Vt) < 0:
        Vt = -Vt
    W = np.array([[0, 1, 0], [-1, 0, 0], [0, 0, 1.0]])
    return U @ W @ Vt, U @ W.T @ Vt, U[:, 2].copy()


def triangulate(P0, P1, x1, x2):
    n = len(x1)
    Q = np.zeros((4, n))
    for i in range(n):
        A = np.stack([x1[i, 0] * P0[2] - P0[0], x1[i, 1] * P0[2] - P0[1],
                      x2[i, 0] * P1[2] - P1[0], x2[i, 1] * P1[2] - P1[1]])
        _, _, vt = np.linalg.svd(A)
        Q[:, i] = vt[3]
    return Q


def cheirality_mask(R, t, x1, x2, dist=RECOVER_POSE_DIST):
    P0 = np.hstack([np.eye(3), np.zeros((3, 1))])
    P1 = np.hstack([R, t.reshape(3, 1)])
    Q = triangulate(P0, P1, x1, x2)
    with np.errstate(divide="ignore", invalid="ignore"):
        ok = Q[2] * Q[3] > 0
        Q = Q / Q[3]
        ok &= Q[2] < dist
        Q2 = P1 @ Q
        ok &= (Q2[2] > 0) & (Q2[2] < dist)
    return ok


def recover_pose(E, x1, x2, mask_in=None):
    """cv::recoverPose on normalised points (App. A.10): (count, R, t, mask)."""
    R1, R2, t = decompose_essential(E)
    best = None
    for (R, tt) in ((R1, t), (R2, t), (R1, -t), (R2, -t)):
        ok = cheirality_mask(R, tt, x1, x2)
        if mask_in is not None:
            ok &= np.asarray(mask_in).astype(bool).ravel()
        c = int(ok.sum())
        if best is None or c > best[0]:
            best = (c, R, tt, ok)
    return best


# ----------------------------------------------------------------------------- host glue
def fundamental_from_essential(E, K):
    Kinv = np.linalg.inv(K)
    return Kinv.T @ E @ Kinv


def epipolar_residual(E, K, p1, p2, mask):
    """visual_odometry.cpp:248-263: mean |x2^T F x1| over the inliers of findEssentialMat, homogeneous
    PIXEL coordinates, F = K^-T E K^-1 (:225-246).  0 when the mask is empty."""
    F = fundamental_from_essential(E, K)
    p1 = np.asarray(p1, np.float64).reshape(-1, 2); p2 = np.asarray(p2, np.float64).reshape(-1, 2)
    m = np.asarray(mask).astype(bool).ravel()
    if not m.any():
        return 0.0
    x1 = np.c_[p1, np.ones(len(p1))][m]; x2 = np.c_[p2, np.ones(len(p2))][m]
    return float(np.abs(np.einsum("ij,jk,ik->i", x2, F, x1)).mean())


def compose_pose(R_wc, t_wc, R, t, scale):
    """visual_odometry.cpp:292-314: invert (R,t) then T_w_c2 = T_w_c1 * T_c1_c2 with scaled t."""
    R_inv = R.T
    t_inv = -R.T @ t
    return R_wc @ R_inv, t_wc + scale * (R_wc @ t_inv)


def rotation_angle_deg(Ra, Rb):
    c = (np.trace(Ra @ Rb.T) - 1.0) / 2.0
    return float(np.degrees(np.arccos(np.clip(c, -1.0, 1.0))))


def direction_angle_deg(ta, tb):
    ta = np.asarray(ta, float).ravel(); tb = np.asarray(tb, float).ravel()
    c = ta @ tb / (np.linalg.norm(ta) * np.linalg.norm(tb))
    return float(np.degrees(np.arccos(np.clip(c, -1.0, 1.0))))
