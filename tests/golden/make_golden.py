#!/usr/bin/env python
"""Generate tests/golden/pair_480x160.npz from the reference's arithmetic (Python cv2 4.13.0).

The reference (evya1/slam-visual-odometry) has no tests or golden vectors of its own
(SURVEY.md section 4), and all of its hot-path arithmetic is inside OpenCV calls
(/root/reference/src/visual_odometry.cpp:75,87,108,113,153,213,268).  This script runs exactly
those calls, with the reference's parameters, through cv2 and stores their outputs on a small
synthetic frame pair, so the oracle (oracle/*.py) and the CUDA path can be pinned without cv2.

    python tests/golden/make_golden.py        # rewrites the .npz (deterministic)
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-visual-odometry_b200"))
from b200vo import synth            # noqa: E402  (input generator only)
from oracle import geometry as G    # noqa: E402  (sample lists + canonical ordering helpers only)
from oracle import orb as O         # noqa: E402

W, H, NF = 480, 160, 500


def main():
    out = {}
    frames, R_wc, C_w, K = synth.make_sequence(W, H, 2)
    frames = frames.numpy()
    out["frames"] = frames
    out["K"] = K
    out["cv2_version"] = np.array(cv2.__version__)
    # A.1 cvtColor
    rng = np.random.default_rng(1)
    bgr = rng.integers(0, 256, (48, 64, 3), dtype=np.uint8)
    out["bgr"] = bgr
    out["gray_from_bgr"] = cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY)
    # A.2 pyramid chain (INTER_LINEAR_EXACT), frame 0
    sizes = O.level_sizes(W, H)
    lev = frames[0]
    pyr = [lev]
    for l in range(1, 8):
        lev = cv2.resize(lev, sizes[l], interpolation=cv2.INTER_LINEAR_EXACT)
        pyr.append(lev)
    out["pyr_sizes"] = np.array(sizes, np.int32)
    out["pyr_flat"] = np.concatenate([p.ravel() for p in pyr])
    # A.3 FAST per level
    fast = cv2.FastFeatureDetector_create(10, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    fl = []
    for l, p in enumerate(pyr):
        for k in fast.detect(p):
            fl.append((l, int(k.pt[0]), int(k.pt[1]), int(k.response)))
    out["fast"] = np.array(fl, np.int32)
    # A.7 blur (float separable path), level 0
    k7 = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
    out["gauss7"] = k7.ravel()
    out["blur_l0"] = np.clip(np.rint(cv2.sepFilter2D(pyr[0], cv2.CV_32F, k7, k7, borderType=cv2.BORDER_REFLECT_101)), 0, 255).astype(np.uint8)
    # ORB::detectAndCompute with the reference's parameters (visual_odometry.cpp:64-85,113)
    orb = cv2.ORB_create(NF, 1.2, 8, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 10)
    feats = []
    for i in range(2):
        k, d = orb.detectAndCompute(frames[i], None)
        c = O.canonical_order_cv2(k, d)
        feats.append(c)
        for key in ("level", "x", "y", "response", "angle", "pt", "size", "desc"):
            out["f%d_%s" % (i, key)] = c[key]
    # fastAtan2 spot values
    xy = rng.normal(0, 1000, (256, 2)).astype(np.float32)
    out["atan_xy"] = xy
    out["atan_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for x, y in xy], np.float32)
    # BFMatcher cross-check (visual_odometry.cpp:87,153)
    m = cv2.BFMatcher_create(cv2.NORM_HAMMING, True).match(feats[0]["desc"], feats[1]["desc"])
    out["match"] = np.array([(x.queryIdx, x.trainIdx, int(x.distance)) for x in m], np.int32)
    # knn (ratio-test mode oracle)
    knn = cv2.BFMatcher_create(cv2.NORM_HAMMING, False).knnMatch(feats[0]["desc"], feats[1]["desc"], k=2)
    out["knn2"] = np.array([[(a.trainIdx, int(a.distance)) for a in row] for row in knn], np.int32)
    # good matches as the reference filters them (visual_odometry.cpp:155-177), canonical tie order
    mm = sorted(m, key=lambda x: (x.distance, x.queryIdx))
    thr = min(max(3.0 * mm[0].distance, 0.7 * mm[len(mm) // 2].distance), 35.0)
    good = [x for x in mm if x.distance < thr]
    out["good"] = np.array([(x.queryIdx, x.trainIdx, int(x.distance)) for x in good], np.int32)
    out["good_thr"] = np.array(thr)
    p1 = np.float32([feats[0]["pt"][x.queryIdx] for x in good])
    p2 = np.float32([feats[1]["pt"][x.trainIdx] for x in good])
    # five-point solution sets (findEssentialMat on exactly 5 points returns all solutions stacked)
    S = G.draw_samples(42, 0, 48, len(good))
    out["samples"] = S
    sols, nsol = [], []
    for s in S:
        E = cv2.findEssentialMat(p1[s], p2[s], K, cv2.RANSAC, 0.999, 2.0)[0]
        E = np.zeros((0, 3)) if E is None else E
        nsol.append(len(E) // 3)
        sols.append(E.reshape(-1, 9))
    out["fivept_nsol"] = np.array(nsol, np.int32)
    out["fivept_E"] = np.concatenate(sols)
    # full findEssentialMat + recoverPose (visual_odometry.cpp:213-216,268)
    E, mask = cv2.findEssentialMat(p1, p2, K, cv2.RANSAC, 0.999, 2.0)
    out["E_cv"] = E
    out["E_mask"] = mask.ravel()
    n, R, t, m2 = cv2.recoverPose(E, p1, p2, K, mask=mask.copy())
    out["rp_count"] = np.array(n)
    out["rp_R"] = R
    out["rp_t"] = t.ravel()
    out["rp_mask"] = (m2.ravel() > 0).astype(np.uint8)
    out["gt_R"], out["gt_t"] = synth.relative_pose_gt(R_wc, C_w, 0)
    path = os.path.join(ROOT, "tests", "golden", "pair_480x160.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(feats[0]["x"]), len(feats[1]["x"]), "kp,",
          len(m), "cross,", len(good), "good, inliers", int(mask.sum()), "valid", n)


if __name__ == "__main__":
    main()
