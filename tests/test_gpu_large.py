"""Larger configurations of BASELINE.json (configs 3 and 5): full-HD / 8000 features against the
oracle; 3840x2160 / 20000 features / 4096 hypotheses against live cv2 (ORB in canonical order, the
cross-checked good-match list) and through size-independent properties (canonical ordering, per-level
budgets, self-match identity, cross-check symmetry, determinism)."""
import numpy as np
import pytest

from oracle import match as M
from oracle import orb as O

pytestmark = pytest.mark.gpu


def _frames(w, h, n):
    from b200vo import synth
    frames, R_wc, C_w, K = synth.make_sequence(w, h, n, device="cuda")
    return frames.cpu().numpy(), K


def test_full_hd_8000_features_bit_exact(built):
    frames, K = _frames(1920, 1080, 2)
    with built.Context(1920, 1080, nfeatures=8000, max_frames=2, ransac_hypotheses=256) as ctx:
        ctx.upload_frames(frames, 0)
        ctx.detect(0, 2)
        kp, desc = ctx.get_features(0)
        ref = O.detect_and_compute(frames[0], 8000)
        assert len(kp) == len(ref["x"])
        assert np.array_equal(kp["lx"], ref["x"]) and np.array_equal(kp["ly"], ref["y"]) and np.array_equal(kp["octave"], ref["level"])
        assert np.array_equal(kp["angle"], ref["angle"]) and np.array_equal(kp["response"], ref["response"])
        assert np.array_equal(desc, ref["desc"])
        kp1, desc1 = ctx.get_features(1)
        good, n_cross, thr = ctx.match(0, 1)
        qi, ti, d, thr_o, nc = M.good_matches(desc, desc1)
        assert n_cross == nc and thr == thr_o and np.array_equal(good["queryIdx"], qi) and np.array_equal(good["trainIdx"], ti)
        rec = ctx.process_pairs(0, 1, 0)[0]
        assert rec["status"] & built.ST_E_FOUND and rec["n_matches"] == len(qi)


def test_4k_20000_features_4096_hypotheses_properties(built):
    frames, K = _frames(3840, 2160, 2)
    nper = O.features_per_level(20000)
    with built.Context(3840, 2160, nfeatures=20000, max_frames=2, ransac_hypotheses=4096) as ctx:
        ctx.upload_frames(frames, 0)
        ctx.detect(0, 2)
        feats = [ctx.get_features(s) for s in (0, 1)]
        for kp, desc in feats:
            assert 15000 <= len(kp) <= ctx.kp_cap and desc.shape == (len(kp), 32)
            key = (kp["octave"].astype(np.int64) << 40) | (kp["ly"].astype(np.int64) << 20) | kp["lx"].astype(np.int64)
            assert (np.diff(key) > 0).all()                                 # canonical (octave, y, x), no duplicates
            cnt = np.bincount(kp["octave"], minlength=8)
            assert all(c >= min(n, 1) for c, n in zip(cnt, nper)) and all(c <= n + max(n // 8, 32) for c, n in zip(cnt, nper))
            for l in range(8):                                              # border filter holds on every level
                w, h, s, _ = ctx.level_info(l)
                sel = kp["octave"] == l
                assert (kp["lx"][sel] >= 31).all() and (kp["lx"][sel] < w - 31).all()
                assert (kp["ly"][sel] >= 31).all() and (kp["ly"][sel] < h - 31).all()
            assert (kp["angle"] >= 0).all() and (kp["angle"] < 360).all()
        # matching a frame against itself: every descriptor's nearest neighbour is at distance 0 and the
        # lowest index among its duplicates
        d0 = feats[0][1]
        nq, dq, nt, dt = ctx.match_descriptors(d0, d0)
        assert (dq[0] == 0).all() and (nq[0] <= np.arange(len(d0))).all() and np.array_equal(nq[0], nt[0])
        good, n_cross, thr = ctx.match(0, 1)
        assert len(good) > 100 and (np.diff(good["distance"]) >= 0).all() and (good["distance"] < thr).all()
        a = ctx.process_pairs(0, 1, 7)[0]
        b = ctx.process_pairs(0, 1, 7)[0]
        assert a.tobytes() == b.tobytes()                                   # deterministic
        assert a["status"] & built.ST_E_FOUND and 0 <= a["best_hyp"] < 4096
        R = a["R"].reshape(3, 3)
        assert abs(np.linalg.det(R) - 1) < 1e-9 and abs(np.linalg.norm(a["t"]) - 1) < 1e-9


def test_4k_20000_features_vs_cv2_live(built):
    """BASELINE config 5 at full size against the reference's own OpenCV calls: cv2.ORB with the
    reference's parameters (visual_odometry.cpp:64-85) at nfeatures = 20000 on 3840x2160 frames, compared in
    canonical (octave, y, x) order -- keypoint positions, Harris responses and angles bit-exact, the
    descriptor mismatch fraction stated (bit-exact on the CPU dispatch the oracle was pinned on; <= 1e-3
    allowed for another OpenCV blur dispatch) -- and get_good_matches_of_features
    (visual_odometry.cpp:145-180) on the GPU's descriptors against cv2.BFMatcher + the reference's threshold."""
    cv2 = pytest.importorskip("cv2")
    from oracle import pipeline as PL
    frames, K = _frames(3840, 2160, 2)
    orb = cv2.ORB_create(20000, 1.2, 8, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 10)
    with built.Context(3840, 2160, nfeatures=20000, max_frames=2, ransac_hypotheses=4096) as ctx:
        ctx.upload_frames(frames, 0)
        ctx.detect(0, 2)
        feats = [ctx.get_features(s) for s in (0, 1)]
        good, n_cross, thr = ctx.match(0, 1)
        rec = ctx.process_pairs(0, 1, 0)[0]
    for s in (0, 1):
        kp, desc = feats[s]
        k, d = orb.detectAndCompute(frames[s], None)
        c = O.canonical_order_cv2(k, d)
        assert len(kp) == len(k) > 15000
        assert np.array_equal(kp["lx"], c["x"]) and np.array_equal(kp["ly"], c["y"]) and np.array_equal(kp["octave"], c["level"])
        assert np.array_equal(kp["response"], c["response"]) and np.array_equal(kp["angle"], c["angle"])
        assert np.array_equal(np.stack([kp["x"], kp["y"]], 1), c["pt"]) and np.array_equal(kp["size"], c["size"])
        mism = (desc != c["desc"]).any(1).mean()
        print("frame %d: %d keypoints, descriptor mismatch fraction vs live cv2 = %g" % (s, len(kp), mism))
        assert mism <= 1e-3
    # the reference's matching on the GPU's descriptors (identical inputs on both sides)
    vo = PL.Cv2VisualOdometry(3840, 2160, 20000)
    ref = vo.good_matches(feats[0][1], feats[1][1])
    want = sorted((m.distance, m.queryIdx, m.trainIdx) for m in ref)            # canonical (distance, queryIdx) order
    got = [(float(g["distance"]), int(g["queryIdx"]), int(g["trainIdx"])) for g in good]
    assert len(ref) > 200 and got == want
    assert int(rec["n_matches"]) == len(ref) and rec["status"] & built.ST_E_FOUND
