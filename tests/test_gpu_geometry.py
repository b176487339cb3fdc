"""GPU parity, geometry stages (K12-K17) through the C ABI, under a shared seeded hypothesis set
(SURVEY D5 / section 8c).  Tolerances: every oracle/cv2 solution that is itself accurate (it holds
the ten cubic constraints to 1e-9) must be reproduced to |dE| <= 1e-6 (max-abs, up to sign) in >= 99 %
of the cases, and >= 90 % of ALL oracle solutions (the remainder are ill-conditioned roots where the
unpolished cv2 / numpy solvers only hold the constraints to ~1e-6, while the GPU solver polishes to
~1e-16); inlier counts equal for >= 99 % of models; relative rotation within 0.1 deg and translation
direction within 0.5 deg (north star)."""
import numpy as np
import pytest

from oracle import geometry as G
from oracle import match as M
from oracle import orb as O
from conftest import HAS_CV2

pytestmark = pytest.mark.gpu

NF, NHYP = 2000, 256


@pytest.fixture(scope="module")
def scene(built, seq_kitti):
    frames, R_wc, C_w, K = seq_kitti
    ctx = built.Context(1241, 376, nfeatures=NF, max_frames=4, ransac_hypotheses=NHYP)
    ctx.upload_frames(frames, 0)
    ctx.detect(0, 2)
    f0, f1 = ctx.get_features(0), ctx.get_features(1)
    good, n_cross, thr = ctx.match(0, 1)
    p1 = np.stack([f0[0]["x"][good["queryIdx"]], f0[0]["y"][good["queryIdx"]]], 1)
    p2 = np.stack([f1[0]["x"][good["trainIdx"]], f1[0]["y"][good["trainIdx"]]], 1)
    x1, x2 = G.normalize_points(p1, K), G.normalize_points(p2, K)
    yield dict(ctx=ctx, K=K, x1=x1, x2=x2, p1=p1, p2=p2, good=good, b=built, gt=(R_wc, C_w))
    ctx.close()


def _set_err(a, b):
    return min(np.abs(a - b).max(), np.abs(a + b).max())


def _cubic_residual(e):
    return max(abs(np.linalg.det(e)), np.abs(2 * e @ e.T @ e - np.trace(e @ e.T) * e).max())


def test_sampler_matches_oracle(built):
    for (seed, pid, m) in ((42, 0, 371), (42, 998, 37), (7, 123456789, 5), (2**63 + 5, 3, 100000)):
        assert np.array_equal(built.draw_samples(seed, pid, 64, m), G.draw_samples(seed, pid, 64, m))


def test_five_point_solution_sets(scene):
    x1, x2 = scene["x1"], scene["x2"]
    S = G.draw_samples(42, 0, NHYP, len(x1))
    E, ns = scene["ctx"].five_point(x1, x2, S)
    errs, count_mismatch, res, ref_ok = [], 0, [], []
    for h, s in enumerate(S):
        ref = G.five_point(x1[s], x2[s])
        got = E[h, :ns[h]].reshape(-1, 3, 3)
        count_mismatch += (len(ref) != len(got))
        for e in ref:
            errs.append(min([_set_err(e, g) for g in got], default=1.0))
            ref_ok.append(_cubic_residual(e) < 1e-9)
        for g in got:
            assert abs(np.linalg.norm(g) - 1.0) < 1e-12
            res.append(max(abs(np.linalg.det(g)), np.abs(2 * g @ g.T @ g - np.trace(g @ g.T) * g).max()))
            X1 = np.c_[x1[s], np.ones(5)]; X2 = np.c_[x2[s], np.ones(5)]
            assert np.abs(((X2 @ g) * X1).sum(1)).max() < 1e-12        # epipolar constraint on the sample
    errs = np.array(errs); ref_ok = np.array(ref_ok)
    assert (errs[ref_ok] <= 1e-6).mean() >= 0.99, "only %.4f of the accurate oracle solutions reproduced" % (errs[ref_ok] <= 1e-6).mean()
    assert (errs <= 1e-6).mean() >= 0.90, "only %.3f of all oracle solutions reproduced to 1e-6" % (errs <= 1e-6).mean()
    assert count_mismatch <= 0.05 * NHYP
    assert np.percentile(res, 99) < 1e-9        # the GPU solutions satisfy the cubic constraints


def test_five_point_vs_cv2_live(scene):
    cv2 = pytest.importorskip("cv2")
    x1, x2, p1, p2, K = scene["x1"], scene["x2"], scene["p1"], scene["p2"], scene["K"]
    S = G.draw_samples(42, 1, 128, len(x1))
    E, ns = scene["ctx"].five_point(x1, x2, S)
    errs, ref_ok = [], []
    for h, s in enumerate(S):
        ref = cv2.findEssentialMat(p1[s].astype(np.float32), p2[s].astype(np.float32), K, cv2.RANSAC, 0.999, 2.0)[0]
        if ref is None:
            continue
        for e in ref.reshape(-1, 3, 3):
            errs.append(min([_set_err(e, g) for g in E[h, :ns[h]].reshape(-1, 3, 3)], default=1.0))
            ref_ok.append(_cubic_residual(e) < 1e-9)
    errs = np.array(errs); ref_ok = np.array(ref_ok)
    assert (errs[ref_ok] <= 1e-6).mean() >= 0.99 and (errs <= 1e-6).mean() >= 0.90


def test_inlier_counts_match_oracle(scene):
    x1, x2, K = scene["x1"], scene["x2"], scene["K"]
    S = G.draw_samples(42, 2, 64, len(x1))
    E, ns = scene["ctx"].five_point(x1, x2, S)
    models = np.concatenate([E[h, :ns[h]] for h in range(len(S))])
    thr = G.normalized_threshold(K)
    counts = scene["ctx"].score_models(x1, x2, models, thr)
    ref = np.array([int(G.inlier_mask(m.reshape(3, 3), x1, x2, thr).sum()) for m in models])
    assert (counts == ref).mean() >= 0.99 and np.abs(counts - ref).max() <= 2


def test_recover_pose_matches_oracle_and_cv2(scene):
    x1, x2, K = scene["x1"], scene["x2"], scene["K"]
    S = G.draw_samples(42, 0, 32, len(x1))
    best = G.ransac_shared(x1, x2, G.normalized_threshold(K), S)
    n, R, t, mask = scene["ctx"].recover_pose(best["E"], x1, x2, best["mask"])
    c, Ro, to, ok = G.recover_pose(best["E"], x1, x2, best["mask"])
    assert n == c and np.array_equal(mask.astype(bool), ok)
    assert np.abs(R - Ro).max() < 1e-9 and np.abs(t - to).max() < 1e-9
    try:
        import cv2
    except Exception:
        return
    nc, Rc, tc, mc = cv2.recoverPose(best["E"], scene["p1"].astype(np.float32), scene["p2"].astype(np.float32), K,
                                     mask=best["mask"].astype(np.uint8).reshape(-1, 1).copy())
    assert n == nc and np.abs(R - Rc).max() < 1e-9 and np.abs(t - tc.ravel()).max() < 1e-9


def test_relative_pose_shared_hypotheses(scene):
    """Full K9-K16 for one pair with an explicit sample list, against the oracle run on the same list."""
    ctx, x1, x2, K = scene["ctx"], scene["x1"], scene["x2"], scene["K"]
    S = G.draw_samples(1234, 0, NHYP, len(x1))
    rec, E, mask = ctx.relative_pose(0, 1, pair_id=0, samples=S, want_mask=True)
    assert rec["n_matches"] == len(x1)
    assert rec["status"] & 1 and rec["status"] & 2
    best = G.ransac_shared(x1, x2, G.normalized_threshold(K), S)
    # arg-max: the GPU's winner must be a maximum under the oracle's scoring too
    assert abs(int(rec["n_inliers"]) - best["count"]) <= 1
    og = G.inlier_mask(E, x1, x2, G.normalized_threshold(K))
    assert abs(int(og.sum()) - best["count"]) <= 1
    if (rec["best_hyp"], rec["best_sol"]) == (best["hyp"], best["sol"]) or _set_err(E, best["E"]) < 1e-6:
        c, Ro, to, ok = G.recover_pose(best["E"], x1, x2, best["mask"])
        R = rec["R"].reshape(3, 3)
        assert G.rotation_angle_deg(R, Ro) < 0.1 and G.direction_angle_deg(rec["t"], to) < 0.5
        assert abs(int(rec["n_valid"]) - c) <= 1
    # F = K^-T E K^-1 (reference visual_odometry.cpp:225-246)
    F = ctx.fundamental_from_essential(E)
    assert np.allclose(F, G.fundamental_from_essential(E, K), rtol=1e-12, atol=1e-18)
    # the reference's per-pair self-check: mean |x2^T F x1| over the inliers of findEssentialMat, pixel
    # coordinates (visual_odometry.cpp:248-263), on the GPU's own E and mask
    res = G.epipolar_residual(E, K, scene["p1"], scene["p2"], mask)
    assert res > 0 and abs(float(rec["epi_residual"]) - res) <= 1e-9 * res
    assert int(mask.sum()) == int(rec["n_inliers"])


def test_match_debug_statistics(scene):
    """Row a4: min / max / mean / median of the unfiltered (cross-checked) match distances, the values of
    the reference's [MatchDebug] line (visual_odometry.cpp:119-143,160-168), are carried by the pose
    record -- exactly (the distances are integers; the mean is one double division)."""
    ctx = scene["ctx"]
    d0, d1 = ctx.get_features(0)[1], ctx.get_features(1)[1]
    qi, ti, d = M.cross_check_match(d0, d1)
    st = M.match_statistics(d)
    rec = ctx.relative_pose(0, 1, pair_id=3)[0]
    assert int(rec["n_cross"]) == len(d)
    assert float(rec["dist_min"]) == st["min"] and float(rec["dist_max"]) == st["max"]
    assert float(rec["dist_median"]) == st["median"] and float(rec["dist_mean"]) == st["mean"]
    thr = min(max(3.0 * st["min"], 0.7 * st["median"]), 35.0)
    assert float(rec["threshold"]) == np.float32(thr)
    if HAS_CV2:
        import cv2
        ml = sorted(cv2.BFMatcher_create(cv2.NORM_HAMMING, True).match(d0, d1), key=lambda m: m.distance)
        dist = [m.distance for m in ml]
        assert (dist[0], dist[-1], dist[len(dist) // 2], sum(dist) / len(dist)) == (st["min"], st["max"], st["median"], st["mean"])


def test_pose_tolerance_is_statistical_and_unconditional(built):
    """North star: under a shared seeded hypothesis set the relative rotation agrees within 0.1 deg and the
    translation direction within 0.5 deg.  20 consecutive pairs of the 1241x376 sequence at N = 3000, ten
    hypothesis seeds each = 200 cases of 96 hypotheses, every one against the numpy oracle run on the same
    sample list.  Asserted for EVERY case: the GPU winner's inlier count ties the oracle's best (+-1).
    Asserted on the fraction of cases where both select the same model (same (hypothesis, root), or the
    same E to 1e-6 up to sign): that fraction is >= 90 %, and on all of it dR < 0.1 deg, dt < 0.5 deg.  In the
    remaining cases two different minimal models tie (+-1) at the float threshold; for those the GPU
    model must still be a maximiser under the oracle's own scoring."""
    from b200vo import synth
    NFR, NSEED, NH = 21, 10, 96
    frames, R_wc, C_w, K = synth.make_sequence(1241, 376, NFR, device="cuda")
    frames = frames.cpu().numpy()
    thr = G.normalized_threshold(K)
    same, total, worst_r, worst_t = 0, 0, 0.0, 0.0
    with built.Context(1241, 376, nfeatures=3000, max_frames=NFR + 1, ransac_hypotheses=NH) as ctx:
        ctx.upload_frames(frames, 0)
        ctx.detect(0, NFR)
        feats = [ctx.get_features(s) for s in range(NFR)]
        for i in range(NFR - 1):
            good, _, _ = ctx.match(i, i + 1)
            k0, k1 = feats[i][0], feats[i + 1][0]
            p1 = np.stack([k0["x"][good["queryIdx"]], k0["y"][good["queryIdx"]]], 1)
            p2 = np.stack([k1["x"][good["trainIdx"]], k1["y"][good["trainIdx"]]], 1)
            x1, x2 = G.normalize_points(p1, K), G.normalize_points(p2, K)
            assert len(x1) >= 50
            for seed in range(NSEED):
                S = G.draw_samples(1000 + seed, i, NH, len(x1))
                rec, E, _ = ctx.relative_pose(i, i + 1, pair_id=i, samples=S)
                best = G.ransac_shared(x1, x2, thr, S)
                total += 1
                assert rec["status"] & 2
                assert abs(int(rec["n_inliers"]) - best["count"]) <= 1, (i, seed, int(rec["n_inliers"]), best["count"])
                assert abs(int(G.inlier_mask(E, x1, x2, thr).sum()) - best["count"]) <= 1
                if (int(rec["best_hyp"]), int(rec["best_sol"])) == (best["hyp"], best["sol"]) or _set_err(E, best["E"]) < 1e-6:
                    same += 1
                    c, Ro, to, ok = G.recover_pose(best["E"], x1, x2, best["mask"])
                    dr = G.rotation_angle_deg(rec["R"].reshape(3, 3), Ro)
                    dt = G.direction_angle_deg(rec["t"], to)
                    worst_r, worst_t = max(worst_r, dr), max(worst_t, dt)
                    assert dr < 0.1 and dt < 0.5, (i, seed, dr, dt)
                    assert abs(int(rec["n_valid"]) - c) <= 1
    frac = same / total
    print("same-model fraction %.3f over %d cases; worst dR %.2e deg, worst dt %.2e deg" % (frac, total, worst_r, worst_t))
    assert total == (NFR - 1) * NSEED and frac >= 0.90, frac


def test_device_sampler_equals_explicit_samples(scene):
    ctx, x1 = scene["ctx"], scene["x1"]
    S = scene["b"].draw_samples(42, 17, NHYP, len(x1))
    a, Ea, _ = ctx.relative_pose(0, 1, pair_id=17, samples=S)
    b, Eb, _ = ctx.relative_pose(0, 1, pair_id=17)
    assert a.tobytes() == b.tobytes() and np.array_equal(Ea, Eb)


def test_too_few_matches_fallback(built):
    """< 10 good matches -> status 0, identity pose (reference fallback a, visual_odometry.cpp:344-345)."""
    rng = np.random.default_rng(0)
    with built.Context(640, 240, nfeatures=500, max_frames=2, ransac_hypotheses=64) as ctx:
        d = rng.integers(0, 256, (6, 32), dtype=np.uint8)
        kp = np.zeros(6, built.KEYPOINT_DTYPE)
        kp["x"] = rng.uniform(40, 600, 6); kp["y"] = rng.uniform(40, 200, 6)
        ctx.set_features(0, kp, d); ctx.set_features(1, kp, d)
        rec = ctx.process_pairs(0, 1, 0)[0]
        assert rec["status"] == 0 and rec["n_matches"] < 10
        assert np.array_equal(rec["R"].reshape(3, 3), np.eye(3)) and not rec["t"].any()
        R, t = built.chain_pose(rec, np.eye(3), np.array([1.0, 2, 3]))
        assert np.array_equal(R, np.eye(3)) and np.array_equal(t, [1, 2, 3])


def test_adaptive_ransac_equals_truncated_fixed(built, seq_kitti):
    """ransac_adaptive=1 applies OpenCV's stopping rule (confidence 0.999) after every block of 128
    hypotheses: the result must equal the fixed-H evaluation of exactly the hypotheses it looked at,
    and the number of blocks must be what the rule gives for the best inlier counts."""
    frames = seq_kitti[0]
    H = 1024
    with built.Context(1241, 376, nfeatures=NF, max_frames=4, ransac_hypotheses=H, ransac_adaptive=1) as ca, \
         built.Context(1241, 376, nfeatures=NF, max_frames=4, ransac_hypotheses=H) as cf:
        for c in (ca, cf):
            c.upload_frames(frames, 0)
            c.detect(0, 2)
        ra, Ea, _ = ca.relative_pose(0, 1, pair_id=5)
        rounds = (int(ra["status"]) >> 16) & 0xFF
        assert 1 <= rounds <= H // 128
        m = int(ra["n_matches"])
        S = built.draw_samples(42, 5, H, m)
        rf, Ef, _ = cf.relative_pose(0, 1, pair_id=5, samples=S[:rounds * 128])
        assert np.array_equal(Ea, Ef)
        for k in ("R", "t", "n_matches", "n_inliers", "n_valid", "best_hyp", "best_sol"):
            assert np.array_equal(ra[k], rf[k]), k
        assert (int(ra["status"]) & 0xFFFF) == (int(rf["status"]) & 0xFFFF)
        # stopping rule: best count after each block, from fixed-H runs on growing prefixes
        best_after = []
        for k in range(rounds):
            r, _, _ = cf.relative_pose(0, 1, pair_id=5, samples=S[:(k + 1) * 128])
            best_after.append(int(r["n_inliers"]))
        assert G.adaptive_blocks(best_after + [best_after[-1]] * 8, m, H) == rounds
        # an explicit sample list is always evaluated completely, even in adaptive mode
        rs, Es, _ = ca.relative_pose(0, 1, pair_id=5, samples=S)
        rfull, Efull, _ = cf.relative_pose(0, 1, pair_id=5, samples=S)
        assert np.array_equal(Es, Efull) and ((int(rs["status"]) >> 16) & 0xFF) == 0


def test_five_point_gpu_equals_host_emulation(scene, built):
    """The cooperative solver is one source for device and host (csrc/fivept_group.cuh; on the host the
    lanes of a group are executed one after another, tests/hostsim).  Same samples -> the same solution
    sets up to floating-point contraction differences (the host build disables FMA contraction)."""
    import ctypes, os
    lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim", "libgeom_host.so"))
    x1, x2 = scene["x1"], scene["x2"]
    S = G.draw_samples(77, 3, 200, len(x1))
    E, ns = scene["ctx"].five_point(x1, x2, S)
    same, errs = 0, []
    for h, s in enumerate(S):
        a, b = np.ascontiguousarray(x1[s]), np.ascontiguousarray(x2[s])
        Eh = np.zeros((10, 9))
        n = lib.hs_five_point_group(a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), Eh.ctypes.data_as(ctypes.c_void_p))
        same += (n == ns[h])
        for e in Eh[:n]:
            errs.append(min([_set_err(e, g) for g in E[h, :ns[h]]], default=1.0))
    errs = np.array(errs)
    assert same >= 0.97 * len(S), same
    assert (errs <= 1e-8).mean() >= 0.97, (errs <= 1e-8).mean()


def test_five_point_is_deterministic(scene):
    """The cooperative solver exchanges data through shared memory under phase barriers; a missing
    barrier would show up as run-to-run differences (compute-sanitizer is not available on this pool).
    Same samples, repeated launches -> byte-identical solution sets."""
    x1, x2 = scene["x1"], scene["x2"]
    S = G.draw_samples(5, 11, NHYP, len(x1))
    ref = None
    for _ in range(4):
        E, ns = scene["ctx"].five_point(x1, x2, S)
        cur = (E.tobytes(), ns.tobytes())
        assert ref is None or cur == ref
        ref = cur


def _synthetic_pairs(kind, n, rng):
    """Noise-free correspondences (normalised coordinates) of a known motion; returns x1, x2, E_true or None."""
    X = np.c_[rng.uniform(-4, 4, n), rng.uniform(-2, 2, n), rng.uniform(4, 20, n)]
    if kind == "planar":
        X[:, 2] = 8.0 + 0.3 * X[:, 0] - 0.2 * X[:, 1]
    w = {"forward": [0.01, -0.02, 0.005], "sideways": [0.0, 0.03, 0.0], "planar": [0.02, 0.01, -0.01],
         "rotation": [0.02, -0.03, 0.01], "general": [0.05, -0.04, 0.03]}[kind]
    t = {"forward": [0.0, 0.0, 1.0], "sideways": [1.0, 0.0, 0.0], "planar": [0.6, -0.2, 0.5],
         "rotation": [0.0, 0.0, 0.0], "general": [0.4, -0.3, 0.8]}[kind]
    w = np.asarray(w); th = np.linalg.norm(w); k = w / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    t = np.asarray(t, float)
    Y = X @ R.T + t
    x1, x2 = X[:, :2] / X[:, 2:], Y[:, :2] / Y[:, 2:]
    if not t.any():
        return x1, x2, None
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    E = tx @ R
    return x1, x2, E / np.linalg.norm(E)


@pytest.mark.parametrize("kind", ["general", "forward", "sideways", "planar"])
def test_five_point_contains_the_true_motion(scene, kind):
    """Noise-free correspondences: the true essential matrix must be among the solutions of (nearly)
    every minimal sample, also for the motions where the constraint system is structured (pure forward
    or sideways translation, points on a plane)."""
    rng = np.random.default_rng({"general": 1, "forward": 2, "sideways": 3, "planar": 4}[kind])
    x1, x2, Et = _synthetic_pairs(kind, 400, rng)
    S = G.draw_samples(9, 1, NHYP, len(x1))
    E, ns = scene["ctx"].five_point(x1, x2, S)
    assert ns.min() >= 0 and ns.max() <= 10
    err = np.array([min([_set_err(Et, g.reshape(3, 3)) for g in E[h, :ns[h]]], default=1.0) for h in range(NHYP)])
    assert (err < 1e-6).mean() >= 0.98, (kind, (err < 1e-6).mean())
    counts = scene["ctx"].score_models(x1, x2, Et.reshape(1, 9), 1e-6)
    assert counts[0] == len(x1)                       # every point is an inlier of the true model


@pytest.mark.parametrize("kind", ["rotation", "repeated", "collinear"])
def test_five_point_degenerate_samples_terminate(scene, kind):
    """Degenerate minimal samples (no translation, repeated points, collinear points) must not hang or
    produce non-finite models: whatever is returned is finite, unit-norm, and the launch is repeatable."""
    rng = np.random.default_rng(11)
    x1, x2, _ = _synthetic_pairs("rotation" if kind == "rotation" else "general", 64, rng)
    if kind == "repeated":
        x1[1::2] = x1[0::2]; x2[1::2] = x2[0::2]
    elif kind == "collinear":
        x1[:, 1] = 0.5 * x1[:, 0] + 0.1
        x2[:, 1] = 0.5 * x2[:, 0] + 0.1
    S = G.draw_samples(3, 2, NHYP, len(x1))
    E, ns = scene["ctx"].five_point(x1, x2, S)
    E2, ns2 = scene["ctx"].five_point(x1, x2, S)
    assert np.array_equal(ns, ns2) and E.tobytes() == E2.tobytes()
    assert ns.min() >= 0 and ns.max() <= 10
    for h in range(NHYP):
        g = E[h, :ns[h]]
        assert np.isfinite(g).all()
        if len(g):
            assert np.abs(np.linalg.norm(g, axis=1) - 1.0).max() < 1e-9
    # the full RANSAC call on the same points returns a record, not an error
    n_valid, R, t, mask = scene["ctx"].recover_pose(np.eye(3).ravel() / np.sqrt(3), x1, x2)
    assert np.isfinite(R).all() and np.isfinite(t).all()
