"""GPU parity, feature stages (K0-K8), through the C ABI: bit-exact against the numpy oracle
(which is itself pinned bit-exactly to cv2, tests/test_oracle_orb.py)."""
import numpy as np
import pytest

from oracle import orb as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx_small(built, seq_small):
    frames = seq_small[0]
    ctx = built.Context(640, 240, nfeatures=1000, max_frames=4, ransac_hypotheses=256)
    ctx.upload_frames(frames, 0)
    ctx.detect(0, len(frames))
    yield ctx
    ctx.close()


def test_level_geometry(ctx_small):
    sizes = O.level_sizes(640, 240)
    nper = O.features_per_level(1000)
    for l in range(8):
        w, h, s, n = ctx_small.level_info(l)
        assert (w, h) == sizes[l] and n == nper[l] and np.float32(s) == O.level_scales()[l]


def test_pyramid_bit_exact(ctx_small, seq_small):
    for slot in range(3):
        pyr = O.build_pyramid(seq_small[0][slot])
        for l in range(8):
            got = ctx_small.get_pyramid_level(slot, l)
            assert np.array_equal(got, pyr[l]), "level %d differs in %d px" % (l, (got != pyr[l]).sum())


def test_fast_candidates_exact(ctx_small, seq_small):
    pyr = O.build_pyramid(seq_small[0][0])
    for l in range(8):
        img = pyr[l]
        xs, ys, sc = O.fast_detect(img)
        h, w = img.shape
        inb = (xs >= 31) & (xs < w - 31) & (ys >= 31) & (ys < h - 31)
        gx, gy, gs = ctx_small.get_fast_candidates(0, l)
        assert np.array_equal(np.stack([gy, gx, gs]), np.stack([ys[inb], xs[inb], sc[inb]])), "level %d" % l


def test_blur_bit_exact(ctx_small, seq_small):
    pyr = O.build_pyramid(seq_small[0][1])
    for l in range(8):
        got = ctx_small.get_pyramid_level(1, l, blurred=True)
        ref = O.gaussian_blur_7x7(pyr[l])
        assert np.array_equal(got, ref), "level %d differs in %d px" % (l, (got != ref).sum())


def _check_features(kp, desc, ref):
    assert len(kp) == len(ref["x"])
    assert np.array_equal(kp["octave"], ref["level"])
    assert np.array_equal(kp["lx"], ref["x"]) and np.array_equal(kp["ly"], ref["y"])
    assert np.array_equal(kp["response"], ref["response"]), "Harris responses not bit-exact"
    assert np.array_equal(kp["angle"], ref["angle"]), "orientations not bit-exact"
    assert np.array_equal(np.stack([kp["x"], kp["y"]], 1), ref["pt"])
    assert np.array_equal(kp["size"], ref["size"])
    assert (kp["class_id"] == -1).all()
    mism = (desc != ref["desc"]).any(1).mean() if len(kp) else 0.0
    assert mism == 0.0, "descriptor mismatch fraction %g" % mism


def test_features_bit_exact_small(ctx_small, seq_small):
    for slot in range(3):
        kp, desc = ctx_small.get_features(slot)
        _check_features(kp, desc, O.detect_and_compute(seq_small[0][slot], 1000))


@pytest.mark.parametrize("nfeatures", [1200, 3000])
def test_features_bit_exact_kitti(built, seq_kitti, nfeatures):
    frames = seq_kitti[0]
    with built.Context(1241, 376, nfeatures=nfeatures, max_frames=2, ransac_hypotheses=64) as ctx:
        ctx.upload_frames(frames, 0)
        ctx.detect(0, 2)
        for slot in range(2):
            kp, desc = ctx.get_features(slot)
            _check_features(kp, desc, O.detect_and_compute(frames[slot], nfeatures))


def test_features_vs_cv2_live(built, seq_kitti):
    cv2 = pytest.importorskip("cv2")
    frames = seq_kitti[0]
    orb = cv2.ORB_create(2000, 1.2, 8, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 10)
    with built.Context(1241, 376, nfeatures=2000, max_frames=2, ransac_hypotheses=64) as ctx:
        ctx.upload_frames(frames, 0)
        ctx.detect(0, 1)
        kp, desc = ctx.get_features(0)
    k, d = orb.detectAndCompute(frames[0], None)
    c = O.canonical_order_cv2(k, d)
    assert len(kp) == len(k)
    assert np.array_equal(kp["lx"], c["x"]) and np.array_equal(kp["ly"], c["y"]) and np.array_equal(kp["octave"], c["level"])
    assert np.array_equal(kp["response"], c["response"]) and np.array_equal(kp["angle"], c["angle"])
    # descriptors: bit-exact on the CPU dispatch the oracle was pinned on; other host CPUs may take a
    # different OpenCV blur dispatch (<= 2.5e-4), so the live check allows 1e-3
    mism = (desc != c["desc"]).any(1).mean()
    assert mism <= 1e-3, "descriptor mismatch fraction vs live cv2: %g" % mism


def test_bgr_input_matches_gray(built, seq_small):
    """3-channel upload goes through the device BGR->gray kernel (reference visual_odometry.cpp:105-111)."""
    rng = np.random.default_rng(7)
    gray = seq_small[0][0]
    bgr = np.stack([np.clip(gray.astype(int) + rng.integers(-20, 20, gray.shape), 0, 255)] * 3, -1).astype(np.uint8)
    bgr[..., 1] = gray
    ref_gray = O.bgr2gray(bgr)
    with built.Context(640, 240, nfeatures=500, max_frames=2, ransac_hypotheses=64) as ctx:
        ctx.upload_frames(bgr[None], 0)
        got = ctx.get_pyramid_level(0, 0)
        assert np.array_equal(got, ref_gray)
        ctx.detect(0, 1)
        kp, desc = ctx.get_features(0)
        _check_features(kp, desc, O.detect_and_compute(ref_gray, 500))


def test_ring_wraparound_and_batch_invariance(built, seq_small):
    """Slots are a ring; results must not depend on slot position or batch composition."""
    frames = seq_small[0]
    with built.Context(640, 240, nfeatures=1000, max_frames=3, ransac_hypotheses=64) as ctx:
        ctx.upload_frames(frames[:2], 2)          # slots 2, 0 (wraps)
        ctx.detect(2, 2)
        a = ctx.get_features(2)
        b = ctx.get_features(0)
        ctx.upload_frames(frames[:1], 1)
        ctx.detect(1, 1)
        c = ctx.get_features(1)
    assert np.array_equal(a[0], c[0]) and np.array_equal(a[1], c[1])
    ref = O.detect_and_compute(frames[1], 1000)
    assert np.array_equal(b[1], ref["desc"])


def test_merged_and_per_level_launches_agree(built, seq_small):
    """Batches of up to 16 frames run FAST and the blur with one launch over all pyramid levels, larger
    batches with one launch per level (common.cuh kMergedLevelsMaxFrames): the same frames through both
    give byte-identical keypoints, descriptors, FAST candidates and blurred levels."""
    frames = np.concatenate([seq_small[0]] * 8)[:20]              # 20 frames: > 16 -> per-level launches
    frames = np.ascontiguousarray(frames)
    frames[3:] = np.roll(frames[3:], 5, axis=2)                    # not all copies identical
    with built.Context(640, 240, nfeatures=1000, max_frames=20, ransac_hypotheses=64) as big, \
         built.Context(640, 240, nfeatures=1000, max_frames=20, ransac_hypotheses=64) as small:
        big.upload_frames(frames, 0)
        big.detect(0, 20)
        small.upload_frames(frames, 0)
        for s in range(0, 20, 5):
            small.detect(s, 5)                                     # <= 16 -> merged launches
        for s in (0, 3, 7, 19):
            a, b = big.get_features(s), small.get_features(s)
            assert a[0].tobytes() == b[0].tobytes() and np.array_equal(a[1], b[1])
            for lv in (0, 4, 5, 7):
                assert np.array_equal(big.get_pyramid_level(s, lv, blurred=True), small.get_pyramid_level(s, lv, blurred=True))
                ca, cb = big.get_fast_candidates(s, lv), small.get_fast_candidates(s, lv)
                assert all(np.array_equal(x, y) for x, y in zip(ca, cb))
        ref = O.detect_and_compute(frames[19], 1000)
        assert np.array_equal(big.get_features(19)[1], ref["desc"])


def test_flat_image_has_no_features(built):
    with built.Context(640, 240, nfeatures=500, max_frames=2, ransac_hypotheses=64) as ctx:
        ctx.upload_frames(np.full((1, 240, 640), 77, np.uint8), 0)
        ctx.detect(0, 1)
        kp, desc = ctx.get_features(0)
        assert len(kp) == 0 and desc.shape == (0, 32)


def test_noise_image_ties_and_budget(built):
    """Dense corners: the per-level budgets bind and retainBest tie-keeping is exercised."""
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (240, 640), dtype=np.uint8)
    img = (img // 32 * 32).astype(np.uint8)        # few grey levels -> many FAST-score ties
    with built.Context(640, 240, nfeatures=800, max_frames=2, ransac_hypotheses=64) as ctx:
        ctx.upload_frames(img[None], 0)
        ctx.detect(0, 1)
        kp, desc = ctx.get_features(0)
    _check_features(kp, desc, O.detect_and_compute(img, 800))


def test_tma_and_word_staging_agree(built, seq_small, monkeypatch):
    """The TMA tile loads (FAST, blur) and the plain word-load staging must give identical features."""
    frames = seq_small[0]
    out = {}
    for tma in ("1", "0"):
        monkeypatch.setenv("B200VO_TMA", tma)
        with built.Context(640, 240, nfeatures=1000, max_frames=4, ransac_hypotheses=64) as ctx:
            ctx.upload_frames(frames, 0)
            ctx.detect(0, 3)
            out[tma] = [ctx.get_features(s) for s in range(3)] + [ctx.get_pyramid_level(1, l, blurred=True) for l in range(8)]
    for a, b in zip(out["1"], out["0"]):
        if isinstance(a, tuple):
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        else:
            assert np.array_equal(a, b)


def test_fused_describe_equals_three_kernel_path(built, seq_kitti, monkeypatch):
    """The fused orientation + keypoint + rBRIEF kernel on per-keypoint TMA boxes (default) and the
    round-1 path (orient_kernel, keypoint_finalize_kernel, brief_kernel reading global memory;
    B200VO_DESCRIBE=0) produce byte-identical keypoints and descriptors at the headline configuration."""
    frames = seq_kitti[0]
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("B200VO_DESCRIBE", mode)
        with built.Context(1241, 376, nfeatures=3000, max_frames=3, ransac_hypotheses=64) as ctx:
            ctx.upload_frames(frames, 0)
            ctx.detect(0, 2)
            out[mode] = [ctx.get_features(s) for s in range(2)]
    for (ka, da), (kb, db) in zip(out["1"], out["0"]):
        assert len(ka) >= 2900 and ka.tobytes() == kb.tobytes() and np.array_equal(da, db)


@pytest.mark.parametrize("W,H,NF", [(100, 80, 300), (333, 222, 700), (1999, 1001, 5000), (64, 64, 100), (4095, 300, 2000)])
def test_odd_sizes_bit_exact(built, W, H, NF):
    """Widths that are not multiples of 4 / 16, the smallest and the widest supported images, a level
    smaller than the border, and a feature budget whose selection kernel needs exactly 48 KB of dynamic
    shared memory (a launch failure found with this case)."""
    from b200vo import synth
    frames = synth.make_sequence(W, H, 2, device="cuda")[0].cpu().numpy()
    with built.Context(W, H, nfeatures=NF, max_frames=3, ransac_hypotheses=64) as ctx:
        ctx.upload_frames(frames, 0)
        ctx.detect(0, 2)
        for s in (0, 1):
            kp, desc = ctx.get_features(s)
            ref = O.detect_and_compute(frames[s], NF)
            assert len(kp) == len(ref["x"])
            assert np.array_equal(kp["lx"], ref["x"]) and np.array_equal(kp["ly"], ref["y"]) and np.array_equal(kp["octave"], ref["level"])
            assert np.array_equal(kp["angle"], ref["angle"]) and np.array_equal(kp["response"], ref["response"])
            assert np.array_equal(desc, ref["desc"])
        rec = ctx.process_pairs(0, 1, 0)[0]
        assert rec["n_matches"] >= 0


@pytest.mark.parametrize("W,H,NF", [(100, 80, 300), (333, 222, 700), (1999, 1001, 5000), (64, 64, 100), (4095, 300, 2000), (1241, 376, 3000)])
def test_large_batch_kernels_at_odd_sizes(built, W, H, NF, monkeypatch):
    """The kernels of large launch batches (warp-strip pyramid and blur, one FAST / blur launch per level, plain launches)
    forced onto two-frame batches with B200VO_SMALL_BATCH_MAX=0, at sizes whose tiles overhang in both directions, whose
    levels are smaller than one tile, and at the bench size: pyramid levels, blurred levels, keypoints and descriptors
    byte-identical to the small-batch kernels (which test_odd_sizes_bit_exact pins against the oracle)."""
    from b200vo import synth
    frames = synth.make_sequence(W, H, 2, device="cuda")[0].cpu().numpy()
    out = {}
    for mode in ("16", "0"):
        monkeypatch.setenv("B200VO_SMALL_BATCH_MAX", mode)
        with built.Context(W, H, nfeatures=NF, max_frames=3, ransac_hypotheses=64) as ctx:
            ctx.upload_frames(frames, 0)
            ctx.detect(0, 2)
            out[mode] = ([ctx.get_features(s) for s in range(2)], [ctx.get_pyramid_level(1, l) for l in range(8)],
                         [ctx.get_pyramid_level(1, l, blurred=True) for l in range(8)], ctx.process_pairs(0, 1, 0))
    for (ka, da), (kb, db) in zip(out["16"][0], out["0"][0]):
        assert ka.tobytes() == kb.tobytes() and np.array_equal(da, db)
    for a, b in zip(out["16"][1] + out["16"][2], out["0"][1] + out["0"][2]):
        assert np.array_equal(a, b)
    assert out["16"][3].tobytes() == out["0"][3].tobytes()


@pytest.mark.parametrize("thr", [1, 40, 127, 128, 200])
def test_fast_threshold_range(built, seq_small, thr):
    """FAST corners for thresholds across the supported range (the byte-SIMD pre-test switches formula at
    t >= 128), against the oracle's score map and live cv2."""
    frames = seq_small[0]
    with built.Context(640, 240, nfeatures=500, max_frames=2, fast_threshold=thr) as ctx:
        ctx.upload_frames(frames[:1], 0)
        ctx.detect(0, 1)
        pyr = O.build_pyramid(frames[0])
        for level in (0, 3):
            x, y, sc = ctx.get_fast_candidates(0, level)
            xs, ys, ss = O.fast_detect(pyr[level], thr)
            h, w = pyr[level].shape
            keep = (xs >= 31) & (xs < w - 31) & (ys >= 31) & (ys < h - 31)
            ref = sorted(zip(ys[keep].tolist(), xs[keep].tolist(), ss[keep].tolist()))
            assert sorted(zip(y.tolist(), x.tolist(), sc.tolist())) == ref, "threshold %d level %d" % (thr, level)
