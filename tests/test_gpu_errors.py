"""Error behaviour at the C ABI on the GPU: hard errors are negative return codes with a message
(never an exception across the boundary, never a silent fallback); the reference's soft failures
(src/visual_odometry.cpp:340-366: too few matches, no model) are status bits in the pose record."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H = 320, 240


@pytest.fixture(scope="module")
def ctx(built):
    c = built.Context(W, H, nfeatures=500, max_frames=4, ransac_hypotheses=64)
    yield c
    c.close()


def _code(built, fn, *a):
    with pytest.raises(built.B200VOError) as e:
        fn(*a)
    assert str(e.value)          # message from b200vo_last_error
    return e.value.code


def test_slot_and_range_checks(built, ctx):
    img = np.zeros((1, H, W), np.uint8)
    assert _code(built, ctx.upload_frames, img, 4) == built.ERR_BAD_ARG            # slot out of range
    assert _code(built, ctx.upload_frames, np.zeros((5, H, W), np.uint8), 0) == built.ERR_BAD_ARG   # more frames than slots
    assert _code(built, ctx.detect, 0, 5) == built.ERR_BAD_ARG
    assert _code(built, ctx.detect, -1, 1) == built.ERR_BAD_ARG
    assert _code(built, ctx.match, 0, 9) == built.ERR_BAD_ARG
    assert _code(built, ctx.process_pairs, 0, 4) == built.ERR_BAD_ARG              # npairs must be <= max_frames - 1
    assert _code(built, ctx.get_pyramid_level, 0, 8) == built.ERR_BAD_ARG


def test_pitch_and_channel_checks(built, ctx):
    buf = np.zeros(H * W * 4, np.uint8)
    lib = ctx.lib
    p = buf.ctypes.data_as(C.c_void_p)
    assert lib.b200vo_upload_frames(ctx.h, p, 1, C.c_size_t(W - 1), C.c_size_t(W * H), 1, 0) == built.ERR_BAD_ARG
    assert b"pitch" in lib.b200vo_last_error()
    assert lib.b200vo_upload_frames(ctx.h, p, 1, C.c_size_t(W * 2), C.c_size_t(W * 2 * H), 2, 0) == built.ERR_BAD_ARG
    assert lib.b200vo_upload_frames(ctx.h, None, 1, C.c_size_t(W), C.c_size_t(W * H), 1, 0) == built.ERR_BAD_ARG


def test_capacity_errors_leave_the_context_usable(built, ctx):
    from b200vo import synth
    frames = synth.make_sequence(W, H, 2)[0].numpy()
    ctx.upload_frames(frames, 0)
    ctx.detect(0, 2)
    kp, desc = ctx.get_features(0)
    assert len(kp) > 50
    small = np.zeros(10, built.KEYPOINT_DTYPE)
    n = C.c_int()
    rc = ctx.lib.b200vo_get_features(ctx.h, 0, small.ctypes.data_as(C.c_void_p), None, 10, C.byref(n))
    assert rc == built.ERR_CAPACITY and b"capacity" in ctx.lib.b200vo_last_error()
    rc = ctx.lib.b200vo_get_features(ctx.h, 0, None, None, 0, C.byref(n))         # count query needs no buffers
    assert rc == 0 and n.value == len(kp)
    out = np.zeros(2, built.DMATCH_DTYPE)
    g, c, thr = C.c_int(), C.c_int(), C.c_double()
    rc = ctx.lib.b200vo_match(ctx.h, 0, 1, out.ctypes.data_as(C.c_void_p), 2, C.byref(g), C.byref(c), C.byref(thr))
    assert rc == built.ERR_CAPACITY
    assert _code(built, ctx.relative_pose, 0, 1, 0, None, 65) == built.ERR_CAPACITY   # more hypotheses than the context was built for
    good, n_cross, _ = ctx.match(0, 1)                                                 # the context still works afterwards
    assert len(good) > 0 and n_cross >= len(good)
    assert _code(built, ctx.set_features, 2, None, np.zeros((ctx.kp_cap + 1, 32), np.uint8)) == built.ERR_CAPACITY


def test_soft_failures_are_status_bits_not_errors(built, ctx):
    """A slot with no features, and a pair with fewer than 10 good matches, are not errors: the record
    carries status 0 / identity and chaining leaves the pose unchanged (reference fallbacks (a), (b))."""
    ctx.upload_frames(np.full((2, H, W), 77, np.uint8), 2)
    ctx.detect(2, 2)
    assert len(ctx.get_features(2)[0]) == 0
    good, n_cross, thr = ctx.match(2, 3)
    assert len(good) == 0 and n_cross == 0
    rec = ctx.process_pairs(2, 1, 0)[0]
    assert rec["status"] == 0 and rec["n_matches"] == 0 and rec["n_inliers"] == 0
    R, t = built.chain_pose(rec, np.eye(3), np.array([1.0, 2.0, 3.0]))
    assert np.array_equal(R, np.eye(3)) and np.array_equal(t, [1.0, 2.0, 3.0])
    # descriptors present on one side only
    rng = np.random.default_rng(3)
    ctx.set_features(3, None, rng.integers(0, 256, (40, 32), dtype=np.uint8))
    good, n_cross, _ = ctx.match(2, 3)
    assert len(good) == 0 and n_cross == 0
    rec = ctx.relative_pose(2, 3)[0]
    assert rec["status"] == 0


def test_raw_matcher_argument_checks(built, ctx):
    q = np.zeros((4, 32), np.uint8)
    assert _code(built, ctx.match_descriptors, q, q, 0) == built.ERR_BAD_ARG
    x = np.zeros((4, 2))
    assert _code(built, ctx.five_point, x, x, np.zeros((1, 5), np.int32)) == built.ERR_BAD_ARG        # fewer than 5 points
    x = np.zeros((8, 2))
    assert _code(built, ctx.five_point, x, x, np.zeros((65, 5), np.int32)) == built.ERR_CAPACITY
