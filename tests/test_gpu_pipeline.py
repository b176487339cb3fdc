"""GPU parity of the whole per-frame-pair path and of the reference-shaped C++ API.

* batched b200vo_process_pairs == per-pair b200vo_relative_pose (batching must not change results);
* results do not depend on which ring slots / launch batch a pair is processed in (the property the
  multi-GPU partition relies on: bit-identical relative poses);
* the C++ VisualOdometry class (include/visual_odometry.h) reproduces process_frame's semantics:
  first frame -> identity, trajectory grows by one per frame, the caller's Frame is moved-from,
  poses equal the chained records, process_sequence == repeated process_frame;
* accuracy against the synthetic ground truth and agreement with the cv2 reference pipeline."""
import ctypes

import numpy as np
import pytest

from oracle import geometry as G
from oracle import pipeline as PL

pytestmark = pytest.mark.gpu

W, H, NF, NHYP, NFRAMES = 1241, 376, 2000, 512, 7


@pytest.fixture(scope="module")
def seq(built):
    from b200vo import synth
    frames, R_wc, C_w, K = synth.make_sequence(W, H, NFRAMES, device="cuda")
    return frames.cpu().numpy(), R_wc, C_w, K


@pytest.fixture(scope="module")
def records(built, seq):
    frames = seq[0]
    with built.Context(W, H, nfeatures=NF, max_frames=NFRAMES + 1, ransac_hypotheses=NHYP) as ctx:
        ctx.upload_frames(frames, 0)
        ctx.detect(0, NFRAMES)
        batched = ctx.process_pairs(0, NFRAMES - 1, 0)
        single = [ctx.relative_pose(i, i + 1, pair_id=i)[0] for i in range(NFRAMES - 1)]
    return batched, single


def test_batched_equals_single(records):
    batched, single = records
    for i, r in enumerate(single):
        assert batched[i].tobytes() == r.tobytes(), "pair %d differs between batched and single launch" % i


def test_partition_invariance(built, seq, records):
    """Process pairs 3.. in a fresh context with a small ring and different batch boundaries."""
    frames = seq[0]
    batched, _ = records
    with built.Context(W, H, nfeatures=NF, max_frames=3, ransac_hypotheses=NHYP) as ctx:
        ctx.upload_frames(frames[3:4], 2)
        ctx.detect(2, 1)
        out = []
        prev = 2
        for i in range(4, NFRAMES):
            cur = (prev + 1) % 3
            ctx.upload_frames(frames[i:i + 1], cur)
            ctx.detect(cur, 1)
            out.append(ctx.process_pairs(prev, 1, i - 1)[0])
            prev = cur
    for k, r in enumerate(out):
        assert r.tobytes() == batched[3 + k].tobytes(), "pair %d depends on partition" % (3 + k)


def test_accuracy_vs_ground_truth_and_cv2(built, seq, records):
    from b200vo import synth
    frames, R_wc, C_w, K = seq
    batched, _ = records
    cv_recs, _ = PL.run_pairs_cv2(list(frames), NF)
    for i, r in enumerate(batched):
        assert r["status"] & built.ST_E_FOUND
        Rg, tg = synth.relative_pose_gt(R_wc, C_w, i)
        R = r["R"].reshape(3, 3)
        assert abs(np.linalg.det(R) - 1) < 1e-9 and abs(np.linalg.norm(r["t"]) - 1) < 1e-9
        assert G.rotation_angle_deg(R, Rg) < 1.0, "pair %d rotation error vs GT" % i
        # same correspondences as the cv2 pipeline (identical good-match count).  cv2 draws its own
        # hypotheses (parity is by tolerance): both results are minimal-sample models
        # within a degree of the ground truth, hence within two degrees of each other
        assert r["n_matches"] == cv_recs[i]["n_matches"]
        assert G.rotation_angle_deg(cv_recs[i]["R"], Rg) < 1.0, "pair %d: cv2 rotation error vs GT" % i
        assert G.rotation_angle_deg(R, cv_recs[i]["R"]) < 2.0


def _host(built):
    lib = ctypes.CDLL(built.HOST_LIB_PATH)
    lib.vo_host_create.restype = ctypes.c_void_p
    lib.vo_host_last_error.restype = ctypes.c_char_p
    lib.vo_host_destroy.argtypes = [ctypes.c_void_p]
    return lib


def test_cpp_visual_odometry_process_frame(built, seq, records):
    frames = seq[0]
    batched, _ = records
    lib = _host(built)
    h = ctypes.c_void_p(lib.vo_host_create(W, H, NF, NHYP, ctypes.c_ulonglong(42), 0, 8))
    assert h.value, lib.vo_host_last_error()
    poses = []
    for i in range(NFRAMES):
        pose = np.zeros(12)
        moved = ctypes.c_int(-1)
        f = np.ascontiguousarray(frames[i])
        n = lib.vo_host_process_frame(h, i, f.ctypes.data_as(ctypes.c_void_p), W, H, 1, pose.ctypes.data_as(ctypes.c_void_p), ctypes.byref(moved))
        assert n == i + 1, lib.vo_host_last_error()         # trajectory grows by one per frame
        assert moved.value == 1                              # frame moved into previous_frame_ (reference :334,:376)
        poses.append(pose.copy())
    assert np.array_equal(poses[0], np.r_[np.eye(3).ravel(), 0, 0, 0])
    traj = built.chain_trajectory(batched)
    assert np.allclose(np.array(poses[1:]), traj, atol=1e-15)
    F = np.zeros(9)
    assert lib.vo_host_last_F(h, F.ctypes.data_as(ctypes.c_void_p)) == 1
    lib.vo_host_destroy(h)


def test_cpp_process_sequence_equals_process_frame(built, seq, records):
    frames = seq[0]
    batched, _ = records
    lib = _host(built)
    h = ctypes.c_void_p(lib.vo_host_create(W, H, NF, NHYP, ctypes.c_ulonglong(42), 0, 3))     # batches of 3: uneven chunks
    out = np.zeros((NFRAMES, 12))
    f = np.ascontiguousarray(frames)
    n = lib.vo_host_process_sequence(h, f.ctypes.data_as(ctypes.c_void_p), NFRAMES, W, H, 1, out.ctypes.data_as(ctypes.c_void_p))
    assert n == NFRAMES, lib.vo_host_last_error()
    assert lib.vo_host_trajectory_size(h) == NFRAMES
    traj = built.chain_trajectory(batched)
    assert np.array_equal(out[0], np.r_[np.eye(3).ravel(), 0, 0, 0]) and np.allclose(out[1:], traj, atol=1e-15)
    lib.vo_host_destroy(h)


def test_python_reference_chain_equals_library(built, records):
    batched, _ = records
    R, t = np.eye(3), np.zeros(3)
    traj = built.chain_trajectory(batched)
    for i, r in enumerate(batched):
        R, t = PL.chain_pose(R, t, dict(status=int(r["status"]), R=r["R"], t=r["t"]))
        assert np.allclose(traj[i][:9].reshape(3, 3), R, atol=1e-14) and np.allclose(traj[i][9:], t, atol=1e-14)


def test_flat_frame_in_sequence_takes_reference_fallback(built, seq, records):
    """A frame without texture yields no keypoints: both pairs touching it have < 10 good matches, the
    pose stays unchanged (reference fallback (a), visual_odometry.cpp:340-345), the other pairs are
    unaffected (byte-identical to the undisturbed run)."""
    frames = seq[0].copy()
    frames[3] = 128
    batched, _ = records
    with built.Context(W, H, nfeatures=NF, max_frames=NFRAMES + 1, ransac_hypotheses=NHYP) as ctx:
        ctx.upload_frames(frames, 0)
        ctx.detect(0, NFRAMES)
        recs = ctx.process_pairs(0, NFRAMES - 1, 0)
    for i in (2, 3):
        assert recs[i]["status"] == 0 and recs[i]["n_matches"] == 0
    for i in (0, 1, 4, 5):
        assert recs[i].tobytes() == batched[i].tobytes()
    poses = built.chain_trajectory(recs)
    assert np.array_equal(poses[2], poses[1]) and np.array_equal(poses[3], poses[1])      # frames 3 and 4 keep frame 2's pose


def test_ratio_mode_pipeline(built, seq):
    """The whole pair path with the ratio-test matcher: its good-match list equals the oracle's on the
    GPU's own descriptors, and the poses are as accurate as in cross-check mode."""
    from oracle import match as M
    frames, R_wc, C_w, K = seq
    with built.Context(W, H, nfeatures=NF, max_frames=NFRAMES + 1, ransac_hypotheses=NHYP,
                       match_mode=built.MATCH_RATIO, ratio=0.8) as ctx:
        ctx.upload_frames(frames, 0)
        ctx.detect(0, NFRAMES)
        recs = ctx.process_pairs(0, NFRAMES - 1, 0)
        d0, d1 = ctx.get_features(0)[1], ctx.get_features(1)[1]
        good, n_pass, thr = ctx.match(0, 1)
    qi, ti, d, thr_o, nc = M.good_matches(d0, d1, ratio=0.8)
    assert n_pass == nc and thr == thr_o and np.array_equal(good["queryIdx"], qi) and np.array_equal(good["trainIdx"], ti)
    assert recs[0]["n_matches"] == len(qi)
    from b200vo import synth
    for i, r in enumerate(recs):
        assert r["status"] & 4, "pair %d has no valid pose in ratio mode" % i
        Rg, tg = synth.relative_pose_gt(R_wc, C_w, i)
        # minimal-sample model without refit on a differently composed match list: 2 degrees
        assert G.rotation_angle_deg(r["R"].reshape(3, 3), Rg) < 2.0


def test_streaming_entry_point_equals_separate_calls(built, seq, records):
    """b200vo_process_frame (one synchronisation) returns what upload + detect + get_features +
    relative_pose return."""
    frames = seq[0]
    batched, _ = records
    with built.Context(W, H, nfeatures=NF, max_frames=2, ransac_hypotheses=NHYP) as ctx, \
         built.Context(W, H, nfeatures=NF, max_frames=2, ransac_hypotheses=NHYP) as ref:
        prev = -1
        for i in range(4):
            cur = i & 1
            kp, desc, rec, E = ctx.process_frame(frames[i], cur, prev, i - 1)
            ref.upload_frames(frames[i:i + 1], cur)
            ref.detect(cur, 1)
            rkp, rdesc = ref.get_features(cur)
            assert kp.tobytes() == rkp.tobytes() and np.array_equal(desc, rdesc)
            if i == 0:
                assert rec is None
            else:
                rrec, rE, _ = ref.relative_pose(prev, cur, pair_id=i - 1)
                assert rec.tobytes() == rrec.tobytes() == batched[i - 1].tobytes() and np.array_equal(E, rE)
            prev = cur
