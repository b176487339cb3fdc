"""Trajectory export and accuracy report -- the headless replacement for what the reference shows in
its Pangolin viewer and prints per frame.

  * save_tum / load_tum      : "timestamp tx ty tz qx qy qz qw" (T_wc), one line per frame; the
                               reference only saves screenshots (src/trajectory_viewer.cpp:155-233)
  * align_similarity, ate_rmse, rpe : the reference applies a FIXED translation scale of 0.3 per
                               successful step (src/visual_odometry.cpp:352), so trajectories are
                               compared after a similarity (Umeyama) alignment
  * epipolar_residual        : mean |x2^T F x1| over the inliers, the self-check the reference prints
                               per frame (src/visual_odometry.cpp:248-263)
Host-side numpy only; nothing here is on the timed path.
"""
import numpy as np


def poses_to_Rt(poses12):
    """(n,12) rows [R_wc row-major | t_wc] -> (n,3,3), (n,3)."""
    p = np.asarray(poses12, np.float64).reshape(-1, 12)
    return p[:, :9].reshape(-1, 3, 3), p[:, 9:]


def _quat_from_R(R):
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    return q / np.linalg.norm(q)


def _R_from_quat(q):
    x, y, z, w = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def save_tum(path, R_wc, t_wc, timestamps=None):
    R_wc = np.asarray(R_wc, np.float64); t_wc = np.asarray(t_wc, np.float64)
    ts = np.arange(len(R_wc), dtype=np.float64) if timestamps is None else np.asarray(timestamps, np.float64)
    with open(path, "w") as f:
        f.write("# timestamp tx ty tz qx qy qz qw   (T_wc: camera -> world, as Pose in include/pose.h)\n")
        for s, R, t in zip(ts, R_wc, t_wc):
            q = _quat_from_R(R)
            f.write("%.9f %.9f %.9f %.9f %.9f %.9f %.9f %.9f\n" % (s, t[0], t[1], t[2], q[0], q[1], q[2], q[3]))


def load_tum(path):
    rows = [l.split() for l in open(path) if l.strip() and not l.startswith("#")]
    a = np.array(rows, np.float64).reshape(-1, 8)
    return a[:, 0], np.stack([_R_from_quat(q) for q in a[:, 4:8]]) if len(a) else np.zeros((0, 3, 3)), a[:, 1:4]


def align_similarity(est, gt):
    """Umeyama: s, R, t minimising sum |gt - (s R est + t)|^2.  est, gt: (n,3)."""
    est = np.asarray(est, np.float64); gt = np.asarray(gt, np.float64)
    me, mg = est.mean(0), gt.mean(0)
    E, Gm = est - me, gt - mg
    U, D, Vt = np.linalg.svd(Gm.T @ E / len(est))
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2, 2] = -1
    R = U @ S @ Vt
    var = (E ** 2).sum() / len(est)
    s = float(np.trace(np.diag(D) @ S) / var) if var > 0 else 1.0
    return s, R, mg - s * R @ me


def ate_rmse(est_positions, gt_positions):
    """Absolute trajectory error (RMSE, in ground-truth units) after similarity alignment."""
    s, R, t = align_similarity(est_positions, gt_positions)
    d = np.asarray(gt_positions) - (s * (R @ np.asarray(est_positions).T).T + t)
    return float(np.sqrt((d ** 2).sum(1).mean())), s


def rotation_angle_deg(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1.0) / 2.0
    return float(np.degrees(np.arccos(np.clip(c, -1.0, 1.0))))


def rpe(R_est, t_est, R_gt, t_gt):
    """Relative pose error between consecutive frames: mean rotation error (deg) and mean angle
    between the translation DIRECTIONS (deg; monocular scale is not observable), over the steps
    where both trajectories actually move."""
    rot, tdir = [], []
    for i in range(len(R_est) - 1):
        dRe = R_est[i].T @ R_est[i + 1]; dRg = R_gt[i].T @ R_gt[i + 1]
        rot.append(rotation_angle_deg(dRe, dRg))
        te = R_est[i].T @ (t_est[i + 1] - t_est[i]); tg = R_gt[i].T @ (t_gt[i + 1] - t_gt[i])
        ne, ng = np.linalg.norm(te), np.linalg.norm(tg)
        if ne > 1e-12 and ng > 1e-12:
            tdir.append(float(np.degrees(np.arccos(np.clip(te @ tg / (ne * ng), -1.0, 1.0)))))
    return (float(np.mean(rot)) if rot else 0.0), (float(np.mean(tdir)) if tdir else 0.0)


def accuracy_report(poses12, R_gt, C_gt):
    """Dict with ATE (similarity-aligned) and RPE of an estimated trajectory against ground truth."""
    R, t = poses_to_Rt(poses12)
    n = min(len(R), len(R_gt))
    ate, scale = ate_rmse(t[:n], np.asarray(C_gt)[:n])
    rr, rt = rpe(R[:n], t[:n], np.asarray(R_gt)[:n], np.asarray(C_gt)[:n])
    return {"frames": int(n), "ate_rmse_aligned": round(ate, 4), "alignment_scale": round(scale, 4),
            "rpe_rot_deg_mean": round(rr, 4), "rpe_trans_dir_deg_mean": round(rt, 4)}


def epipolar_residual(F, p1, p2, mask=None):
    """mean |x2^T F x1| over the (inlier) correspondences, pixel coordinates (visual_odometry.cpp:248-263)."""
    p1 = np.asarray(p1, np.float64).reshape(-1, 2); p2 = np.asarray(p2, np.float64).reshape(-1, 2)
    x1 = np.c_[p1, np.ones(len(p1))]; x2 = np.c_[p2, np.ones(len(p2))]
    r = np.abs(((x2 @ np.asarray(F, np.float64)) * x1).sum(1))
    if mask is not None:
        r = r[np.asarray(mask).astype(bool).ravel()]
    return float(r.mean()) if len(r) else 0.0
