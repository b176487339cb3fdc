"""ORACLE -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.  Never import this from the product path.

Python port of the reference's per-frame loop, VisualOdometry::process_frame
(/root/reference/src/visual_odometry.cpp:323-378), calling the SAME OpenCV functions the
reference calls, through Python cv2 (the reference's C++ cannot be built here: no OpenCV /
Pangolin / Eigen dev packages -- SURVEY.md section 8c):

  cv2.ORB_create(nf,1.2,8,31,0,2,HARRIS,31,10).detectAndCompute   visual_odometry.cpp:64-85,113
  cv2.BFMatcher(NORM_HAMMING, crossCheck=True).match              visual_odometry.cpp:87,153
  sort / median / min(max(3*min,0.7*median),35) / d < thr         visual_odometry.cpp:155-177
  cv2.findEssentialMat(p1,p2,K,RANSAC,0.999,2.0)                  visual_odometry.cpp:213-216
  cv2.recoverPose(E,p1,p2,K,mask)                                 visual_odometry.cpp:268
  fallbacks and pose chaining                                     visual_odometry.cpp:292-314,340-366

Used by bench.py's cpu_baseline / --impl reference legs and by the parity tests.
"""
import numpy as np

MIN_MATCHES_FOR_POSE = 10      # visual_odometry.cpp:344
MIN_VALID_POINTS = 10          # visual_odometry.cpp:270
MIN_INLIERS = 12               # visual_odometry.cpp:271
SCALE_GOOD = 0.3               # visual_odometry.cpp:352

# status bits shared with include/b200vo.h
ST_MATCHED = 1        # >= 10 good matches, pose estimation attempted
ST_E_FOUND = 2        # essential matrix found (R,t valid)
ST_POSE_OK = 4        # valid_points >= 10 and inliers >= 12  -> scale 0.3, else 0.0


def camera_matrix(width, height):
    """visual_odometry.cpp:90-98."""
    return np.array([[float(width), 0, width / 2.0], [0, float(width), height / 2.0], [0, 0, 1.0]])


class Cv2VisualOdometry:
    """The reference pipeline on cv2 (CPU)."""

    def __init__(self, width, height, nfeatures=1200):
        import cv2
        self.cv2 = cv2
        self.orb = cv2.ORB_create(nfeatures, 1.2, 8, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 10)
        self.bf = cv2.BFMatcher_create(cv2.NORM_HAMMING, True)
        self.K = camera_matrix(width, height)
        self.prev = None
        self.R_wc = np.eye(3)
        self.t_wc = np.zeros(3)
        self.trajectory = []
        self.last_F = None

    def detect(self, img):
        cv2 = self.cv2
        gray = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY) if img.ndim == 3 else img
        return self.orb.detectAndCompute(gray, None)

    def good_matches(self, d1, d2):
        if d1 is None or d2 is None or len(d1) == 0 or len(d2) == 0:
            return []
        m = sorted(self.bf.match(d1, d2), key=lambda x: x.distance)
        if not m:
            return m
        thr = min(max(3.0 * m[0].distance, 0.7 * m[len(m) // 2].distance), 35.0)
        return [x for x in m if x.distance < thr]

    def relative_pose(self, k1, k2, matches):
        """estimate_relative_pose (visual_odometry.cpp:182-281): (status, R, t, n_inl, n_valid)."""
        cv2 = self.cv2
        if len(matches) < 8:
            return 0, None, None, 0, 0
        p1 = np.float32([k1[m.queryIdx].pt for m in matches])
        p2 = np.float32([k2[m.trainIdx].pt for m in matches])
        E, mask = cv2.findEssentialMat(p1, p2, self.K, cv2.RANSAC, 0.999, 2.0)
        if E is None or E.size == 0:
            return 0, None, None, 0, 0
        E = E[:3]
        Kinv = np.linalg.inv(self.K)
        self.last_F = Kinv.T @ E @ Kinv
        n_inl = int(np.count_nonzero(mask))
        n_valid, R, t, _ = cv2.recoverPose(E, p1, p2, self.K, mask=mask)
        st = ST_E_FOUND | (ST_POSE_OK if (n_valid >= MIN_VALID_POINTS and n_inl >= MIN_INLIERS) else 0)
        return st, R, t.ravel(), n_inl, n_valid

    def process_frame(self, img):
        kp, desc = self.detect(img)
        rec = dict(status=0, R=np.eye(3), t=np.zeros(3), n_matches=0, n_inliers=0, n_valid=0)
        if self.prev is not None:
            pk, pd = self.prev
            good = self.good_matches(pd, desc)
            rec["n_matches"] = len(good)
            if len(good) >= MIN_MATCHES_FOR_POSE:
                st, R, t, ni, nv = self.relative_pose(pk, kp, good)
                rec.update(status=ST_MATCHED | st, n_inliers=ni, n_valid=nv)
                if R is not None:
                    rec.update(R=R, t=t)
            self.R_wc, self.t_wc = chain_pose(self.R_wc, self.t_wc, rec)
        self.trajectory.append((self.R_wc.copy(), self.t_wc.copy()))
        self.prev = (kp, desc)
        return rec


def chain_pose(R_wc, t_wc, rec):
    """process_frame's pose update (visual_odometry.cpp:340-366) from one relative-pose record."""
    if not (rec["status"] & ST_E_FOUND):
        return R_wc, t_wc                          # fallback (a)/(b): pose unchanged
    scale = SCALE_GOOD if (rec["status"] & ST_POSE_OK) else 0.0     # fallback (c): rotation only
    R = np.asarray(rec["R"], float).reshape(3, 3)
    t = np.asarray(rec["t"], float).reshape(3)
    R_inv = R.T
    t_inv = -R.T @ t
    return R_wc @ R_inv, t_wc + scale * (R_wc @ t_inv)


def run_pairs_cv2(frames, nfeatures, width=None, height=None):
    """Run the reference pipeline over a list/array of frames; returns per-pair records."""
    h, w = frames[0].shape[:2]
    vo = Cv2VisualOdometry(width or w, height or h, nfeatures)
    recs = []
    for i, f in enumerate(frames):
        r = vo.process_frame(np.ascontiguousarray(f))
        if i > 0:
            recs.append(r)
    return recs, vo.trajectory
