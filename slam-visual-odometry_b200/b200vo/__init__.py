"""ctypes binding of the b200vo C ABI (include/b200vo.h) -- the Python face of the drop-in boundary.

Everything here forwards to libb200vo.so (hand-written sm_100a kernels).  There is no CPU
fallback: if the library is missing, or no B200 is present, construction raises.
Reference mapping (what each call replaces) is documented in include/b200vo.h.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG, "lib", "libb200vo.so")
HOST_LIB_PATH = os.path.join(_PKG, "lib", "libb200vo_host.so")

NLEVELS = 8
ST_MATCHED, ST_E_FOUND, ST_POSE_OK, ST_OVERFLOW = 1, 2, 4, 256
MATCH_CROSSCHECK, MATCH_RATIO = 0, 1
ERR_BAD_ARG, ERR_CAPACITY, ERR_CUDA, ERR_NOT_READY, ERR_ALLOC = -1, -2, -3, -4, -5


class Options(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("nfeatures", C.c_int32),
                ("fast_threshold", C.c_int32), ("max_frames", C.c_int32), ("ransac_hypotheses", C.c_int32),
                ("match_mode", C.c_int32), ("ratio", C.c_float), ("ransac_adaptive", C.c_int32),
                ("ransac_seed", C.c_uint64), ("ransac_threshold_px", C.c_double), ("max_hamming", C.c_double),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("min_matches_for_pose", C.c_int32), ("min_valid_points", C.c_int32), ("min_inliers", C.c_int32),
                ("reserved1", C.c_int32)]


KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                           ("octave", "<i4"), ("class_id", "<i4"), ("lx", "<i2"), ("ly", "<i2")])
DMATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")])
POSE_RECORD_DTYPE = np.dtype([("R", "<f8", (9,)), ("t", "<f8", (3,)), ("n_matches", "<i4"), ("n_inliers", "<i4"),
                              ("n_valid", "<i4"), ("status", "<i4"), ("n_cross", "<i4"), ("best_hyp", "<i4"),
                              ("best_sol", "<i4"), ("threshold", "<f4"),
                              ("dist_mean", "<f8"), ("epi_residual", "<f8"), ("dist_min", "<f4"), ("dist_max", "<f4"),
                              ("dist_median", "<f4"), ("reserved", "<i4")])
RECORD_BYTES = POSE_RECORD_DTYPE.itemsize
ABI_VERSION = 3
assert KEYPOINT_DTYPE.itemsize == 32 and DMATCH_DTYPE.itemsize == 16 and RECORD_BYTES == 160


class StageTimes(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("upload_ms", "gray_ms", "pyramid_ms", "fast_ms", "select_ms", "orient_ms",
                                         "blur_ms", "brief_ms", "match_ms", "filter_ms", "ransac_ms", "pose_ms",
                                         "download_ms", "fivept_ms", "score_ms", "describe_ms")]


EXPORTS = [
    "b200vo_abi_version", "b200vo_last_error", "b200vo_default_options", "b200vo_create", "b200vo_destroy",
    "b200vo_synchronize", "b200vo_level_info", "b200vo_max_keypoints", "b200vo_upload_frames",
    "b200vo_upload_frames_device", "b200vo_detect", "b200vo_get_features", "b200vo_get_pyramid_level",
    "b200vo_get_fast_candidates", "b200vo_set_features", "b200vo_match_descriptors",
    "b200vo_match_descriptors_device", "b200vo_knn_match_descriptors", "b200vo_match", "b200vo_relative_pose", "b200vo_process_frame", "b200vo_process_frame_begin", "b200vo_process_frame_end", "b200vo_process_pairs",
    "b200vo_process_pairs_device", "b200vo_five_point", "b200vo_score_models", "b200vo_recover_pose",
    "b200vo_draw_samples", "b200vo_chain_pose", "b200vo_chain_trajectory", "b200vo_measure_popc_rate", "b200vo_fundamental_from_essential", "b200vo_enable_timing",
    "b200vo_get_stage_times", "b200vo_kernel_launches", "b200vo_stream",
    "b200vo_group_create", "b200vo_group_destroy", "b200vo_group_size", "b200vo_group_ctx", "b200vo_partition_pairs",
    "b200vo_group_process_sequence", "b200vo_group_allgather_records", "b200vo_nccl_version",
]

_lib = None


def load_library():
    """Load libb200vo.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s not found: build it first (python -c 'import __graft_entry__ as g; g.build()')" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.b200vo_last_error.restype = C.c_char_p
    lib.b200vo_kernel_launches.restype = C.c_int64
    lib.b200vo_kernel_launches.argtypes = [C.c_void_p]
    lib.b200vo_stream.restype = C.c_void_p
    lib.b200vo_stream.argtypes = [C.c_void_p]
    lib.b200vo_destroy.argtypes = [C.c_void_p]
    lib.b200vo_destroy.restype = None
    lib.b200vo_group_destroy.argtypes = [C.c_void_p]
    lib.b200vo_group_destroy.restype = None
    lib.b200vo_group_process_sequence.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int64, C.c_void_p]
    _lib = lib
    return lib


class B200VOError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200vo error %d: %s" % (code, msg))
        self.code = code


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def default_options(width, height, **kw):
    lib = load_library()
    o = Options()
    rc = lib.b200vo_default_options(C.byref(o), int(width), int(height))
    if rc:
        raise B200VOError(rc, lib.b200vo_last_error().decode())
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def draw_samples(seed, pair_id, nhyp, m):
    lib = load_library()
    out = np.zeros((nhyp, 5), np.int32)
    rc = lib.b200vo_draw_samples(C.c_uint64(seed), C.c_int64(pair_id), int(nhyp), int(m), _p(out))
    if rc:
        raise B200VOError(rc, lib.b200vo_last_error().decode())
    return out


def chain_pose(rec, R_wc, t_wc):
    """process_frame's pose update from one record (host arithmetic inside the library)."""
    lib = load_library()
    r = np.ascontiguousarray(np.asarray(rec).reshape(1), POSE_RECORD_DTYPE)
    Rin = np.ascontiguousarray(R_wc, np.float64).reshape(9)
    tin = np.ascontiguousarray(t_wc, np.float64).reshape(3)
    Ro = np.zeros(9); to = np.zeros(3)
    rc = lib.b200vo_chain_pose(_p(r), _p(Rin), _p(tin), _p(Ro), _p(to))
    if rc:
        raise B200VOError(rc, lib.b200vo_last_error().decode())
    return Ro.reshape(3, 3), to


def chain_trajectory(recs, R0=None, t0=None):
    """Poses (n,12) after applying each record in turn to (R0,t0) -- process_frame's accumulation."""
    lib = load_library()
    r = np.ascontiguousarray(recs, POSE_RECORD_DTYPE)
    R0 = np.eye(3) if R0 is None else R0
    t0 = np.zeros(3) if t0 is None else t0
    Rin = np.ascontiguousarray(R0, np.float64).reshape(9)
    tin = np.ascontiguousarray(t0, np.float64).reshape(3)
    out = np.zeros((len(r), 12))
    rc = lib.b200vo_chain_trajectory(_p(r), len(r), _p(Rin), _p(tin), _p(out))
    if rc:
        raise B200VOError(rc, lib.b200vo_last_error().decode())
    return out


class Context:
    """One b200vo context = one GPU.  Not thread-safe (mirrors the C ABI)."""

    def __init__(self, width, height, device=0, **options):
        self.lib = load_library()
        self.opt = default_options(width, height, **options)
        h = C.c_void_p()
        rc = self.lib.b200vo_create(C.byref(self.opt), int(device), C.byref(h))
        if rc:
            raise B200VOError(rc, self.lib.b200vo_last_error().decode())
        self.h = h
        self.width, self.height = int(width), int(height)
        self.kp_cap = self.lib.b200vo_max_keypoints(self.h)

    def _chk(self, rc):
        if rc:
            raise B200VOError(rc, self.lib.b200vo_last_error().decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200vo_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- geometry
    def level_info(self, level):
        w, h, n = C.c_int(), C.c_int(), C.c_int()
        s = C.c_float()
        self._chk(self.lib.b200vo_level_info(self.h, level, C.byref(w), C.byref(h), C.byref(s), C.byref(n)))
        return w.value, h.value, s.value, n.value

    # ---- frames
    def upload_frames(self, frames, first_slot=0):
        """frames: (n,H,W) u8 gray or (n,H,W,3) u8 BGR host array."""
        f = np.ascontiguousarray(frames, np.uint8)
        ch = 3 if f.ndim == 4 else 1
        n = f.shape[0]
        pitch = f.shape[2] * ch
        self._chk(self.lib.b200vo_upload_frames(self.h, _p(f), n, C.c_size_t(pitch), C.c_size_t(pitch * f.shape[1]), ch, first_slot))
        self.synchronize()        # `f` may be a temporary

    def upload_frames_ptr(self, host_ptr, n, pitch, frame_stride, channels, first_slot):
        """Asynchronous upload from caller-owned (ideally pinned) host memory."""
        self._chk(self.lib.b200vo_upload_frames(self.h, C.c_void_p(host_ptr), n, C.c_size_t(pitch), C.c_size_t(frame_stride), channels, first_slot))

    def upload_frames_device(self, dev_ptr, n, pitch, frame_stride, channels, first_slot):
        self._chk(self.lib.b200vo_upload_frames_device(self.h, C.c_void_p(dev_ptr), n, C.c_size_t(pitch), C.c_size_t(frame_stride), channels, first_slot))

    def detect(self, first_slot, n):
        self._chk(self.lib.b200vo_detect(self.h, first_slot, n))

    def get_features(self, slot):
        kp = np.zeros(self.kp_cap, KEYPOINT_DTYPE)
        desc = np.zeros((self.kp_cap, 32), np.uint8)
        n = C.c_int()
        self._chk(self.lib.b200vo_get_features(self.h, slot, _p(kp), _p(desc), self.kp_cap, C.byref(n)))
        return kp[:n.value].copy(), desc[:n.value].copy()

    def feature_count(self, slot):
        """Number of keypoints of a slot (no copy of the features)."""
        n = C.c_int()
        self._chk(self.lib.b200vo_get_features(self.h, slot, None, None, 0, C.byref(n)))
        return n.value

    def set_features(self, slot, kp, desc):
        desc = np.ascontiguousarray(desc, np.uint8)
        kp = None if kp is None else np.ascontiguousarray(kp, KEYPOINT_DTYPE)
        self._chk(self.lib.b200vo_set_features(self.h, slot, _p(kp), _p(desc), len(desc)))

    def get_pyramid_level(self, slot, level, blurred=False):
        w, h, _, _ = self.level_info(level)
        out = np.zeros((h, w), np.uint8)
        self._chk(self.lib.b200vo_get_pyramid_level(self.h, slot, level, int(blurred), _p(out), C.c_size_t(w)))
        return out

    def get_fast_candidates(self, slot, level):
        n = C.c_int()
        self._chk(self.lib.b200vo_get_fast_candidates(self.h, slot, level, None, 0, C.byref(n)))
        out = np.zeros(max(n.value, 1), np.uint32)
        self._chk(self.lib.b200vo_get_fast_candidates(self.h, slot, level, _p(out), len(out), C.byref(n)))
        out = np.sort(out[:n.value])
        return ((out >> 8) & 0xFFF).astype(np.int32), (out >> 20).astype(np.int32), (out & 0xFF).astype(np.int32)

    # ---- matching
    def match_descriptors(self, q, t, npairs=1):
        q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
        n1 = q.shape[-2] if q.size else 0
        n2 = t.shape[-2] if t.size else 0
        nq = np.zeros((npairs, n1), np.int32); dq = np.zeros((npairs, n1), np.int32)
        nt = np.zeros((npairs, n2), np.int32); dt = np.zeros((npairs, n2), np.int32)
        if q.size == 0:
            q = np.zeros(1, np.uint8)
        if t.size == 0:
            t = np.zeros(1, np.uint8)
        self._chk(self.lib.b200vo_match_descriptors(self.h, _p(q), n1, _p(t), n2, npairs, _p(nq), _p(dq), _p(nt), _p(dt)))
        return nq, dq, nt, dt

    def knn2_match_descriptors(self, q, t, npairs=1):
        """BFMatcher(NORM_HAMMING, False).knnMatch(k=2): (idx, dist), each (npairs, n1, 2); -1 = no neighbour."""
        q = np.ascontiguousarray(q, np.uint8); t = np.ascontiguousarray(t, np.uint8)
        n1 = q.shape[-2] if q.size else 0
        n2 = t.shape[-2] if t.size else 0
        idx = np.zeros((npairs, n1, 2), np.int32); dist = np.zeros((npairs, n1, 2), np.int32)
        if n1 == 0:
            return idx, dist
        if t.size == 0:
            t = np.zeros(1, np.uint8)
        self._chk(self.lib.b200vo_knn_match_descriptors(self.h, _p(q), n1, _p(t), n2, npairs, _p(idx), _p(dist)))
        return idx, dist

    def match_descriptors_device(self, q_ptr, n1, t_ptr, n2, npairs, nnq_ptr, nnt_ptr):
        ms = C.c_float()
        self._chk(self.lib.b200vo_match_descriptors_device(self.h, C.c_void_p(q_ptr), n1, C.c_void_p(t_ptr), n2, npairs,
                                                            C.c_void_p(nnq_ptr), C.c_void_p(nnt_ptr), C.byref(ms)))
        return ms.value

    def match(self, slot1, slot2):
        out = np.zeros(self.kp_cap, DMATCH_DTYPE)
        g, c = C.c_int(), C.c_int()
        thr = C.c_double()
        self._chk(self.lib.b200vo_match(self.h, slot1, slot2, _p(out), self.kp_cap, C.byref(g), C.byref(c), C.byref(thr)))
        return out[:g.value].copy(), c.value, thr.value

    # ---- geometry
    def relative_pose(self, slot1, slot2, pair_id=0, samples=None, nhyp=0, want_mask=False):
        rec = np.zeros(1, POSE_RECORD_DTYPE)
        E = np.zeros(9)
        mask = np.zeros(self.kp_cap, np.uint8) if want_mask else None
        if samples is not None:
            samples = np.ascontiguousarray(samples, np.int32)
            nhyp = len(samples)
        self._chk(self.lib.b200vo_relative_pose(self.h, slot1, slot2, C.c_int64(pair_id), _p(samples), int(nhyp), _p(rec), _p(E),
                                                _p(mask), self.kp_cap if want_mask else 0))
        r = rec[0]
        return r, E.reshape(3, 3), (mask[:r["n_matches"]].copy() if want_mask else None)

    def process_frame(self, image, slot, prev_slot=-1, pair_id=0):
        """One streaming step (b200vo_process_frame): returns (keypoints, descriptors, record or None, E or None)."""
        im = np.ascontiguousarray(image, np.uint8)
        ch = 3 if im.ndim == 3 else 1
        kp = np.zeros(self.kp_cap, KEYPOINT_DTYPE)
        desc = np.zeros((self.kp_cap, 32), np.uint8)
        n = C.c_int()
        rec = np.zeros(1, POSE_RECORD_DTYPE)
        E = np.zeros(9)
        self._chk(self.lib.b200vo_process_frame(self.h, _p(im), C.c_size_t(im.shape[1] * ch), ch, int(slot), int(prev_slot), C.c_int64(pair_id),
                                                _p(kp), _p(desc), self.kp_cap, C.byref(n), _p(rec), _p(E)))
        if prev_slot < 0:
            return kp[:n.value].copy(), desc[:n.value].copy(), None, None
        return kp[:n.value].copy(), desc[:n.value].copy(), rec[0], E.reshape(3, 3)

    def process_pairs(self, first_slot, npairs, first_pair_id=0):
        out = np.zeros(npairs, POSE_RECORD_DTYPE)
        self._chk(self.lib.b200vo_process_pairs(self.h, first_slot, npairs, C.c_int64(first_pair_id), _p(out)))
        return out

    def process_pairs_device(self, first_slot, npairs, first_pair_id=0):
        ptr = C.c_void_p()
        self._chk(self.lib.b200vo_process_pairs_device(self.h, first_slot, npairs, C.c_int64(first_pair_id), C.byref(ptr)))
        return ptr.value

    def five_point(self, x1, x2, samples):
        x1 = np.ascontiguousarray(x1, np.float64); x2 = np.ascontiguousarray(x2, np.float64)
        samples = np.ascontiguousarray(samples, np.int32)
        nh = len(samples)
        E = np.zeros((nh, 10, 9)); ns = np.zeros(nh, np.int32)
        self._chk(self.lib.b200vo_five_point(self.h, _p(x1), _p(x2), len(x1), _p(samples), nh, _p(E), _p(ns)))
        return E, ns

    def score_models(self, x1, x2, models, threshold):
        x1 = np.ascontiguousarray(x1, np.float64); x2 = np.ascontiguousarray(x2, np.float64)
        models = np.ascontiguousarray(models, np.float64).reshape(-1, 9)
        counts = np.zeros(len(models), np.int32)
        self._chk(self.lib.b200vo_score_models(self.h, _p(x1), _p(x2), len(x1), _p(models), len(models), C.c_double(threshold), _p(counts)))
        return counts

    def recover_pose(self, E, x1, x2, mask_in=None):
        x1 = np.ascontiguousarray(x1, np.float64); x2 = np.ascontiguousarray(x2, np.float64)
        E = np.ascontiguousarray(E, np.float64).reshape(9)
        R = np.zeros(9); t = np.zeros(3); nv = C.c_int32()
        mo = np.zeros(len(x1), np.uint8)
        mi = None if mask_in is None else np.ascontiguousarray(mask_in, np.uint8)
        self._chk(self.lib.b200vo_recover_pose(self.h, _p(E), _p(x1), _p(x2), len(x1), _p(mi), _p(R), _p(t), C.byref(nv), _p(mo)))
        return nv.value, R.reshape(3, 3), t, mo

    def fundamental_from_essential(self, E):
        E = np.ascontiguousarray(E, np.float64).reshape(9)
        F = np.zeros(9)
        self._chk(self.lib.b200vo_fundamental_from_essential(self.h, _p(E), _p(F)))
        return F.reshape(3, 3)

    # ---- misc
    def synchronize(self):
        self._chk(self.lib.b200vo_synchronize(self.h))

    def enable_timing(self, on=True):
        self._chk(self.lib.b200vo_enable_timing(self.h, int(on)))

    def stage_times(self):
        st = StageTimes()
        self._chk(self.lib.b200vo_get_stage_times(self.h, C.byref(st)))
        return {n: getattr(st, n) for n, _ in StageTimes._fields_}

    def measure_popc_rate(self):
        v = C.c_double()
        self._chk(self.lib.b200vo_measure_popc_rate(self.h, C.byref(v)))
        return v.value

    def kernel_launches(self):
        return int(self.lib.b200vo_kernel_launches(self.h))

    def stream(self):
        return self.lib.b200vo_stream(self.h)


class Group:
    """A device group (b200vo_group): one sequence partitioned over several GPUs of this node, records
    all-gathered with NCCL inside the library."""

    def __init__(self, width, height, devices, **options):
        self.lib = load_library()
        self.opt = default_options(width, height, **options)
        dev = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        rc = self.lib.b200vo_group_create(C.byref(self.opt), dev, len(devices), C.byref(h))
        if rc:
            raise B200VOError(rc, self.lib.b200vo_last_error().decode())
        self.h = h
        self.width, self.height = int(width), int(height)

    def process_sequence(self, frames, batch_frames, first_pair_id=0):
        f = np.ascontiguousarray(frames, np.uint8)
        ch = 3 if f.ndim == 4 else 1
        pitch = f.shape[2] * ch
        out = np.zeros(len(f) - 1, POSE_RECORD_DTYPE)
        rc = self.lib.b200vo_group_process_sequence(self.h, _p(f), len(f), C.c_size_t(pitch), C.c_size_t(pitch * f.shape[1]), ch,
                                                    int(batch_frames), C.c_int64(first_pair_id), _p(out))
        if rc:
            raise B200VOError(rc, self.lib.b200vo_last_error().decode())
        return out

    def close(self):
        if getattr(self, "h", None):
            self.lib.b200vo_group_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def partition_pairs(total_pairs, world, rank):
    lib = load_library()
    lo, hi = C.c_int(), C.c_int()
    rc = lib.b200vo_partition_pairs(int(total_pairs), int(world), int(rank), C.byref(lo), C.byref(hi))
    if rc:
        raise B200VOError(rc, lib.b200vo_last_error().decode())
    return lo.value, hi.value
