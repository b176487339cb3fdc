"""The C-ABI boundary without a GPU: the library loads, exports every symbol include/b200vo.h
declares, struct layouts agree with the ctypes mirror, host-only entry points work, and device
entry points FAIL LOUDLY when no B200 is present (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import geometry as G
from oracle import pipeline as PL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "b200vo.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200vo_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported(built):
    lib = built.load_library()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(built.EXPORTS) == names
    assert lib.b200vo_abi_version() == 3


def test_struct_layouts(built):
    o = built.default_options(1241, 376)
    assert o.struct_size == ctypes.sizeof(built.Options)
    assert built.KEYPOINT_DTYPE.itemsize == 32 and built.DMATCH_DTYPE.itemsize == 16 and built.POSE_RECORD_DTYPE.itemsize == 160


def test_default_options_are_the_reference_constants(built):
    o = built.default_options(1241, 376)
    # /root/reference/src/visual_odometry.cpp:65,73,90-98,147,215,270-271,344
    assert (o.nfeatures, o.fast_threshold) == (1200, 10)
    assert (o.fx, o.fy, o.cx, o.cy) == (1241.0, 1241.0, 620.5, 188.0)
    assert (o.max_hamming, o.ransac_threshold_px) == (35.0, 2.0)
    assert (o.min_matches_for_pose, o.min_valid_points, o.min_inliers) == (10, 10, 12)
    with pytest.raises(built.B200VOError):
        built.default_options(0, 10)


def test_create_fails_loudly_without_gpu(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(built.B200VOError) as e:
        built.Context(640, 240)
    assert e.value.code == built.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_bad_options_rejected(built):
    lib = built.load_library()
    o = built.default_options(640, 240)
    o.struct_size = 8
    h = ctypes.c_void_p()
    assert lib.b200vo_create(ctypes.byref(o), 0, ctypes.byref(h)) == built.ERR_BAD_ARG
    assert b"struct_size" in lib.b200vo_last_error()
    for kw in (dict(width=32), dict(nfeatures=0), dict(nfeatures=30842), dict(max_frames=1), dict(ransac_hypotheses=0)):
        o = built.default_options(640, 240)
        for k, v in kw.items():
            setattr(o, k, v)
        assert lib.b200vo_create(ctypes.byref(o), 0, ctypes.byref(h)) == built.ERR_BAD_ARG


def test_host_only_entry_points(built):
    assert np.array_equal(built.draw_samples(42, 5, 100, 77), G.draw_samples(42, 5, 100, 77))
    with pytest.raises(built.B200VOError):
        built.draw_samples(1, 0, 4, 3)
    rng = np.random.default_rng(0)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3))); q *= np.sign(np.linalg.det(q))
    t = rng.normal(size=3); t /= np.linalg.norm(t)
    R0, _ = np.linalg.qr(rng.normal(size=(3, 3))); t0 = rng.normal(size=3)
    for status, scale in ((0, None), (1, None), (3, 0.0), (7, 0.3)):
        rec = np.zeros(1, built.POSE_RECORD_DTYPE)
        rec["R"] = q.ravel(); rec["t"] = t; rec["status"] = status
        R, tt = built.chain_pose(rec[0], R0, t0)
        Ro, to = PL.chain_pose(R0, t0, dict(status=status, R=q, t=t))
        assert np.allclose(R, Ro, atol=1e-15) and np.allclose(tt, to, atol=1e-15)
        if scale is None:
            assert np.array_equal(R, R0) and np.array_equal(tt, t0)
        else:
            Rg, tg = G.compose_pose(R0, t0, q, t, scale)
            assert np.allclose(R, Rg, atol=1e-15) and np.allclose(tt, tg, atol=1e-15)


def test_host_layer_library_loads(built):
    lib = ctypes.CDLL(built.HOST_LIB_PATH)
    for n in ("vo_host_create", "vo_host_destroy", "vo_host_process_frame", "vo_host_process_sequence",
              "vo_host_trajectory_size", "vo_host_last_F", "vo_host_last_error"):
        assert hasattr(lib, n)
    import torch
    if not torch.cuda.is_available():
        lib.vo_host_create.restype = ctypes.c_void_p
        lib.vo_host_last_error.restype = ctypes.c_char_p
        h = lib.vo_host_create(640, 240, 0, 0, ctypes.c_ulonglong(42), 0, 0)
        assert h is None and b"no CPU fallback" in lib.vo_host_last_error()


def test_product_sources_never_touch_the_oracle():
    """The oracle is test infrastructure: nothing under the package or include/ may reference it."""
    bad = []
    for base in ("slam-visual-odometry_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                    txt = open(os.path.join(d, f), errors="ignore").read()
                    if re.search(r"(from|import)\s+oracle|oracle/|import cv2|hostsim", txt) and "tests/hostsim" not in txt:
                        bad.append(os.path.join(d, f))
    assert not bad, bad
