#!/usr/bin/env python
"""Per-frame streaming throughput of the drop-in C++ API: VisualOdometry::process_frame called once per
frame exactly as the reference's main loop does (src/main.cpp:128-155), through the ctypes shim.
   python tools/bench_streaming.py [--frames 200] [--features 1200]"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-visual-odometry_b200"))
import b200vo  # noqa: E402
from b200vo import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=200)
ap.add_argument("--features", type=int, default=1200)      # the reference's hard-coded value (visual_odometry.cpp:65)
a = ap.parse_args()
W, H = 1241, 376
tex = synth.make_texture()
R, C = synth.make_trajectory(a.frames)
frames = synth.render_frames(W, H, R, C, tex, device="cuda:0", chunk=16).cpu().numpy()
lib = ctypes.CDLL(b200vo.HOST_LIB_PATH)
lib.vo_host_create.restype = ctypes.c_void_p
lib.vo_host_last_error.restype = ctypes.c_char_p
h = lib.vo_host_create(W, H, a.features, 0, ctypes.c_ulonglong(42), 0, 0)
assert h, lib.vo_host_last_error()
h = ctypes.c_void_p(h)
pose = np.zeros(12)
moved = ctypes.c_int()
t = []
for i, f in enumerate(frames):
    f = np.ascontiguousarray(f)
    t0 = time.perf_counter()
    rc = lib.vo_host_process_frame(h, i, f.ctypes.data_as(ctypes.c_void_p), W, H, 1, pose.ctypes.data_as(ctypes.c_void_p), ctypes.byref(moved))
    t.append(time.perf_counter() - t0)
    assert rc > 0, lib.vo_host_last_error()
lib.vo_host_destroy(h)
t = np.array(t[10:])
print(json.dumps({"api": "VisualOdometry::process_frame (per-frame streaming, host image in, keypoints/descriptors/pose out)",
                  "frames": len(t), "features": a.features, "frames_per_s": round(1.0 / t.mean(), 1), "ms_per_frame_mean": round(t.mean() * 1e3, 3),
                  "ms_per_frame_p50": round(float(np.median(t)) * 1e3, 3), "ms_per_frame_p99": round(float(np.percentile(t, 99)) * 1e3, 3),
                  "end_position": pose[9:].round(3).tolist()}))
