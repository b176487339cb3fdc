#!/usr/bin/env python
"""In-library multi-GPU path on every visible GPU: b200vo_group_process_sequence (one host thread per device, NCCL
all-gather inside the library) against a single context, byte for byte, plus its throughput at config 3's frame shape.
   python tools/group_check.py [--frames 129] [--width 1920 --height 1080 --features 8000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-visual-odometry_b200"))
import torch  # noqa: E402
import b200vo  # noqa: E402
from b200vo import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=129)
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--features", type=int, default=8000)
ap.add_argument("--batch", type=int, default=16)
a = ap.parse_args()
ndev = torch.cuda.device_count()
frames = synth.make_sequence(a.width, a.height, a.frames, device="cuda")[0].cpu().numpy()
out = {"devices": ndev, "frames": a.frames, "shape": [a.width, a.height], "features": a.features, "nccl_version": b200vo.load_library().b200vo_nccl_version()}
res = {}
for devs in ([0], list(range(ndev))):
    with b200vo.Group(a.width, a.height, devs, nfeatures=a.features, max_frames=a.batch + 1, ransac_hypotheses=1024) as g:
        g.process_sequence(frames, a.batch)                       # warm-up
        t0 = time.perf_counter()
        recs = g.process_sequence(frames, a.batch)
        dt = time.perf_counter() - t0
    res[len(devs)] = recs
    out["pairs_per_s_%d_gpu" % len(devs)] = round((a.frames - 1) / dt, 1)
out["records_identical"] = bool(res[1].tobytes() == res[ndev].tobytes())
out["pairs_pose_ok"] = int(((res[ndev]["status"] & 4) != 0).sum())
print(json.dumps(out))
