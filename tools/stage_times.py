#!/usr/bin/env python
"""Per-stage CUDA-event times of one launch batch under a set of environment knobs -- the A/B harness for
kernel variants (B200VO_FPG, B200VO_DESCRIBE, B200VO_TMA, B200VO_MATCH_CSA ...):
   python tools/stage_times.py --frames 251 --features 3000 --set B200VO_FPG=0,1,2,3,4
prints one JSON line per value (median of --reps instrumented passes)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-visual-odometry_b200"))
import b200vo  # noqa: E402
from b200vo import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=251)
ap.add_argument("--features", type=int, default=3000)
ap.add_argument("--hypotheses", type=int, default=1024)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--width", type=int, default=1241)
ap.add_argument("--height", type=int, default=376)
ap.add_argument("--set", default="", help="NAME=v1,v2,... one context per value")
a = ap.parse_args()
W, H, F = a.width, a.height, a.frames
tex = synth.make_texture()
R, C = synth.make_trajectory(F)
d = synth.render_frames(W, H, R, C, tex, device="cuda:0", chunk=16)
name, vals = (a.set.split("=")[0], a.set.split("=")[1].split(",")) if a.set else ("", [""])
for v in vals:
    if name:
        os.environ[name] = v
    with b200vo.Context(W, H, nfeatures=a.features, max_frames=F + 1, ransac_hypotheses=a.hypotheses) as ctx:
        ctx.enable_timing(True)
        rows = []
        for rep in range(a.reps + 1):
            ctx.upload_frames_device(d.data_ptr(), F, W, W * H, 1, 0)
            ctx.detect(0, F)
            st = dict(ctx.stage_times())
            ctx.process_pairs_device(0, F - 1, 0)
            st2 = ctx.stage_times()
            for k in ("match_ms", "filter_ms", "ransac_ms", "pose_ms", "fivept_ms", "score_ms"):
                st[k] = st2[k]
            if rep:
                rows.append(st)
        med = {k: round(float(np.median([r[k] for r in rows])), 4) for k in rows[0] if rows[0][k] or k in ("describe_ms",)}
        rec = ctx.process_pairs(0, 4, 0)
        print(json.dumps({"knob": "%s=%s" % (name, v) if name else "default", "frames": F, "features": a.features, "ms": med,
                          "check": [int(rec["n_inliers"].sum()), int(rec["n_valid"].sum()), int(rec["best_hyp"].sum())]}), flush=True)
