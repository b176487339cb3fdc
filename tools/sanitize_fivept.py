#!/usr/bin/env python
"""Small driver for compute-sanitizer (racecheck / memcheck) runs of the geometry kernels:
   compute-sanitizer --tool racecheck python tools/sanitize_fivept.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-visual-odometry_b200"))
import b200vo  # noqa: E402
from b200vo import synth  # noqa: E402

frames = synth.make_sequence(640, 240, 2)[0].numpy()
with b200vo.Context(640, 240, nfeatures=600, max_frames=4, ransac_hypotheses=96) as ctx:
    ctx.upload_frames(frames, 0)
    ctx.detect(0, 2)
    rec = ctx.process_pairs(0, 1, 0)[0]
    print("status", int(rec["status"]), "matches", int(rec["n_matches"]), "inliers", int(rec["n_inliers"]), "valid", int(rec["n_valid"]))
