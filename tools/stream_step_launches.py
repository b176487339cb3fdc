#!/usr/bin/env python
"""Four streaming steps (b200vo_process_frame) at the reference's 1200 features: the command whose ncu
launch list shows the dependent kernel sequence of one per-frame step.
   python tools/stream_step_launches.py
   ncu --metrics gpu__time_duration.sum --clock-control none -k regex:... --csv --log-file out.csv python tools/stream_step_launches.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-visual-odometry_b200"))
import b200vo  # noqa: E402
from b200vo import synth  # noqa: E402

frames = synth.make_sequence(1241, 376, 4)[0].numpy()
with b200vo.Context(1241, 376, nfeatures=1200, max_frames=2) as ctx:
    prev = -1
    for i, f in enumerate(frames):
        l0 = ctx.kernel_launches()
        kp, desc, rec, E = ctx.process_frame(f, i & 1, prev, i - 1)
        print("frame %d: %d keypoints, %d kernel launches, status %s" % (i, len(kp), ctx.kernel_launches() - l0, None if rec is None else int(rec["status"])))
        prev = i & 1
