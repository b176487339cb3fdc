#!/usr/bin/env python
"""Experiment: ONE GPU, the pair range of a sequence split over K contexts (K streams) whose launch batches are
enqueued round-robin, so the latency-bound kernels and the tail waves of one context's batch overlap the
throughput-bound kernels of another's.  Wall time per step between device synchronisations, inputs resident in HBM.
   python tools/two_contexts.py [--frames 1000] [--features 3000] [--batch 500] [--contexts 1,2,3]"""
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
ap.add_argument("--frames", type=int, default=1000)
ap.add_argument("--features", type=int, default=3000)
ap.add_argument("--batch", default="500,250")
ap.add_argument("--contexts", default="1,2,3")
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
W, H, F = 1241, 376, a.frames
tex = synth.make_texture()
R, C = synth.make_trajectory(F)
d = synth.render_frames(W, H, R, C, tex, device="cuda:0", chunk=16)
fb = W * H
ref = None
for K in [int(x) for x in a.contexts.split(",")]:
    for B in [int(x) for x in a.batch.split(",")]:
        # context k owns pairs [lo_k, hi_k): frames lo_k .. hi_k
        cuts = [round(i * (F - 1) / K) for i in range(K + 1)]
        ctxs = [b200vo.Context(W, H, nfeatures=a.features, max_frames=B + 1) for _ in range(K)]
        recs = [None] * K

        def step():
            work = []
            for k in range(K):
                lo, hi = cuts[k], cuts[k + 1]
                frames = hi - lo + 1
                items, done_f, done_p, prev = [], 0, 0, 0
                while done_f < frames:
                    n = min(B, frames - done_f)
                    if frames - done_f - n == 1:
                        n -= 1
                    first = 0 if done_f == 0 else (prev + 1) % (B + 1)
                    p0, np_ = (first, n - 1) if done_f == 0 else (prev, n)
                    items.append((k, lo, first, n, done_f, p0, np_, done_p))
                    prev = (first + n - 1) % (B + 1)
                    done_f += n
                    done_p += np_
                work.append(items)
            out = [[] for _ in range(K)]
            for i in range(max(len(w) for w in work)):          # round-robin over the contexts
                for k in range(K):
                    if i < len(work[k]):
                        _, lo, first, n, done_f, p0, np_, done_p = work[k][i]
                        c = ctxs[k]
                        c.upload_frames_device(d.data_ptr() + (lo + done_f) * fb, n, W, fb, 1, first)
                        c.detect(first, n)
                        if np_ > 0:
                            out[k].append(c.process_pairs(p0, np_, lo + done_p) if False else (c.process_pairs_device(p0, np_, lo + done_p), np_))
            return out

        for _ in range(2):
            step()
            torch.cuda.synchronize()
        t = []
        for _ in range(a.reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            step()
            torch.cuda.synchronize()
            t.append(time.perf_counter() - t0)
        ms = float(np.median(t)) * 1e3
        print(json.dumps({"contexts": K, "batch": B, "ms_per_step": round(ms, 3), "pairs_per_s": round((F - 1) / ms * 1e3, 1)}), flush=True)
        for c in ctxs:
            c.close()
