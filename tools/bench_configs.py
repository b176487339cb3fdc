#!/usr/bin/env python
"""Single-GPU timings of BASELINE.json's OTHER configurations (they are parity-test cases, not bench
lines; this script only records what they cost on one B200):
   config 3: 1920x1080, 8000 features        config 5: 3840x2160, 20000 features, H = 4096
   config 4: brute-force Hamming stress, 32768 x 32768 descriptors per pair, batched 64 pairs
python tools/bench_configs.py [--frames F]   -> one JSON line per configuration on stdout"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-visual-odometry_b200"))


def sequence_config(name, W, H, NF, HYP, F, B):
    import torch
    import b200vo
    from b200vo import synth
    tex = synth.make_texture()
    R, C = synth.make_trajectory(F)
    d = synth.render_frames(W, H, R, C, tex, device="cuda:0", chunk=4)
    fb = W * H
    with b200vo.Context(W, H, nfeatures=NF, max_frames=B + 1, ransac_hypotheses=HYP) as ctx:
        stream = torch.cuda.ExternalStream(ctx.stream())

        def step():
            done, prev = 0, 0
            ok = 0
            while done < F:
                n = min(B, F - done)
                first = 0 if done == 0 else (prev + 1) % (B + 1)
                ctx.upload_frames_device(d.data_ptr() + done * fb, n, W, fb, 1, first)
                ctx.detect(first, n)
                p0, np_ = (first, n - 1) if done == 0 else (prev, n)
                if np_ > 0:
                    ctx.process_pairs_device(p0, np_, done - (0 if done == 0 else 1))
                prev = (first + n - 1) % (B + 1)
                done += n
        for _ in range(2):
            step()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            step()
        e1.record(stream)
        stream.synchronize()
        ms = e0.elapsed_time(e1) / 3
        ctx.enable_timing(True)
        ctx.upload_frames_device(d.data_ptr(), min(B, F), W, fb, 1, 0)
        ctx.detect(0, min(B, F))
        st = ctx.stage_times()
        recs = ctx.process_pairs(0, min(B, F) - 1, 0)
        st2 = ctx.stage_times()
        ctx.enable_timing(False)
        kp = len(ctx.get_features(0)[0])
    stages = {k[:-3]: round(st[k], 3) for k in ("pyramid_ms", "fast_ms", "select_ms", "orient_ms", "blur_ms", "brief_ms")}
    stages.update({k[:-3]: round(st2[k], 3) for k in ("match_ms", "filter_ms", "ransac_ms", "pose_ms")})
    return {"config": name, "frames": F, "batch": B, "pairs_per_s": round((F - 1) / (ms * 1e-3), 1), "ms_per_step": round(ms, 3),
            "keypoints_frame0": kp, "pairs_pose_ok_first_batch": int(((recs["status"] & 4) != 0).sum()), "pairs_first_batch": len(recs),
            "stage_ms_first_batch": stages}


def matcher_stress(npairs=64, n=32768):
    import torch
    import b200vo
    g = torch.Generator(device="cuda").manual_seed(7)
    q = torch.randint(0, 256, (npairs, n, 32), dtype=torch.uint8, device="cuda", generator=g)
    t = torch.randint(0, 256, (npairs, n, 32), dtype=torch.uint8, device="cuda", generator=g)
    nnq = torch.empty((npairs, n), dtype=torch.int32, device="cuda"); nnt = torch.empty_like(nnq)
    with b200vo.Context(640, 240, nfeatures=500, max_frames=2, ransac_hypotheses=64) as ctx:
        rate = ctx.measure_popc_rate()
        ctx.match_descriptors_device(q.data_ptr(), n, t.data_ptr(), n, npairs, nnq.data_ptr(), nnt.data_ptr())
        ms = min(ctx.match_descriptors_device(q.data_ptr(), n, t.data_ptr(), n, npairs, nnq.data_ptr(), nnt.data_ptr()) for _ in range(3))
    torch.cuda.synchronize()
    pairs = float(npairs) * n * n
    return {"config": "4: Hamming stress %d pairs of %dx%d 256-bit descriptors, cross-check" % (npairs, n, n), "ms_per_batch": round(ms, 2),
            "descriptor_pairs_per_s": round(pairs / (ms * 1e-3), 0), "executed_Gpopc32_per_s": round(pairs * 5 / (ms * 1e-3) / 1e9, 1),
            "popc_pipe_peak_Gpopc32_per_s": round(rate / 1e9, 1), "frac_of_popc_peak": round(pairs * 5 / (ms * 1e-3) / rate, 4)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=33)
    a = ap.parse_args()
    print(json.dumps(matcher_stress()), flush=True)
    print(json.dumps(sequence_config("3: 1920x1080, 8000 features, H=1024", 1920, 1080, 8000, 1024, a.frames, 16)), flush=True)
    print(json.dumps(sequence_config("5: 3840x2160, 20000 features, H=4096", 3840, 2160, 20000, 4096, max(9, a.frames // 2), 8)), flush=True)
