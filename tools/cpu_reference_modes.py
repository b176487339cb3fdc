#!/usr/bin/env python
"""CPU reference path (cv2 = the OpenCV calls the reference makes) timed three ways on this host,
following SURVEY.md section 8(d)'s CPU baseline timing plan:

  (i)   cv2.setNumThreads(1), one process: per-pair latency with the stage split
  (ii)  cv2.setNumThreads(os.cpu_count()): OpenCV-internal parallelism (only the matcher scales)
  (iii) throughput: one process per core, cv2.setNumThreads(1), contiguous pair ranges

Config 1 (one 1241x376 pair, 2000 features) is the reference's own CPU-runnable case; the same
three modes are repeated at config 2's 3000 features.  Measurement tool only: it runs oracle/ code
(the cv2 restatement of VisualOdometry::process_frame) and never touches the GPU library.

  python tools/cpu_reference_modes.py [--reps 10] [--pairs 128] > profiles/cpu_modes_rNN.jsonl
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "slam-visual-odometry_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pair_latency(f0, f1, nfeatures, nthreads, reps, warmup=3):
    """Median per-stage milliseconds of one reference pair step: detect (the new frame only, as in a
    sequence), match + threshold, findEssentialMat, recoverPose."""
    import cv2
    from oracle import pipeline as PL
    cv2.setNumThreads(nthreads)
    h, w = f0.shape[:2]
    vo = PL.Cv2VisualOdometry(w, h, nfeatures)
    k0, d0 = vo.detect(f0)
    rows = []
    for it in range(warmup + reps):
        t0 = time.perf_counter()
        k1, d1 = vo.detect(f1)
        t1 = time.perf_counter()
        good = vo.good_matches(d0, d1)
        t2 = time.perf_counter()
        p1 = np.float32([k0[m.queryIdx].pt for m in good])
        p2 = np.float32([k1[m.trainIdx].pt for m in good])
        E, mask = cv2.findEssentialMat(p1, p2, vo.K, cv2.RANSAC, 0.999, 2.0)
        t3 = time.perf_counter()
        nv = cv2.recoverPose(E[:3], p1, p2, vo.K, mask=mask)[0] if E is not None and E.size else 0
        t4 = time.perf_counter()
        if it >= warmup:
            rows.append([(t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t4 - t0) * 1e3])
    med = np.median(np.asarray(rows), axis=0)
    best = np.min(np.asarray(rows), axis=0)
    return {"threads": nthreads, "reps": reps, "keypoints": len(k1), "good_matches": len(good), "valid_points": int(nv),
            "median_ms": dict(zip(("detect", "match", "essential", "recover_pose", "total"), (round(float(v), 3) for v in med))),
            "best_total_ms": round(float(best[4]), 3), "pairs_per_s": round(1e3 / float(med[4]), 2)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--pairs", type=int, default=128, help="pairs in the throughput mode (iii)")
    ap.add_argument("--width", type=int, default=1241)
    ap.add_argument("--height", type=int, default=376)
    a = ap.parse_args()
    import cv2
    import bench
    from b200vo import synth
    ncpu = os.cpu_count() or 1
    nfr = a.pairs + 1
    frames = synth.make_sequence(a.width, a.height, nfr)[0].numpy()
    try:
        cpu = [l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        cpu = "unknown"
    for cfg, nf in (("1: single 1241x376 pair, 2000 features (reference CPU path)", 2000),
                    ("2: 1241x376, 3000 features (the bench workload)", 3000)):
        out = {"config": cfg, "size": [a.width, a.height], "cv2": cv2.__version__, "host_cpu": cpu, "cores": ncpu,
               "one_thread": pair_latency(frames[0], frames[1], nf, 1, a.reps),
               "all_threads": pair_latency(frames[0], frames[1], nf, ncpu, a.reps)}
        rate, pairs, dt = bench.cpu_reference_throughput(frames, nf, a.pairs, ncpu)
        out["throughput"] = {"processes": min(ncpu, a.pairs), "threads_each": 1, "pairs": pairs, "seconds": round(dt, 3),
                             "pairs_per_s": round(rate, 2)}
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
