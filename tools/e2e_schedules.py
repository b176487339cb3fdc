#!/usr/bin/env python
"""How the launch-batch schedule of the e2e leg (host frames, H2D inside the timed region) trades the exposed first
upload against the efficiency of small batches: ms per 1000-frame step for several schedules, device-resident and e2e.
   python tools/e2e_schedules.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-visual-odometry_b200"))
import torch  # noqa: E402
import b200vo  # noqa: E402
from b200vo import synth  # noqa: E402
import bench  # noqa: E402

W, H, F, B = 1241, 376, 1000, 500
tex = synth.make_texture()
R, C = synth.make_trajectory(F)
d, h = bench.render(torch, synth, W, H, R, C, tex, "cuda:0", 32)
run = bench.SeqRunner(torch, b200vo, W, H, 3000, 1024, B, 0, d, h, 0)
host = torch.empty(999 * b200vo.RECORD_BYTES, dtype=torch.uint8, pin_memory=True)


def schedule_step(sizes, host_input, d2h):
    done_f, done_p, prev = 0, 0, 0
    for n in sizes:
        first = 0 if done_f == 0 else (prev + 1) % (B + 1)
        p0, np_ = (first, n - 1) if done_f == 0 else (prev, n)
        if host_input:
            run.ctx.upload_frames_ptr(h.data_ptr() + done_f * W * H, n, W, W * H, 1, first)
        else:
            run.ctx.upload_frames_device(d.data_ptr() + done_f * W * H, n, W, W * H, 1, first)
        run.ctx.detect(first, n)
        if np_ > 0:
            run.ctx.process_pairs_device(p0, np_, done_p)
        prev = (first + n - 1) % (B + 1)
        done_f += n
        done_p += np_
    if d2h:
        with torch.cuda.stream(run.stream):
            host[:128].copy_(run.recs_dev[:128], non_blocking=True)
        run.stream.synchronize()


def ramp(start, growth):
    out, r, done = [], start, 0
    while done < F:
        n = min(B, F - done, r)
        r = int(r * growth) + 1
        out.append(n)
        done += n
    return out


scheds = {"500+500": [500, 500], "100+500+400": [100, 500, 400], "250x4": [250] * 4, "125x8": [125] * 8,
          "ramp41x1.4": ramp(41, 1.4), "ramp83x2": ramp(83, 2.0), "ramp125x1.6": ramp(125, 1.6)}
for name, sizes in scheds.items():
    row = {"schedule": name, "sizes": sizes}
    for label, host_input, d2h in (("device", False, False), ("e2e_async", True, False), ("e2e_sync_per_step", True, True)):
        ms, _, _ = bench.timed_steps(torch, None, 1, run.stream, lambda: schedule_step(sizes, host_input, d2h), 4, 3, "cuda:0")
        row[label + "_ms"] = round(ms, 3)
    print(json.dumps(row), flush=True)
run.close()
del d, h, host
bench.close_contexts(torch)
