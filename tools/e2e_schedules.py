#!/usr/bin/env python
"""How the launch-batch schedule of the e2e leg (host frames, H2D inside the timed region) trades the exposed first
upload against the efficiency of small batches: ms per 1000-frame step for several schedules, device-resident and e2e,
with the device-side ingest done by kernel or by cudaMemcpy3DAsync (B200VO_REPACK_KERNEL).  Under torchrun all ranks
run the same schedule at the same time (they share the host's memory system); times are the max over ranks.
   python tools/e2e_schedules.py        torchrun --nproc-per-node 8 tools/e2e_schedules.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "slam-visual-odometry_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import b200vo  # noqa: E402
from b200vo import synth  # noqa: E402
import bench  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = "cuda:%d" % local
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, F, B = 1241, 376, 1000, 500
tex = synth.make_texture()
R, C = synth.make_trajectory(F)
d, h = bench.render(torch, synth, W, H, R, C, tex, dev, 32)
host = torch.empty(999 * b200vo.RECORD_BYTES, dtype=torch.uint8, pin_memory=True)


def ramp(start, growth):
    out, r, done = [], start, 0
    while done < F:
        n = min(B, F - done, int(r))
        r = r * growth + 1
        out.append(n)
        done += n
    return out


scheds = {"500+500": [500, 500], "100+500+400": [100, 500, 400], "250x4": [250] * 4, "200x5": [200] * 5, "167x6": [167] * 5 + [165],
          "ramp41x1.4": ramp(41, 1.4), "ramp83x2": ramp(83, 2.0), "ramp125x1.6": ramp(125, 1.6), "ramp100x1.3": ramp(100, 1.3)}
for mode in ("1", "0"):
    os.environ["B200VO_REPACK_KERNEL"] = mode
    run = bench.SeqRunner(torch, b200vo, W, H, 3000, 1024, B, local, d, h, 0)

    def fn(sizes, host_input, sync):
        run.e2e_sizes = sizes
        if host_input:
            run.step(True)
        else:
            saved = run.B
            for first, n, done_f, p0, np_, done_p in run.batches(short_first=True):     # the same sizes, device-resident
                run.ctx.upload_frames_device(d.data_ptr() + done_f * W * H, n, W, W * H, 1, first)
                run.ctx.detect(first, n)
                if np_ > 0:
                    run.ctx.process_pairs_device(p0, np_, done_p)
        if sync:
            with torch.cuda.stream(run.stream):
                host.copy_(run.recs_dev, non_blocking=True)
            run.stream.synchronize()

    for name, sizes in scheds.items():
        row = {"repack_kernel": mode, "ranks": world, "schedule": name, "sizes": sizes}
        for label, host_input, sync in (("device", False, False), ("e2e_sync_per_step", True, True)):
            ms, _, _ = bench.timed_steps(torch, dist if world > 1 else None, world, run.stream, lambda: fn(sizes, host_input, sync), 4, 3, dev)
            row[label + "_ms"] = round(ms, 3)
        if rank == 0:
            print(json.dumps(row), flush=True)
    run.close()
del d, h, host
bench.close_contexts(torch)
if world > 1:
    dist.destroy_process_group()
