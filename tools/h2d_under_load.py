#!/usr/bin/env python
"""Pinned host -> device copy rate of this process's GPU (a) with the GPU idle and (b) while the device-resident
pipeline step runs on another stream -- is the upload of the e2e leg slowed down by the kernels it overlaps with?
Under torchrun every rank measures at the same time (all ranks share the host's memory system and PCIe fabric).
   python tools/h2d_under_load.py            torchrun --nproc-per-node 8 tools/h2d_under_load.py"""
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

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H, F, B = 1241, 376, 501, 500
tex = synth.make_texture()
R, C = synth.make_trajectory(F)
d = synth.render_frames(W, H, R, C, tex, device="cuda:%d" % local, chunk=32)
h = torch.empty(tuple(d.shape), dtype=torch.uint8, pin_memory=True)
h.copy_(d)
dst = torch.empty_like(d)
ctx = b200vo.Context(W, H, device=local, nfeatures=3000, max_frames=B + 1, ransac_hypotheses=1024)
s2 = torch.cuda.Stream()


def compute(nsteps):
    for _ in range(nsteps):
        ctx.upload_frames_device(d.data_ptr(), B, W, W * H, 1, 0)
        ctx.detect(0, B)
        ctx.process_pairs_device(0, B - 1, 0)


def copy_rate(chunks, reps):
    """Rate of the FIRST pass over the frames (all ranks are still copying then: the contended rate) and of all passes."""
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    n = F // chunks
    with torch.cuda.stream(s2):
        ev[0].record()
        for r in range(reps):
            for c in range(chunks):
                dst[c * n:(c + 1) * n].copy_(h[c * n:(c + 1) * n], non_blocking=True)
            ev[r + 1].record()
    torch.cuda.synchronize()
    per = chunks * n * W * H / 1e9
    return per / (ev[0].elapsed_time(ev[1]) * 1e-3), reps * per / (ev[0].elapsed_time(ev[reps]) * 1e-3)


compute(1)
torch.cuda.synchronize()
out = {}
for chunks in (1, 24):
    if world > 1:
        dist.barrier()
    out["idle_first_pass_%dchunks" % chunks], out["idle_all_passes_%dchunks" % chunks] = copy_rate(chunks, 4)
    if world > 1:
        dist.barrier()
    compute(8)                       # ~240 ms of queued kernels; the copies below overlap them
    out["under_compute_first_pass_%dchunks" % chunks], out["under_compute_all_passes_%dchunks" % chunks] = copy_rate(chunks, 4)
    torch.cuda.synchronize()
vals = torch.tensor([out[k] for k in sorted(out)], device="cuda:%d" % local, dtype=torch.float64)
allv = [torch.zeros_like(vals) for _ in range(world)]
if world > 1:
    dist.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    print(json.dumps({"ranks": world, "GBps_per_rank": {k: [round(float(v[i]), 2) for v in allv] for i, k in enumerate(sorted(out))}}))
del d, h, dst, vals, allv
torch.cuda.synchronize()
ctx.close()
if world > 1:
    dist.destroy_process_group()
