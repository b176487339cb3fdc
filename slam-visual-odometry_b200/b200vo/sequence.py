"""Multi-GPU plumbing for one sequence: contiguous frame-pair partition, record gather, chaining.

The reference processes a sequence strictly in order on one thread (src/main.cpp:128-194).  Frame
pairs are independent until pose chaining, so rank r of G owns a contiguous range of pairs, runs the
whole per-pair path on its own GPU, and the only exchange is one all_gather of the fixed-size
160-byte pose records (b200vo_pose_record); rank 0 then chains them on the host exactly as
process_frame would (src/visual_odometry.cpp:340-374).  torch.distributed is plumbing only: NCCL
over NVLink on GPUs, gloo in the CPU tests.
"""
import numpy as np

from . import POSE_RECORD_DTYPE


def partition_pairs(total_pairs, world, rank):
    """Contiguous range [lo, hi) of pair indices owned by `rank` (first ranks get the remainder)."""
    base, rem = divmod(total_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def frames_for_pairs(lo, hi):
    """Half-open frame range needed for pairs [lo, hi): frames lo .. hi inclusive."""
    return lo, hi + 1


def gather_records(mine, total_pairs, world, rank, device=None):
    """all_gather of per-rank record arrays (padded to the longest range) -> (total_pairs,) records."""
    import torch
    import torch.distributed as dist
    per = -(-total_pairs // world)
    buf = np.zeros(per, POSE_RECORD_DTYPE)
    buf[:len(mine)] = mine
    t = torch.from_numpy(buf.view(np.uint8).copy())
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * per * POSE_RECORD_DTYPE.itemsize, dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t)
    allb = out.cpu().numpy().view(POSE_RECORD_DTYPE).reshape(world, per)
    parts = []
    for r in range(world):
        lo, hi = partition_pairs(total_pairs, world, r)
        parts.append(allb[r, :hi - lo])
    return np.concatenate(parts)
