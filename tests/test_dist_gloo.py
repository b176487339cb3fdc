"""Multi-rank logic on CPU (gloo, world_size 2): the contiguous frame-pair partition, the gather
of fixed-size 160-byte pose records and the host-side chaining.  The records here are synthetic
(no GPU): what is tested is that partition + all_gather + chain reproduces the single-process
trajectory bit for bit, with ranges of unequal length and with failed pairs inside."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make_records(b200vo, n, seed=0):
    rng = np.random.default_rng(seed)
    recs = np.zeros(n, b200vo.POSE_RECORD_DTYPE)
    for i in range(n):
        a = rng.normal(size=3) * 0.02
        K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
        R = np.eye(3) + K + K @ K / 2
        u, _, vt = np.linalg.svd(R)
        recs[i]["R"] = (u @ vt).ravel()
        t = rng.normal(size=3); recs[i]["t"] = t / np.linalg.norm(t)
        recs[i]["status"] = [7, 7, 7, 3, 1, 0][rng.integers(0, 6)]      # ok / rotation-only / failed pairs
        recs[i]["n_matches"] = rng.integers(0, 900)
    return recs


def _worker(rank, world, port, total_pairs, out_path):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-visual-odometry_b200"))
    import b200vo
    from b200vo import sequence as SQ
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    allrecs = _make_records(b200vo, total_pairs)
    lo, hi = SQ.partition_pairs(total_pairs, world, rank)
    mine = allrecs[lo:hi]                                   # what this rank's GPU would have produced
    gathered = SQ.gather_records(mine, total_pairs, world, rank)
    if rank == 0:
        traj = b200vo.chain_trajectory(gathered)
        np.save(out_path, traj)
        assert gathered.tobytes() == allrecs.tobytes()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("total_pairs", [999, 7, 2])
def test_partition_gather_chain_world2(built, tmp_path, total_pairs):
    from b200vo import sequence as SQ
    world = 2
    covered = []
    for r in range(world):
        lo, hi = SQ.partition_pairs(total_pairs, world, r)
        covered += list(range(lo, hi))
    assert covered == list(range(total_pairs))              # contiguous, disjoint, complete
    out = str(tmp_path / "traj.npy")
    port = 29500 + (os.getpid() + total_pairs) % 2000
    mp.spawn(_worker, args=(world, port, total_pairs, out), nprocs=world, join=True)
    ref = built.chain_trajectory(_make_records(built, total_pairs))
    assert np.array_equal(np.load(out), ref)


def test_frames_needed_by_rank(built):
    from b200vo import sequence as SQ
    # rank r needs frames [lo, hi] inclusive: boundary frames are detected by both neighbours
    assert SQ.frames_for_pairs(0, 500) == (0, 501)
    assert SQ.frames_for_pairs(500, 999) == (500, 1000)
