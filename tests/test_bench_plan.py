"""CPU tests of bench.py's launch-batch planning (no GPU): the ring-of-slots batch plan every runner follows, the
launch-batch schedules of the e2e leg, and the split of a rank's pair range over several contexts of one GPU."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("F,B,sizes", [(1000, 500, []), (1000, 500, [100]), (1000, 500, [83, 167, 335, 415]), (501, 501, []), (501, 500, []),
                                       (256, 16, [3]), (129, 17, [2, 5, 10]), (17, 16, []), (2, 500, []), (1000, 501, [31, 279]), (7, 3, [2])])
def test_batch_plan_covers_every_frame_and_pair_once(F, B, sizes):
    slots = B + 1
    frames_seen, pairs_seen = [], []
    slot_of = {}
    prev_last = None
    for first, n, done_f, p0, np_, done_p in bench.batch_plan(F, B, sizes):
        assert 1 <= n <= B and 0 <= first < slots
        assert done_f == len(frames_seen) and done_p == len(pairs_seen)
        for i in range(n):
            slot_of[done_f + i] = (first + i) % slots
        frames_seen += list(range(done_f, done_f + n))
        # pairs of this batch: (p0 + i, p0 + i + 1) in slot space must be frames (done_p + i, done_p + i + 1)
        for i in range(np_):
            a, b = done_p + i, done_p + i + 1
            assert slot_of[a] == (p0 + i) % slots and slot_of[b] == (p0 + i + 1) % slots
            pairs_seen.append(a)
        # a batch never overwrites the previous batch's last frame (the first image of its first pair)
        if prev_last is not None:
            assert slot_of[prev_last] not in [(first + i) % slots for i in range(n)]
        prev_last = done_f + n - 1
        if F > 2:
            assert not (n == 1 and done_f + n == F and done_f > 0), "a single trailing frame got a batch of its own"
    assert frames_seen == list(range(F)) and pairs_seen == list(range(F - 1))


def test_e2e_schedules_are_valid_plans():
    for F, B in ((1000, 500), (501, 501), (500, 65), (128, 17), (64, 16)):
        for sizes in bench.e2e_schedule_candidates(F, B):
            assert all(1 <= s <= B for s in sizes)
            got = list(bench.batch_plan(F, B, sizes))
            assert sum(n for _, n, *_ in got) == F


def test_context_cuts_partition_the_pair_range():
    for npairs in (999, 500, 255, 16, 9):
        for K in (1, 2, 3):
            cuts = [round(k * npairs / K) for k in range(K + 1)]
            assert cuts[0] == 0 and cuts[-1] == npairs and all(b > a for a, b in zip(cuts, cuts[1:]))
            # context k owns pairs [cuts[k], cuts[k+1]) and therefore frames cuts[k] .. cuts[k+1]: one boundary frame per cut is shared
            frames = sum(cuts[k + 1] - cuts[k] + 1 for k in range(K))
            assert frames == npairs + 1 + (K - 1)
