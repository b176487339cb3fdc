#!/usr/bin/env python
"""e2e step time of the headline workload with two contexts per GPU under explicit per-context launch-batch schedules
(the experiment behind MultiRunner.e2e_candidates in bench.py):  python tools/e2e_two_contexts.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-visual-odometry_b200"))
import torch, b200vo, bench
from b200vo import synth
W, H, F, B = 1241, 376, 1000, 500
tex = synth.make_texture(); R, C = synth.make_trajectory(F)
d, h = bench.render(torch, synth, W, H, R, C, tex, "cuda:0", 32)
run = bench.MultiRunner(torch, b200vo, W, H, 3000, 1024, B, 0, d, h, 0, K=2)
RB = b200vo.RECORD_BYTES
recs_host = torch.empty(999 * RB, dtype=torch.uint8, pin_memory=True)
def fn():
    run.step(True)
    with torch.cuda.stream(run.stream):
        recs_host.copy_(run.recs_dev, non_blocking=True)
    run.stream.synchronize()
ms, _, _ = bench.timed_steps(torch, None, 1, run.stream, lambda: run.step(False), 4, 3, "cuda:0")
print(json.dumps({"device_resident_ms": round(ms, 3)}), flush=True)
cands = [[100], [[31], [100]], [[31, 279], [100]], [[31, 250], [100, 401]], [[25, 150], [75, 300]], [[31, 120, 350], [62, 250]],
         [[16, 100, 385], [50, 200, 250]], [[31, 200], [100, 300]]]
for chained in (False,):
    run.chained = chained
    for sizes in cands:
        run.e2e_sizes = sizes
        ms, _, _ = bench.timed_steps(torch, None, 1, run.stream, fn, 4, 2, "cuda:0")
        print(json.dumps({"chained": chained, "sizes": sizes, "e2e_ms": round(ms, 3)}), flush=True)
run.close(); del d, h, recs_host
bench.close_contexts(torch)
