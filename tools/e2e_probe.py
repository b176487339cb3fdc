#!/usr/bin/env python
"""Where do the ~2.5 ms between the device-resident step and the end-to-end step go?  The e2e step of the headline workload
with (a) pinned host frames, (b) the SAME staged upload path fed from a device buffer (no PCIe transfer: what is left is the
pipeline's own structure -- staging, events between the copy and the compute stream, extra launch batches)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-visual-odometry_b200"))
import torch, b200vo, bench
from b200vo import synth
W, H, F, B = 1241, 376, 1000, 500
tex = synth.make_texture(); R, C = synth.make_trajectory(F)
d, h = bench.render(torch, synth, W, H, R, C, tex, "cuda:0", 32)
d2 = d.clone()
RB = b200vo.RECORD_BYTES
recs_host = torch.empty(999 * RB, dtype=torch.uint8, pin_memory=True)
for K in (1, 2):
    for name, src in (("pinned_host", h), ("device_buffer", d2)):
        run = bench.MultiRunner(torch, b200vo, W, H, 3000, 1024, B, 0, d, src, 0, K=K) if K > 1 else bench.SeqRunner(torch, b200vo, W, H, 3000, 1024, B, 0, d, src, 0)
        def fn():
            run.step(True)
            with torch.cuda.stream(run.stream):
                recs_host.copy_(run.recs_dev, non_blocking=True)
            run.stream.synchronize()
        ms0, _, _ = bench.timed_steps(torch, None, 1, run.stream, lambda: run.step(False), 4, 3, "cuda:0")
        for sizes in ([], [100], [250], [31, 100, 320]):
            run.e2e_sizes = sizes
            ms, _, _ = bench.timed_steps(torch, None, 1, run.stream, fn, 4, 3, "cuda:0")
            print(json.dumps({"contexts": K, "source": name, "sizes": sizes, "device_resident_ms": round(ms0, 3), "e2e_ms": round(ms, 3)}), flush=True)
        run.close()
del d, h, d2, recs_host
bench.close_contexts(torch)
