import json, os, sys, time
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "slam-visual-odometry_b200"))
import torch, b200vo, bench
from b200vo import synth
W, H, F, B = 1241, 376, 1000, 500
tex = synth.make_texture(); R, C = synth.make_trajectory(F)
d, h = bench.render(torch, synth, W, H, R, C, tex, "cuda:0", 32)
run = bench.SeqRunner(torch, b200vo, W, H, 3000, 1024, B, 0, d, h, 0)
RB = b200vo.RECORD_BYTES
recs_host = torch.empty(999 * RB, dtype=torch.uint8, pin_memory=True)
def mk(level):
    def fn():
        t0 = time.perf_counter()
        run.step(True)
        t1 = time.perf_counter()
        if level >= 1:
            with torch.cuda.stream(run.stream):
                recs_host.copy_(run.recs_dev, non_blocking=True)
        run.stream.synchronize()
        t2 = time.perf_counter()
        if level >= 2:
            b200vo.chain_trajectory(recs_host.numpy().view(b200vo.POSE_RECORD_DTYPE).copy())
        t3 = time.perf_counter()
        fn.t = (t1 - t0, t2 - t1, t3 - t2)
    return fn
for sizes in ([100, 500, 400], [41, 124, 373, 462], None):
    run.e2e_sizes = sizes
    for level in (0, 1, 2):
        f = mk(level)
        ms, _, _ = bench.timed_steps(torch, None, 1, run.stream, f, 4, 3, "cuda:0")
        print(json.dumps({"sizes": sizes, "level": level, "ms": round(ms, 3), "host_enqueue_ms": round(f.t[0]*1e3, 3), "wait_ms": round(f.t[1]*1e3, 3), "chain_ms": round(f.t[2]*1e3, 3)}), flush=True)
run.close(); del d, h, recs_host
bench.close_contexts(torch)
