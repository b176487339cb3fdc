#!/usr/bin/env python
"""Stage times and launch-sequence wall time of ONE frame / ONE pair (the streaming regime of
VisualOdometry::process_frame): python tools/single_frame_times.py"""
import sys, time, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/slam-visual-odometry_b200')
import b200vo
from b200vo import synth
frames=synth.make_sequence(1241,376,3,device="cuda")[0].cpu().numpy()
for NF in (1200,3000):
    with b200vo.Context(1241,376,nfeatures=NF,max_frames=3) as ctx:
        ctx.upload_frames(frames,0); ctx.detect(0,3); ctx.process_pairs(0,2,0)
        ctx.enable_timing(True)
        ctx.detect(1,1); st=ctx.stage_times(); ctx.process_pairs(0,1,0); st2=ctx.stage_times()
        ctx.enable_timing(False)
        d={k:round(st[k],3) for k in ("pyramid_ms","fast_ms","select_ms","orient_ms","blur_ms","brief_ms")}
        d.update({k:round(st2[k],3) for k in ("match_ms","filter_ms","ransac_ms","pose_ms","fivept_ms","score_ms")})
        print(NF, d, "sum", round(sum(v for k,v in d.items() if k not in ("fivept_ms","score_ms")),3))
        # wall time of the pure launch sequence without sync in between
        ctx.synchronize(); t0=time.perf_counter()
        for _ in range(50):
            ctx.detect(1,1); ctx.process_pairs_device(0,1,0)
        ctx.synchronize(); print("  back-to-back detect+pairs per frame: %.3f ms"%((time.perf_counter()-t0)/50*1e3))
        t0=time.perf_counter()
        for _ in range(50):
            ctx.process_frame(frames[1],1,0,0)
        print("  process_frame (python) per frame: %.3f ms"%((time.perf_counter()-t0)/50*1e3))
