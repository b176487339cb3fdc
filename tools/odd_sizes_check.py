import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/slam-visual-odometry_b200')
import b200vo
from b200vo import synth
from oracle import orb as O
for (W,H,NF) in ((100,80,300),(333,222,700),(1999,1001,5000),(64,64,100),(4095,300,2000)):
    frames=synth.make_sequence(W,H,2,device="cuda")[0].cpu().numpy()
    with b200vo.Context(W,H,nfeatures=NF,max_frames=3,ransac_hypotheses=64) as ctx:
        ctx.upload_frames(frames,0); ctx.detect(0,2)
        ok=True
        for s in (0,1):
            kp,desc=ctx.get_features(s); ref=O.detect_and_compute(frames[s],NF)
            same=(len(kp)==len(ref["x"]) and np.array_equal(kp["lx"],ref["x"]) and np.array_equal(kp["ly"],ref["y"]) and np.array_equal(kp["octave"],ref["level"]) and np.array_equal(kp["angle"],ref["angle"]) and np.array_equal(kp["response"],ref["response"]) and np.array_equal(desc,ref["desc"]))
            ok&=same
        rec=ctx.process_pairs(0,1,0)[0]
        print(W,H,NF,"features exact:",ok,"n",len(kp),"status",int(rec["status"]),"matches",int(rec["n_matches"]))
