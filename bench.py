#!/usr/bin/env python
"""bench.py -- frame-pairs/sec of the VO front end (detect + match + E-RANSAC + pose).

Contract (see DESIGN.md "Measurement"):
  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
One STEP = one pass of the hot path over the whole synthetic sequence of this rank:
  F frames of 1241x376 (BASELINE.json configs[1]: 1000 frames, 3000 ORB features) -> F-1 pairs,
  processed in launch batches of --batch frames (default 500): upload -> b200vo_detect -> b200vo_process_pairs.
`value`  : whole-job pairs/s with the frames already resident in HBM (device->device ingest).
`e2e`    : the same through the C ABI with HOST (pinned) frames: H2D copy of every frame and D2H
           read of the pose records + host-side pose chaining inside the timed region.
Multi-GPU (torchrun, one process per GPU): the global sequence is partitioned into contiguous
frame-pair ranges (weak scaling: each rank gets its own F-1 pairs); the only exchange is one NCCL
all_gather of the 160-byte pose records, then rank 0 chains the trajectory on the host.
--impl reference: the reference's own CPU path (the same OpenCV calls through cv2, oracle/pipeline.py)
on the host cores, one worker process per core over contiguous pair ranges.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "slam-visual-odometry_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

W_IMG, H_IMG = 1241, 376
MATCH_DRAM_BYTES_PER_LAUNCH = 62631936      # profiles/ncu_r01_match_fullconfig.txt: 60.03 MB read + 2.60 MB written, 500-frame batch, 5.65 ms
METRIC = "frame-pairs/sec (detect+match+E-RANSAC+pose)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=1000, help="frames per rank (F-1 pairs per step)")
    ap.add_argument("--features", type=int, default=3000)
    ap.add_argument("--hypotheses", type=int, default=1024)
    ap.add_argument("--batch", type=int, default=500, help="frames per launch batch (measured: 125 -> 26.6k, 250 -> 28.6k, 500 -> 29.5k, 1000 -> 30.1k pairs/s)")
    ap.add_argument("--cpu-pairs", type=int, default=0, help="pairs in the bounded CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    frames, nfeatures, nthreads = args
    import cv2
    cv2.setNumThreads(nthreads)
    from oracle import pipeline as PL
    t0 = time.perf_counter()
    recs, _ = PL.run_pairs_cv2(frames, nfeatures)
    return len(recs), time.perf_counter() - t0


def cpu_reference_throughput(frames_host, nfeatures, pairs_total, nproc):
    """The reference path (cv2) over `pairs_total` pairs split into contiguous ranges, one process
    per core with cv2.setNumThreads(1).  Returns (pairs/s, pairs, seconds)."""
    import multiprocessing as mp
    nproc = max(1, min(nproc, pairs_total))
    per = pairs_total // nproc
    jobs = []
    for w in range(nproc):
        s = w * per
        jobs.append((frames_host[s:s + per + 1], nfeatures, 1))
    ctx = mp.get_context("spawn")
    with ctx.Pool(nproc) as pool:
        pool.map(_cpu_worker, [(frames_host[:2], nfeatures, 1)] * nproc)      # warm-up: import cv2, page in
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, jobs)
        dt = time.perf_counter() - t0
    pairs = sum(r[0] for r in res)
    return pairs / dt, pairs, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from b200vo import synth
    ncpu = os.cpu_count() or 1
    per = args.cpu_pairs or 8
    pairs = ncpu * per
    tex = synth.make_texture()
    R_wc, C_w = synth.make_trajectory(args.frames)
    nfr = min(args.frames, pairs + 1)
    # rendered on the CPU: nothing of this arm touches a GPU
    frames = synth.render_frames(W_IMG, H_IMG, R_wc[:nfr], C_w[:nfr], tex, device="cpu").numpy()
    pairs = min(pairs, nfr - 1)
    vals = []
    for i in range(args.warmup + args.steps):
        v, n, dt = cpu_reference_throughput(frames, args.features, pairs, ncpu)
        if i >= args.warmup:
            vals.append((v, n, dt))
    v = float(np.mean([x[0] for x in vals]))
    ms = float(np.mean([x[2] for x in vals]) * 1e3)
    import cv2
    sample = "%d pairs of the same synthetic %dx%d frames (N=%d) per step, %d processes x cv2.setNumThreads(1), cv2 %s" % (
        vals[0][1], W_IMG, H_IMG, args.features, min(ncpu, pairs), cv2.__version__)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/f64", "data": "synthetic",
            "config": {"workload": "%d-frame synthetic KITTI-shaped sequence %dx%d, %d ORB features (bounded sample)" % (args.frames, W_IMG, H_IMG, args.features)},
            "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": min(ncpu, pairs), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


# --------------------------------------------------------------------------- B200 arm
class DevView:
    """Zero-copy torch view of a raw device pointer."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPU cores NVML reports as local to its GPU BEFORE the pinned frame buffer is
    allocated (first touch decides the NUMA node): with eight ranks uploading 467 MB per step each, host
    memory on the wrong socket makes the e2e path PCIe/UPI-bound.  Best effort; silently skipped."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = None
        try:                                   # CUDA ordinal -> NVML handle through the UUID (CUDA_VISIBLE_DEVICES may reorder)
            import torch
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def algorithmic_bytes(ctx, nfeatures):
    """algorithmic bytes per frame of the HBM-bound stages."""
    px = [ctx.level_info(l)[0] * ctx.level_info(l)[1] for l in range(8)]
    return {
        "pyramid": sum(px[l - 1] + px[l] for l in range(1, 8)),
        "fast": sum(px),
        "blur": 2 * sum(px),
        "brief": sum(px) + 32 * nfeatures,
        "orient": 749 * nfeatures,
    }


def accuracy(traj, R_gt, C_gt):
    """Chained trajectory against the ground truth of the synthetic sequence: similarity-aligned ATE
    (the reference applies a fixed 0.3 step scale) and mean relative rotation / translation-direction
    error per pair.  Frame 0 is the identity pose the chain starts from."""
    from b200vo import trajectory as T
    poses = np.concatenate([np.concatenate([np.eye(3).ravel(), np.zeros(3)])[None], np.asarray(traj)], 0)
    n = len(poses)
    # ground truth relative to its own first frame (the estimate starts at identity)
    R0, C0 = R_gt[0], C_gt[0]
    Rg = np.stack([R0.T @ R for R in R_gt[:n]])
    Cg = np.stack([R0.T @ (C - C0) for C in C_gt[:n]])
    return {"vs_ground_truth": T.accuracy_report(poses, Rg, Cg)}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import b200vo
    from b200vo import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    F, NF, B = args.frames, args.features, min(args.batch, args.frames)
    npairs = F - 1
    # global sequence: rank r owns pairs [r*(F-1), (r+1)*(F-1)) -> frames [r*(F-1), r*(F-1)+F)
    tex = synth.make_texture()
    R_all, C_all = synth.make_trajectory(world * npairs + 1)
    f0 = rank * npairs
    d_frames = synth.render_frames(W_IMG, H_IMG, R_all[f0:f0 + F], C_all[f0:f0 + F], tex, device="cuda:%d" % local, chunk=32)
    h_frames = torch.empty((F, H_IMG, W_IMG), dtype=torch.uint8, pin_memory=True)
    h_frames.copy_(d_frames)
    torch.cuda.synchronize()

    ctx = b200vo.Context(W_IMG, H_IMG, device=local, nfeatures=NF, max_frames=B + 1, ransac_hypotheses=args.hypotheses)
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    RB = b200vo.RECORD_BYTES
    recs_dev = torch.zeros(npairs * RB, dtype=torch.uint8, device="cuda:%d" % local)
    recs_all = torch.zeros(world * npairs * RB, dtype=torch.uint8, device="cuda:%d" % local) if world > 1 else recs_dev
    recs_host = torch.empty(world * npairs * RB, dtype=torch.uint8, pin_memory=True)
    frame_bytes = W_IMG * H_IMG
    state = {"traj": None}

    def step(host_input):
        done_frames, done_pairs, prev_slot = 0, 0, 0
        while done_frames < F:
            n = min(B, F - done_frames)
            if host_input and done_frames == 0:
                n = min(n, max(B // 5, 2))      # e2e: a short first batch, so only its H2D copy is not hidden behind compute
            first = 0 if done_frames == 0 else (prev_slot + 1) % (B + 1)
            if host_input:
                ctx.upload_frames_ptr(h_frames.data_ptr() + done_frames * frame_bytes, n, W_IMG, frame_bytes, 1, first)
            else:
                ctx.upload_frames_device(d_frames.data_ptr() + done_frames * frame_bytes, n, W_IMG, frame_bytes, 1, first)
            ctx.detect(first, n)
            if done_frames == 0:
                p0, np_ = first, n - 1
            else:
                p0, np_ = prev_slot, n
            if np_ > 0:
                ptr = ctx.process_pairs_device(p0, np_, f0 + done_pairs)
                with torch.cuda.stream(stream):
                    src = torch.as_tensor(DevView(ptr, np_ * RB), device="cuda:%d" % local)
                    recs_dev[done_pairs * RB:(done_pairs + np_) * RB].copy_(src, non_blocking=True)
            prev_slot = (first + n - 1) % (B + 1)
            done_frames += n
            done_pairs += np_
        with torch.cuda.stream(stream):
            if world > 1:
                dist.all_gather_into_tensor(recs_all, recs_dev)     # the one exchange step: 160 B per pair
            if host_input:
                recs_host.copy_(recs_all, non_blocking=True)
        if host_input:
            stream.synchronize()
            if rank == 0:   # host-side chaining of the whole trajectory (reference visual_odometry.cpp:340-374)
                state["traj"] = b200vo.chain_trajectory(recs_host.numpy().view(b200vo.POSE_RECORD_DTYPE).copy())

    def timed(host_input, steps, warmup):
        for _ in range(warmup):
            step(host_input)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ctx.kernel_launches()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            step(host_input)
        e1.record(stream)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda:%d" % local)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, wall / steps, ctx.kernel_launches() - l0

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, wall_dev, launches = timed(False, args.steps, args.warmup)
    ms_e2e, wall_e2e, _ = timed(True, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None

    # ---- extra: the same workload with OpenCV's adaptive RANSAC stopping rule (ransac_adaptive=1),
    # i.e. what the reference's findEssentialMat does; reported beside the fixed-H headline, not as it
    main_ctx = ctx
    ctx = b200vo.Context(W_IMG, H_IMG, device=local, nfeatures=NF, max_frames=B + 1, ransac_hypotheses=args.hypotheses, ransac_adaptive=1)
    stream_main = stream
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    ms_dev_ad, _, _ = timed(False, args.steps, args.warmup)
    ms_e2e_ad, _, _ = timed(True, args.steps, max(args.warmup, 3))
    rounds_ad = None
    if rank == 0:
        rr = recs_host.numpy().view(b200vo.POSE_RECORD_DTYPE).copy()
        rounds_ad = float((((rr["status"] >> 16) & 0xFF) * 128).mean())
        del rr
    stream.synchronize()
    ad_ctx, ad_stream = ctx, stream          # closed at teardown, after the tensors that touched its stream
    ctx, stream = main_ctx, stream_main
    step(True)                      # restore the fixed-H records / trajectory for the result check
    stream.synchronize()

    # ---- per-stage breakdown (separate instrumented pass; timing mode synchronises per call)
    stage = {}
    if rank == 0:
        ctx.enable_timing(True)
        tot = {}
        done_frames, done_pairs, prev_slot = 0, 0, 0
        while done_frames < F:
            n = min(B, F - done_frames)
            first = 0 if done_frames == 0 else (prev_slot + 1) % (B + 1)
            ctx.upload_frames_device(d_frames.data_ptr() + done_frames * frame_bytes, n, W_IMG, frame_bytes, 1, first)
            ctx.detect(first, n)
            st = ctx.stage_times()
            for k in ("pyramid_ms", "fast_ms", "select_ms", "orient_ms", "blur_ms", "brief_ms"):
                tot[k] = tot.get(k, 0.0) + st[k]
            p0, np_ = (first, n - 1) if done_frames == 0 else (prev_slot, n)
            if np_ > 0:
                ctx.process_pairs_device(p0, np_, f0 + done_pairs)
                done_pairs += np_
                st = ctx.stage_times()
                for k in ("match_ms", "filter_ms", "ransac_ms", "pose_ms", "fivept_ms", "score_ms"):
                    tot[k] = tot.get(k, 0.0) + st[k]
            prev_slot = (first + n - 1) % (B + 1)
            done_frames += n
        ctx.enable_timing(False)
        stage = tot

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        del recs_dev, recs_all, d_frames, recs_host, h_frames, stream, ad_stream, stream_main
        import gc
        gc.collect()
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        ad_ctx.close()
        ctx.close()
        return 0

    hbm_peak, peak_src = peaks()
    popc_rate = ctx.measure_popc_rate()
    alg = algorithmic_bytes(ctx, NF)
    stages_out = {}
    for name, key in (("pyramid", "pyramid_ms"), ("fast", "fast_ms"), ("blur", "blur_ms"), ("brief", "brief_ms"), ("orient", "orient_ms")):
        ms = stage.get(key, 0.0)
        gbs = alg[name] * F / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        stages_out[name] = {"ms": round(ms, 3), "bound": "hbm", "algorithmic_bytes_per_frame": alg[name], "achieved_GBps": round(gbs, 1), "frac": round(gbs / hbm_peak, 4)}
    # matcher: N1*N2 descriptor pairs; algorithmically 8 popc32 each.  The kernel executes 5 (three
    # carry-save adders move the rest onto the LOP3 pipe), so the POPC-pipe roofline is taken on the
    # EXECUTED population counts against the rate measured by the XOR+POPC+ADD microbenchmark.
    kp_pairs = float(NF) * NF * npairs
    mm = stage.get("match_ms", 0.0)
    popc_per = {0: 8, 1: 6, 2: 4, 3: 5}[int(os.environ.get("B200VO_MATCH_CSA", "3"))]
    popc_exec = kp_pairs * popc_per / (mm * 1e-3) if mm > 0 else 0.0
    stages_out["match"] = {"ms": round(mm, 3), "bound": "int_pipe(popc)", "algorithmic_popc32_per_pair": 8, "executed_popc32_per_pair": popc_per,
                           "achieved_Gpopc_s": round(popc_exec / 1e9, 1), "peak_Gpopc_s": round(popc_rate / 1e9, 1),
                           "frac": round(popc_exec / popc_rate, 4), "descriptor_pairs_per_s": round(kp_pairs / (mm * 1e-3), 0) if mm > 0 else 0}
    for name in ("select", "filter", "ransac", "pose", "fivept", "score"):
        stages_out[name] = {"ms": round(stage.get(name + "_ms", 0.0), 3)}
    # FP64 work of the scoring kernel: models x points x ~20 DFMA against 64 DFMA/clk/SM
    kernels = {k: v for k, v in stage.items() if k not in ("ransac_ms",)}
    total_stage_ms = sum(kernels.values())
    dom = max(kernels, key=kernels.get) if kernels else "match_ms"
    if dom == "match_ms":
        roof = {"kernel": "match_kernel<4,%d>" % int(os.environ.get("B200VO_MATCH_CSA", "3")), "bound": "int_pipe", "achieved": round(popc_exec / 1e9, 1), "peak": round(popc_rate / 1e9, 1),
                "unit": "Gpopc32/s", "frac": round(popc_exec / popc_rate, 4),
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch (500-frame batch) from the ncu --set full
                # capture profiles/ncu_r01_match_fullconfig.txt; algorithmic: 500 x 3001 x 32 B = 48 MB with every
                # descriptor read once (each slot is the query side of one pair and the train side of the next)
                "traffic": MATCH_DRAM_BYTES_PER_LAUNCH if (F, NF, B) == (1000, 3000, 500) else None, "traffic_unit": "bytes/launch",
                "peak_source": "XOR+POPC+ADD microbenchmark measured in this run (b200vo_measure_popc_rate); not an HBM- or tensor-bound kernel "
                               "(63 MB of DRAM traffic per 5.6 ms launch)",
                "share_of_step": round(mm / total_stage_ms, 3)}
    else:
        nm = dom[:-3]
        sdict = stages_out.get(nm, {})
        roof = {"kernel": nm, "bound": sdict.get("bound", "issue"), "achieved": sdict.get("achieved_GBps"), "peak": hbm_peak, "unit": "GB/s",
                "frac": sdict.get("frac"), "traffic": None, "peak_source": peak_src, "share_of_step": round(stage[dom] / total_stage_ms, 3)}
    hb = max(("pyramid", "fast", "blur", "brief"), key=lambda n: stages_out[n]["ms"])
    roof_hbm = {"kernel": hb, "bound": "hbm", "achieved": stages_out[hb]["achieved_GBps"], "peak": hbm_peak, "unit": "GB/s",
                "frac": stages_out[hb]["frac"], "traffic": None, "peak_source": peak_src}

    total_pairs = world * npairs
    value = total_pairs / (ms_dev * 1e-3)
    e2e = total_pairs / (ms_e2e * 1e-3)

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            ncpu = os.cpu_count() or 1
            pairs = ncpu * (args.cpu_pairs or 8)
            fr = h_frames[:min(F, pairs + 1)].numpy().copy()      # a copy: no view may outlive the pinned tensor
            v, n, dt = cpu_reference_throughput(fr, NF, min(pairs, len(fr) - 1), ncpu)
            import cv2
            cpu = {"value": round(v, 2), "unit": "pairs/s", "cores": min(ncpu, n), "kind": "port",
                   "sample": "%d pairs of the same frames, %d processes x cv2.setNumThreads(1) (cv2 %s calling the reference's OpenCV functions), %.1f s" % (n, min(ncpu, n), cv2.__version__, dt)}
        except Exception as ex:      # the CPU leg must never take the GPU numbers down with it
            cpu = {"value": None, "unit": "pairs/s", "cores": 0, "kind": "port", "sample": "failed: %s" % ex}

    traj = state["traj"]
    ok = recs_host.numpy().view(b200vo.POSE_RECORD_DTYPE).copy()
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_dev, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/i32 (features, matching), f64 (geometry)", "data": "synthetic",
        "config": {"workload": "%d-frame synthetic KITTI-shaped sequence %dx%d per GPU, %d ORB features, %d five-point hypotheses/pair, launch batches of %d frames" % (F, W_IMG, H_IMG, NF, args.hypotheses, B),
                   "pairs_per_step": total_pairs, "cache": "inputs larger than L2 (%.0f MB of frames per step)" % (F * frame_bytes / 1e6),
                   "partition": "contiguous frame-pair ranges, one per GPU; one NCCL all_gather of 160-B pose records" if world > 1 else "single GPU"},
        "e2e": {"value": round(e2e, 1), "unit": "pairs/s", "ms_per_step": round(ms_e2e, 3), "h2d_bytes_per_step": world * F * frame_bytes,
                "d2h_bytes_per_step": world * npairs * RB, "wall_ms_per_step": round(wall_e2e * 1e3, 3)},
        "gpu_launches": int(launches) * world,
        "clocks": clocks,
        "roofline": roof, "roofline_hbm": roof_hbm, "stages": stages_out,
        "adaptive_ransac": {"note": "same workload with ransac_adaptive=1 (OpenCV confidence-0.999 stopping rule, checked every 128 hypotheses)",
                            "value": round(total_pairs / (ms_dev_ad * 1e-3), 1), "e2e": round(total_pairs / (ms_e2e_ad * 1e-3), 1),
                            "mean_hypotheses_per_pair": rounds_ad},
        "cpu_baseline": cpu,
        "result_check": dict({"pairs_pose_ok": int(((ok["status"] & 4) != 0).sum()), "pairs_total": int(len(ok)),
                              "trajectory_end": [round(float(x), 3) for x in traj[-1][9:12]] if traj is not None else None},
                             **(accuracy(traj, R_all, C_all) if traj is not None else {})),
        "wall_ms_per_step": round(wall_dev * 1e3, 3),
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    # torch tensors that were used on the context's stream must be released before the stream dies
    del recs_dev, recs_all, d_frames, recs_host, h_frames, ok, stream, ad_stream, stream_main
    import gc
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    ad_ctx.close()
    ctx.close()
    return 0


_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else any library prints (NCCL's version
    banner, torchrun notices) was redirected to stderr, so stdout carries exactly this line."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    args = parse_args()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                   # fd 1 -> stderr for the duration of the run (C libraries included)
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    rc = main()
    sys.stdout.flush()
    sys.stderr.flush()
    sys.exit(rc)
