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
Multi-GPU (torchrun, one process per GPU): the headline is weak scaling (each rank owns its own F-1 pairs of
one long global sequence); the only exchange is one NCCL all_gather of the 160-byte pose records, then rank 0
chains the trajectory on the host.  After the headline the same run adds
  `config3` : ONE 1000-frame 1920x1080 sequence, 8000 features, partitioned over the N ranks by contiguous
              frame-pair ranges (strong scaling; BASELINE.json configs[2]), and
  `config5` : ONE 256-frame 3840x2160 sequence, 20000 features, 4096 hypotheses per pair (configs[4]),
each with the whole sequence also run on rank 0 alone, so the line carries the strong-scaling efficiency and a
byte-comparison of the partitioned records against the one-GPU records.
Further keys: `streaming` (the reference's real calling pattern: VisualOdometry::process_frame once per frame
through libb200vo_host.so), `h2d` (per-rank upload bandwidth in the e2e leg and of a plain pinned copy),
`keypoints_per_frame`, per-stage roofline table.
--impl reference: the reference's own CPU path (the same OpenCV calls through cv2, oracle/pipeline.py)
on the host cores, one worker process per core over contiguous pair ranges.
"""
import argparse
import ctypes
import itertools
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
# dram__bytes_read.sum + dram__bytes_write.sum of one match_kernel launch (500-frame batch) from the ncu --set full capture
# named in profiles/README.md
MATCH_DRAM_BYTES_PER_LAUNCH = 61690000      # profiles/ncu_r02_match_fullconfig.txt: 60.1 MB read + 1.4-1.8 MB written per 5.65 ms launch
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
    ap.add_argument("--cpu-pairs", type=int, default=0, help="pairs per core in the bounded CPU sample (0 = 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline only: skip the config3 / config5 / streaming / adaptive legs")
    ap.add_argument("--c3-frames", type=int, default=1000)
    ap.add_argument("--c5-frames", type=int, default=256)
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


# --------------------------------------------------------------------------- accuracy helpers
def pair_errors(recs, R_gt, C_gt, first_pair=0):
    """Per-pair relative-pose error against the synthetic ground truth over the pairs with a valid pose:
    rotation angle and angle between the translation directions (deg).  recs: sequence of mappings with
    status / R / t (b200vo records or the cv2 pipeline's dicts)."""
    from b200vo import synth
    from b200vo import trajectory as T
    rot, tdir = [], []
    for i, r in enumerate(recs):
        if not (int(r["status"]) & 4):
            continue
        Rg, tg = synth.relative_pose_gt(R_gt, C_gt, first_pair + i)
        R = np.asarray(r["R"], np.float64).reshape(3, 3)
        t = np.asarray(r["t"], np.float64).ravel()
        rot.append(T.rotation_angle_deg(R, Rg))
        tdir.append(float(np.degrees(np.arccos(np.clip(t @ tg / max(np.linalg.norm(t), 1e-300), -1.0, 1.0)))))
    if not rot:
        return {"pairs": 0}
    return {"pairs": len(rot), "rot_deg_mean": round(float(np.mean(rot)), 4), "rot_deg_median": round(float(np.median(rot)), 4),
            "trans_dir_deg_mean": round(float(np.mean(tdir)), 3), "trans_dir_deg_median": round(float(np.median(tdir)), 3)}


def accuracy(traj, R_gt, C_gt):
    """Chained trajectory against the ground truth of the synthetic sequence: similarity-aligned ATE
    (the reference applies a fixed 0.3 step scale) and mean relative rotation / translation-direction
    error per pair.  Frame 0 is the identity pose the chain starts from."""
    from b200vo import trajectory as T
    poses = np.concatenate([np.concatenate([np.eye(3).ravel(), np.zeros(3)])[None], np.asarray(traj)], 0)
    n = len(poses)
    R0, C0 = R_gt[0], C_gt[0]
    Rg = np.stack([R0.T @ R for R in R_gt[:n]])
    Cg = np.stack([R0.T @ (C - C0) for C in C_gt[:n]])
    return T.accuracy_report(poses, Rg, Cg)


# --------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    frames, nfeatures, nthreads = args
    import cv2
    cv2.setNumThreads(nthreads)
    from oracle import pipeline as PL
    t0 = time.perf_counter()
    recs, _ = PL.run_pairs_cv2(frames, nfeatures)
    dt = time.perf_counter() - t0
    slim = [{"status": int(r["status"]), "R": np.asarray(r["R"], np.float64).reshape(9), "t": np.asarray(r["t"], np.float64).reshape(3)} for r in recs]
    return len(recs), dt, slim


def cpu_reference_throughput(frames_host, nfeatures, pairs_total, nproc):
    """The reference path (cv2) over `pairs_total` pairs split into contiguous ranges, one process
    per core with cv2.setNumThreads(1).  Returns (pairs/s, pairs, seconds, records in pair order)."""
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
    recs = [x for r in res for x in r[2]]
    return pairs / dt, pairs, dt, recs


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
    recs = []
    for i in range(args.warmup + args.steps):
        v, n, dt, recs = cpu_reference_throughput(frames, args.features, pairs, ncpu)
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
            "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "result_check": {"vs_ground_truth_pairs": pair_errors(recs, R_wc, C_w)}}
    emit(line)
    return 0


# --------------------------------------------------------------------------- B200 arm
class DevView:
    """Zero-copy torch view of a raw device pointer."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPU cores NVML reports as local to its GPU BEFORE the pinned frame buffer is
    allocated (first touch decides the NUMA node).  Best effort; returns (cpus bound, cpus NVML named)."""
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
        named = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus = named & set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus), len(named)
    except Exception:
        return 0, 0


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


def batch_plan(F, B, sizes):
    """Launch batches of an F-frame sequence in a ring of B + 1 slots: (first slot, frames in the batch, frames done
    before, first pair slot, pairs, pairs done before).  `sizes` gives the first batch sizes; after it full batches."""
    done_f, done_p, prev = 0, 0, 0
    for n in itertools.chain(sizes, itertools.repeat(B)):
        n = min(n, F - done_f)
        if n <= 0:
            break
        if F - done_f - n == 1:                                  # never leave a single trailing frame to a batch of its own
            n += 1 if n < B else -1
        first = 0 if done_f == 0 else (prev + 1) % (B + 1)
        p0, np_ = (first, n - 1) if done_f == 0 else (prev, n)
        yield first, n, done_f, p0, np_, done_p
        prev = (first + n - 1) % (B + 1)
        done_f += n
        done_p += np_


COPY_BY_KERNEL = os.environ.get("B200VO_BENCH_COPY_KERNEL", "0") != "0"      # experiment knob: no effect measured (profiles/README.md)


class SeqRunner:
    """One rank's share of a sequence through the C ABI: frames [0, F) of `d_frames` (device) / `h_frames`
    (pinned host) in launch batches of B frames; pair ids start at pair0.  Records land in self.recs_dev."""

    def __init__(self, torch, b200vo, W, H, NF, hyp, B, local, d_frames, h_frames, pair0, recs_dev=None, **opts):
        self.torch, self.b = torch, b200vo
        self.W, self.H, self.F, self.B = W, H, int(d_frames.shape[0]), max(2, min(B, int(d_frames.shape[0])))
        self.npairs = self.F - 1
        self.pair0 = pair0
        self.dev = "cuda:%d" % local
        self.d_frames, self.h_frames = d_frames, h_frames
        self.ctx = b200vo.Context(W, H, device=local, nfeatures=NF, max_frames=self.B + 1, ransac_hypotheses=hyp, **opts)
        self.stream = torch.cuda.ExternalStream(self.ctx.stream(), device=torch.device("cuda", local))
        self.RB = b200vo.RECORD_BYTES
        self.recs_dev = recs_dev if recs_dev is not None else torch.zeros(max(self.npairs, 1) * self.RB, dtype=torch.uint8, device=self.dev)
        self.fb = W * H
        self.e2e_sizes = None

    def kernel_launches(self):
        return self.ctx.kernel_launches()

    def e2e_candidates(self):
        return e2e_schedule_candidates(self.F, self.B)

    def batches(self, short_first=False):
        """short_first (the e2e leg) follows self.e2e_sizes, the schedule pick_e2e_schedule chose."""
        return batch_plan(self.F, self.B, self.e2e_sizes if (short_first and self.e2e_sizes) else [])

    def step(self, host_input):
        for _ in self.step_iter(host_input):
            pass

    def step_iter(self, host_input):
        """Enqueue the step batch by batch (a generator: MultiRunner interleaves several of them)."""
        torch, ctx = self.torch, self.ctx
        for first, n, done_f, p0, np_, done_p in self.batches(short_first=host_input):
            if host_input:
                ctx.upload_frames_ptr(self.h_frames.data_ptr() + done_f * self.fb, n, self.W, self.fb, 1, first)
            else:
                ctx.upload_frames_device(self.d_frames.data_ptr() + done_f * self.fb, n, self.W, self.fb, 1, first)
            ctx.detect(first, n)
            if np_ > 0:
                ptr = ctx.process_pairs_device(p0, np_, self.pair0 + done_p)
                with torch.cuda.stream(self.stream):
                    src = torch.as_tensor(DevView(ptr, np_ * self.RB), device=self.dev)
                    dst = self.recs_dev[done_p * self.RB:(done_p + np_) * self.RB]
                    if COPY_BY_KERNEL:
                        torch.bitwise_or(src, 0, out=dst)      # an elementwise kernel: a device-to-device cudaMemcpyAsync may be served by
                    else:                                      # a copy engine and then queues behind the upload that is in flight
                        dst.copy_(src, non_blocking=True)
            yield

    def close(self):
        """Drop the tensors; the context itself (and with it the CUDA stream torch has seen) is destroyed only at
        the very end of the run (close_contexts): torch's pinned-memory allocator records an event on every stream a
        pinned tensor was used on when that tensor is freed, so no stream may die before the last such tensor."""
        self.stream.synchronize()
        self.stream = None
        self.recs_dev = None
        self.d_frames = self.h_frames = None
        _OPEN_CONTEXTS.append(self.ctx)


class MultiRunner:
    """The same share of a sequence over K contexts (K CUDA streams) of ONE GPU: context k owns the contiguous pair range
    [cuts[k], cuts[k+1]) -- the partitioning the multi-GPU path uses, applied inside a GPU -- and the launch batches of the
    contexts are enqueued round-robin.  Every batch ends in latency-bound kernels (filter, deferred five-point, recoverPose:
    a few CTAs each) and every kernel in a partially filled last wave; with a second stream those overlap the other
    context's throughput-bound kernels (tools/two_contexts.py: 28.02 -> 27.33 ms per 999 pairs with K = 2; 27.28 with 3).
    One boundary frame per cut is detected twice.  Fork / join: the other streams wait on an event of the main stream at
    the start of a step and the main stream waits on theirs at its end, so CUDA events on the main stream time the whole
    step and everything enqueued on it afterwards (gather, D2H) sees all records."""

    def __init__(self, torch, b200vo, W, H, NF, hyp, B, local, d_frames, h_frames, pair0, K=2, **opts):
        F = int(d_frames.shape[0])
        K = max(1, min(K, (F - 1) // 8 if F > 8 else 1))
        self.torch, self.K = torch, K
        self.F, self.npairs = F, F - 1
        self.dev = "cuda:%d" % local
        self.RB = b200vo.RECORD_BYTES
        self.recs_dev = torch.zeros(max(self.npairs, 1) * self.RB, dtype=torch.uint8, device=self.dev)
        self.cuts = [round(k * self.npairs / K) for k in range(K + 1)]
        self.sub = []
        for k in range(K):
            lo, hi = self.cuts[k], self.cuts[k + 1]
            # a launch batch may take the K - 1 boundary frames on top of B, so an even split of K * B frames stays one batch per context
            self.sub.append(SeqRunner(torch, b200vo, W, H, NF, hyp, B + K - 1, local, d_frames[lo:hi + 1], h_frames[lo:hi + 1], pair0 + lo,
                                      recs_dev=self.recs_dev[lo * self.RB:hi * self.RB], **opts))
        self.B = self.sub[0].B
        self.ctx, self.stream = self.sub[0].ctx, self.sub[0].stream
        self.ev_fork = torch.cuda.Event()
        self.ev_join = [torch.cuda.Event() for _ in range(K)]
        self.chained = False        # e2e option: context k starts after context k-1 (its first upload still runs ahead on its copy stream)
        self.chained_dev = False    # the same for the device-resident step: the contexts one after the other (= one stream)

    @property
    def e2e_sizes(self):
        return getattr(self, "_sizes", None)

    @e2e_sizes.setter
    def e2e_sizes(self, sizes):
        per_ctx = bool(sizes) and isinstance(sizes[0], (list, tuple))       # one schedule for all, or one per context
        for k, r in enumerate(self.sub):
            r.e2e_sizes = list(sizes[k]) if per_ctx else sizes
        self._sizes = sizes

    def e2e_candidates(self):
        """The single-context candidates for every context alike, plus staggered ones: the batches reach the copy engine
        in round-robin order (context 0's first, context 1's first, context 0's second, ...), each hidden behind the
        compute of its predecessors only if it is at most ~3x as long (8.5 us of upload vs 29 us of compute per frame
        at the headline configuration), and only the very first upload is exposed -- so context 0 opens with a small
        batch and context 1 with one about three times as large."""
        Fs, B = max(r.F for r in self.sub), self.B
        cands = list(e2e_schedule_candidates(Fs, B))
        if self.K == 2:
            for a0, a1 in ((B // 16, B // 5), (B // 12, B // 4), (B // 24, B // 8)):
                a0, a1 = max(2, a0), max(3, a1)
                cands.append([[a0, min(B, 9 * a0)], [a1]])
                cands.append([[a0], [a1]])
        return cands

    def kernel_launches(self):
        return sum(r.ctx.kernel_launches() for r in self.sub)

    def step(self, host_input):
        if (host_input and self.chained) or (not host_input and self.chained_dev):
            # one context after the other on the device; what remains of the split is that the first upload of every later
            # context is hidden behind its predecessor's compute
            for k, r in enumerate(self.sub):
                if k:
                    r.stream.wait_event(self.ev_join[k - 1])
                r.step(host_input)
                self.ev_join[k].record(r.stream)
            self.stream.wait_event(self.ev_join[self.K - 1])
            return
        self.ev_fork.record(self.stream)
        for r in self.sub[1:]:
            r.stream.wait_event(self.ev_fork)
        its = [r.step_iter(host_input) for r in self.sub]
        while its:
            its = [it for it in its if next(it, StopIteration) is not StopIteration]
        for k, r in enumerate(self.sub[1:], 1):
            self.ev_join[k].record(r.stream)
            self.stream.wait_event(self.ev_join[k])

    def close(self):
        for r in self.sub:
            r.close()
        self.sub = []
        self.ctx = self.stream = None
        self.recs_dev = None


def make_runner(torch, b200vo, W, H, NF, hyp, B, local, d_frames, h_frames, pair0, **opts):
    """B200VO_BENCH_STREAMS contexts per GPU (default 2); 1 = the plain single-stream runner."""
    K = int(os.environ.get("B200VO_BENCH_STREAMS", "2"))
    if K <= 1 or int(d_frames.shape[0]) < 17:
        return SeqRunner(torch, b200vo, W, H, NF, hyp, B, local, d_frames, h_frames, pair0, **opts)
    return MultiRunner(torch, b200vo, W, H, NF, hyp, B, local, d_frames, h_frames, pair0, K=K, **opts)


_OPEN_CONTEXTS = []


def close_contexts(torch):
    import gc
    gc.collect()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    try:
        torch._C._host_emptyCache()          # return the pinned blocks now, while every stream they were used on still exists
    except Exception:
        pass
    while _OPEN_CONTEXTS:
        _OPEN_CONTEXTS.pop().close()


def e2e_schedule_candidates(F, B):
    """Launch-batch schedules for the e2e leg.  The upload of batch k+1 is enqueued on the copy stream right behind batch
    k's, so it hides behind the compute of batch k when it is not longer; what stays exposed is the first upload, every
    positive (copy(k+1) - compute(k)), and a fixed cost per batch (launch tails: 0.6 ms at the headline configuration).
    Which schedule wins depends on the upload rate the host gives this rank (55 GB/s alone, 20-35 GB/s with eight ranks
    copying): `B/5, B, ...` for a fast link, ramps of x2 from B/6 or x1.6 from B/4 for a slow one
    (tools/e2e_schedules.py, profiles/e2e_schedules_r02_{1,8}gpu.jsonl).  bench.py times each candidate on untimed
    warm-up steps and keeps the fastest (max over ranks, so every rank keeps the same one)."""
    def ramp(start, growth):
        sizes, r, done = [], float(max(2, start)), 0
        while done < F:
            n = min(B, F - done, int(r))
            if F - done - n == 1:              # never leave a single trailing frame to a batch of its own
                n += 1 if n < B else -1
            sizes.append(n)
            done += n
            r = r * growth + 1
        return sizes
    cands = [ramp(B // 5, 1e9), ramp(B // 6, 2.0), ramp(B // 4, 1.6), ramp(B, 1.0)]
    out = []
    for c in cands:
        if c not in out:
            out.append(c)
    return out


def pick_dev_mode(torch, dist, world, runner, step_fn, dev):
    """Device-resident step of a MultiRunner: contexts concurrent or one after the other, whichever is faster on untimed
    steps (the overlap of two streams pays when the launch batches are small -- config 3 / 5 -- and is a wash at the
    headline configuration since the kernels run at higher occupancy)."""
    if not hasattr(runner, "chained_dev"):
        return
    t = {False: [], True: []}
    for mode in (False, True, False, True):               # ABAB: the two modes differ by ~1 %, box noise is of that order
        runner.chained_dev = mode
        ms, _, _ = timed_steps(torch, dist, world, runner.stream, step_fn, 3, 1, dev)
        t[mode].append(ms)
    runner.chained_dev = min(t[True]) <= min(t[False])


def pick_e2e_schedule(torch, dist, world, runner, step_fn, dev):
    """Time every candidate schedule for two steps after one settling step; keep the fastest."""
    best = None
    modes = (False, True) if hasattr(runner, "chained") else (False,)
    for chained in modes:
        if len(modes) > 1:
            runner.chained = chained
        for sizes in runner.e2e_candidates():
            runner.e2e_sizes = sizes
            ms, _, _ = timed_steps(torch, dist, world, runner.stream, step_fn, 2, 1, dev)
            if best is None or ms < best[0]:
                best = (ms, sizes, chained)
    runner.e2e_sizes = best[1]
    if len(modes) > 1:
        runner.chained = best[2]
    return best[1]


def timed_steps(torch, dist, world, runner_stream, fn, steps, warmup, dev, launches_of=None):
    """warm-up, barrier + synchronize, CUDA events on the launching stream around exactly `steps` calls,
    barrier + synchronize, MAX over ranks.  Returns (ms per step, wall s per step, kernel launches)."""
    for _ in range(warmup):
        fn()
    runner_stream.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = launches_of.kernel_launches() if launches_of is not None else 0
    t0 = time.perf_counter()
    e0.record(runner_stream)
    for _ in range(steps):
        fn()
    e1.record(runner_stream)
    runner_stream.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms / steps, wall / steps, (launches_of.kernel_launches() - l0) if launches_of is not None else 0


def h2d_probe(torch, dist, world, d_like, h_frames):
    """GB/s of 3 back-to-back pinned -> device copies of h_frames on this rank, all ranks copying at once."""
    dst = torch.empty_like(d_like)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        dst.copy_(h_frames, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbps = 3 * h_frames.numel() / (e0.elapsed_time(e1) * 1e-3) / 1e9
    del dst
    return gbps


def render(torch, synth, W, H, R, C, tex, dev, chunk):
    d = synth.render_frames(W, H, R, C, tex, device=dev, chunk=chunk)
    h = torch.empty(tuple(d.shape), dtype=torch.uint8, pin_memory=True)
    h.copy_(d)
    torch.cuda.synchronize()
    return d, h


def strong_scaling_leg(torch, dist, b200vo, synth, tex, name, W, H, NF, hyp, B, F, rank, world, local, steps, warmup):
    """ONE F-frame sequence partitioned over the ranks by contiguous frame-pair ranges (b200vo/sequence.py), one
    all_gather of the records; then the same sequence on rank 0 alone.  Returns the leg's JSON object (rank 0)."""
    from b200vo import sequence as SQ
    dev = "cuda:%d" % local
    RB = b200vo.RECORD_BYTES
    P = F - 1
    R_all, C_all = synth.make_trajectory(F)
    lo, hi = SQ.partition_pairs(P, world, rank)
    f0, f1 = SQ.frames_for_pairs(lo, hi)
    chunk = max(1, (32 * 1241 * 376) // (W * H))
    d_fr, h_fr = render(torch, synth, W, H, R_all[f0:f1], C_all[f0:f1], tex, dev, chunk)
    run = make_runner(torch, b200vo, W, H, NF, hyp, B, local, d_fr, h_fr, lo)
    per = -(-P // world)
    send = torch.zeros(per * RB, dtype=torch.uint8, device=dev)
    gath = torch.zeros(world * per * RB, dtype=torch.uint8, device=dev) if world > 1 else send
    host = torch.empty(world * per * RB, dtype=torch.uint8, pin_memory=True)

    def step(host_input):
        run.step(host_input)
        with torch.cuda.stream(run.stream):
            send[:(hi - lo) * RB].copy_(run.recs_dev[:(hi - lo) * RB], non_blocking=True)
            if world > 1:
                dist.all_gather_into_tensor(gath, send)
            if host_input:
                host.copy_(gath, non_blocking=True)
        if host_input:
            run.stream.synchronize()

    pick_dev_mode(torch, dist, world, run, lambda: step(False), dev)
    ms_dev, _, launches = timed_steps(torch, dist, world, run.stream, lambda: step(False), steps, warmup, dev, run)
    pick_e2e_schedule(torch, dist, world, run, lambda: step(True), dev)
    ms_e2e, _, _ = timed_steps(torch, dist, world, run.stream, lambda: step(True), steps, warmup, dev)
    out, recs = None, None
    if rank == 0:
        allb = host.numpy().view(b200vo.POSE_RECORD_DTYPE).reshape(world, per)
        recs = np.concatenate([allb[r, :SQ.partition_pairs(P, world, r)[1] - SQ.partition_pairs(P, world, r)[0]] for r in range(world)]).copy()
        del allb
        out = {"workload": "%s: ONE %d-frame %dx%d sequence, %d features, %d hypotheses/pair, launch batches of %d frames, partitioned over %d GPU(s) by contiguous frame-pair ranges" % (name, F, W, H, NF, hyp, run.B, world),
               "pairs": P, "scaling": "strong", "value": round(P / (ms_dev * 1e-3), 1), "e2e": round(P / (ms_e2e * 1e-3), 1), "unit": "pairs/s",
               "ms_per_step": round(ms_dev, 3), "e2e_ms_per_step": round(ms_e2e, 3), "steps": steps, "warmup": warmup,
               "h2d_bytes_per_step": (P + world) * W * H, "d2h_bytes_per_step": world * per * RB, "e2e_batch_sizes_rank0": run.e2e_sizes, "streams_per_gpu": getattr(run, "K", 1), "e2e_contexts_chained": bool(getattr(run, "chained", False)),
               "device_resident_contexts_chained": bool(getattr(run, "chained_dev", False)),
               "boundary_frames_detected_twice": world - 1, "boundary_overhead_frac": round((world - 1) / float(P + 1), 5),
               "gpu_launches": int(launches) * world,
               "pairs_pose_ok": int(((recs["status"] & 4) != 0).sum()), "pairs_overflow": int(((recs["status"] & 256) != 0).sum()),
               "vs_ground_truth_pairs": pair_errors(recs, R_all, C_all)}
    run.close()
    del d_fr, h_fr, send, gath
    # the same sequence on ONE GPU (rank 0; the other ranks wait): denominator of the strong-scaling efficiency and
    # the byte-identity check of the partition
    if world > 1:
        if rank == 0:
            d_all, h_all = render(torch, synth, W, H, R_all, C_all, tex, dev, chunk)
            solo = make_runner(torch, b200vo, W, H, NF, hyp, B, local, d_all, h_all, 0)
            pick_dev_mode(torch, None, 1, solo, lambda: solo.step(False), dev)
            ms1, _, _ = timed_steps(torch, None, 1, solo.stream, lambda: solo.step(False), steps, warmup, dev)
            host1 = torch.empty(P * RB, dtype=torch.uint8, pin_memory=True)

            def e2e1():
                solo.step(True)
                with torch.cuda.stream(solo.stream):
                    host1.copy_(solo.recs_dev[:P * RB], non_blocking=True)
                solo.stream.synchronize()
            pick_e2e_schedule(torch, None, 1, solo, e2e1, dev)
            ms1e, _, _ = timed_steps(torch, None, 1, solo.stream, e2e1, steps, warmup, dev)
            recs1 = host1.numpy().view(b200vo.POSE_RECORD_DTYPE).copy()
            out["single_gpu"] = {"value": round(P / (ms1 * 1e-3), 1), "e2e": round(P / (ms1e * 1e-3), 1), "ms_per_step": round(ms1, 3)}
            out["efficiency_vs_single_gpu"] = {"value": round(ms1 / (world * ms_dev), 4), "e2e": round(ms1e / (world * ms_e2e), 4)}
            out["records_identical_to_single_gpu"] = bool(recs1.tobytes() == recs.tobytes())
            solo.close()
            del d_all, h_all, host1
        dist.barrier()
    del host
    return out


def streaming_leg(b200vo, frames_host, nfeatures, nframes):
    """The reference's real calling pattern (src/main.cpp:155): VisualOdometry::process_frame once per frame,
    host image in, keypoints / descriptors / pose out, through the C++ class in libb200vo_host.so."""
    lib = ctypes.CDLL(b200vo.HOST_LIB_PATH)
    lib.vo_host_create.restype = ctypes.c_void_p
    lib.vo_host_last_error.restype = ctypes.c_char_p
    lib.vo_host_destroy.argtypes = [ctypes.c_void_p]
    H, W = frames_host.shape[1:3]
    h = lib.vo_host_create(W, H, nfeatures, 0, ctypes.c_ulonglong(42), 0, 0)
    if not h:
        return {"error": lib.vo_host_last_error().decode()}
    h = ctypes.c_void_p(h)
    pose = np.zeros(12)
    moved = ctypes.c_int()
    t = []
    for i in range(nframes):
        f = np.ascontiguousarray(frames_host[i])
        t0 = time.perf_counter()
        rc = lib.vo_host_process_frame(h, i, f.ctypes.data_as(ctypes.c_void_p), W, H, 1, pose.ctypes.data_as(ctypes.c_void_p), ctypes.byref(moved))
        t.append(time.perf_counter() - t0)
        if rc <= 0:
            lib.vo_host_destroy(h)
            return {"error": lib.vo_host_last_error().decode()}
    lib.vo_host_destroy(h)
    t = np.array(t[10:])
    return {"frames": int(len(t)), "frames_per_s": round(1.0 / float(t.mean()), 1), "ms_per_frame_mean": round(float(t.mean()) * 1e3, 4),
            "ms_per_frame_p50": round(float(np.median(t)) * 1e3, 4), "ms_per_frame_p99": round(float(np.percentile(t, 99)) * 1e3, 4)}


def run_b200(args):
    import torch
    import torch.distributed as dist
    import b200vo
    from b200vo import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    F, NF, B = args.frames, args.features, min(args.batch, args.frames)
    npairs = F - 1
    RB = b200vo.RECORD_BYTES
    # global sequence: rank r owns pairs [r*(F-1), (r+1)*(F-1)) -> frames [r*(F-1), r*(F-1)+F)
    tex = synth.make_texture()
    R_all, C_all = synth.make_trajectory(world * npairs + 1)
    f0 = rank * npairs
    d_frames, h_frames = render(torch, synth, W_IMG, H_IMG, R_all[f0:f0 + F], C_all[f0:f0 + F], tex, dev, 32)
    frame_bytes = W_IMG * H_IMG

    run = make_runner(torch, b200vo, W_IMG, H_IMG, NF, args.hypotheses, B, local, d_frames, h_frames, f0)
    ctx, stream = run.ctx, run.stream
    streams_per_gpu = getattr(run, "K", 1)
    run_mode = run
    gathered = {"t": torch.zeros(world * npairs * RB, dtype=torch.uint8, device=dev) if world > 1 else None}
    recs_host = torch.empty(world * npairs * RB, dtype=torch.uint8, pin_memory=True)
    state = {"traj": None}

    def make_step(r, host_input):
        def fn():
            r.step(host_input)
            with torch.cuda.stream(r.stream):
                if world > 1:
                    dist.all_gather_into_tensor(gathered["t"], r.recs_dev)     # the one exchange step: 160 B per pair
                if host_input:
                    recs_host.copy_(gathered["t"] if world > 1 else r.recs_dev, non_blocking=True)
            if host_input:
                r.stream.synchronize()
                if rank == 0:   # host-side chaining of the whole trajectory (reference visual_odometry.cpp:340-374)
                    state["traj"] = b200vo.chain_trajectory(recs_host.numpy().view(b200vo.POSE_RECORD_DTYPE).copy())
        return fn

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    pick_dev_mode(torch, dist, world, run, make_step(run, False), dev)
    ms_dev, wall_dev, launches = timed_steps(torch, dist, world, stream, make_step(run, False), args.steps, args.warmup, dev, run)
    # ---- per-rank upload bandwidth of a plain pinned -> device copy of the same frames with all ranks copying at the same
    # time (the ceiling the host's PCIe / memory system gives this box)
    probe_gbps = h2d_probe(torch, dist, world, d_frames, h_frames)
    pick_e2e_schedule(torch, dist, world, run, make_step(run, True), dev)
    ms_e2e, wall_e2e, _ = timed_steps(torch, dist, world, stream, make_step(run, True), args.steps, max(args.warmup, 3), dev)
    clocks = sampler.stop() if rank == 0 else None
    ok = recs_host.numpy().view(b200vo.POSE_RECORD_DTYPE).copy() if rank == 0 else None
    traj = state["traj"]

    h2d_vals = torch.tensor([F * frame_bytes / (ms_e2e * 1e-3) / 1e9, probe_gbps,
                             float(numa[0]), float(numa[1])], device=dev, dtype=torch.float64)
    h2d_all = [torch.zeros_like(h2d_vals) for _ in range(world)]
    if world > 1:
        dist.all_gather(h2d_all, h2d_vals)
    else:
        h2d_all = [h2d_vals]
    h2d = {"e2e_GBps_per_rank": [round(float(v[0]), 2) for v in h2d_all], "plain_pinned_copy_GBps_per_rank_concurrent": [round(float(v[1]), 2) for v in h2d_all],
           "cpus_bound_per_rank": [int(v[2]) for v in h2d_all], "cpus_nvml_affinity_per_rank": [int(v[3]) for v in h2d_all],
           "e2e_batch_sizes_rank0": run.e2e_sizes, "e2e_contexts_chained": bool(getattr(run, "chained", False)),
           "note": "e2e: frame bytes of a rank / its e2e step time (the copy is overlapped with compute, so this is a lower bound of the copy rate); plain: 3 back-to-back pinned->device copies of the same frames, all ranks at once, measured before the e2e leg; the e2e launch-batch schedule is the fastest of a few candidates timed on warm-up steps (pick_e2e_schedule)"}
    del h2d_all, h2d_vals

    # ---- per-stage breakdown (separate instrumented pass; timing mode synchronises per call) + keypoints per frame
    stage, kp_counts = {}, []
    if rank == 0:
        ctx.enable_timing(True)
        for first, n, done_f, p0, np_, done_p in batch_plan(F, run.B, []):
            ctx.upload_frames_device(d_frames.data_ptr() + done_f * frame_bytes, n, W_IMG, frame_bytes, 1, first)
            ctx.detect(first, n)
            st = ctx.stage_times()
            for k in ("pyramid_ms", "fast_ms", "select_ms", "orient_ms", "blur_ms", "brief_ms", "describe_ms"):
                stage[k] = stage.get(k, 0.0) + st[k]
            if np_ > 0:
                ctx.process_pairs_device(p0, np_, f0 + done_p)
                st = ctx.stage_times()
                for k in ("match_ms", "filter_ms", "ransac_ms", "pose_ms", "fivept_ms", "score_ms"):
                    stage[k] = stage.get(k, 0.0) + st[k]
            for s in range(first, first + n, max(1, n // 64)):           # a sample of the batch's slots
                try:
                    kp_counts.append(ctx.feature_count(s % (run.B + 1)))
                except b200vo.B200VOError:
                    kp_counts.append(-1)
        ctx.enable_timing(False)
    popc_rate = ctx.measure_popc_rate() if rank == 0 else 0.0
    alg = algorithmic_bytes(ctx, NF) if rank == 0 else None

    # ---- CPU baseline on rank 0 at N=1: bounded sample of the same frames (the first 8 x cores pairs)
    cpu, cpu_acc = None, None
    if world == 1 and not args.no_cpu_baseline:
        try:
            ncpu = os.cpu_count() or 1
            pairs = ncpu * (args.cpu_pairs or 8)
            fr = h_frames[:min(F, pairs + 1)].numpy().copy()      # a copy: no view may outlive the pinned tensor
            v, n, dt, crecs = cpu_reference_throughput(fr, NF, min(pairs, len(fr) - 1), ncpu)
            import cv2
            cpu = {"value": round(v, 2), "unit": "pairs/s", "cores": min(ncpu, n), "kind": "port",
                   "sample": "%d pairs of the same frames, %d processes x cv2.setNumThreads(1) (cv2 %s calling the reference's OpenCV functions), %.1f s" % (n, min(ncpu, n), cv2.__version__, dt)}
            cpu_acc = {"pairs_sampled": n, "cv2_reference_path": pair_errors(crecs, R_all, C_all, f0), "this_path_same_pairs": pair_errors(ok[:n], R_all, C_all, f0)}
            del fr
        except Exception as ex:      # the CPU leg must never take the GPU numbers down with it
            cpu = {"value": None, "unit": "pairs/s", "cores": 0, "kind": "port", "sample": "failed: %s" % ex}

    # ---- release the headline runner before the extra legs allocate
    run.close()
    gathered["t"] = None
    del d_frames
    extras = {}
    if not args.no_extras:
        # the same workload with OpenCV's adaptive RANSAC stopping rule (what the reference's findEssentialMat does);
        # reported beside the fixed-H headline, not as it
        d2 = torch.empty((F, H_IMG, W_IMG), dtype=torch.uint8, device=dev)
        d2.copy_(h_frames)
        ad = make_runner(torch, b200vo, W_IMG, H_IMG, NF, args.hypotheses, B, local, d2, h_frames, f0, ransac_adaptive=1)
        if world > 1:
            gathered["t"] = torch.zeros(world * npairs * RB, dtype=torch.uint8, device=dev)
        pick_dev_mode(torch, dist, world, ad, make_step(ad, False), dev)
        ms_dev_ad, _, _ = timed_steps(torch, dist, world, ad.stream, make_step(ad, False), args.steps, args.warmup, dev)
        ms_e2e_ad, _, _ = timed_steps(torch, dist, world, ad.stream, make_step(ad, True), args.steps, max(args.warmup, 3), dev)
        if rank == 0:
            rr = recs_host.numpy().view(b200vo.POSE_RECORD_DTYPE).copy()
            extras["adaptive_ransac"] = {"note": "same workload with ransac_adaptive=1 (OpenCV confidence-0.999 stopping rule, checked every 128 hypotheses)",
                                         "value": round(world * npairs / (ms_dev_ad * 1e-3), 1), "e2e": round(world * npairs / (ms_e2e_ad * 1e-3), 1),
                                         "mean_hypotheses_per_pair": float((((rr["status"] >> 16) & 0xFF) * 128).mean())}
            del rr
        ad.close()
        gathered["t"] = None
        del d2
        if rank == 0:
            fr = h_frames[:min(F, 210)].numpy().copy()
            extras["streaming"] = {"api": "VisualOdometry::process_frame per frame through libb200vo_host.so (host image in, keypoints + descriptors + pose out, one synchronisation per frame)",
                                   "features_1200": streaming_leg(b200vo, fr, 1200, len(fr)), "features_3000": streaming_leg(b200vo, fr, 3000, len(fr))}
            del fr
        if world > 1:
            dist.barrier()
        ls, lw = max(2, min(args.steps, 3)), 2
        extras["config3"] = strong_scaling_leg(torch, dist, b200vo, synth, tex, "BASELINE configs[2]", 1920, 1080, 8000, 1024, 64, args.c3_frames, rank, world, local, ls, lw)
        extras["config5"] = strong_scaling_leg(torch, dist, b200vo, synth, tex, "BASELINE configs[4]", 3840, 2160, 20000, 4096, 16, args.c5_frames, rank, world, local, ls, lw)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        del recs_host, h_frames
        close_contexts(torch)
        return 0

    hbm_peak, peak_src = peaks()
    stages_out = {}
    fused = stage.get("describe_ms", 0.0) > 0
    rows = [("pyramid", "pyramid_ms", "pyramid"), ("fast", "fast_ms", "fast"), ("blur", "blur_ms", "blur")]
    rows += [("describe", "describe_ms", None)] if fused else [("brief", "brief_ms", "brief"), ("orient", "orient_ms", "orient")]
    for name, key, algk in rows:
        ms = stage.get(key, 0.0)
        ab = alg[algk] if algk else alg["orient"] + 32 * NF + 39 * 39 * NF        # fused: disc + pattern window + descriptor per keypoint
        gbs = ab * F / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        stages_out[name] = {"ms": round(ms, 3), "bound": "hbm", "algorithmic_bytes_per_frame": ab, "achieved_GBps": round(gbs, 1), "frac": round(gbs / hbm_peak, 4)}
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
    kernels = {k: v for k, v in stage.items() if k not in ("ransac_ms",)}
    total_stage_ms = sum(kernels.values())
    dom = max(kernels, key=kernels.get) if kernels else "match_ms"
    if dom == "match_ms":
        roof = {"kernel": "match_kernel<4,%d,%d>" % (int(os.environ.get("B200VO_MATCH_CSA", "3")), int(os.environ.get("B200VO_MATCH_V", "3"))), "bound": "int_pipe", "achieved": round(popc_exec / 1e9, 1), "peak": round(popc_rate / 1e9, 1),
                "unit": "Gpopc32/s", "frac": round(popc_exec / popc_rate, 4),
                # algorithmic: 500 x 3001 x 32 B = 48 MB with every descriptor read once (each slot is the query side of one pair and the
                # train side of the next)
                "traffic": MATCH_DRAM_BYTES_PER_LAUNCH if (F, NF, B) == (1000, 3000, 500) else None, "traffic_unit": "bytes/launch",
                "peak_source": "XOR+POPC+ADD microbenchmark measured in this run (b200vo_measure_popc_rate); not an HBM- or tensor-bound kernel "
                               "(62 MB of DRAM traffic per 5.6 ms launch)",
                "share_of_step": round(mm / total_stage_ms, 3)}
    else:
        nm = dom[:-3]
        sdict = stages_out.get(nm, {})
        roof = {"kernel": nm, "bound": sdict.get("bound", "issue"), "achieved": sdict.get("achieved_GBps"), "peak": hbm_peak, "unit": "GB/s",
                "frac": sdict.get("frac"), "traffic": None, "peak_source": peak_src, "share_of_step": round(stage[dom] / total_stage_ms, 3)}
    hb = max((n for n in ("pyramid", "fast", "blur", "brief", "describe") if n in stages_out), key=lambda n: stages_out[n]["ms"])
    roof_hbm = {"kernel": hb, "bound": "hbm", "achieved": stages_out[hb]["achieved_GBps"], "peak": hbm_peak, "unit": "GB/s",
                "frac": stages_out[hb]["frac"], "traffic": None, "peak_source": peak_src}

    total_pairs = world * npairs
    value = total_pairs / (ms_dev * 1e-3)
    e2e = total_pairs / (ms_e2e * 1e-3)
    kc = [c for c in kp_counts if c >= 0]
    found = (ok["status"] & 2) != 0
    result_check = {"pairs_pose_ok": int(((ok["status"] & 4) != 0).sum()), "pairs_overflow": int(((ok["status"] & 256) != 0).sum()), "pairs_total": int(len(ok)),
                    "trajectory_end": [round(float(x), 3) for x in traj[-1][9:12]] if traj is not None else None,
                    "mean_epipolar_residual": float(ok["epi_residual"][found].mean()) if found.any() else None,
                    "vs_ground_truth": accuracy(traj, R_all, C_all) if traj is not None else None,
                    "vs_ground_truth_pairs": pair_errors(ok, R_all, C_all)}
    if cpu_acc:
        result_check["same_pairs_as_cpu_baseline"] = cpu_acc
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(ms_dev, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/i32 (features, matching), f64 (geometry)", "data": "synthetic",
        "config": {"workload": "%d-frame synthetic KITTI-shaped sequence %dx%d per GPU, %d ORB features, %d five-point hypotheses/pair, launch batches of %d frames" % (F, W_IMG, H_IMG, NF, args.hypotheses, B),
                   "pairs_per_step": total_pairs, "cache": "inputs larger than L2 (%.0f MB of frames per step)" % (F * frame_bytes / 1e6),
                   "partition": ("contiguous frame-pair ranges, one per GPU; one NCCL all_gather of 160-B pose records" if world > 1 else "single GPU") +
                                "; inside a GPU the range is split over %d contexts (CUDA streams), enqueued round-robin or one after the other, whichever was faster on the warm-up steps (device-resident: %s)" % (streams_per_gpu, "one after the other" if getattr(run_mode, "chained_dev", False) else "round-robin")},
        "e2e": {"value": round(e2e, 1), "unit": "pairs/s", "ms_per_step": round(ms_e2e, 3), "h2d_bytes_per_step": world * F * frame_bytes,
                "d2h_bytes_per_step": world * npairs * RB, "wall_ms_per_step": round(wall_e2e * 1e3, 3)},
        "gpu_launches": int(launches) * world,
        "clocks": clocks,
        "roofline": roof, "roofline_hbm": roof_hbm, "stages": stages_out,
        "keypoints_per_frame": {"mean": round(float(np.mean(kc)), 1) if kc else None, "min": int(min(kc)) if kc else None, "frames_sampled": len(kp_counts), "nfeatures": NF},
        "h2d": h2d,
        "cpu_baseline": cpu,
        "result_check": result_check,
        "wall_ms_per_step": round(wall_dev * 1e3, 3),
    }
    line.update(extras)
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    # release the last torch tensors, then destroy the contexts, then leave through the normal interpreter exit
    del recs_host, h_frames, ok
    close_contexts(torch)
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
