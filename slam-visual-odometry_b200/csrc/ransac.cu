// K13 batched five-point hypotheses, K14 Sampson scoring, K15 best-model selection,
// K16 recoverPose (decomposition, triangulation, cheirality vote).
//
// What it replaces: cv::findEssentialMat(p1, p2, K, RANSAC, 0.999, 2.0, mask) and
// cv::recoverPose(E, p1, p2, K, R, t, mask) as called by estimate_relative_pose
// (reference src/visual_odometry.cpp:213-216, :265-277).  The math is in geom_core.cuh.
//
// OpenCV draws its samples from a private RNG and stops adaptively; here every pair evaluates a
// fixed, seeded set of H hypotheses (a pure function of (seed, pair id, hypothesis index), never
// of the rank or batch), the whole batch of pairs per launch:
//   five_point_group_kernel    : 10 lanes per hypothesis (fivept_group.cuh) -> up to 10 essential matrices
//   five_point_deferred_kernel : the few hypotheses that need the exact root path, densely packed
//   score_kernel               : CTA per (pair, 64 hypotheses), correspondences staged in shared memory,
//                                one warp per model; ballot + popc counts inliers; the best (count, first
//                                index) of the pair is kept with one 64-bit atomicMax on (count << 32 | ~index)
//   finalize_kernel            : winning E, its inlier mask, record fields
//   pose_decompose / pose_vote / pose_select : recoverPose; the four (R,t) candidates are voted on by all points
// "First arg-max wins" is exactly OpenCV's rule (a model replaces the best only if strictly
// better), realised order-independently by the packed atomicMax.
#include "common.cuh"
#include "geom_core.cuh"
#include "fivept_group.cuh"
#include <algorithm>
#include <cstdlib>

namespace b200vo {

struct RansacArgs {
    const int32_t* samples;     // [pairs][nhyp][5] or NULL (draw on device)
    uint64_t seed;
    int64_t first_pair_id;
    int nhyp;                   // total hypotheses per pair (cap in adaptive mode)
    int h0, nround;             // this launch covers hypotheses [h0, h0 + nround)
    int adaptive;               // 1: pairs whose stopping rule is met (pv.done) are skipped
    float thr2f;                // (float)(thr*thr), thr in normalised units
    int min_matches;            // reference kMinMatchesForPose
};

// ---- cooperative five-point solver (fivept_group.cuh) ----------------------------------------
// One hypothesis per group of 10 lanes, three groups per warp, 24 per CTA; the whole solver
// (null space -> constraints -> elimination -> det B(z) -> Sturm isolation -> per-root finish) is ONE
// kernel: the 10x20 matrix stays in the registers of the group, every exchange goes through the
// group's 2.7 KB block of shared memory, nothing is staged through global or local memory.
// (round-1 history: a thread-per-hypothesis solver ran with 10 of 32 lanes active; a three-phase
// version with stream compaction fixed the divergence but left 5.6 KB of local memory per thread and
// 85 % long-scoreboard stalls -- profiles/ncu_r01_final_kernels.txt.)
// Phases A-E run per group (fivept_group.cuh).  Then the root brackets of the CTA (~4.85 real roots per
// hypothesis: half of the lane slots) are re-packed onto dense lanes for the per-root finish (refine,
// back-substitute, polish), which reads the owning group's setup data from shared memory; only as
// many warps as there are roots keep running.
// EXACT = false (main launch): hypothesis = blockIdx; one whose Sturm chain is unreliable (2-6 %,
// near-multiple roots) is appended to a queue instead of taking the long exact-root path, which would
// make its warp the straggler of the CTA (measured: 42 % of the warp time was spent waiting for it).
// EXACT = true (deferred launch): the queued hypotheses, packed densely -- every group takes the exact path.
template <int THREADS, bool EXACT>
__device__ __forceinline__ void fpg_solve(const PairView& pv, const RansacArgs& ra, geom::GroupShared* groups, int* s_nroot, int* s_gid,
                                          unsigned short* s_root, bool in_range, int pair, int h) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane / geom::kGroupLanes;               // lanes 30, 31 (g == 3) have no group: they only keep the barriers full
    const int r = lane - geom::kGroupLanes * g;
    const int slot = warp * 3 + min(g, 2);
    geom::GroupShared& gs = groups[slot];
    const int gid = pair * pv.nhyp_cap + h;
    const int m = in_range ? pv.n_good[pair] : 0;
    const bool solve = in_range && m >= ra.min_matches && m >= 5;        // uniform over the group
    if (threadIdx.x == 0) *s_nroot = 0;
    if (g < 3 && r == 0) s_gid[slot] = gid;
    double pts[20];
#pragma unroll
    for (int k = 0; k < 20; ++k) pts[k] = 0.0;
    if (solve) {
        int idx[5];
        if (ra.samples) {
#pragma unroll
            for (int k = 0; k < 5; ++k) idx[k] = ra.samples[((size_t)pair * ra.nhyp + h) * 5 + k];
        } else {
            geom::draw_sample(geom::pair_stream(ra.seed, ra.first_pair_id + pair), h, m, idx);
        }
        const size_t base = (size_t)pair * pv.kp_cap;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int i = min(max(idx[k], 0), m - 1);
            pts[2 * k] = pv.x1x[base + i]; pts[2 * k + 1] = pv.x1y[base + i];
            pts[10 + 2 * k] = pv.x2x[base + i]; pts[10 + 2 * k + 1] = pv.x2y[base + i];
        }
    }
    geom::DeviceGroupExec ex{r, solve};
    __syncthreads();
    const int nfinal = geom::five_point_group<false, EXACT>(ex, gs, pts, (double (*)[9]) nullptr, (bool*)nullptr);
    if (in_range) {
        pv.model_ok[(size_t)gid * 10 + r] = 0;
        if (r == 0) {
            pv.nsol[gid] = max(nfinal, 0);
            if (!EXACT && nfinal < 0) pv.fb_gid[atomicAdd(pv.fb_count, 1u)] = gid;
        }
        if (r < nfinal) s_root[atomicAdd(s_nroot, 1)] = (unsigned short)(slot * 10 + r);
    }
    __syncthreads();
    const int nroot = *s_nroot;
    for (int i = threadIdx.x; i < nroot; i += THREADS) {
        const int e = s_root[i], sl = e / 10, k = e - sl * 10;
        const geom::GroupShared& go = groups[sl];
        geom::RootBracket br;
        br.a = go.v.fin.a[k]; br.b = go.v.fin.b[k]; br.reversed = go.v.fin.rev[k];
        double E[9];
        if (geom::five_point_finish(go.S, br, E)) {
            const size_t gd = (size_t)s_gid[sl];
            double* out = pv.models + gd * 90 + k * 9;
#pragma unroll
            for (int q = 0; q < 9; ++q) out[q] = E[q];
            pv.model_ok[gd * 10 + k] = 1;
        }
    }
}

// INLINE_EXACT (small batches, i.e. the per-frame streaming step): a group whose Sturm chain is unreliable runs the
// exact path right here instead of being queued -- the same function on the same inputs as the deferred launch, so
// the result is identical; with one pair in flight a straggler warp costs nothing, a second dependent launch ~90 us.
template <int THREADS, int MINB, bool INLINE_EXACT>
__global__ void __launch_bounds__(THREADS, MINB)
five_point_group_kernel(PairView pv, RansacArgs ra) {
    pdl_grid_sync();
    constexpr int GROUPS = 3 * (THREADS / 32);
    extern __shared__ __align__(16) unsigned char fpg_smem[];
    __shared__ int s_nroot;
    __shared__ int s_gid[GROUPS];
    __shared__ unsigned short s_root[GROUPS * 10];
    const int pair = blockIdx.y;
    if (ra.adaptive && pv.done[pair]) return;             // uniform over the CTA
    const int lane = threadIdx.x & 31, slot = (threadIdx.x >> 5) * 3 + min(lane / geom::kGroupLanes, 2);
    const int h = ra.h0 + blockIdx.x * GROUPS + slot;
    const bool in_range = (lane < 3 * geom::kGroupLanes) && (h < ra.h0 + ra.nround);
    fpg_solve<THREADS, INLINE_EXACT>(pv, ra, reinterpret_cast<geom::GroupShared*>(fpg_smem), &s_nroot, s_gid, s_root, in_range, pair, h);
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
five_point_deferred_kernel(PairView pv, RansacArgs ra) {
    pdl_grid_sync();
    constexpr int GROUPS = 3 * (THREADS / 32);
    extern __shared__ __align__(16) unsigned char fpg_smem[];
    __shared__ int s_nroot;
    __shared__ int s_gid[GROUPS];
    __shared__ unsigned short s_root[GROUPS * 10];
    const unsigned nq = *pv.fb_count;
    const int lane = threadIdx.x & 31, slot = (threadIdx.x >> 5) * 3 + min(lane / geom::kGroupLanes, 2);
    const unsigned qbase = blockIdx.x * GROUPS;            // the grid covers the worst case (every hypothesis queued);
    if (qbase >= nq) return;                               // CTAs beyond the queue exit at once
    const bool in_range = (lane < 3 * geom::kGroupLanes) && (qbase + slot < nq);
    const int q = in_range ? pv.fb_gid[qbase + slot] : 0;
    const int pair = q / pv.nhyp_cap;
    fpg_solve<THREADS, true>(pv, ra, reinterpret_cast<geom::GroupShared*>(fpg_smem), &s_nroot, s_gid, s_root, in_range, pair, q - pair * pv.nhyp_cap);
}

// Inlier count of one model by one warp.  The Sampson test is evaluated in float first (one 16-byte
// load and ~25 FFMA per point on the full-rate FP32 pipe); only points within a 0.3 % band of the
// threshold -- where float rounding could flip the decision -- are re-evaluated with the exact
// double-precision expression, so the count is identical to the all-double evaluation.
// `floor_count`: the scan stops as soon as even all remaining points could not lift the model to
// floor_count inliers (it then can neither win nor tie; the returned count is a lower bound).
template <bool STAGED>
__device__ __forceinline__ int warp_count_inliers(const PairView& pv, int pair, int m, const double* E, const float* __restrict__ Ef,
                                                  const float4* __restrict__ xf, const geom::SampsonThreshold& thr, int lane, int floor_count) {
    const size_t base = (size_t)pair * pv.kp_cap;
    float e[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) e[k] = Ef[k];
    const float tmid = (float)thr.tmid;
    int cnt = 0;
    for (int i0 = 0; i0 < m; i0 += 32) {
        const int j = i0 + lane;
        bool in = false;
        if (j < m) {
            const float4 p = xf[j];
            const float a0 = fmaf(e[0], p.x, fmaf(e[1], p.y, e[2]));
            const float a1 = fmaf(e[3], p.x, fmaf(e[4], p.y, e[5]));
            const float a2 = fmaf(e[6], p.x, fmaf(e[7], p.y, e[8]));
            const float b0 = fmaf(e[0], p.z, fmaf(e[3], p.w, e[6]));
            const float b1 = fmaf(e[1], p.z, fmaf(e[4], p.w, e[7]));
            const float num = fmaf(p.z, a0, fmaf(p.w, a1, a2));
            const float den = fmaf(a0, a0, fmaf(a1, a1, fmaf(b0, b0, b1 * b1)));
            const float lhs = num * num, rhs = tmid * den;
            // float error of num is ~1e-7 absolute (cancellation of O(1) terms): |num| ~ 1.6e-3 at the
            // threshold, so num^2 is good to ~2e-4 relative; 3e-3 leaves a 10x margin
            if (fabsf(lhs - rhs) > 3e-3f * rhs) in = lhs < rhs;
            else in = geom::sampson_inlier_fast(E, pv.x1x[base + j], pv.x1y[base + j], pv.x2x[base + j], pv.x2y[base + j], thr);
        }
        cnt += __popc(__ballot_sync(0xffffffffu, in));
        if (cnt + (m - i0 - 32) < floor_count) break;          // warp-uniform
    }
    return cnt;
}

// K14: one CTA per (pair, block of SC_HB hypotheses).  The pair's correspondences are staged once in shared
// memory (float4 per point); the valid models of the block (~4.85 of the 10 root slots per hypothesis) are
// compacted and copied to shared memory with coalesced loads, so the scoring loop itself touches no global
// memory (round 1 walked the root slots one by one -- a dependent model_ok / model load per slot: 69 % of the
// stall samples were long-scoreboard, issue utilisation 26 %).
// K15: arg-max with "first model wins" through one 64-bit atomicMax on (count << 32 | ~model index).
// Models are abandoned as soon as they cannot reach the best count known so far (exact: such a model can neither
// win nor tie).  The floor is the CTA's own best (shared memory), seeded with the pair's global best at CTA start --
// a stale floor only abandons later, never wrongly.  E is kept twice in shared memory: double for the exact
// re-evaluation inside the threshold band, float (converted once at staging) for the scoring loop -- the per-model
// double->float conversions were 13 % of the kernel's instructions.
constexpr int SC_THREADS = 256;
constexpr int SC_HB = 32;
constexpr int SC_MAXPTS = 1536;                            // staging capacity: 24 KB of float4 (pairs with more good matches read global memory); with the float models 36 KB: 5 CTAs per SM
constexpr int SC_SLOTS = SC_HB * kMaxSolutions;
constexpr int SC_MAXPTS_LIMIT = 4096;                      // largest staging capacity the kernel may be launched with (B200VO_SCORE_CAP)

__global__ void __launch_bounds__(SC_THREADS, 5)       // 48 registers: the fifth CTA the shared memory allows (0.301 -> 0.288 ms per 250 pairs)
score_kernel(PairView pv, RansacArgs ra, int stage_cap) {
    pdl_grid_sync();
    extern __shared__ __align__(16) unsigned char sc_smem[];
    __shared__ int s_n, s_best;
    __shared__ unsigned short s_slot[SC_SLOTS];
    const int pair = blockIdx.y;
    if (ra.adaptive && pv.done[pair]) return;
    const int m = pv.n_good[pair];
    if (m < ra.min_matches || m < 5) return;
    const int hb0 = ra.h0 + blockIdx.x * SC_HB;
    const int hb1 = min(hb0 + SC_HB, ra.h0 + ra.nround);
    const int tid = threadIdx.x;
    // the models as float (converted once); the exact re-evaluation inside the threshold band (0.3 % of the points) reads the
    // double model from global memory: keeping a double copy in shared memory too (23 KB) held the kernel at 3 CTAs per SM
    float* s_Ef = reinterpret_cast<float*>(sc_smem);                       // [SC_SLOTS][9]
    float4* s_pts = reinterpret_cast<float4*>(sc_smem + SC_SLOTS * 9 * sizeof(float));
    const float4* xf = pv.xf + (size_t)pair * pv.kp_cap;
    const bool staged = m <= stage_cap;                    // uniform over the CTA
    if (tid == 0) { s_n = 0; s_best = pv.counts ? 0 : (int)(pv.best[pair] >> 32); }
    __syncthreads();
    const size_t slot0 = ((size_t)pair * pv.nhyp_cap + hb0) * kMaxSolutions;
    const int nslots = (hb1 - hb0) * kMaxSolutions;
    for (int i = tid; i < nslots; i += SC_THREADS)
        if (pv.model_ok[slot0 + i]) s_slot[atomicAdd(&s_n, 1)] = (unsigned short)i;
    if (staged) {
        for (int i = tid; i < m; i += SC_THREADS) s_pts[i] = xf[i];
        xf = s_pts;
    }
    __syncthreads();
    const int n = s_n;
    for (int i = tid; i < n * 9; i += SC_THREADS) {
        const int j = i / 9, k = i - j * 9;
        s_Ef[i] = (float)pv.models[(slot0 + s_slot[j]) * 9 + k];
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    const geom::SampsonThreshold thr = geom::make_sampson_threshold(ra.thr2f);
    for (int j = warp; j < n; j += SC_THREADS / 32) {
        const int slot = s_slot[j];
        // the floor: the best count this CTA knows of -- the pair's global best when the CTA started and its own models
        // since.  No global read inside the loop: the round-2 profile of the first staged version showed the per-model
        // read of the global best (L2 round trip, consumed one model later) as 62 % of all stall samples.
        const int floor_count = pv.counts ? 0 : *(volatile int*)&s_best;
        const double* Ed = pv.models + (slot0 + slot) * 9;
        const int cnt = staged ? warp_count_inliers<true>(pv, pair, m, Ed, s_Ef + j * 9, xf, thr, lane, floor_count)
                               : warp_count_inliers<false>(pv, pair, m, Ed, s_Ef + j * 9, xf, thr, lane, floor_count);
        if (lane == 0) {
            if (pv.counts) pv.counts[slot0 + slot] = cnt;
            if (cnt >= floor_count) {
                const uint32_t model = (uint32_t)(hb0 * kMaxSolutions + slot);            // = h * 10 + s
                atomicMax(&pv.best[pair], ((unsigned long long)(uint32_t)cnt << 32) | (unsigned long long)(0xFFFFFFFFu - model));   // no return value used: a RED, nobody waits
                atomicMax(&s_best, cnt);
            }
        }
    }
}

template <int THREADS, int MINB>
static int launch_five_point_as(b200vo_ctx* ctx, const RansacArgs& ra, int npairs) {
    constexpr int GROUPS = 3 * (THREADS / 32);
    const size_t smem = GROUPS * sizeof(geom::GroupShared);
    if (THREADS == 256 && MINB == 2 && npairs <= ctx->small_batch_max) {
        launch_k(ctx, five_point_group_kernel<256, 2, true>, dim3(dim3(div_up(ra.nround, GROUPS), npairs)), dim3(THREADS), smem, ctx->pairs, ra);
        B200VO_LAUNCH_CHECK(ctx);
        return B200VO_OK;
    }
    launch_k(ctx, five_point_group_kernel<THREADS, MINB, false>, dim3(dim3(div_up(ra.nround, GROUPS), npairs)), dim3(THREADS), smem, ctx->pairs, ra);
    B200VO_LAUNCH_CHECK(ctx);
    // the queue length is only known on the device (measured share: 0.5-6 %): the grid covers the worst
    // case, CTAs beyond the queue exit at once
    launch_k(ctx, five_point_deferred_kernel<THREADS, MINB>, dim3(div_up(npairs * ra.nround, GROUPS)), dim3(THREADS), smem, ctx->pairs, ra);
    B200VO_LAUNCH_CHECK(ctx);
    return B200VO_OK;
}

static int launch_five_point(b200vo_ctx* ctx, const RansacArgs& ra, int npairs) {
    // launch shape (threads x CTAs/SM -> register cap), measured on B200 at 1000 frames, H = 1024 -- see DESIGN.md section 4
    { const int rcf = device_fill32(ctx, ctx->pairs.fb_count, 0u, 1); if (rcf) return rcf; }
    switch (ctx->fpg_variant) {
        // measured per 250 pairs x 1024 hypotheses (profiles/fpg_variants_r02.jsonl): 256x2 1.42 ms; 192x2 (168 registers, a
        // third of the spills) 1.64; 256x3 (80 registers) 1.75; 128x4 1.68; 160x3 1.70
        case 1: return launch_five_point_as<192, 2>(ctx, ra, npairs);
        case 3: return launch_five_point_as<128, 4>(ctx, ra, npairs);
        case 5: return launch_five_point_as<320, 2>(ctx, ra, npairs);      // 102 registers, 20 warps/SM
        case 6: return launch_five_point_as<288, 2>(ctx, ra, npairs);      // 113 registers, 18 warps/SM
        default: return launch_five_point_as<256, 2>(ctx, ra, npairs);     // 128 registers, 16 warps/SM
    }
}

static int launch_score(b200vo_ctx* ctx, const RansacArgs& ra, int npairs) {
    const dim3 grid(div_up(ra.nround, SC_HB), npairs);
    int cap = std::min(ctx->pairs.kp_cap, SC_MAXPTS);
    if (const char* v = getenv("B200VO_SCORE_CAP")) { const int c = atoi(v); if (c >= 0 && c <= SC_MAXPTS_LIMIT) cap = std::min(ctx->pairs.kp_cap, c); }
    launch_k(ctx, score_kernel, dim3(grid), dim3(SC_THREADS), SC_SLOTS * 9 * sizeof(float) + (size_t)cap * sizeof(float4), ctx->pairs, ra, cap);
    B200VO_LAUNCH_CHECK(ctx);
    return B200VO_OK;
}

int init_ransac_kernels() {
    cudaError_t e = cudaFuncSetAttribute(score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SC_SLOTS * 9 * sizeof(float) + SC_MAXPTS_LIMIT * sizeof(float4)));
    auto fpg_attr = [&](auto main_k, auto def_k, int threads) {
        const int bytes = 3 * (threads / 32) * (int)sizeof(geom::GroupShared);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(main_k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(def_k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    };
    fpg_attr(five_point_group_kernel<256, 2, false>, five_point_deferred_kernel<256, 2>, 256);
    fpg_attr(five_point_group_kernel<256, 2, true>, five_point_deferred_kernel<256, 2>, 256);
    fpg_attr(five_point_group_kernel<192, 2, false>, five_point_deferred_kernel<192, 2>, 192);
    fpg_attr(five_point_group_kernel<128, 4, false>, five_point_deferred_kernel<128, 4>, 128);
    fpg_attr(five_point_group_kernel<320, 2, false>, five_point_deferred_kernel<320, 2>, 320);
    fpg_attr(five_point_group_kernel<288, 2, false>, five_point_deferred_kernel<288, 2>, 288);
    if (e != cudaSuccess) { set_error("ransac kernel smem attribute: %s", cudaGetErrorString(e)); return B200VO_ERR_CUDA; }
    return B200VO_OK;
}

__global__ void __launch_bounds__(256)
finalize_kernel(PairView pv, RansacArgs ra) {
    pdl_grid_sync();
    const int pair = blockIdx.x;
    const int tid = threadIdx.x;
    const int m = pv.n_good[pair];
    const unsigned long long best = pv.best[pair];
    const int cnt = (int)(best >> 32);
    const bool found = (m >= ra.min_matches) && (m >= 8) && cnt >= 5;
    const uint32_t model = 0xFFFFFFFFu - (uint32_t)(best & 0xFFFFFFFFu);
    const int h = found ? (int)(model / 10) : -1, s = found ? (int)(model % 10) : -1;
    double E[9];
    if (found) {
        const double* src = pv.models + ((size_t)pair * pv.nhyp_cap + h) * 90 + s * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) E[k] = src[k];
    } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) E[k] = 0.0;
    }
    const size_t base = (size_t)pair * pv.kp_cap;
    // inlier mask of the winner, and the reference's self-check: mean |x2^T F x1| over the inliers in pixel
    // coordinates (src/visual_odometry.cpp:248-263).  With F = K^-T E K^-1 that is x2n^T E x1n on the
    // normalised points.  Fixed summation order (strided partial sums, then a tree): the record is byte-identical
    // whatever the batch.
    __shared__ double s_sum[256];
    double part = 0.0;
    for (int i = tid; i < m; i += blockDim.x) {
        bool in = false;
        if (found) {
            const double x1 = pv.x1x[base + i], y1 = pv.x1y[base + i], x2 = pv.x2x[base + i], y2 = pv.x2y[base + i];
            in = geom::sampson_inlier(E, x1, y1, x2, y2, ra.thr2f);
            if (in) part += fabs(x2 * (E[0] * x1 + E[1] * y1 + E[2]) + y2 * (E[3] * x1 + E[4] * y1 + E[5]) + (E[6] * x1 + E[7] * y1 + E[8]));
        }
        pv.inl[base + i] = in ? (uint8_t)1 : (uint8_t)0;
    }
    s_sum[tid] = part;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) s_sum[tid] += s_sum[tid + o];
        __syncthreads();
    }
    if (tid == 0) {
        for (int k = 0; k < 9; ++k) pv.E[(size_t)pair * 9 + k] = E[k];
        b200vo_pose_record r;
        for (int k = 0; k < 9; ++k) r.R[k] = (k % 4 == 0) ? 1.0 : 0.0;
        r.t[0] = r.t[1] = r.t[2] = 0.0;
        r.n_matches = m;
        r.n_inliers = found ? cnt : 0;
        r.n_valid = 0;
        r.status = ((m >= ra.min_matches) ? B200VO_ST_MATCHED : 0) | (found ? B200VO_ST_E_FOUND : 0) |
                   (pv.ovf[pair] ? B200VO_ST_OVERFLOW : 0);
        r.n_cross = pv.n_cross[pair];
        r.best_hyp = h;
        r.best_sol = s;
        r.threshold = (float)pv.thr[pair];
        const double* st = pv.mstat + (size_t)pair * 4;
        r.dist_min = (float)st[0]; r.dist_max = (float)st[1]; r.dist_median = (float)st[2]; r.dist_mean = st[3];
        r.epi_residual = (found && cnt > 0) ? s_sum[0] / (double)cnt : 0.0;
        r.reserved = 0;
        pv.rec[pair] = r;
    }
}

struct PoseArgs { int min_valid, min_inliers; uint8_t* mask_out; /* [pairs][kp_cap] or NULL */ };

// recoverPose in three flat launches (a CTA-per-pair kernel with thread 0 decomposing E left the
// SMs idle: 12 % occupancy, 35 % barrier stalls in the v1 profile):
//   pose_decompose_kernel  thread / pair           SVD of E -> R1, R2, t
//   pose_vote_kernel       thread / (pair, point)  4 triangulations + cheirality, warp-reduced votes
//   pose_select_kernel     thread / pair           first maximum wins -> record
__global__ void __launch_bounds__(64)
pose_decompose_kernel(PairView pv, int npairs) {
    pdl_grid_sync();
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= npairs) return;
    int* cnt = pv.pose_cnt + pair * 4;
    cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0;
    if (!(pv.rec[pair].status & B200VO_ST_E_FOUND)) return;
    double* c = pv.pose_cand + (size_t)pair * 24;
    geom::decompose_essential(pv.E + (size_t)pair * 9, c, c + 9, c + 18);
}

// thread = (point, candidate): 4x the threads, a quarter of the serial latency per thread
__global__ void __launch_bounds__(128)
pose_vote_kernel(PairView pv) {
    pdl_grid_sync();
    const int pair = blockIdx.y;
    if (!(pv.rec[pair].status & B200VO_ST_E_FOUND)) return;
    const int m = pv.n_good[pair];
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (blockIdx.x * blockDim.x >= 4 * m) return;
    const int i = t >> 2, k = t & 3;              // candidate order of cv::recoverPose: (R1,t), (R2,t), (R1,-t), (R2,-t)
    const size_t base = (size_t)pair * pv.kp_cap;
    int ok = 0;
    if (i < m && pv.inl[base + i]) {
        const double* cand = pv.pose_cand + (size_t)pair * 24;
        double R[9], tt[3];
        const double sg = (k >= 2) ? -1.0 : 1.0;
#pragma unroll
        for (int j = 0; j < 9; ++j) R[j] = cand[9 * (k & 1) + j];
#pragma unroll
        for (int j = 0; j < 3; ++j) tt[j] = sg * cand[18 + j];
        ok = geom::cheirality_ok(R, tt, pv.x1x[base + i], pv.x1y[base + i], pv.x2x[base + i], pv.x2y[base + i], 50.0);
    }
    // lanes with equal (lane & 3) vote for the same candidate
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    const int lane = threadIdx.x & 31;
    if (lane < 4) {
        const int w = __popc(bal & (0x11111111u << lane));
        if (w) atomicAdd(&pv.pose_cnt[pair * 4 + lane], w);
    }
}

__global__ void __launch_bounds__(64)
pose_select_kernel(PairView pv, PoseArgs pa, int npairs) {
    pdl_grid_sync();
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= npairs) return;
    b200vo_pose_record* rec = pv.rec + pair;
    if (!(rec->status & B200VO_ST_E_FOUND)) return;
    const int* cnt = pv.pose_cnt + pair * 4;
    int bi = 0;
    for (int k = 1; k < 4; ++k)
        if (cnt[k] > cnt[bi]) bi = k;              // first maximum wins
    const double* cand = pv.pose_cand + (size_t)pair * 24;
    const double* R = cand + 9 * (bi & 1);
    const double sg = (bi >= 2) ? -1.0 : 1.0;
    for (int k = 0; k < 9; ++k) rec->R[k] = R[k];
    for (int k = 0; k < 3; ++k) rec->t[k] = sg * cand[18 + k];
    rec->n_valid = cnt[bi];
    if (cnt[bi] >= pa.min_valid && rec->n_inliers >= pa.min_inliers) rec->status |= B200VO_ST_POSE_OK;
}

// parity hook only: cv::recoverPose overwrites the mask with the winning candidate's test
__global__ void pose_mask_kernel(PairView pv, PoseArgs pa) {
    pdl_grid_sync();
    const int pair = blockIdx.y;
    const b200vo_pose_record* rec = pv.rec + pair;
    if (!(rec->status & B200VO_ST_E_FOUND)) return;
    const int m = pv.n_good[pair];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const size_t base = (size_t)pair * pv.kp_cap;
    pa.mask_out[base + i] = pv.inl[base + i] &&
        geom::cheirality_ok(rec->R, rec->t, pv.x1x[base + i], pv.x1y[base + i], pv.x2x[base + i], pv.x2y[base + i], 50.0);
}

int launch_recover_pose_masked(b200vo_ctx* ctx, int npairs, uint8_t* d_mask_out) {
    PoseArgs pa{ctx->opt.min_valid_points, ctx->opt.min_inliers, d_mask_out};
    launch_k(ctx, pose_decompose_kernel, dim3(div_up(npairs, 64)), dim3(64), 0, ctx->pairs, npairs);
    B200VO_LAUNCH_CHECK(ctx);
    launch_k(ctx, pose_vote_kernel, dim3(dim3(div_up(4 * ctx->pairs.kp_cap, 128), npairs)), dim3(128), 0, ctx->pairs);
    B200VO_LAUNCH_CHECK(ctx);
    launch_k(ctx, pose_select_kernel, dim3(div_up(npairs, 64)), dim3(64), 0, ctx->pairs, pa, npairs);
    B200VO_LAUNCH_CHECK(ctx);
    if (d_mask_out) {
        launch_k(ctx, pose_mask_kernel, dim3(dim3(div_up(ctx->pairs.kp_cap, 128), npairs)), dim3(128), 0, ctx->pairs, pa);
        B200VO_LAUNCH_CHECK(ctx);
    }
    return B200VO_OK;
}

int launch_recover_pose(b200vo_ctx* ctx, int npairs) { return launch_recover_pose_masked(ctx, npairs, nullptr); }

// parity hook: solve the given samples for pair 0 (points staged by the caller)
int launch_five_point_only(b200vo_ctx* ctx, const int32_t* d_samples, int nhyp) {
    RansacArgs ra;
    ra.samples = d_samples; ra.seed = 0; ra.first_pair_id = 0; ra.nhyp = nhyp; ra.thr2f = 0.f; ra.min_matches = 5;
    ra.h0 = 0; ra.nround = nhyp; ra.adaptive = 0;
    return launch_five_point(ctx, ra, 1);
}

// cv::RANSACUpdateNumIters (OpenCV modules/calib3d/src/ptsetreg.cpp), as findEssentialMat applies it
// after every improvement of the best model: p = confidence, ep = outlier ratio of the best model.
__device__ __forceinline__ int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = fmin(fmax(p, 0.0), 1.0);
    ep = fmin(fmax(ep, 0.0), 1.0);
    double num = fmax(1.0 - p, 2.2250738585072014e-308);
    double denom = 1.0 - pow(1.0 - ep, (double)model_points);
    if (denom < 2.2250738585072014e-308) return 0;
    num = log(num);
    denom = log(denom);
    return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : __double2int_rn(num / denom);
}

// adaptive mode: after each block of hypotheses decide per pair whether OpenCV's loop would have ended
__global__ void adaptive_update_kernel(PairView pv, int npairs, int evaluated, int max_iters, double confidence) {
    pdl_grid_sync();
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair >= npairs || pv.done[pair]) return;
    pv.rounds[pair] += 1;
    const int m = pv.n_good[pair];
    const int cnt = (int)(pv.best[pair] >> 32);
    int niters = max_iters;
    if (m > 0 && cnt >= 5) niters = ransac_update_num_iters(confidence, (double)(m - cnt) / m, 5, max_iters);
    if (evaluated >= niters || m < 5) pv.done[pair] = 1;
}

__global__ void stamp_rounds_kernel(PairView pv, int npairs) {
    pdl_grid_sync();
    const int pair = blockIdx.x * blockDim.x + threadIdx.x;
    if (pair < npairs) pv.rec[pair].status |= (pv.rounds[pair] & 0xFF) << B200VO_ST_ROUNDS_SHIFT;
}

int launch_ransac(b200vo_ctx* ctx, int npairs, int64_t first_pair_id, const int32_t* d_samples, int nhyp) {
    if (nhyp > ctx->pairs.nhyp_cap) { set_error("nhyp %d exceeds capacity %d", nhyp, ctx->pairs.nhyp_cap); return B200VO_ERR_CAPACITY; }
    RansacArgs ra;
    ra.samples = d_samples; ra.seed = ctx->opt.ransac_seed; ra.first_pair_id = first_pair_id; ra.nhyp = nhyp;
    const double thr = ctx->opt.ransac_threshold_px / ((ctx->K[0] + ctx->K[1]) / 2.0);
    ra.thr2f = (float)(thr * thr);
    ra.min_matches = ctx->opt.min_matches_for_pose;
    // an explicit sample list (parity hooks) is always evaluated completely
    ra.adaptive = (ctx->opt.ransac_adaptive != 0 && d_samples == nullptr) ? 1 : 0;
    { const int rcf = device_fill32(ctx, ctx->pairs.best, 0u, (size_t)npairs * 2); if (rcf) return rcf; }
    const int block = ra.adaptive ? 128 : nhyp;
    if (ra.adaptive) {
        { int rcf; if ((rcf = device_fill32(ctx, ctx->pairs.done, 0u, (size_t)npairs))) return rcf;
          if ((rcf = device_fill32(ctx, ctx->pairs.rounds, 0u, (size_t)npairs))) return rcf; }
    }
    for (int h0 = 0; h0 < nhyp; h0 += block) {
        ra.h0 = h0; ra.nround = std::min(block, nhyp - h0);
        const bool tm = ctx->timing && !ra.adaptive;        // events 8..10 bracket the two kernels
        if (tm) cudaEventRecord(ctx->ev[8], ctx->stream);
        { const int rc = launch_five_point(ctx, ra, npairs); if (rc) return rc; }
        if (tm) cudaEventRecord(ctx->ev[9], ctx->stream);
        { const int rc = launch_score(ctx, ra, npairs); if (rc) return rc; }
        if (tm) cudaEventRecord(ctx->ev[10], ctx->stream);
        if (ra.adaptive) {
            launch_k(ctx, adaptive_update_kernel, dim3(div_up(npairs, 128)), dim3(128), 0, ctx->pairs, npairs, h0 + ra.nround, nhyp, 0.999);
            B200VO_LAUNCH_CHECK(ctx);
        }
    }
    launch_k(ctx, finalize_kernel, dim3(npairs), dim3(256), 0, ctx->pairs, ra);
    B200VO_LAUNCH_CHECK(ctx);
    if (ra.adaptive) {
        launch_k(ctx, stamp_rounds_kernel, dim3(div_up(npairs, 128)), dim3(128), 0, ctx->pairs, npairs);
        B200VO_LAUNCH_CHECK(ctx);
    }
    return B200VO_OK;
}

// ------------------------------------------------------------------ parity hooks (device side)
__global__ void score_models_kernel(PairView pv, const double* models, int nmodels, float thr2f, int* counts) {
    pdl_grid_sync();
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (w >= nmodels) return;
    double E[9];
    for (int k = 0; k < 9; ++k) E[k] = models[(size_t)w * 9 + k];
    const geom::SampsonThreshold thr = geom::make_sampson_threshold(thr2f);
    float Ef[9];
    for (int k = 0; k < 9; ++k) Ef[k] = (float)E[k];
    const int cnt = warp_count_inliers<false>(pv, 0, pv.n_good[0], E, Ef, pv.xf, thr, lane, 0);      // points staged in pair 0
    if (lane == 0) counts[w] = cnt;
}

int launch_score_models(b200vo_ctx* ctx, const double* d_models, int nmodels, float thr2f, int* d_counts) {
    launch_k(ctx, score_models_kernel, dim3(div_up(nmodels, 8)), dim3(256), 0, ctx->pairs, d_models, nmodels, thr2f, d_counts);
    B200VO_LAUNCH_CHECK(ctx);
    return B200VO_OK;
}

}  // namespace b200vo
