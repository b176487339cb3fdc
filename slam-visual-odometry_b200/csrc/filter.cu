// K10 cross-check (or ratio test) + K11 sort / statistics / adaptive threshold + K12 point gather & K^-1.
//
// What it replaces: the cross-check half of cv::BFMatcher::match and the reference's own
// get_good_matches_of_features (src/visual_odometry.cpp:145-180): std::sort by distance,
// min = front, median = element size/2, thr = min(max(3*min, 0.7*median), 35), keep d < thr;
// then the point gather of estimate_relative_pose (:198-210) and the normalisation
// x^ = (u - cx)/fx, y^ = (v - cy)/fy that cv::findEssentialMat applies.
//
// One CTA per pair.  Mutual nearest neighbours are packed (dist<<16 | queryIdx); ascending key order is
// the canonical (distance, queryIdx) order (the reference's std::sort leaves ties unordered).  Minimum,
// median (element m/2 of the sorted list) and the number of survivors only depend on the distance
// histogram (257 bins), so they are read off the histogram, the threshold arithmetic is done in double
// exactly as the reference writes it, and only the survivors (d < thr, typically a quarter of the
// matches) are compacted and bitonic-sorted.  More than FL_FAST_CAP survivors: the whole list is sorted
// and the survivors are its prefix (the first version of this kernel).
#include "common.cuh"
#include <algorithm>

namespace b200vo {

constexpr int FL_THREADS = 512;
constexpr int FL_FAST_CAP = kFilterFastCap;

struct FilterArgs {
    int first_slot, slot_b_override;
    double max_hamming;
    double fx, fy, cx, cy;
    int mode;                  // B200VO_MATCH_CROSSCHECK | B200VO_MATCH_RATIO
    float ratio;
    int fast_cap;              // <= FL_FAST_CAP; 0 forces the full sort (tests)
};

// One thread per compare-exchange; steps at distances <= 32 stay inside a warp's own aligned block of 64 elements, so
// they need only a warp barrier (see bitonic_sort_u64 in select.cu).
__device__ __forceinline__ void bitonic_sort_u32(uint32_t* keys, int P, int tid) {
    const int half = P >> 1;
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < half; t += FL_THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const uint32_t x = keys[i], y = keys[p];
                const bool up = (i & k) == 0;
                if ((x > y) == up) { keys[i] = y; keys[p] = x; }
            }
            const int next_j = (j > 1) ? (j >> 1) : k;
            if (j > 32 || next_j > 32) __syncthreads(); else __syncwarp();
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(FL_THREADS)
filter_kernel(StoreView sv, PairView pv, FilterArgs fa, int P /* pow2 >= kp_cap */) {
    pdl_grid_sync();
    extern __shared__ uint32_t keys[];             // [P] all candidates | [FL_FAST_CAP] compacted survivors
    uint32_t* sub = keys + P;
    __shared__ int s_hist[288];                    // distance histogram (0..256), padded to 32 lanes x 9 bins
    __shared__ int s_m, s_g, s_cnt;
    __shared__ double s_thr;
    const int pair = blockIdx.x;
    const int tid = threadIdx.x;
    const int sa = (fa.first_slot + pair) % sv.slots;
    const int sb = (fa.slot_b_override >= 0) ? fa.slot_b_override : (fa.first_slot + pair + 1) % sv.slots;
    const int n1 = sv.kp_count[sa], n2 = sv.kp_count[sb];
    const uint32_t* nnq = pv.nnq + (size_t)pair * pv.kp_cap;
    const uint32_t* nnt = pv.nnt + (size_t)pair * pv.kp_cap;
    if (tid < 288) s_hist[tid] = 0;
    if (tid == 0) s_cnt = 0;
    __syncthreads();

    for (int i = tid; i < n1; i += FL_THREADS) {
        uint32_t key = 0xFFFFFFFFu;
        if (n2 > 0) {
            const uint32_t v = nnq[i];
            const uint32_t t = v & 0xFFFFu;
            bool keep;
            if (fa.mode == B200VO_MATCH_RATIO) {
                // Lowe: best.distance < ratio * second.distance on DMatch's float distances (one
                // rounded float product); a query with fewer than two neighbours is rejected
                const uint32_t v2 = pv.nnq2[(size_t)pair * pv.kp_cap + i];
                keep = (v != 0xFFFFFFFFu) && (v2 != 0xFFFFFFFFu) && ((float)(v >> 16) < __fmul_rn(fa.ratio, (float)(v2 >> 16)));
            } else {
                keep = (v != 0xFFFFFFFFu) && ((nnt[t] & 0xFFFFu) == (uint32_t)i);
            }
            if (keep) { key = (v & 0xFFFF0000u) | (uint32_t)i; atomicAdd(&s_hist[min(v >> 16, 256u)], 1); }
        }
        keys[i] = key;
    }
    __syncthreads();

    // warp 0: m, minimum, median and threshold from the histogram (lane l owns bins 9l .. 9l+8)
    if (tid < 32) {
        const int lane = tid;
        int c[9], sum = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) { c[k] = s_hist[lane * 9 + k]; sum += c[k]; }
        int incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { const int v = __shfl_up_sync(0xFFFFFFFFu, incl, off); if (lane >= off) incl += v; }
        const int excl = incl - sum;
        const int m = __shfl_sync(0xFFFFFFFFu, incl, 31);
        double thr = 0.0, mn = 0.0, mx = 0.0, med = 0.0, mean = 0.0;
        if (m > 0) {
            const unsigned nz = __ballot_sync(0xFFFFFFFFu, sum > 0);
            const int rank = m / 2;
            const unsigned md = __ballot_sync(0xFFFFFFFFu, excl <= rank && rank < incl);
            int first = 0, last = 0, at_rank = 0, cum = excl, wsum = 0;
            bool found_first = false, found_rank = false;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                if (!found_first && c[k] > 0) { first = lane * 9 + k; found_first = true; }
                if (c[k] > 0) last = lane * 9 + k;
                wsum += (lane * 9 + k) * c[k];
                cum += c[k];
                if (!found_rank && cum > rank) { at_rank = lane * 9 + k; found_rank = true; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) wsum += __shfl_xor_sync(0xFFFFFFFFu, wsum, off);
            mn = (double)__shfl_sync(0xFFFFFFFFu, first, __ffs(nz) - 1);
            mx = (double)__shfl_sync(0xFFFFFFFFu, last, 31 - __clz(nz));
            med = (double)__shfl_sync(0xFFFFFFFFu, at_rank, __ffs(md) - 1);
            mean = (double)wsum / (double)m;                   // the distances are integers: the sum is exact in any order
            thr = fmin(fmax(3.0 * mn, 0.7 * med), fa.max_hamming);
        }
        if (lane == 0) {                                        // [MatchDebug] statistics (reference :119-129,160-168)
            double* st = pv.mstat + (size_t)pair * 4;
            st[0] = mn; st[1] = mx; st[2] = med; st[3] = mean;
            pv.ovf[pair] = sv.flags[sa] | sv.flags[sb];
        }
        int gl = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) if ((double)(lane * 9 + k) < thr) gl += c[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) gl += __shfl_xor_sync(0xFFFFFFFFu, gl, off);
        if (lane == 0) { s_m = m; s_g = gl; s_thr = thr; }
    }
    __syncthreads();
    const int m = s_m, g = s_g;
    const double thr = s_thr;

    const uint32_t* sorted;
    if (g <= fa.fast_cap) {
        // compact the survivors (any order), pad to a power of two, sort
        for (int i = tid; i < n1; i += FL_THREADS) {
            const uint32_t key = keys[i];
            if (key != 0xFFFFFFFFu && (double)(key >> 16) < thr) sub[atomicAdd(&s_cnt, 1)] = key;
        }
        int P2 = 1;
        while (P2 < g) P2 <<= 1;
        for (int i = g + tid; i < P2; i += FL_THREADS) sub[i] = 0xFFFFFFFFu;
        __syncthreads();
        bitonic_sort_u32(sub, P2, tid);
        sorted = sub;
    } else {
        for (int i = n1 + tid; i < P; i += FL_THREADS) keys[i] = 0xFFFFFFFFu;
        __syncthreads();
        bitonic_sort_u32(keys, P, tid);
        sorted = keys;
    }

    const b200vo_keypoint* ka = sv.kp + (size_t)sa * sv.kp_cap;
    const b200vo_keypoint* kb = sv.kp + (size_t)sb * sv.kp_cap;
    b200vo_dmatch* good = pv.good + (size_t)pair * pv.kp_cap;
    double* x1x = pv.x1x + (size_t)pair * pv.kp_cap;
    double* x1y = pv.x1y + (size_t)pair * pv.kp_cap;
    double* x2x = pv.x2x + (size_t)pair * pv.kp_cap;
    double* x2y = pv.x2y + (size_t)pair * pv.kp_cap;
    for (int i = tid; i < g; i += FL_THREADS) {
        const uint32_t key = sorted[i];
        const int qi = key & 0xFFFFu;
        const int ti = nnq[qi] & 0xFFFFu;
        b200vo_dmatch dm;
        dm.queryIdx = qi; dm.trainIdx = ti; dm.imgIdx = 0; dm.distance = (float)(key >> 16);
        good[i] = dm;
        const b200vo_keypoint a = ka[qi], b = kb[ti];
        x1x[i] = ((double)a.x - fa.cx) / fa.fx;
        x1y[i] = ((double)a.y - fa.cy) / fa.fy;
        x2x[i] = ((double)b.x - fa.cx) / fa.fx;
        x2y[i] = ((double)b.y - fa.cy) / fa.fy;
        pv.xf[(size_t)pair * pv.kp_cap + i] = make_float4((float)x1x[i], (float)x1y[i], (float)x2x[i], (float)x2y[i]);
    }
    if (tid == 0) {
        pv.n_good[pair] = g;
        pv.n_cross[pair] = m;
        pv.thr[pair] = thr;
    }
}

int launch_filter(b200vo_ctx* ctx, int first_slot, int npairs, int slot_b_override) {
    int P = 1;
    while (P < ctx->store.kp_cap) P <<= 1;
    const size_t smem = ((size_t)P + FL_FAST_CAP) * sizeof(uint32_t);
    // the attribute is per function AND per device and the last value set wins, so it only ever grows;
    // static shared memory counts against the 48 KB default too: always opt in
    static size_t configured_per_device[64] = {};
    size_t& configured = configured_per_device[ctx->device & 63];
    if (configured < smem) {
        B200VO_CUDA_CHECK(cudaFuncSetAttribute(filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    FilterArgs fa;
    fa.first_slot = first_slot; fa.slot_b_override = slot_b_override;
    fa.max_hamming = ctx->opt.max_hamming;
    fa.mode = ctx->opt.match_mode; fa.ratio = ctx->opt.ratio;
    fa.fast_cap = std::min(std::max(ctx->filter_fast_cap, 0), FL_FAST_CAP);
    fa.fx = ctx->K[0]; fa.fy = ctx->K[1]; fa.cx = ctx->K[2]; fa.cy = ctx->K[3];
    launch_k(ctx, filter_kernel, dim3(npairs), dim3(FL_THREADS), smem, ctx->store, ctx->pairs, fa, P);
    B200VO_LAUNCH_CHECK(ctx);
    return B200VO_OK;
}

}  // namespace b200vo
