// K10 cross-check (or ratio test) + K11 sort / statistics / adaptive threshold + K12 point gather & K^-1.
//
// What it replaces: the cross-check half of cv::BFMatcher::match and the reference's own
// get_good_matches_of_features (src/visual_odometry.cpp:145-180): std::sort by distance,
// min = front, median = element size/2, thr = min(max(3*min, 0.7*median), 35), keep d < thr;
// then the point gather of estimate_relative_pose (:198-210) and the normalisation
// x^ = (u - cx)/fx, y^ = (v - cy)/fy that cv::findEssentialMat applies.
//
// One CTA per pair.  Mutual nearest neighbours are packed (dist<<16 | queryIdx) and bitonic-sorted
// in shared memory, which yields the canonical (distance, queryIdx) order (the reference's
// std::sort leaves ties unordered); the threshold arithmetic is done in double exactly as the
// reference writes it; survivors are a prefix of the sorted list.
#include "common.cuh"

namespace b200vo {

constexpr int FL_THREADS = 512;

struct FilterArgs {
    int first_slot, slot_b_override;
    double max_hamming;
    double fx, fy, cx, cy;
    int mode;                  // B200VO_MATCH_CROSSCHECK | B200VO_MATCH_RATIO
    float ratio;
};

__global__ void __launch_bounds__(FL_THREADS)
filter_kernel(StoreView sv, PairView pv, FilterArgs fa, int P /* pow2 >= kp_cap */) {
    extern __shared__ uint32_t keys[];
    __shared__ int s_m, s_g;
    __shared__ double s_thr;
    const int pair = blockIdx.x;
    const int tid = threadIdx.x;
    const int sa = (fa.first_slot + pair) % sv.slots;
    const int sb = (fa.slot_b_override >= 0) ? fa.slot_b_override : (fa.first_slot + pair + 1) % sv.slots;
    const int n1 = sv.kp_count[sa], n2 = sv.kp_count[sb];
    const uint32_t* nnq = pv.nnq + (size_t)pair * pv.kp_cap;
    const uint32_t* nnt = pv.nnt + (size_t)pair * pv.kp_cap;
    if (tid == 0) { s_m = 0; s_g = 0; }
    __syncthreads();

    int local = 0;
    for (int i = tid; i < P; i += FL_THREADS) {
        uint32_t key = 0xFFFFFFFFu;
        if (i < n1 && n2 > 0) {
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
            if (keep) { key = (v & 0xFFFF0000u) | (uint32_t)i; ++local; }
        }
        keys[i] = key;
    }
    if (local) atomicAdd(&s_m, local);
    __syncthreads();
    const int m = s_m;

    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < P; i += FL_THREADS) {
                const int p = i ^ j;
                if (p > i) {
                    const uint32_t x = keys[i], y = keys[p];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { keys[i] = y; keys[p] = x; }
                }
            }
            __syncthreads();
        }
    }

    if (tid == 0) {
        double thr = 0.0;
        if (m > 0) {
            const double mn = (double)(keys[0] >> 16);
            const double med = (double)(keys[m / 2] >> 16);
            thr = fmin(fmax(3.0 * mn, 0.7 * med), fa.max_hamming);
        }
        s_thr = thr;
    }
    __syncthreads();
    const double thr = s_thr;

    // survivors: sorted prefix with (double)d < thr
    local = 0;
    for (int i = tid; i < m; i += FL_THREADS) local += ((double)(keys[i] >> 16) < thr);
    if (local) atomicAdd(&s_g, local);
    __syncthreads();
    const int g = s_g;

    const b200vo_keypoint* ka = sv.kp + (size_t)sa * sv.kp_cap;
    const b200vo_keypoint* kb = sv.kp + (size_t)sb * sv.kp_cap;
    b200vo_dmatch* good = pv.good + (size_t)pair * pv.kp_cap;
    double* x1x = pv.x1x + (size_t)pair * pv.kp_cap;
    double* x1y = pv.x1y + (size_t)pair * pv.kp_cap;
    double* x2x = pv.x2x + (size_t)pair * pv.kp_cap;
    double* x2y = pv.x2y + (size_t)pair * pv.kp_cap;
    for (int i = tid; i < g; i += FL_THREADS) {
        const uint32_t key = keys[i];
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
    const size_t smem = (size_t)P * sizeof(uint32_t);
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
    fa.fx = ctx->K[0]; fa.fy = ctx->K[1]; fa.cx = ctx->K[2]; fa.cy = ctx->K[3];
    filter_kernel<<<npairs, FL_THREADS, smem, ctx->stream>>>(ctx->store, ctx->pairs, fa, P);
    B200VO_LAUNCH_CHECK(ctx);
    return B200VO_OK;
}

}  // namespace b200vo
