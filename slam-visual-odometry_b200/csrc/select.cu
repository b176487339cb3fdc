// K3-K5: retainBest(2 N_l) by FAST score, Harris response, retainBest(N_l) by Harris.
//
// What it replaces: the per-level keypoint selection inside cv::ORB::detectAndCompute
// (reference src/visual_odometry.cpp:113; HARRIS_SCORE, nfeatures at :65-85).  Semantics from
// retainBest(n) keeps EVERY element >= the n-th largest (ties kept, so a
// level may return more than N_l); Harris = 7x7 block of 3x3 Sobel sums, k = 0.04f, scaled by
// (1/(4*7*255))^4, evaluated in float WITHOUT fused multiply-add, in OpenCV's operation order.
//
// One CTA per (slot, level).  The CTA (1) finds the FAST-score cut from the 256-bin histogram the
// FAST kernel produced, (2) gathers the survivors into shared memory and bitonic-sorts them by
// packed (y, x) so every later step is in raster order, (3) computes Harris, (4) finds the N_l-th
// largest response with a second in-shared-memory sort, (5) compacts order-preservingly.
#include "common.cuh"

namespace b200vo {

constexpr int SEL_THREADS = 512;

__device__ __forceinline__ uint32_t float_key_desc(float f) {
    // monotone: larger float -> smaller key (so an ascending sort yields descending responses)
    uint32_t b = __float_as_uint(f);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ~b;
}

__device__ void bitonic_sort_u32(uint32_t* a, int n /* power of two */) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const uint32_t x = a[i], y = a[p];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { a[i] = y; a[p] = x; }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ float harris_at(const uint8_t* __restrict__ img, int pitch, int x, int y) {
    // rolling three rows of 9 pixels; Ix/Iy are 3x3 Sobel-like sums (App. A.5)
    int a = 0, b = 0, c = 0;
    int r0[9], r1[9], r2[9];
    const uint8_t* p = img + (size_t)(y - 4) * pitch + (x - 4);
#pragma unroll
    for (int i = 0; i < 9; ++i) { r0[i] = __ldg(p + i); r1[i] = __ldg(p + pitch + i); }
#pragma unroll
    for (int row = 0; row < 7; ++row) {
        const uint8_t* q = p + (size_t)(row + 2) * pitch;
#pragma unroll
        for (int i = 0; i < 9; ++i) r2[i] = __ldg(q + i);
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            const int ix = (r1[i + 1] - r1[i - 1]) * 2 + (r0[i + 1] - r0[i - 1]) + (r2[i + 1] - r2[i - 1]);
            const int iy = (r2[i] - r0[i]) * 2 + (r2[i - 1] - r0[i - 1]) + (r2[i + 1] - r0[i + 1]);
            a += ix * ix; b += iy * iy; c += ix * iy;
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) { r0[i] = r1[i]; r1[i] = r2[i]; }
    }
    const float scale = 1.0f / (4 * 7 * 255.0f);
    const float s4 = __fmul_rn(__fmul_rn(__fmul_rn(scale, scale), scale), scale);
    const float af = (float)a, bf = (float)b, cf = (float)c;
    const float s = __fadd_rn(af, bf);
    const float t3 = __fmul_rn(__fmul_rn(0.04f, s), s);
    return __fmul_rn(__fsub_rn(__fsub_rn(__fmul_rn(af, bf), __fmul_rn(cf, cf)), t3), s4);
}

__global__ void __launch_bounds__(SEL_THREADS)
select_kernel(StoreView sv, int first_slot) {
    extern __shared__ uint32_t smem[];
    const int level = blockIdx.x;
    const int slot = (first_slot + blockIdx.y) % sv.slots;
    const LevelGeom g = sv.lv[level];
    uint32_t* keys = smem;                               // [sel_cap]
    float* resp = reinterpret_cast<float*>(smem + g.sel_cap);   // [sel_cap]
    uint32_t* work = smem + 2 * g.sel_cap;               // [sel_cap]
    __shared__ int s_cut, s_m, s_scan[SEL_THREADS + 1];
    __shared__ uint32_t s_kth;

    const int tid = threadIdx.x;
    const int total = sv.cand_count[slot * kLevels + level];
    const int cnt = min(total, g.cand_cap);
    if (tid == 0 && total > g.cand_cap) atomicOr(&sv.flags[slot], 1);
    const int n2 = 2 * g.nfeat;

    // (1) FAST-score cut: largest T with count(score >= T) >= 2 N_l
    if (tid < 32) {
        const int* hist = sv.fast_hist + ((size_t)slot * kLevels + level) * 256;
        int cut = 0;
        if (total > n2 && n2 > 0) {
            // lane L owns bins [8L, 8L+8); suffix sums from the top
            int local[8], sum = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { local[i] = hist[tid * 8 + i]; sum += local[i]; }
            int above = 0;   // sum over lanes > tid
            for (int l = 0; l < 32; ++l) {
                const int v = __shfl_sync(0xffffffffu, sum, l);
                if (l > tid) above += v;
            }
            int best = -1, run = above;
#pragma unroll
            for (int i = 7; i >= 0; --i) {
                run += local[i];
                if (run >= n2 && best < 0) best = tid * 8 + i;
            }
            // the cut is the maximum candidate bin over lanes
            for (int o = 16; o > 0; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
            cut = max(best, 0);
        }
        if (tid == 0) { s_cut = (n2 > 0) ? cut : 256; s_m = 0; }
    }
    __syncthreads();

    // (2) gather survivors, sort by (y, x)
    const uint32_t* cand = sv.cand + (size_t)slot * sv.cand_total + sv.cand_off[level];
    const uint32_t cut = (uint32_t)s_cut;
    for (int i = tid; i < cnt; i += SEL_THREADS) {
        const uint32_t k = cand[i];
        if ((k & 0xFFu) >= cut) {
            const int j = atomicAdd(&s_m, 1);
            if (j < g.sel_cap) keys[j] = k;
        }
    }
    __syncthreads();
    int m = s_m;
    if (m > g.sel_cap) { if (tid == 0) atomicOr(&sv.flags[slot], 2); m = g.sel_cap; }
    int P = 1;
    while (P < m) P <<= 1;
    for (int i = m + tid; i < P; i += SEL_THREADS) keys[i] = 0xFFFFFFFFu;
    __syncthreads();
    bitonic_sort_u32(keys, P);

    // (3) Harris
    const uint8_t* img = sv.pyr + (size_t)slot * sv.slot_bytes + g.offset;
    for (int i = tid; i < m; i += SEL_THREADS) {
        const uint32_t k = keys[i];
        resp[i] = harris_at(img, g.pitch, (int)((k >> 8) & 0xFFFu), (int)(k >> 20));
    }
    __syncthreads();

    // (4) N_l-th largest response (ties kept)
    uint32_t kth = 0xFFFFFFFFu;    // keep all
    if (m > g.nfeat) {
        for (int i = tid; i < P; i += SEL_THREADS) work[i] = (i < m) ? float_key_desc(resp[i]) : 0xFFFFFFFFu;
        __syncthreads();
        bitonic_sort_u32(work, P);
        if (tid == 0) s_kth = work[g.nfeat - 1];
        __syncthreads();
        kth = s_kth;
    }

    // (5) order-preserving compaction
    const int per = (m + SEL_THREADS - 1) / SEL_THREADS;
    const int lo = min(tid * per, m), hi = min(lo + per, m);
    int c = 0;
    for (int i = lo; i < hi; ++i) c += (float_key_desc(resp[i]) <= kth);
    s_scan[tid + 1] = c;
    if (tid == 0) s_scan[0] = 0;
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i <= SEL_THREADS; ++i) s_scan[i] += s_scan[i - 1];
    }
    __syncthreads();
    int o = s_scan[tid];
    uint32_t* okey = sv.lkp_key + (size_t)slot * sv.lkp_total + sv.lkp_off[level];
    float* oresp = sv.lkp_resp + (size_t)slot * sv.lkp_total + sv.lkp_off[level];
    for (int i = lo; i < hi; ++i) {
        if (float_key_desc(resp[i]) <= kth) {
            if (o < g.kp_cap) { okey[o] = keys[i]; oresp[o] = resp[i]; }
            ++o;
        }
    }
    if (tid == 0) {
        const int outn = s_scan[SEL_THREADS];
        if (outn > g.kp_cap) atomicOr(&sv.flags[slot], 4);
        sv.lkp_count[slot * kLevels + level] = min(outn, g.kp_cap);
    }
}

int launch_select(b200vo_ctx* ctx, int first_slot, int n) {
    int max_cap = 0;
    for (int l = 0; l < kLevels; ++l) max_cap = max(max_cap, ctx->store.lv[l].sel_cap);
    const size_t smem = (size_t)max_cap * 12;
    static size_t configured = 0;
    if (smem > 48 * 1024 && configured < smem) {
        B200VO_CUDA_CHECK(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid(kLevels, n);
    select_kernel<<<grid, SEL_THREADS, smem, ctx->stream>>>(ctx->store, first_slot);
    B200VO_LAUNCH_CHECK(ctx);
    return B200VO_OK;
}

}  // namespace b200vo
