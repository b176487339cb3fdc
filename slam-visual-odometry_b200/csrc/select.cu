// K3-K5: retainBest(2 N_l) by FAST score, Harris response, retainBest(N_l) by Harris.
//
// What it replaces: the per-level keypoint selection inside cv::ORB::detectAndCompute
// (reference src/visual_odometry.cpp:113; HARRIS_SCORE, nfeatures at :65-85).  Semantics from
// retainBest(n) keeps EVERY element >= the n-th largest (ties kept, so a
// level may return more than N_l); Harris = 7x7 block of 3x3 Sobel sums, k = 0.04f, scaled by
// (1/(4*7*255))^4, evaluated in float WITHOUT fused multiply-add, in OpenCV's operation order.
//
// One CTA per (slot, level).  The CTA (1) finds the FAST-score cut from the 256-bin histogram the
// FAST kernel produced, (2) gathers the survivors into shared memory, (3) computes Harris, (4) finds
// the N_l-th largest response by radix selection (four 256-bin histogram passes), (5) keeps every
// response >= it and bitonic-sorts only those (key, response) pairs into raster order.  (v1 sorted
// the 2 N_l survivors twice -- 79 % of the kernel's instructions.)
#include "common.cuh"
#include <algorithm>

namespace b200vo {

constexpr int SEL_THREADS = 512;

__device__ __forceinline__ uint32_t float_key_desc(float f) {
    // monotone: larger float -> smaller key (so an ascending sort yields descending responses)
    uint32_t b = __float_as_uint(f);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ~b;
}

// One thread per compare-exchange (n / 2 of them per step) instead of one per element with half of them idle.  A warp's 32
// consecutive pairs at distance j <= 32 lie inside ONE aligned block of 64 elements that no other warp touches in that
// step, so consecutive steps at such distances need only a warp barrier: 15 CTA barriers instead of 55 for n = 1024
// (barrier stalls were 27 % of the kernel's stall samples, profiles/ncu_r02_lines_select_kernel.txt).
__device__ void bitonic_sort_u64(unsigned long long* a, int n /* power of two */) {
    const int half = n >> 1;
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < half; t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));      // the lower index of the t-th pair at distance j
                const int p = i | j;
                const unsigned long long x = a[i], y = a[p];
                const bool up = (i & k) == 0;
                if ((x > y) == up) { a[i] = y; a[p] = x; }
            }
            const int next_j = (j > 1) ? (j >> 1) : k;                    // the next stage (2k) opens at distance k
            if (j > 32 || next_j > 32) __syncthreads(); else __syncwarp();
        }
    }
    __syncthreads();
}

// k-th smallest (1-based) of the keys d_i = float_key_desc(resp[i]), i < m, by most-significant-byte-first
// radix selection: four passes of a 256-bin shared-memory histogram instead of a full sort.
__device__ uint32_t radix_select_kth(const float* resp, int m, int k, int* hist /* [256] shared */, int* s_pick /* [2] shared */) {
    uint32_t prefix = 0, mask = 0;
    for (int pass = 3; pass >= 0; --pass) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < m; i += blockDim.x) {
            const uint32_t d = float_key_desc(resp[i]);
            if ((d & mask) == prefix) atomicAdd(&hist[(d >> (8 * pass)) & 0xFFu], 1);
        }
        __syncthreads();
        if (threadIdx.x < 32) {                       // lane L owns bins [8L, 8L+8): first bin whose running count reaches k
            const int lane = threadIdx.x;
            int local[8], sum = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) { local[i] = hist[lane * 8 + i]; sum += local[i]; }
            int incl = sum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += v;
            }
            int run = incl - sum;                      // count in the bins below this lane's
            if (run < k && incl >= k) {                // the k-th element falls into one of this lane's bins
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (run < k && run + local[i] >= k) { s_pick[0] = lane * 8 + i; s_pick[1] = run; }
                    run += local[i];
                }
            }
        }
        __syncthreads();
        prefix |= (uint32_t)s_pick[0] << (8 * pass);
        mask |= 0xFFu << (8 * pass);
        k -= s_pick[1];
        __syncthreads();
    }
    return prefix;
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
        // separable form of the same integer sums: vertical [1 2 1] smoothing / difference per column once, then the
        // horizontal difference / smoothing per output pixel (69 instead of ~105 instructions per patch row)
        int vs[9], vd[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { vs[i] = r0[i] + 2 * r1[i] + r2[i]; vd[i] = r2[i] - r0[i]; }
#pragma unroll
        for (int i = 1; i < 8; ++i) {
            const int ix = vs[i + 1] - vs[i - 1];
            const int iy = vd[i - 1] + 2 * vd[i] + vd[i + 1];
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

__global__ void __launch_bounds__(SEL_THREADS, 3)      // 40 registers: the separable Harris sums need 46 (59 before), three CTAs per SM pay now (0.240 -> 0.232 ms)
select_kernel(StoreView sv, int first_slot) {
    pdl_grid_sync();
    extern __shared__ __align__(8) unsigned char smem_raw[];
    const int level = blockIdx.x;
    const int slot = ring_slot(first_slot, blockIdx.y, sv.slots);
    const LevelGeom g = sv.lv[level];
    int P2 = 64;                                         // power of two >= kp_cap of this level
    while (P2 < g.kp_cap) P2 <<= 1;
    unsigned long long* pairs = reinterpret_cast<unsigned long long*>(smem_raw);       // [P2] (key << 32 | response bits)
    uint32_t* keys = reinterpret_cast<uint32_t*>(pairs + P2);                            // [sel_cap]
    float* resp = reinterpret_cast<float*>(keys + g.sel_cap);                            // [sel_cap]
    __shared__ int s_cut, s_m, s_out, s_hist[256], s_pick[2];

    const int tid = threadIdx.x, lane = tid & 31;
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
            // retainBest keeps every tie at the cut score.  If they do not fit the working set (saturated or
            // binary imagery), drop the whole tie bin -- count(score > cut) < 2 N_l always fits -- so the
            // truncation is deterministic, and flag the slot (B200VO_ST_OVERFLOW in the pair record).
            int ge = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) ge += (tid * 8 + i >= cut) ? local[i] : 0;
            for (int o = 16; o > 0; o >>= 1) ge += __shfl_xor_sync(0xffffffffu, ge, o);
            if (ge > g.sel_cap) { cut += 1; if (tid == 0) atomicOr(&sv.flags[slot], 2); }
        }
        if (tid == 0) { s_cut = (n2 > 0) ? cut : 256; s_m = 0; s_out = 0; }
    }
    __syncthreads();

    // (2) gather the survivors (warp-aggregated append; the order does not matter yet)
    const uint32_t* cand = sv.cand + (size_t)slot * sv.cand_total + sv.cand_off[level];
    const uint32_t cut = (uint32_t)s_cut;
    for (int i0 = 0; i0 < cnt; i0 += SEL_THREADS) {
        const int i = i0 + tid;
        const uint32_t k = (i < cnt) ? cand[i] : 0u;
        const bool keep = (i < cnt) && ((k & 0xFFu) >= cut);
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_m, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            const int j = base + __popc(bal & ((1u << lane) - 1u));
            if (keep && j < g.sel_cap) keys[j] = k;
        }
    }
    __syncthreads();
    int m = s_m;
    if (m > g.sel_cap) { if (tid == 0) atomicOr(&sv.flags[slot], 2); m = g.sel_cap; }

    // (3) Harris
    const uint8_t* img = sv.pyr + (size_t)slot * sv.slot_bytes + g.offset;
    for (int i = tid; i < m; i += SEL_THREADS) {
        const uint32_t k = keys[i];
        resp[i] = harris_at(img, g.pitch, (int)((k >> 8) & 0xFFFu), (int)(k >> 20));
    }
    __syncthreads();

    // (4) N_l-th largest response (ties kept): radix selection on the order-preserving key
    uint32_t kth = 0xFFFFFFFFu;    // keep all
    if (m > g.nfeat && g.nfeat > 0) kth = radix_select_kth(resp, m, g.nfeat, s_hist, s_pick);

    // (5) keep every response >= the k-th, then sort the survivors by (y, x): raster order is the
    //     canonical order of the keypoint list (the packed key is y << 20 | x << 8 | score)
    for (int i0 = 0; i0 < m; i0 += SEL_THREADS) {
        const int i = i0 + tid;
        const bool keep = (i < m) && (float_key_desc(resp[i]) <= kth);
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        if (bal) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_out, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            const int j = base + __popc(bal & ((1u << lane) - 1u));
            if (keep && j < P2) pairs[j] = ((unsigned long long)keys[i] << 32) | (unsigned long long)__float_as_uint(resp[i]);
        }
    }
    __syncthreads();
    const int outn = s_out;
    const int nsort = min(outn, P2);
    const unsigned long long* sorted = pairs;
    // Raster order by a counting sort on the row: the keys are (y << 20 | x << 8 | score) with a handful of keypoints per
    // row, so a row histogram, its prefix sum, a scatter and an insertion sort inside each row order the list in seven
    // barriers and O(n) work (the bitonic network: 55 steps for 1024 pairs).  The working arrays live in the key / response
    // arrays, dead by now.  Levels whose rows do not fit there, or with a crowded row, take the bitonic sort.
    int* rowcnt = reinterpret_cast<int*>(keys);                    // [h + 1]
    int* rowpos = rowcnt + (g.h + 1);                              // [h + 1] start of the row in `out`, then its fill pointer
    unsigned long long* out = reinterpret_cast<unsigned long long*>(rowpos + (g.h + 1));      // 2 (h + 1) ints behind an 8-byte aligned base: aligned
    const bool fits = 2 * (g.h + 1) + 2 * nsort + 2 <= 2 * g.sel_cap && g.h + 1 <= 16 * SEL_THREADS;
    bool counted = false;
    if (fits && nsort > 64) {
        __shared__ int s_wsum[SEL_THREADS / 32], s_crowded;
        for (int i = tid; i <= g.h; i += SEL_THREADS) rowcnt[i] = 0;
        if (tid == 0) s_crowded = 0;
        __syncthreads();
        for (int i = tid; i < nsort; i += SEL_THREADS) atomicAdd(&rowcnt[(uint32_t)(pairs[i] >> 32) >> 20], 1);
        __syncthreads();
        // exclusive prefix sum over the rows: a thread owns E consecutive rows
        const int E = (g.h + 1 + SEL_THREADS - 1) / SEL_THREADS;
        int local = 0, mx = 0;
        for (int e = 0; e < E; ++e) { const int r = tid * E + e; if (r <= g.h) { local += rowcnt[r]; mx = max(mx, rowcnt[r]); } }
        int incl = local;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
        if (lane == 31) s_wsum[tid >> 5] = incl;
        if (mx > 32) s_crowded = 1;
        __syncthreads();
        int pre = incl - local;
        for (int w = 0; w < (tid >> 5); ++w) pre += s_wsum[w];
        for (int e = 0; e < E; ++e) { const int r = tid * E + e; if (r <= g.h) { rowpos[r] = pre; pre += rowcnt[r]; } }
        const bool crowded = s_crowded != 0;
        __syncthreads();
        if (!crowded) {
            for (int i = tid; i < nsort; i += SEL_THREADS) {
                const unsigned long long v = pairs[i];
                out[atomicAdd(&rowpos[(uint32_t)(v >> 32) >> 20], 1)] = v;
            }
            __syncthreads();
            for (int r = tid; r <= g.h; r += SEL_THREADS) {              // rowpos[r] is the END of row r now
                const int c = rowcnt[r];
                if (c > 1) {
                    unsigned long long* a = out + rowpos[r] - c;
                    for (int i = 1; i < c; ++i) {
                        const unsigned long long v = a[i];
                        int j = i - 1;
                        while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; --j; }
                        a[j + 1] = v;
                    }
                }
            }
            __syncthreads();
            sorted = out;
            counted = true;
        }
    }
    if (!counted) {
        int P = 1;
        while (P < nsort) P <<= 1;
        for (int i = nsort + tid; i < P; i += SEL_THREADS) pairs[i] = ~0ull;
        __syncthreads();
        bitonic_sort_u64(pairs, P);
    }
    uint32_t* okey = sv.lkp_key + (size_t)slot * sv.lkp_total + sv.lkp_off[level];
    float* oresp = sv.lkp_resp + (size_t)slot * sv.lkp_total + sv.lkp_off[level];
    for (int i = tid; i < min(nsort, g.kp_cap); i += SEL_THREADS) {
        const unsigned long long v = sorted[i];
        okey[i] = (uint32_t)(v >> 32);
        oresp[i] = __uint_as_float((uint32_t)v);
    }
    if (tid == 0) {
        if (outn > g.kp_cap) atomicOr(&sv.flags[slot], 4);
        sv.lkp_count[slot * kLevels + level] = min(outn, g.kp_cap);
    }
}

int launch_select(b200vo_ctx* ctx, int first_slot, int n) {
    size_t smem = 0;
    for (int l = 0; l < kLevels; ++l) {
        int p2 = 64;
        while (p2 < ctx->store.lv[l].kp_cap) p2 <<= 1;
        smem = std::max(smem, (size_t)ctx->store.lv[l].sel_cap * 8 + (size_t)p2 * 8);
    }
    // the attribute is per function AND per device and the last value set wins, so it only ever grows;
    // static shared memory counts against the 48 KB default too: always opt in
    static size_t configured_per_device[64] = {};
    size_t& configured = configured_per_device[ctx->device & 63];
    if (configured < smem) {
        B200VO_CUDA_CHECK(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid(kLevels, n);
    launch_k(ctx, select_kernel, dim3(grid), dim3(SEL_THREADS), smem, ctx->store, first_slot);
    B200VO_LAUNCH_CHECK(ctx);
    return B200VO_OK;
}

}  // namespace b200vo
