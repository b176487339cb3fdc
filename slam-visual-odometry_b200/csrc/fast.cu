// K2 FAST-9/16 corner detection + corner score + 3x3 non-max suppression + border filter.
//
// What it replaces: the FAST stage inside cv::ORB::detectAndCompute
// (reference src/visual_odometry.cpp:113; ORB created with fastThreshold 10, edgeThreshold 31 at
// :64-85).  Semantics follow SURVEY.md App. A.3/A.4: segment test on the 16-pixel Bresenham ring,
// score = (max over 9-arcs of the arc minimum of |ring - centre|) - 1, strict-greater 3x3 NMS,
// then keep 31 <= x < w-31, 31 <= y < h-31.  Because of that border filter every pixel the kernel
// touches lies inside the level image, so there is no boundary handling at all.
//
// Output: an unordered candidate list per (slot, level), packed (y<<20 | x<<8 | score), plus a
// 256-bin score histogram used by the selection kernel to find the retainBest(2 N_l) cut.
#include "common.cuh"

namespace b200vo {

constexpr int FT_W = 64, FT_H = 32;             // output tile
constexpr int FT_SW = FT_W + 8, FT_SH = FT_H + 8;   // pixel tile (halo 3 ring + 1 NMS)
constexpr int FT_CW = FT_W + 2, FT_CH = FT_H + 2;   // score tile (halo 1)

__device__ __forceinline__ bool has_arc9(uint32_t m) {
    uint32_t r = m | (m << 16);
    r &= r >> 1;
    r &= r >> 2;
    r &= r >> 4;
    r &= (m | (m << 16)) >> 8;
    return (r & 0xFFFFu) != 0;
}

// p points at the centre pixel inside the shared tile (row stride FT_SW).
__device__ __forceinline__ int fast_score(const uint8_t* p, int t) {
    const int S = FT_SW;
    const int c = p[0];
    const int hi = c + t, lo = c - t;
    // every 9-arc contains one pixel of each opposite pair: cheap rejection on (0,8) and (4,12)
    const int r0 = p[3 * S], r8 = p[-3 * S];
    if (!((r0 > hi) | (r8 > hi) | (r0 < lo) | (r8 < lo))) return 0;
    const int r4 = p[3], r12 = p[-3];
    if (!((r4 > hi) | (r12 > hi) | (r4 < lo) | (r12 < lo))) return 0;
    int v[16];
    v[0] = r0;            v[1] = p[3 * S + 1];   v[2] = p[2 * S + 2];   v[3] = p[S + 3];
    v[4] = r4;            v[5] = p[-S + 3];      v[6] = p[-2 * S + 2];  v[7] = p[-3 * S + 1];
    v[8] = r8;            v[9] = p[-3 * S - 1];  v[10] = p[-2 * S - 2]; v[11] = p[-S - 3];
    v[12] = r12;          v[13] = p[S - 3];      v[14] = p[2 * S - 2];  v[15] = p[3 * S - 1];
    uint32_t mb = 0, md = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        mb |= (uint32_t)(v[i] > hi) << i;
        md |= (uint32_t)(v[i] < lo) << i;
    }
    const bool bright = has_arc9(mb);
    if (!bright && !has_arc9(md)) return 0;
    // exactly one polarity can hold a 9-arc; its arc-min maximum is the corner's best threshold
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = bright ? (v[i] - c) : (c - v[i]);
    int m2[16], m4[16], best = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) m2[i] = min(v[i], v[(i + 1) & 15]);
#pragma unroll
    for (int i = 0; i < 16; ++i) m4[i] = min(m2[i], m2[(i + 2) & 15]);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int m8 = min(m4[i], m4[(i + 4) & 15]);
        best = max(best, min(m8, v[(i + 8) & 15]));
    }
    return best - 1;    // best > t is implied by the mask test
}

__global__ void __launch_bounds__(256)
fast_kernel(StoreView sv, int level, int first_slot) {
    __shared__ __align__(16) uint8_t s_img[FT_SH * FT_SW];
    __shared__ uint8_t s_score[FT_CH * FT_CW];
    __shared__ uint32_t s_list[FT_W * FT_H / 4];
    __shared__ int s_n, s_base;

    const LevelGeom g = sv.lv[level];
    const int slot = (first_slot + blockIdx.z) % sv.slots;
    const uint8_t* __restrict__ img = sv.pyr + (size_t)slot * sv.slot_bytes + g.offset;
    const int ox = kEdge + blockIdx.x * FT_W;       // first output pixel of the tile
    const int oy = kEdge + blockIdx.y * FT_H;
    const int tid = threadIdx.x;
    if (tid == 0) s_n = 0;

    // stage the pixel tile: rows oy-4 .. oy+FT_H+3, cols ox-4 .. ox+FT_W+3 (clamped: clamped
    // pixels only feed outputs beyond the border filter, which are discarded)
    for (int i = tid; i < FT_SH * FT_SW; i += 256) {
        const int ty = i / FT_SW, tx = i - ty * FT_SW;
        const int gy = min(oy - 4 + ty, g.h - 1), gx = min(ox - 4 + tx, g.w - 1);
        s_img[i] = __ldg(img + (size_t)gy * g.pitch + gx);
    }
    __syncthreads();

    const int t = sv.fast_threshold;
    for (int i = tid; i < FT_CH * FT_CW; i += 256) {
        const int cy = i / FT_CW, cx = i - cy * FT_CW;
        s_score[i] = (uint8_t)fast_score(&s_img[(cy + 3) * FT_SW + (cx + 3)], t);
    }
    __syncthreads();

    const int xmax = g.w - kEdge, ymax = g.h - kEdge;
    for (int i = tid; i < FT_W * FT_H; i += 256) {
        const int ty = i / FT_W, tx = i - ty * FT_W;
        const int gx = ox + tx, gy = oy + ty;
        const uint8_t* s = &s_score[(ty + 1) * FT_CW + (tx + 1)];
        const int sc = s[0];
        if (sc == 0 || gx >= xmax || gy >= ymax) continue;
        const bool keep = sc > s[-1] && sc > s[1] && sc > s[-FT_CW - 1] && sc > s[-FT_CW] && sc > s[-FT_CW + 1] &&
                          sc > s[FT_CW - 1] && sc > s[FT_CW] && sc > s[FT_CW + 1];
        if (keep) {
            const int k = atomicAdd(&s_n, 1);
            s_list[k] = ((uint32_t)gy << 20) | ((uint32_t)gx << 8) | (uint32_t)sc;
            atomicAdd(&sv.fast_hist[((size_t)slot * kLevels + level) * 256 + sc], 1);
        }
    }
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;
    if (tid == 0) s_base = atomicAdd(&sv.cand_count[slot * kLevels + level], n);
    __syncthreads();
    const int base = s_base;
    uint32_t* out = sv.cand + (size_t)slot * sv.cand_total + sv.cand_off[level];
    for (int i = tid; i < n; i += 256)
        if (base + i < g.cand_cap) out[base + i] = s_list[i];
}

__global__ void clear_detect_state_kernel(StoreView sv, int first_slot, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = kLevels * 256;
    if (i < n * per) {
        const int slot = (first_slot + i / per) % sv.slots;
        sv.fast_hist[(size_t)slot * per + (i % per)] = 0;
    }
    if (i < n * kLevels) {
        const int slot = (first_slot + i / kLevels) % sv.slots;
        sv.cand_count[slot * kLevels + (i % kLevels)] = 0;
        sv.lkp_count[slot * kLevels + (i % kLevels)] = 0;
    }
    if (i < n) {
        const int slot = (first_slot + i) % sv.slots;
        sv.kp_count[slot] = 0;
        sv.flags[slot] = 0;
    }
}

int launch_fast(b200vo_ctx* ctx, int first_slot, int n) {
    {
        const int total = n * kLevels * 256;
        clear_detect_state_kernel<<<div_up(total, 256), 256, 0, ctx->stream>>>(ctx->store, first_slot, n);
        B200VO_LAUNCH_CHECK(ctx);
    }
    for (int l = 0; l < kLevels; ++l) {
        const LevelGeom& g = ctx->store.lv[l];
        const int iw = g.w - 2 * kEdge, ih = g.h - 2 * kEdge;
        if (iw <= 0 || ih <= 0) continue;
        dim3 grid(div_up(iw, FT_W), div_up(ih, FT_H), n);
        fast_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->store, l, first_slot);
        B200VO_LAUNCH_CHECK(ctx);
    }
    return B200VO_OK;
}

}  // namespace b200vo
