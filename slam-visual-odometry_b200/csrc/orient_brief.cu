// K6 intensity-centroid orientation and K8 rotated-BRIEF 256-bit descriptors.
//
// What it replaces: ICAngles + computeOrbDescriptors inside cv::ORB::detectAndCompute
// (reference src/visual_odometry.cpp:113).  Semantics from SURVEY.md App. A.6/A.7:
//   * moments m10, m01 over the radius-15 disc (749 px) of the UNBLURRED level, integers;
//     angle = fastAtan2((float)m01, (float)m10) -- OpenCV's degree-7 polynomial, float32, no FMA;
//   * descriptor on the BLURRED level: theta = angle * (float)(pi/180); a = (float)cos((double)theta),
//     b = (float)sin((double)theta); each of the 512 pattern points is rotated with separate float
//     multiplies/adds (no FMA) and rounded half-to-even; bit k of byte i = I(p[16i+2k]) < I(p[16i+2k+1]).
//
// One warp per keypoint.  Orientation: the 31 rows of the disc are read as aligned 32-bit words
// (9 per row); a lane takes (row, word) items and folds four pixels per instruction with dp4a
// against precomputed per-byte weights (the x offset inside the disc, 0 outside) and masks, then an
// integer redux.sync reduction -- the moments are integers, so any summation order is exact.  Descriptor: lane = descriptor byte; the lane keeps its 16 pattern points
// in registers for the lifetime of the warp; the 32 bytes go out as one coalesced 32-B store.
// The kernel also assembles the frame's final keypoint array in canonical (octave, y, x) order.
#include "common.cuh"
#include "tma.cuh"
#include "brief_pattern.h"

namespace b200vo {

__device__ __align__(16) int8_t d_pattern[1024];
// intensity-centroid tables: [alignment of the row start (0..3)][item = row * 9 + word] ->
// .x = four signed byte weights (u of each pixel, 0 outside the disc), .y = four 0/1 masks
constexpr int kIcWords = 9, kIcItems = 31 * kIcWords;
__device__ __align__(8) uint2 d_ic_table[4][kIcItems];

__device__ __forceinline__ int dp4a_u8_s8(uint32_t a, uint32_t b, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float deg = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, deg), p3 = __fmul_rn(-0.3258083974640975f, deg),
                p5 = __fmul_rn(0.1555786518463281f, deg), p7 = __fmul_rn(-0.04432655554792128f, deg);
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, 2.2204460492503131e-16f));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, 2.2204460492503131e-16f));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.0f,
                      __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0.0f) a = __fsub_rn(180.0f, a);
    if (y < 0.0f) a = __fsub_rn(360.0f, a);
    return a;
}

// grid: (ceil(max level kp / warps per block), kLevels, n); block: 256 threads = 8 warps
__global__ void __launch_bounds__(256)
orient_kernel(StoreView sv, int first_slot) {
    pdl_grid_sync();
    const int level = blockIdx.y;
    const int slot = ring_slot(first_slot, blockIdx.z, sv.slots);
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int* counts = sv.lkp_count + slot * kLevels;
    const int cnt = counts[level];
    int base = 0, total = 0;
#pragma unroll
    for (int l = 0; l < kLevels; ++l) {
        const int c = counts[l];
        if (l < level) base += c;
        total += c;
    }
    if (level == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        if (total > sv.kp_cap) atomicOr(&sv.flags[slot], 8);
        sv.kp_count[slot] = min(total, sv.kp_cap);
    }
    if (i >= cnt || base + i >= sv.kp_cap) return;
    const LevelGeom g = sv.lv[level];
    const uint32_t key = sv.lkp_key[(size_t)slot * sv.lkp_total + sv.lkp_off[level] + i];
    const int x = (key >> 8) & 0xFFF, y = key >> 20;
    // row v of the disc starts at byte (x - 15) of image row (y + v); rows are pitched to 128 bytes, so
    // every row has the same misalignment al and spans words 0..8 of the aligned address below
    const uint8_t* __restrict__ img = sv.pyr + (size_t)slot * sv.slot_bytes + g.offset;
    const int xs = x - kHalfPatch, al = xs & 3;
    const uint8_t* __restrict__ wb = img + (size_t)(y - kHalfPatch) * g.pitch + (xs - al);
    const uint2* __restrict__ tab = d_ic_table[al];
    int m10 = 0, m01 = 0;
#pragma unroll
    for (int it = 0; it < (kIcItems + 31) / 32; ++it) {
        const int item = lane + 32 * it;
        if (item < kIcItems) {
            const int row = item / kIcWords, k = item - row * kIcWords;
            const uint32_t px = __ldg(reinterpret_cast<const uint32_t*>(wb + row * g.pitch) + k);
            const uint2 wm = __ldg(tab + item);
            m10 = dp4a_u8_s8(px, wm.x, m10);
            m01 += (row - kHalfPatch) * dp4a_u8_s8(px, wm.y, 0);
        }
    }
    m10 = __reduce_add_sync(0xffffffffu, m10);
    m01 = __reduce_add_sync(0xffffffffu, m01);
    if (lane == 0) sv.kp_mom[(size_t)slot * sv.kp_cap + base + i] = make_int2(m10, m01);
}

// One THREAD per keypoint: angle = fastAtan2(m01, m10), the rotation (cos, sin) the descriptor uses,
// and the final keypoint record in canonical (octave, y, x) order.  (As the tail of the warp-per-
// keypoint kernels these ~300 scalar instructions -- a division, a double-precision sincos -- cost a
// full warp instruction stream per keypoint; here 32 keypoints share one.)
__global__ void __launch_bounds__(256)
keypoint_finalize_kernel(StoreView sv, int first_slot) {
    pdl_grid_sync();
    const int slot = ring_slot(first_slot, blockIdx.y, sv.slots);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int* counts = sv.lkp_count + slot * kLevels;
    int level = 0, base = 0, total = 0;
#pragma unroll
    for (int l = 0; l < kLevels; ++l) {
        const int c = counts[l];
        if (i >= total + c) { level = l + 1; base = total + c; }
        total += c;
    }
    if (i >= min(total, sv.kp_cap)) return;
    const int li = i - base;
    const float scale = sv.lv[level].scale;
    const uint32_t key = sv.lkp_key[(size_t)slot * sv.lkp_total + sv.lkp_off[level] + li];
    const int x = (key >> 8) & 0xFFF, y = key >> 20;
    const int2 mom = sv.kp_mom[(size_t)slot * sv.kp_cap + i];
    b200vo_keypoint kp;
    kp.x = __fmul_rn((float)x, scale);
    kp.y = __fmul_rn((float)y, scale);
    kp.size = __fmul_rn(31.0f, scale);
    kp.angle = fast_atan2_deg((float)mom.y, (float)mom.x);
    kp.response = sv.lkp_resp[(size_t)slot * sv.lkp_total + sv.lkp_off[level] + li];
    kp.octave = level;
    kp.class_id = -1;
    kp.lx = (int16_t)x;
    kp.ly = (int16_t)y;
    sv.kp[(size_t)slot * sv.kp_cap + i] = kp;
    // descriptor rotation: theta = angle * (float)(pi/180); a = (float)cos((double)theta), b = (float)sin(...)
    const float theta = __fmul_rn(kp.angle, (float)(3.14159265358979323846 / 180.0));
    double sd, cd;
    sincos((double)theta, &sd, &cd);
    sv.kp_rot[(size_t)slot * sv.kp_cap + i] = make_float2((float)cd, (float)sd);
}

// grid: (ceil(kp_cap / 8), n); one warp per keypoint
__global__ void __launch_bounds__(256)
brief_kernel(StoreView sv, int first_slot) {
    pdl_grid_sync();
    const int slot = ring_slot(first_slot, blockIdx.y, sv.slots);
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= sv.kp_count[slot]) return;
    // this lane's 8 tests = pattern rows 8*lane .. 8*lane+7
    int8_t pat[32];
    {
        const int4* p = reinterpret_cast<const int4*>(d_pattern + lane * 32);
        *reinterpret_cast<int4*>(&pat[0]) = p[0];
        *reinterpret_cast<int4*>(&pat[16]) = p[1];
    }
    const b200vo_keypoint kp = sv.kp[(size_t)slot * sv.kp_cap + i];
    const LevelGeom g = sv.lv[kp.octave];
    const uint8_t* __restrict__ c = sv.blur + (size_t)slot * sv.slot_bytes + g.offset + (size_t)kp.ly * g.pitch + kp.lx;
    const float2 rot = sv.kp_rot[(size_t)slot * sv.kp_cap + i];       // (cos, sin) from keypoint_finalize_kernel
    const float a = rot.x, b = rot.y;
    uint32_t byte = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float x0 = (float)pat[4 * k], y0 = (float)pat[4 * k + 1];
        const float x1 = (float)pat[4 * k + 2], y1 = (float)pat[4 * k + 3];
        const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
        const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
        const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int t0 = __ldg(c + iy0 * g.pitch + ix0);
        const int t1 = __ldg(c + iy1 * g.pitch + ix1);
        byte |= (uint32_t)(t0 < t1) << k;
    }
    sv.desc[((size_t)slot * sv.kp_cap + i) * kDescBytes + lane] = (uint8_t)byte;
}

// ---------------------------------------------------------------------------------------------------------
// Fused descriptor stage (default): orientation + keypoint record + rBRIEF in ONE kernel on TMA-staged
// shared-memory boxes.  A CTA takes DK_KP consecutive keypoints of one (slot, level) -- the level lists are in
// raster order, so they are spatially adjacent -- and for each of them TMA-loads
//   * the 31 rows x 48 bytes of the UNBLURRED level that cover the radius-15 disc  (box start 16-byte aligned), and
//   * the 39 rows x 64 bytes of the BLURRED level that cover every rotated pattern point (|coordinate| <= 18)
// into shared memory (2 x 16 `cp.async.bulk.tensor.3d` per CTA, two mbarriers).  Phase A: warp per keypoint,
// the 749-pixel moments by dp4a on shared-memory words.  Phase B: thread per keypoint, fastAtan2 + the double-
// precision sincos + the cv::KeyPoint record (the scalar tail costs one instruction stream per 16 keypoints).
// Phase C: warp per keypoint, lane = descriptor byte, the 512 taps read shared memory.  The moments and the
// rotation never touch global memory (round 1: three kernels, kp_mom / kp_rot round trips, 512 scattered global
// byte loads per keypoint -- long-scoreboard 43 % in brief_kernel).  The blurred box lands while phases A and B run.
constexpr int DK_KP = 8, DK_THREADS = 128;          // 20 KB of boxes per CTA: 9 CTAs = 36 warps per SM (registers).  Measured per 251 frames: 8 keypoints
                                                    // with separate boxes (32 KB, 6-7 CTAs) 0.578 ms; 8 keypoints, shared box 0.490 ms; 16 keypoints / 256 threads 0.600 ms
constexpr int DK_OBOX = (kOrientBoxW * kOrientBoxH + 127) / 128 * 128;      // 1536: TMA destinations are 128-byte aligned
constexpr int DK_BBOX = (kBriefBoxW * kBriefBoxH + 127) / 128 * 128;        // 2560
constexpr int DK_SMEM = DK_KP * DK_BBOX;               // the unblurred box of a keypoint is dead after phase A: the blurred box lands in the same 2.5 KB

__global__ void __launch_bounds__(DK_THREADS)
describe_kernel(StoreView sv, int first_slot, const __grid_constant__ LevelMaps maps_pyr, const __grid_constant__ LevelMaps maps_blur) {
    pdl_grid_sync();
    extern __shared__ __align__(128) uint8_t dk_smem[];
    __shared__ __align__(8) uint64_t s_bar[2];
    __shared__ int s_x[DK_KP], s_y[DK_KP];
    __shared__ int2 s_mom[DK_KP];
    __shared__ float2 s_rot[DK_KP];
    const int level = blockIdx.y;
    int slot = first_slot + blockIdx.z;                    // blockIdx.z < slots
    if (slot >= sv.slots) slot -= sv.slots;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* counts = sv.lkp_count + slot * kLevels;
    const int cnt = counts[level];
    const int i0 = blockIdx.x * DK_KP;
    if (i0 >= cnt && !(level == 0 && blockIdx.x == 0)) return;      // most CTAs of the (kp_cap-sized) grid are empty: leave before any arithmetic
    int base = 0, total = 0;
#pragma unroll
    for (int l = 0; l < kLevels; ++l) {
        const int c = counts[l];
        if (l < level) base += c;
        total += c;
    }
    if (level == 0 && blockIdx.x == 0 && tid == 0) {
        if (total > sv.kp_cap) atomicOr(&sv.flags[slot], 8);
        sv.kp_count[slot] = min(total, sv.kp_cap);
    }
    const int nk = min(DK_KP, min(cnt - i0, sv.kp_cap - (base + i0)));          // uniform over the CTA
    if (nk <= 0) return;
    const LevelGeom g = sv.lv[level];
    uint8_t* s_obox = dk_smem;                              // keypoint j: [j * DK_BBOX, + DK_OBOX) during phase A ...
    uint8_t* s_bbox = dk_smem;                              // ... then [j * DK_BBOX, + DK_BBOX) holds the blurred box
    const size_t lk = (size_t)slot * sv.lkp_total + sv.lkp_off[level] + i0;
    if (tid == 0) {
        tma::mbar_init(&s_bar[0], 1); tma::mbar_init(&s_bar[1], 1); tma::fence_mbar_init();
        tma::mbar_arrive_expect_tx(&s_bar[0], nk * kOrientBoxW * kOrientBoxH);
        tma::mbar_arrive_expect_tx(&s_bar[1], nk * kBriefBoxW * kBriefBoxH);
    }
    __syncthreads();
    if (tid < nk) {
        const uint32_t key = sv.lkp_key[lk + tid];
        const int x = (key >> 8) & 0xFFF, y = key >> 20;
        s_x[tid] = x; s_y[tid] = y;
        tma::load_3d(s_obox + tid * DK_BBOX, &maps_pyr.m[level], &s_bar[0], (x - kHalfPatch) & ~15, y - kHalfPatch, slot);
    }
    __syncthreads();
    tma::mbar_wait(&s_bar[0], 0);

    // ---- phase A: moments, warp per keypoint
    for (int j = warp; j < nk; j += DK_THREADS / 32) {
        const int sh = (s_x[j] - kHalfPatch) & 15;                                // byte shift of the disc's first column in the box
        const uint32_t* __restrict__ b32 = reinterpret_cast<const uint32_t*>(s_obox + j * DK_BBOX) + (sh >> 2);
        const uint2* __restrict__ tab = d_ic_table[sh & 3];
        int m10 = 0, m01 = 0;
#pragma unroll
        for (int it = 0; it < (kIcItems + 31) / 32; ++it) {
            const int item = lane + 32 * it;
            if (item < kIcItems) {
                const int row = item / kIcWords, k = item - row * kIcWords;
                const uint32_t px = b32[row * (kOrientBoxW / 4) + k];
                const uint2 wm = __ldg(tab + item);
                m10 = dp4a_u8_s8(px, wm.x, m10);
                m01 += (row - kHalfPatch) * dp4a_u8_s8(px, wm.y, 0);
            }
        }
        m10 = __reduce_add_sync(0xffffffffu, m10);
        m01 = __reduce_add_sync(0xffffffffu, m01);
        if (lane == 0) s_mom[j] = make_int2(m10, m01);
    }
    __syncthreads();

    // ---- phase B: angle, rotation, final keypoint record (canonical (octave, y, x) order), thread per keypoint
    if (tid < nk) {
        const int x = s_x[tid], y = s_y[tid];
        // the blurred box overwrites the unblurred one (every warp is past phase A: the barrier above); it lands while this
        // thread computes the angle.  Generic-proxy reads before, async-proxy writes after: proxy fence in between.
        tma::fence_proxy_async();
        tma::load_3d(s_bbox + tid * DK_BBOX, &maps_blur.m[level], &s_bar[1], (x - kBriefReach) & ~15, y - kBriefReach, slot);
        const int2 mom = s_mom[tid];
        b200vo_keypoint kp;
        kp.x = __fmul_rn((float)x, g.scale);
        kp.y = __fmul_rn((float)y, g.scale);
        kp.size = __fmul_rn(31.0f, g.scale);
        kp.angle = fast_atan2_deg((float)mom.y, (float)mom.x);
        kp.response = sv.lkp_resp[lk + tid];
        kp.octave = level;
        kp.class_id = -1;
        kp.lx = (int16_t)x;
        kp.ly = (int16_t)y;
        sv.kp[(size_t)slot * sv.kp_cap + base + i0 + tid] = kp;
        const float theta = __fmul_rn(kp.angle, (float)(3.14159265358979323846 / 180.0));
        double sd, cd;
        sincos((double)theta, &sd, &cd);
        s_rot[tid] = make_float2((float)cd, (float)sd);
    }
    __syncthreads();
    tma::mbar_wait(&s_bar[1], 0);

    // ---- phase C: rBRIEF, warp per keypoint, lane = descriptor byte (pattern rows 8*lane .. 8*lane+7)
    float pat[32];                                         // the lane's 16 pattern points as floats, converted once
    {
        const int4* p = reinterpret_cast<const int4*>(d_pattern + lane * 32);
        const int4 w0 = p[0], w1 = p[1];
        const int wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int q = 0; q < 32; ++q) pat[q] = (float)(int)(int8_t)((uint32_t)wv[q >> 2] >> (8 * (q & 3)));
    }
    for (int j = warp; j < nk; j += DK_THREADS / 32) {
        const int sh = (s_x[j] - kBriefReach) & 15;
        const uint8_t* __restrict__ c = s_bbox + j * DK_BBOX + kBriefReach * kBriefBoxW + sh + kBriefReach;     // the keypoint itself
        const float a = s_rot[j].x, b = s_rot[j].y;
        uint32_t byte = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float x0 = pat[4 * k], y0 = pat[4 * k + 1];
            const float x1 = pat[4 * k + 2], y1 = pat[4 * k + 3];
            const int ix0 = __float2int_rn(__fsub_rn(__fmul_rn(x0, a), __fmul_rn(y0, b)));
            const int iy0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
            const int ix1 = __float2int_rn(__fsub_rn(__fmul_rn(x1, a), __fmul_rn(y1, b)));
            const int iy1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
            const int t0 = c[iy0 * kBriefBoxW + ix0];
            const int t1 = c[iy1 * kBriefBoxW + ix1];
            byte |= (uint32_t)(t0 < t1) << k;
        }
        sv.desc[((size_t)slot * sv.kp_cap + base + i0 + j) * kDescBytes + lane] = (uint8_t)byte;
    }
}

int upload_brief_pattern() {
    cudaError_t e = cudaMemcpyToSymbol(d_pattern, kBriefPattern31, sizeof(kBriefPattern31));
    if (e != cudaSuccess) { set_error("pattern upload failed: %s", cudaGetErrorString(e)); return B200VO_ERR_CUDA; }
    // intensity-centroid weights: OpenCV's umax table for patch size 31 (half-widths of the disc rows)
    static const int umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};
    static uint2 tab[4][kIcItems];
    for (int al = 0; al < 4; ++al)
        for (int row = 0; row < 31; ++row)
            for (int k = 0; k < kIcWords; ++k) {
                uint32_t w = 0, m = 0;
                const int v = row - kHalfPatch, lim = umax[v < 0 ? -v : v];
                for (int j = 0; j < 4; ++j) {
                    const int u = 4 * k + j - al - kHalfPatch;
                    if (u >= -lim && u <= lim) { w |= (uint32_t)(uint8_t)(int8_t)u << (8 * j); m |= 1u << (8 * j); }
                }
                tab[al][row * kIcWords + k] = make_uint2(w, m);
            }
    e = cudaMemcpyToSymbol(d_ic_table, tab, sizeof(tab));
    if (e != cudaSuccess) { set_error("orientation table upload failed: %s", cudaGetErrorString(e)); return B200VO_ERR_CUDA; }
    return B200VO_OK;
}

int launch_describe(b200vo_ctx* ctx, int first_slot, int n) {
    static bool configured_per_device[64] = {};
    bool& configured = configured_per_device[ctx->device & 63];
    if (!configured) {
        B200VO_CUDA_CHECK(cudaFuncSetAttribute(describe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DK_SMEM));
        configured = true;
    }
    int maxkp = 1;
    for (int l = 0; l < kLevels; ++l) maxkp = max(maxkp, ctx->store.lv[l].kp_cap);
    dim3 grid(div_up(maxkp, DK_KP), kLevels, n);
    launch_k(ctx, describe_kernel, dim3(grid), dim3(DK_THREADS), DK_SMEM, ctx->store, first_slot, *reinterpret_cast<const LevelMaps*>(ctx->tmap_orient),
                                                               *reinterpret_cast<const LevelMaps*>(ctx->tmap_brief));
    B200VO_LAUNCH_CHECK(ctx);
    return B200VO_OK;
}

// the round-1 three-kernel path (B200VO_DESCRIBE=0): kept for A/B measurements and as a cross-check of the fused kernel
int launch_orient(b200vo_ctx* ctx, int first_slot, int n) {
    int maxkp = 1;
    for (int l = 0; l < kLevels; ++l) maxkp = max(maxkp, ctx->store.lv[l].kp_cap);
    dim3 grid(div_up(maxkp, 8), kLevels, n);
    launch_k(ctx, orient_kernel, dim3(grid), dim3(256), 0, ctx->store, first_slot);
    B200VO_LAUNCH_CHECK(ctx);
    launch_k(ctx, keypoint_finalize_kernel, dim3(dim3(div_up(ctx->store.kp_cap, 256), n)), dim3(256), 0, ctx->store, first_slot);
    B200VO_LAUNCH_CHECK(ctx);
    return B200VO_OK;
}

int launch_brief(b200vo_ctx* ctx, int first_slot, int n) {
    dim3 grid(div_up(ctx->store.kp_cap, 8), n);
    launch_k(ctx, brief_kernel, dim3(grid), dim3(256), 0, ctx->store, first_slot);
    B200VO_LAUNCH_CHECK(ctx);
    return B200VO_OK;
}

}  // namespace b200vo
