// K0 BGR->gray and K1 image pyramid.
//
// What it replaces: cv::cvtColor(BGR2GRAY) (reference src/visual_odometry.cpp:105-111) and the
// pyramid construction inside cv::ORB::detectAndCompute (src/visual_odometry.cpp:113): 8 levels,
// scale 1.2^l, level l = resize(level l-1, INTER_LINEAR_EXACT) -- a dependent chain of
// fixed-point bilinear resamplings.  Integer arithmetic only, so the
// result is bit-exact by construction: 8.8 horizontal taps, 16.16 vertical taps, +32768 >> 16.
//
// Layout: every level of every slot lives in one HBM buffer, rows pitched to 128 B.  One launch
// resamples level l for ALL frames of the batch (grid.z = frame), so a 1000-frame sequence costs
// 7 launches, not 7000.
#include "common.cuh"

namespace b200vo {

// (3735 B + 19235 G + 9798 R + 2^14) >> 15   
__global__ void bgr2gray_kernel(StoreView sv, const uint8_t* __restrict__ bgr, size_t pitch, size_t frame_stride,
                                int first_slot) {
    const LevelGeom g = sv.lv[0];
    const int slot = (first_slot + blockIdx.z) % sv.slots;
    const int y = blockIdx.y;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (x >= g.w) return;
    const uint8_t* src = bgr + blockIdx.z * frame_stride + (size_t)y * pitch + (size_t)x * 3;
    uint8_t* dst = sv.pyr + (size_t)slot * sv.slot_bytes + g.offset + (size_t)y * g.pitch + x;
    uint32_t packed = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (x + i < g.w) {
            const int b = src[3 * i], gg = src[3 * i + 1], r = src[3 * i + 2];
            const uint32_t v = (uint32_t)(3735 * b + 19235 * gg + 9798 * r + 16384) >> 15;
            packed |= v << (8 * i);
        }
    }
    *reinterpret_cast<uint32_t*>(dst) = packed;   // rows are 128-B pitched, x multiple of 4
}

// One thread produces 4 horizontally adjacent output pixels (one 32-bit store).
// tab: xofs[dwp], xc1[dwp], yofs[dhp], yc1[dhp]  (dwp/dhp = sizes padded to 4)
__global__ void __launch_bounds__(256)
resize_kernel(StoreView sv, const int* __restrict__ tab, int level, int first_slot) {
    const LevelGeom gs = sv.lv[level - 1];
    const LevelGeom gd = sv.lv[level];
    const int dwp = (gd.w + 3) & ~3, dhp = (gd.h + 3) & ~3;
    const int* __restrict__ xofs = tab + gd.tab_off;
    const int* __restrict__ xc1 = xofs + dwp;
    const int* __restrict__ yofs = xc1 + dwp;
    const int* __restrict__ yc1 = yofs + dhp;

    const int x = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= gd.w || y >= gd.h) return;
    const int slot = (first_slot + blockIdx.z) % sv.slots;
    const uint8_t* __restrict__ src = sv.pyr + (size_t)slot * sv.slot_bytes + gs.offset;
    uint8_t* dst = sv.pyr + (size_t)slot * sv.slot_bytes + gd.offset + (size_t)y * gd.pitch + x;

    const int y0 = yofs[y];
    const int cy1 = yc1[y];
    const int cy0 = 256 - cy1;
    const int y1 = min(y0 + 1, gs.h - 1);
    const uint8_t* __restrict__ r0 = src + (size_t)y0 * gs.pitch;
    const uint8_t* __restrict__ r1 = src + (size_t)y1 * gs.pitch;
    const int4 xo = *reinterpret_cast<const int4*>(xofs + x);
    const int4 xc = *reinterpret_cast<const int4*>(xc1 + x);
    const int xos[4] = {xo.x, xo.y, xo.z, xo.w};
    const int xcs[4] = {xc.x, xc.y, xc.z, xc.w};
    uint32_t packed = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int o = xos[i];
        const int o1 = min(o + 1, gs.w - 1);
        const uint32_t c1 = xcs[i], c0 = 256 - c1;
        const uint32_t h0 = __ldg(r0 + o) * c0 + __ldg(r0 + o1) * c1;     // 8.8
        const uint32_t h1 = __ldg(r1 + o) * c0 + __ldg(r1 + o1) * c1;
        const uint32_t v = (h0 * cy0 + h1 * cy1 + 32768u) >> 16;           // 16.16 -> u8
        packed |= v << (8 * i);
    }
    *reinterpret_cast<uint32_t*>(dst) = packed;
}

int launch_bgr2gray(b200vo_ctx* ctx, const uint8_t* d_bgr, size_t pitch, size_t frame_stride, int first_slot, int n) {
    const LevelGeom& g = ctx->store.lv[0];
    dim3 block(64), grid(div_up(div_up(g.w, 4), 64), g.h, n);
    bgr2gray_kernel<<<grid, block, 0, ctx->stream>>>(ctx->store, d_bgr, pitch, frame_stride, first_slot);
    B200VO_LAUNCH_CHECK(ctx);
    return B200VO_OK;
}

int launch_pyramid(b200vo_ctx* ctx, int first_slot, int n) {
    for (int l = 1; l < kLevels; ++l) {
        const LevelGeom& g = ctx->store.lv[l];
        dim3 block(32, 8), grid(div_up(g.w, 128), div_up(g.h, 8), n);
        resize_kernel<<<grid, block, 0, ctx->stream>>>(ctx->store, ctx->d_tab, l, first_slot);
        B200VO_LAUNCH_CHECK(ctx);
    }
    return B200VO_OK;
}

}  // namespace b200vo
