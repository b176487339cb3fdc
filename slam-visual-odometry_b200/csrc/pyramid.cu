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
#include "tma.cuh"

namespace b200vo {

// (3735 B + 19235 G + 9798 R + 2^14) >> 15   
__global__ void bgr2gray_kernel(StoreView sv, const uint8_t* __restrict__ bgr, size_t pitch, size_t frame_stride,
                                int first_slot) {
    pdl_grid_sync();
    const LevelGeom g = sv.lv[0];
    const int slot = ring_slot(first_slot, blockIdx.z, sv.slots);
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

// Gray frames already on the device (dense or arbitrarily pitched) -> level 0 of the frame store (128-B pitch).  A kernel
// instead of cudaMemcpy3DAsync: a pitched device-to-device copy of 1241-byte rows runs on a copy engine row by row, and
// during the e2e pipeline it queued behind the host->device transfer of the NEXT batch on that engine; 16 output bytes
// per thread from unaligned source rows through aligned 32-bit loads and funnel shifts.
__global__ void __launch_bounds__(256)
repack_gray_kernel(StoreView sv, const uint8_t* __restrict__ src, size_t pitch, size_t frame_stride, int first_slot) {
    pdl_grid_sync();
    const LevelGeom g = sv.lv[0];
    const int slot = ring_slot(first_slot, blockIdx.z, sv.slots);
    const int y = blockIdx.y;
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 16;
    if (x >= g.w) return;
    const uint8_t* row = src + blockIdx.z * frame_stride + (size_t)y * pitch + x;
    uint8_t* dst = sv.pyr + (size_t)slot * sv.slot_bytes + g.offset + (size_t)y * g.pitch + x;
    const uintptr_t a = reinterpret_cast<uintptr_t>(row);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const int off = (int)(a & 3), sh = off * 8;
    // only the aligned words that hold a needed byte are read: at the two ends of the source that touches at most the rest
    // of an aligned 4-byte word, which cannot leave an allocation (device allocations are at least 4-byte granular)
    const int nbytes = min(16, g.w - x);
    uint32_t v[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) v[i] = (4 * i < off + nbytes) ? __ldg(w + i) : 0u;
    uint4 o;
    o.x = __funnelshift_r(v[0], v[1], sh); o.y = __funnelshift_r(v[1], v[2], sh);
    o.z = __funnelshift_r(v[2], v[3], sh); o.w = __funnelshift_r(v[3], v[4], sh);
    if (nbytes == 16) { *reinterpret_cast<uint4*>(dst) = o; return; }
    const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
    for (int i = 0; i < nbytes; ++i) dst[i] = (uint8_t)(ow[i >> 2] >> (8 * (i & 3)));
}

__global__ void __launch_bounds__(256)
fill32_kernel(uint32_t* __restrict__ p, uint32_t value, size_t words) {
    pdl_grid_sync();
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i + 4 <= words && (reinterpret_cast<uintptr_t>(p + i) & 15) == 0) { *reinterpret_cast<uint4*>(p + i) = make_uint4(value, value, value, value); return; }
    for (size_t k = i; k < words && k < i + 4; ++k) p[k] = value;
}

int device_fill32(b200vo_ctx* ctx, void* p, uint32_t value, size_t words) {
    if (words == 0) return B200VO_OK;
    if (!ctx->fill_kernel) {
        B200VO_CUDA_CHECK(cudaMemsetAsync(p, (int)(value & 0xFF), words * sizeof(uint32_t), ctx->stream));
        return B200VO_OK;
    }
    const size_t threads = (words + 3) / 4;
    launch_k(ctx, fill32_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, static_cast<uint32_t*>(p), value, words);
    B200VO_LAUNCH_CHECK(ctx);
    return B200VO_OK;
}

int launch_repack_gray(b200vo_ctx* ctx, const uint8_t* d_src, size_t pitch, size_t frame_stride, int first_slot, int n) {
    const LevelGeom& g = ctx->store.lv[0];
    dim3 grid(div_up(div_up(g.w, 16), 256), g.h, n);
    launch_k(ctx, repack_gray_kernel, grid, dim3(256), 0, ctx->store, d_src, pitch, frame_stride, first_slot);
    B200VO_LAUNCH_CHECK(ctx);
    return B200VO_OK;
}

// One thread produces 4 horizontally adjacent output pixels (one 32-bit store).
// tab: xofs[dwp], xc1[dwp], yofs[dhp], yc1[dhp]  (dwp/dhp = sizes padded to 4)
__global__ void __launch_bounds__(256)
resize_kernel(StoreView sv, const int* __restrict__ tab, int level, int first_slot) {
    pdl_grid_sync();
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
    const int slot = ring_slot(first_slot, blockIdx.z, sv.slots);
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

// TMA variant: a 128x16 output tile.  Its source footprint (<= 171 x 22 px at scale ~1.2, start aligned
// down to 16 bytes) arrives as one cp.async.bulk.tensor box; the horizontal 8.8 pass is done once per
// source row into shared memory (u16), the vertical 16.16 pass reads two of those rows per output.
// Per-thread column taps live in registers for the whole tile.  Same integer arithmetic, bit-exact.
constexpr int RZ_W = 128, RZ_H = 16;

__global__ void __launch_bounds__(256)
resize_tma_kernel(StoreView sv, const int* __restrict__ tab, int level, int first_slot, const __grid_constant__ CUtensorMap tmap) {
    pdl_grid_sync();
    __shared__ __align__(128) uint8_t s_src[kResizeBoxH * kResizeBoxW];
    __shared__ __align__(16) uint16_t s_h[kResizeBoxH][RZ_W];
    __shared__ __align__(8) uint64_t s_bar;
    const LevelGeom gs = sv.lv[level - 1];
    const LevelGeom gd = sv.lv[level];
    const int dwp = (gd.w + 3) & ~3, dhp = (gd.h + 3) & ~3;
    const int* __restrict__ xofs = tab + gd.tab_off;
    const int* __restrict__ xc1 = xofs + dwp;
    const int* __restrict__ yofs = xc1 + dwp;
    const int* __restrict__ yc1 = yofs + dhp;
    const int slot = ring_slot(first_slot, blockIdx.z, sv.slots);
    const int x0 = blockIdx.x * RZ_W, y0 = blockIdx.y * RZ_H;
    const int tid = threadIdx.x;
    const int sx0 = xofs[x0] & ~15;
    const int sy0 = yofs[y0];
    const int ylast = min(y0 + RZ_H, gd.h) - 1;
    const int nrows = min(min(yofs[ylast] + 1, gs.h - 1) - sy0 + 1, kResizeBoxH);

    if (tid == 0) { tma::mbar_init(&s_bar, 1); tma::fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
        tma::mbar_arrive_expect_tx(&s_bar, kResizeBoxH * kResizeBoxW);
        tma::load_3d(s_src, &tmap, &s_bar, sx0, sy0, slot);
    }
    // this thread's four output columns: tap offsets relative to the box, in registers
    const int cx = (tid & 31) * 4;
    int o0[4], o1[4], c1[4];
    {
        const int4 xo = *reinterpret_cast<const int4*>(xofs + x0 + cx);
        const int4 xc = *reinterpret_cast<const int4*>(xc1 + x0 + cx);
        const int xos[4] = {xo.x, xo.y, xo.z, xo.w}, xcs[4] = {xc.x, xc.y, xc.z, xc.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool valid = (x0 + cx + i) < gd.w;
            o0[i] = valid ? xos[i] - sx0 : 0;
            o1[i] = valid ? min(xos[i] + 1, gs.w - 1) - sx0 : 0;
            c1[i] = xcs[i];
        }
    }
    tma::mbar_wait(&s_bar, 0);

    // horizontal pass, 8.8 fixed point: h = I[s]*(256-c1) + I[s+1]*c1
    for (int r = tid >> 5; r < nrows; r += 8) {
        const uint8_t* row = s_src + r * kResizeBoxW;
        ushort4 h;
        h.x = (uint16_t)(row[o0[0]] * (256 - c1[0]) + row[o1[0]] * c1[0]);
        h.y = (uint16_t)(row[o0[1]] * (256 - c1[1]) + row[o1[1]] * c1[1]);
        h.z = (uint16_t)(row[o0[2]] * (256 - c1[2]) + row[o1[2]] * c1[2]);
        h.w = (uint16_t)(row[o0[3]] * (256 - c1[3]) + row[o1[3]] * c1[3]);
        *reinterpret_cast<ushort4*>(&s_h[r][cx]) = h;
    }
    __syncthreads();

    // vertical pass, 16.16 fixed point, +0.5, >> 16
    if (x0 + cx >= gd.w) return;
#pragma unroll
    for (int k = 0; k < RZ_H / 8; ++k) {
        const int y = y0 + (tid >> 5) + 8 * k;
        if (y >= gd.h) break;
        const int ys = yofs[y];
        const int r0 = ys - sy0, r1 = min(ys + 1, gs.h - 1) - sy0;
        const uint32_t cy1 = yc1[y], cy0 = 256 - cy1;
        const ushort4 a = *reinterpret_cast<const ushort4*>(&s_h[r0][cx]);
        const ushort4 b = *reinterpret_cast<const ushort4*>(&s_h[r1][cx]);
        const uint32_t v0 = (a.x * cy0 + b.x * cy1 + 32768u) >> 16, v1 = (a.y * cy0 + b.y * cy1 + 32768u) >> 16;
        const uint32_t v2 = (a.z * cy0 + b.z * cy1 + 32768u) >> 16, v3 = (a.w * cy0 + b.w * cy1 + 32768u) >> 16;
        uint8_t* dst = sv.pyr + (size_t)slot * sv.slot_bytes + gd.offset + (size_t)y * gd.pitch + x0 + cx;
        *reinterpret_cast<uint32_t*>(dst) = v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);
    }
}

// Large batches: a warp walks down a strip of R output rows x 128 columns and keeps the horizontal 8.8 results of the two
// current source rows in registers.  resize_tma_kernel spends ~50 thread-instructions per output pixel (ncu, round 2):
// the u16 tile's shared-memory round trip, the per-row table reads from global memory and the setup of a thread that
// produces only 8 pixels.  Here a source row is filtered horizontally exactly once per lane (8 byte loads, 8 IMAD) as
// the walk reaches it, the vertical 16.16 pass is 8 IMAD, and the four result bytes are picked with three PRMT
// (v >> 16 is byte 2 of v): ~45 instructions per 4 output pixels.  The tile is 8 strips tall (128 x 8R outputs); its
// source footprint arrives as ONE TMA box of 176 x kResizeStripBoxH bytes; the row taps of the tile are staged in shared
// memory while the box is in flight.  Same integer arithmetic as above, bit-exact.
template <int R>
__global__ void __launch_bounds__(256, 6)       // 40 registers, no spills: 6 CTAs / SM instead of 5 at 47 registers (0.275 -> 0.257 ms per 251 frames)
resize_strip_kernel(StoreView sv, const int* __restrict__ tab, int level, int first_slot, const __grid_constant__ CUtensorMap tmap) {
    pdl_grid_sync();
    constexpr int TH = 8 * R;
    __shared__ __align__(128) uint8_t s_src[kResizeStripBoxH * kResizeBoxW];
    __shared__ int s_yo[TH], s_yc[TH];
    __shared__ __align__(8) uint64_t s_bar;
    const LevelGeom gs = sv.lv[level - 1];
    const LevelGeom gd = sv.lv[level];
    const int dwp = (gd.w + 3) & ~3, dhp = (gd.h + 3) & ~3;
    const int* __restrict__ xofs = tab + gd.tab_off;
    const int* __restrict__ xc1 = xofs + dwp;
    const int* __restrict__ yofs = xc1 + dwp;
    const int* __restrict__ yc1 = yofs + dhp;
    const int slot = ring_slot(first_slot, blockIdx.z, sv.slots);
    const int x0 = blockIdx.x * RZ_W, y0 = blockIdx.y * TH;
    const int tid = threadIdx.x;
    const int sx0 = xofs[x0] & ~15;
    const int sy0 = yofs[y0];

    if (tid == 0) { tma::mbar_init(&s_bar, 1); tma::fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
        tma::mbar_arrive_expect_tx(&s_bar, kResizeStripBoxH * kResizeBoxW);
        tma::load_3d(s_src, &tmap, &s_bar, sx0, sy0, slot);
    }
    for (int i = tid; i < TH; i += 256) {
        const int y = y0 + i;
        if (y < gd.h) { s_yo[i] = yofs[y] - sy0; s_yc[i] = yc1[y]; }
    }
    // this lane's four output columns: tap offsets relative to the box and weights, in registers
    const int cx = (tid & 31) * 4;
    int o0[4], o1[4];
    uint32_t c0[4], c1[4];
    {
        const int4 xo = *reinterpret_cast<const int4*>(xofs + x0 + cx);
        const int4 xc = *reinterpret_cast<const int4*>(xc1 + x0 + cx);
        const int xos[4] = {xo.x, xo.y, xo.z, xo.w}, xcs[4] = {xc.x, xc.y, xc.z, xc.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool valid = (x0 + cx + i) < gd.w;
            o0[i] = valid ? xos[i] - sx0 : 0;
            o1[i] = valid ? min(xos[i] + 1, gs.w - 1) - sx0 : 0;
            c1[i] = (uint32_t)xcs[i];
            c0[i] = 256u - c1[i];
        }
    }
    __syncthreads();
    tma::mbar_wait(&s_bar, 0);

    const int ty0 = (tid >> 5) * R;                              // warp = strip
    const int nrows = min(R, gd.h - (y0 + ty0));
    if (nrows <= 0 || x0 + cx >= gd.w) return;
    const int rmax = gs.h - 1 - sy0;                             // last source row, box-relative
    auto hpass = [&](int r, uint32_t* h) {                       // 8.8 fixed point: h = I[s]*(256-c1) + I[s+1]*c1
        const uint8_t* row = s_src + r * kResizeBoxW;
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = row[o0[i]] * c0[i] + row[o1[i]] * c1[i];
    };
    int cur = s_yo[ty0];
    uint32_t a[4], b[4];
    hpass(cur, a);
    hpass(min(cur + 1, rmax), b);
    uint8_t* dst = sv.pyr + (size_t)slot * sv.slot_bytes + gd.offset + (size_t)(y0 + ty0) * gd.pitch + x0 + cx;
    for (int i = 0; i < nrows; ++i) {
        const int ys = s_yo[ty0 + i];
        const uint32_t cy1 = (uint32_t)s_yc[ty0 + i], cy0 = 256u - cy1;
#pragma unroll 1
        while (cur < ys) {                                       // warp-uniform: one or two steps at scale 1.2
#pragma unroll
            for (int k = 0; k < 4; ++k) a[k] = b[k];
            ++cur;
            hpass(min(cur + 1, rmax), b);
        }
        // vertical pass, 16.16 fixed point, +0.5, >> 16: the result is byte 2 of each sum
        const uint32_t v0 = a[0] * cy0 + b[0] * cy1 + 32768u, v1 = a[1] * cy0 + b[1] * cy1 + 32768u;
        const uint32_t v2 = a[2] * cy0 + b[2] * cy1 + 32768u, v3 = a[3] * cy0 + b[3] * cy1 + 32768u;
        const uint32_t lo = __byte_perm(v0, v1, 0x0062), hi = __byte_perm(v2, v3, 0x0062);      // (v0 >> 16) & 0xFF | ((v1 >> 16) & 0xFF) << 8
        *reinterpret_cast<uint32_t*>(dst) = __byte_perm(lo, hi, 0x5410);
        dst += gd.pitch;
    }
}

int launch_bgr2gray(b200vo_ctx* ctx, const uint8_t* d_bgr, size_t pitch, size_t frame_stride, int first_slot, int n) {
    const LevelGeom& g = ctx->store.lv[0];
    dim3 block(64), grid(div_up(div_up(g.w, 4), 64), g.h, n);
    launch_k(ctx, bgr2gray_kernel, dim3(grid), dim3(block), 0, ctx->store, d_bgr, pitch, frame_stride, first_slot);
    B200VO_LAUNCH_CHECK(ctx);
    return B200VO_OK;
}

int launch_pyramid(b200vo_ctx* ctx, int first_slot, int n) {
    for (int l = 1; l < kLevels; ++l) {
        const LevelGeom& g = ctx->store.lv[l];
        if (ctx->use_tma && ctx->resize_strip && ctx->resize_strip_ok[l] && n > ctx->small_batch_max) {
            dim3 grid(div_up(g.w, RZ_W), div_up(g.h, 8 * kResizeStripRows), n);
            launch_k(ctx, resize_strip_kernel<kResizeStripRows>, dim3(grid), dim3(256), 0, ctx->store, ctx->d_tab, l, first_slot, ctx->tmap_resize_strip[l - 1]);
        } else if (ctx->use_tma) {
            dim3 grid(div_up(g.w, RZ_W), div_up(g.h, RZ_H), n);
            launch_k(ctx, resize_tma_kernel, dim3(grid), dim3(256), 0, ctx->store, ctx->d_tab, l, first_slot, ctx->tmap_resize[l - 1]);
        } else {
            dim3 block(32, 8), grid(div_up(g.w, 128), div_up(g.h, 8), n);
            launch_k(ctx, resize_kernel, dim3(grid), dim3(block), 0, ctx->store, ctx->d_tab, l, first_slot);
        }
        B200VO_LAUNCH_CHECK(ctx);
    }
    return B200VO_OK;
}

}  // namespace b200vo
