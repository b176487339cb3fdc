// K7 7x7 sigma=2 Gaussian blur of every pyramid level (input of the rBRIEF tests).
//
// What it replaces: the GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) inside
// cv::ORB::detectAndCompute (reference src/visual_odometry.cpp:113).  OpenCV runs it through its
// float32 separable filter; descriptor parity needs the SAME float operation order
// (SURVEY.md App. A.7):
//   row pass   (u8 -> f32): s = k[0]*p[x-3];  s = fma(p[x-3+i], k[i], s), i = 1..6
//   column pass(f32 -> u8): s = k[3]*r[y];    s = fma(r[y+j] + r[y-j], k[3+j], s), j = 1..3
//   out = saturate_u8(round-half-even(s))
// Every rounding is written explicitly (__fmul_rn/__fmaf_rn/__fadd_rn) so the compiler's FMA
// contraction setting cannot change the result.
#include "common.cuh"
#include "tma.cuh"

namespace b200vo {

constexpr int BL_W = 128, BL_H = 32;
constexpr int BL_HALO = 16;                             // 3 needed; 16 keeps the TMA box start 16-byte aligned
constexpr int BL_SWW = (BL_W + 2 * BL_HALO) / 4;        // 40 words per staged row
constexpr int BL_SH = BL_H + 6;                         // 38 staged rows
static_assert(BL_SWW * 4 == kBlurBoxW && BL_SH == kBlurBoxH, "TMA box must match the staged tile");

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

// Each thread produces 4 horizontally adjacent pixels in both passes: the row pass reads 3 aligned
// shared-memory words (10 bytes) for 4 outputs, the column pass 7 float4.
template <bool MERGED>
__global__ void __launch_bounds__(256)
blur_kernel(StoreView sv, const __grid_constant__ LevelTiles lt, int level_arg, int first_slot, const __grid_constant__ LevelMaps maps, int use_tma) {
    pdl_grid_sync();
    __shared__ __align__(128) uint32_t s_in[BL_SH][BL_SWW];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(16) float s_row[BL_SH][BL_W];
    // kernel taps as the float32 bit patterns OpenCV's getGaussianKernel(7, 2, CV_32F) yields
    const float k0 = __uint_as_float(0x3d8fafb1u), k1 = __uint_as_float(0x3e06387eu),
                k2 = __uint_as_float(0x3e434a39u), k3 = __uint_as_float(0x3e5d4ae0u);
    int level = level_arg, bx = blockIdx.x, by = blockIdx.y, frame = blockIdx.z;
    if (MERGED) { const TileCoord tc = tile_coord(lt, blockIdx.x); level = tc.level; bx = tc.bx; by = tc.by; frame = tc.frame; }
    const LevelGeom g = sv.lv[level];
    const int slot = ring_slot(first_slot, frame, sv.slots);
    const uint8_t* __restrict__ img = sv.pyr + (size_t)slot * sv.slot_bytes + g.offset;
    uint8_t* out = sv.blur + (size_t)slot * sv.slot_bytes + g.offset;
    const int ox = bx * BL_W, oy = by * BL_H;
    const int tid = threadIdx.x;

    // stage rows oy-3 .. oy+BL_H+2, bytes ox-16 .. ox+BL_W+15 (BORDER_REFLECT_101 at the image edges)
    const bool border = (ox == 0) || (ox + BL_W + 3 > g.w) || (oy < 3) || (oy + BL_H + 3 > g.h);
    if (use_tma) {
        if (tid == 0) { tma::mbar_init(&s_bar, 1); tma::fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) {
            tma::mbar_arrive_expect_tx(&s_bar, BL_SH * BL_SWW * 4);
            tma::load_3d(&s_in[0][0], &maps.m[level], &s_bar, ox - BL_HALO, oy - 3, slot);
        }
        tma::mbar_wait(&s_bar, 0);
        if (border) {
            // the hardware zero-fills outside the image; OpenCV reflects (101).  Every reflected
            // source pixel lies inside this tile, so the halo is patched from shared memory:
            // columns first (all staged rows), then whole rows.
            uint8_t* sb = reinterpret_cast<uint8_t*>(&s_in[0][0]);
            const int rowb = BL_SWW * 4;
            __syncthreads();
            for (int i = tid; i < BL_SH * 6; i += 256) {
                const int ty = i / 6, k = i - ty * 6;
                const int x = (k < 3) ? (-1 - k) : (g.w + (k - 3));
                const int xs = (k < 3) ? (1 + k) : (g.w - 2 - (k - 3));
                const int c = x - (ox - BL_HALO), cs = xs - (ox - BL_HALO);
                if (c >= 0 && c < rowb && cs >= 0 && cs < rowb) sb[ty * rowb + c] = sb[ty * rowb + cs];
            }
            __syncthreads();
            for (int i = tid; i < 6 * BL_SWW; i += 256) {
                const int k = i / BL_SWW, wx = i - k * BL_SWW;
                const int y = (k < 3) ? (-1 - k) : (g.h + (k - 3));
                const int ys = (k < 3) ? (1 + k) : (g.h - 2 - (k - 3));
                const int r = y - (oy - 3), rs = ys - (oy - 3);
                if (r >= 0 && r < BL_SH && rs >= 0 && rs < BL_SH) s_in[r][wx] = s_in[rs][wx];
            }
        }
    } else {
        for (int i = tid; i < BL_SH * BL_SWW; i += 256) {
            const int ty = i / BL_SWW, wx = i - ty * BL_SWW;
            const int gy = reflect101(min(max(oy - 3 + ty, -3), g.h + 2), g.h);
            const int gx = ox - BL_HALO + 4 * wx;
            const uint8_t* row = img + (size_t)gy * g.pitch;
            uint32_t v;
            if (gx >= 0 && gx + 3 < g.w) {
                v = __ldg(reinterpret_cast<const uint32_t*>(row + gx));
            } else {
                v = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int x = reflect101(min(max(gx + b, -3), g.w + 2), g.w);
                    v |= (uint32_t)__ldg(row + x) << (8 * b);
                }
            }
            s_in[ty][wx] = v;
        }
    }
    __syncthreads();

    // row pass: s = k0*p[-3]; s = fma(p[-3+i], k[i], s), i = 1..6   (sequential, as OpenCV)
    for (int i = tid; i < BL_SH * (BL_W / 4); i += 256) {
        const int ty = i / (BL_W / 4), q = i - ty * (BL_W / 4);
        // outputs x = 4q..4q+3 need staged bytes 4q+13 .. 4q+22 -> words q+3, q+4, q+5
        const uint32_t w0 = s_in[ty][q + 3], w1 = s_in[ty][q + 4], w2 = s_in[ty][q + 5];
        float p[10];
        p[0] = (float)((w0 >> 8) & 0xFF); p[1] = (float)((w0 >> 16) & 0xFF); p[2] = (float)(w0 >> 24);
        p[3] = (float)(w1 & 0xFF); p[4] = (float)((w1 >> 8) & 0xFF); p[5] = (float)((w1 >> 16) & 0xFF); p[6] = (float)(w1 >> 24);
        p[7] = (float)(w2 & 0xFF); p[8] = (float)((w2 >> 8) & 0xFF); p[9] = (float)((w2 >> 16) & 0xFF);
        float4 r;
        float* rr = &r.x;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float s = __fmul_rn(k0, p[c]);
            s = __fmaf_rn(p[c + 1], k1, s);
            s = __fmaf_rn(p[c + 2], k2, s);
            s = __fmaf_rn(p[c + 3], k3, s);
            s = __fmaf_rn(p[c + 4], k2, s);
            s = __fmaf_rn(p[c + 5], k1, s);
            s = __fmaf_rn(p[c + 6], k0, s);
            rr[c] = s;
        }
        *reinterpret_cast<float4*>(&s_row[ty][4 * q]) = r;
    }
    __syncthreads();

    // column pass: s = k3*r[y]; s = fma(r[y+j] + r[y-j], k[3+j], s), j = 1..3; round-half-even, saturate
    for (int i = tid; i < BL_H * (BL_W / 4); i += 256) {
        const int ty = i / (BL_W / 4), q = i - ty * (BL_W / 4);
        const int gy = oy + ty, gx = ox + 4 * q;
        if (gy >= g.h || gx >= g.w) continue;
        float4 r[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) r[j] = *reinterpret_cast<const float4*>(&s_row[ty + j][4 * q]);
        uint32_t packed = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float v0 = (&r[0].x)[c], v1 = (&r[1].x)[c], v2 = (&r[2].x)[c], v3 = (&r[3].x)[c],
                        v4 = (&r[4].x)[c], v5 = (&r[5].x)[c], v6 = (&r[6].x)[c];
            float s = __fmul_rn(k3, v3);
            s = __fmaf_rn(__fadd_rn(v4, v2), k2, s);
            s = __fmaf_rn(__fadd_rn(v5, v1), k1, s);
            s = __fmaf_rn(__fadd_rn(v6, v0), k0, s);
            int qv = __float2int_rn(s);
            qv = min(max(qv, 0), 255);
            packed |= (uint32_t)qv << (8 * c);
        }
        *reinterpret_cast<uint32_t*>(out + (size_t)gy * g.pitch + gx) = packed;   // pitch >= round_up(w,4)
    }
}

// Large batches: the same arithmetic, organised so that the row-pass results never leave registers.
//
// Per-line profile of blur_kernel at the bench configuration (profiles/ncu_r02_lines_blur_kernel.txt): 38 thread-instructions
// per pixel at 80 % issue utilisation -- issue-bound; of the 153 instructions a thread spends per 4 pixels only 56 are the
// float operations the result needs, the rest is the float tile's shared-memory round trip (one 16-byte store and seven
// 16-byte loads per 4 pixels), index arithmetic of two strided loops and the barrier between the passes.  Here a warp owns
// a strip of R output rows x 128 columns and walks down it: a lane row-filters 4 adjacent pixels of the next staged row
// (3 aligned words in, 28 float operations), keeps the last seven results in a register ring (the loop is unrolled by 7
// so every ring index is a compile-time constant) and emits one output word per step: ~85 instructions per step.
// The tile is 8 strips tall (128 x 8R outputs, one TMA box of 160 x (8R + 6) bytes); the row pass is repeated for the
// 6 halo rows of every strip, (R + 6) / R of the row work.  Same operation order, bit-exact (tests/test_gpu_features.py).
__device__ __forceinline__ float4 blur_row4(const uint32_t* __restrict__ s, float k0, float k1, float k2, float k3) {
    // outputs x = 4q..4q+3 need staged bytes 4q+13 .. 4q+22: s points at word q+3
    const uint32_t w0 = s[0], w1 = s[1], w2 = s[2];
    float p[10];
    p[0] = (float)((w0 >> 8) & 0xFF); p[1] = (float)((w0 >> 16) & 0xFF); p[2] = (float)(w0 >> 24);
    p[3] = (float)(w1 & 0xFF); p[4] = (float)((w1 >> 8) & 0xFF); p[5] = (float)((w1 >> 16) & 0xFF); p[6] = (float)(w1 >> 24);
    p[7] = (float)(w2 & 0xFF); p[8] = (float)((w2 >> 8) & 0xFF); p[9] = (float)((w2 >> 16) & 0xFF);
    float4 r;
    float* rr = &r.x;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float v = __fmul_rn(k0, p[c]);
        v = __fmaf_rn(p[c + 1], k1, v);
        v = __fmaf_rn(p[c + 2], k2, v);
        v = __fmaf_rn(p[c + 3], k3, v);
        v = __fmaf_rn(p[c + 4], k2, v);
        v = __fmaf_rn(p[c + 5], k1, v);
        v = __fmaf_rn(p[c + 6], k0, v);
        rr[c] = v;
    }
    return r;
}
__device__ __forceinline__ uint32_t blur_col4(const float4& a0, const float4& a1, const float4& a2, const float4& a3,
                                              const float4& a4, const float4& a5, const float4& a6, float k0, float k1, float k2, float k3) {
    uint32_t packed = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float v0 = (&a0.x)[c], v1 = (&a1.x)[c], v2 = (&a2.x)[c], v3 = (&a3.x)[c], v4 = (&a4.x)[c], v5 = (&a5.x)[c], v6 = (&a6.x)[c];
        float v = __fmul_rn(k3, v3);
        v = __fmaf_rn(__fadd_rn(v4, v2), k2, v);
        v = __fmaf_rn(__fadd_rn(v5, v1), k1, v);
        v = __fmaf_rn(__fadd_rn(v6, v0), k0, v);
        uint32_t qv;                                             // round-half-even + saturate to 0..255 in one conversion
        asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(qv) : "f"(v));
        packed |= qv << (8 * c);
    }
    return packed;
}

template <int R>
__global__ void __launch_bounds__(256)          // 64 registers, 4 CTAs / SM; capping at 48 / 40 registers (5 / 6 CTAs) spills the ring: 0.448 -> 0.481 / 0.513 ms
blur_strip_kernel(StoreView sv, int level, int first_slot, const __grid_constant__ CUtensorMap tmap) {
    pdl_grid_sync();
    constexpr int TH = 8 * R, SH = TH + 6;
    __shared__ __align__(128) uint32_t s_in[SH][BL_SWW];
    __shared__ __align__(8) uint64_t s_bar;
    const float k0 = __uint_as_float(0x3d8fafb1u), k1 = __uint_as_float(0x3e06387eu),
                k2 = __uint_as_float(0x3e434a39u), k3 = __uint_as_float(0x3e5d4ae0u);
    const LevelGeom g = sv.lv[level];
    const int slot = ring_slot(first_slot, blockIdx.z, sv.slots);
    uint8_t* out = sv.blur + (size_t)slot * sv.slot_bytes + g.offset;
    const int ox = blockIdx.x * BL_W, oy = blockIdx.y * TH;
    const int tid = threadIdx.x;

    if (tid == 0) { tma::mbar_init(&s_bar, 1); tma::fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
        tma::mbar_arrive_expect_tx(&s_bar, SH * BL_SWW * 4);
        tma::load_3d(&s_in[0][0], &tmap, &s_bar, ox - BL_HALO, oy - 3, slot);
    }
    tma::mbar_wait(&s_bar, 0);
    if ((ox == 0) || (ox + BL_W + 3 > g.w) || (oy < 3) || (oy + TH + 3 > g.h)) {
        // BORDER_REFLECT_101 halo patched from the tile itself (every reflected source pixel lies inside it):
        // columns first (all staged rows), then whole rows
        uint8_t* sb = reinterpret_cast<uint8_t*>(&s_in[0][0]);
        const int rowb = BL_SWW * 4;
        __syncthreads();
        for (int i = tid; i < SH * 6; i += 256) {
            const int ty = i / 6, k = i - ty * 6;
            const int x = (k < 3) ? (-1 - k) : (g.w + (k - 3));
            const int xs = (k < 3) ? (1 + k) : (g.w - 2 - (k - 3));
            const int c = x - (ox - BL_HALO), cs = xs - (ox - BL_HALO);
            if (c >= 0 && c < rowb && cs >= 0 && cs < rowb) sb[ty * rowb + c] = sb[ty * rowb + cs];
        }
        __syncthreads();
        for (int i = tid; i < 6 * BL_SWW; i += 256) {
            const int k = i / BL_SWW, wx = i - k * BL_SWW;
            const int y = (k < 3) ? (-1 - k) : (g.h + (k - 3));
            const int ys = (k < 3) ? (1 + k) : (g.h - 2 - (k - 3));
            const int r = y - (oy - 3), rs = ys - (oy - 3);
            if (r >= 0 && r < SH && rs >= 0 && rs < SH) s_in[r][wx] = s_in[rs][wx];
        }
        __syncthreads();
    }

    const int q = tid & 31, ty0 = (tid >> 5) * R;               // lane = word column, warp = strip
    const int gx = ox + 4 * q;
    const int nrows = min(R, g.h - (oy + ty0));                 // output rows of this strip inside the level
    if (nrows <= 0 || gx >= g.w) return;
    const uint32_t* src = &s_in[ty0][q + 3];                    // staged row ty0 + t feeds step t; output row ty0 + t - 6
    uint8_t* dst = out + (size_t)(oy + ty0) * g.pitch + gx;     // pitch >= round_up(w, 4): whole words can be stored
    float4 ring[7];
#pragma unroll
    for (int u = 0; u < 6; ++u) ring[u] = blur_row4(src + u * BL_SWW, k0, k1, k2, k3);
    for (int t0 = 6; t0 < nrows + 6; t0 += 7) {
#pragma unroll
        for (int u = 0; u < 7; ++u) {
            const int t = t0 + u;                                // t % 7 == (6 + u) % 7
            if (t < nrows + 6) {
                ring[(6 + u) % 7] = blur_row4(src + t * BL_SWW, k0, k1, k2, k3);
                const uint32_t packed = blur_col4(ring[u % 7], ring[(u + 1) % 7], ring[(u + 2) % 7], ring[(u + 3) % 7],
                                                  ring[(u + 4) % 7], ring[(u + 5) % 7], ring[(u + 6) % 7], k0, k1, k2, k3);
                *reinterpret_cast<uint32_t*>(dst + (size_t)(t - 6) * g.pitch) = packed;
            }
        }
    }
}

int launch_blur(b200vo_ctx* ctx, int first_slot, int n) {
    const LevelMaps& maps = *reinterpret_cast<const LevelMaps*>(ctx->tmap_blur);
    LevelTiles lt;
    lt.first[0] = 0;
    for (int l = 0; l < kLevels; ++l) level_tiles_add(lt, l, div_up(ctx->store.lv[l].w, BL_W), div_up(ctx->store.lv[l].h, BL_H), n);
    if (n <= ctx->small_batch_max) {
        // one launch for all eight levels
        launch_k(ctx, blur_kernel<true>, dim3(lt.first[kLevels]), dim3(256), 0, ctx->store, lt, 0, first_slot, maps, ctx->use_tma ? 1 : 0);
        B200VO_LAUNCH_CHECK(ctx);
        return B200VO_OK;
    }
    for (int l = 0; l < kLevels; ++l) {
        const LevelGeom& g = ctx->store.lv[l];
        if (ctx->use_tma && ctx->blur_rows == 16) {
            dim3 grid(div_up(g.w, BL_W), div_up(g.h, 8 * 16), n);
            launch_k(ctx, blur_strip_kernel<16>, dim3(grid), dim3(256), 0, ctx->store, l, first_slot, ctx->tmap_blur16[l]);
        } else if (ctx->use_tma && ctx->blur_rows == 24) {
            dim3 grid(div_up(g.w, BL_W), div_up(g.h, 8 * 24), n);
            launch_k(ctx, blur_strip_kernel<24>, dim3(grid), dim3(256), 0, ctx->store, l, first_slot, ctx->tmap_blur24[l]);
        } else {
            dim3 grid(div_up(g.w, BL_W), div_up(g.h, BL_H), n);
            launch_k(ctx, blur_kernel<false>, dim3(grid), dim3(256), 0, ctx->store, lt, l, first_slot, maps, ctx->use_tma ? 1 : 0);
        }
        B200VO_LAUNCH_CHECK(ctx);
    }
    return B200VO_OK;
}

}  // namespace b200vo
