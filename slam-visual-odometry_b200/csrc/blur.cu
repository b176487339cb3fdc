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

namespace b200vo {

constexpr int BL_W = 128, BL_H = 32;
constexpr int BL_HALO = 4;                              // 3 needed; 4 keeps the staged rows word-aligned
constexpr int BL_SWW = (BL_W + 2 * BL_HALO) / 4;        // 34 words per staged row
constexpr int BL_SH = BL_H + 6;                         // 38 staged rows

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

// Each thread produces 4 horizontally adjacent pixels in both passes: the row pass reads 3 aligned
// shared-memory words (10 bytes) for 4 outputs, the column pass 7 float4.
__global__ void __launch_bounds__(256)
blur_kernel(StoreView sv, int level, int first_slot) {
    __shared__ uint32_t s_in[BL_SH][BL_SWW];
    __shared__ __align__(16) float s_row[BL_SH][BL_W];
    // kernel taps as the float32 bit patterns OpenCV's getGaussianKernel(7, 2, CV_32F) yields
    const float k0 = __uint_as_float(0x3d8fafb1u), k1 = __uint_as_float(0x3e06387eu),
                k2 = __uint_as_float(0x3e434a39u), k3 = __uint_as_float(0x3e5d4ae0u);
    const LevelGeom g = sv.lv[level];
    const int slot = (first_slot + blockIdx.z) % sv.slots;
    const uint8_t* __restrict__ img = sv.pyr + (size_t)slot * sv.slot_bytes + g.offset;
    uint8_t* out = sv.blur + (size_t)slot * sv.slot_bytes + g.offset;
    const int ox = blockIdx.x * BL_W, oy = blockIdx.y * BL_H;
    const int tid = threadIdx.x;

    // stage rows oy-3 .. oy+BL_H+2, bytes ox-4 .. ox+BL_W+3 (BORDER_REFLECT_101 at the image edges)
    for (int i = tid; i < BL_SH * BL_SWW; i += 256) {
        const int ty = i / BL_SWW, wx = i - ty * BL_SWW;
        const int gy = reflect101(min(oy - 3 + ty, g.h + 2), g.h);
        const int gx = ox - BL_HALO + 4 * wx;
        const uint8_t* row = img + (size_t)gy * g.pitch;
        uint32_t v;
        if (gx >= 0 && gx + 3 < g.w) {
            v = __ldg(reinterpret_cast<const uint32_t*>(row + gx));
        } else {
            v = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int x = reflect101(min(gx + b, g.w + 2), g.w);
                v |= (uint32_t)__ldg(row + x) << (8 * b);
            }
        }
        s_in[ty][wx] = v;
    }
    __syncthreads();

    // row pass: s = k0*p[-3]; s = fma(p[-3+i], k[i], s), i = 1..6   (sequential, as OpenCV)
    for (int i = tid; i < BL_SH * (BL_W / 4); i += 256) {
        const int ty = i / (BL_W / 4), q = i - ty * (BL_W / 4);
        // outputs x = 4q..4q+3 need staged bytes 4q+1 .. 4q+10 -> words q, q+1, q+2
        const uint32_t w0 = s_in[ty][q], w1 = s_in[ty][q + 1], w2 = s_in[ty][q + 2];
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

int launch_blur(b200vo_ctx* ctx, int first_slot, int n) {
    for (int l = 0; l < kLevels; ++l) {
        const LevelGeom& g = ctx->store.lv[l];
        dim3 grid(div_up(g.w, BL_W), div_up(g.h, BL_H), n);
        blur_kernel<<<grid, 256, 0, ctx->stream>>>(ctx->store, l, first_slot);
        B200VO_LAUNCH_CHECK(ctx);
    }
    return B200VO_OK;
}

}  // namespace b200vo
