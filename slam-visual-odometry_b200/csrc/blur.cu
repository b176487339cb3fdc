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

constexpr int BL_W = 64, BL_H = 32;
constexpr int BL_SW = BL_W + 6, BL_SH = BL_H + 6;

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return i;
}

__global__ void __launch_bounds__(256)
blur_kernel(StoreView sv, int level, int first_slot) {
    __shared__ uint8_t s_in[BL_SH][BL_SW + 2];
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

    for (int i = tid; i < BL_SH * BL_SW; i += 256) {
        const int ty = i / BL_SW, tx = i - ty * BL_SW;
        // clamp first so that far-outside tile coordinates stay reflectable, then reflect
        const int gy = reflect101(min(oy - 3 + ty, g.h + 2), g.h);
        const int gx = reflect101(min(ox - 3 + tx, g.w + 2), g.w);
        s_in[ty][tx] = __ldg(img + (size_t)gy * g.pitch + gx);
    }
    __syncthreads();

    for (int i = tid; i < BL_SH * BL_W; i += 256) {
        const int ty = i / BL_W, tx = i - ty * BL_W;
        const uint8_t* p = &s_in[ty][tx];
        float s = __fmul_rn(k0, (float)p[0]);
        s = __fmaf_rn((float)p[1], k1, s);
        s = __fmaf_rn((float)p[2], k2, s);
        s = __fmaf_rn((float)p[3], k3, s);
        s = __fmaf_rn((float)p[4], k2, s);
        s = __fmaf_rn((float)p[5], k1, s);
        s = __fmaf_rn((float)p[6], k0, s);
        s_row[ty][tx] = s;
    }
    __syncthreads();

    const int tx4 = (tid & 15) * 4;
    for (int ty = tid >> 4; ty < BL_H; ty += 16) {
        const int gy = oy + ty, gx = ox + tx4;
        if (gy >= g.h || gx >= g.w) continue;
        float4 r[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) r[j] = *reinterpret_cast<const float4*>(&s_row[ty + j][tx4]);
        uint32_t packed = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float v0 = (&r[0].x)[c], v1 = (&r[1].x)[c], v2 = (&r[2].x)[c], v3 = (&r[3].x)[c],
                        v4 = (&r[4].x)[c], v5 = (&r[5].x)[c], v6 = (&r[6].x)[c];
            float s = __fmul_rn(k3, v3);
            s = __fmaf_rn(__fadd_rn(v4, v2), k2, s);
            s = __fmaf_rn(__fadd_rn(v5, v1), k1, s);
            s = __fmaf_rn(__fadd_rn(v6, v0), k0, s);
            int q = __float2int_rn(s);
            q = min(max(q, 0), 255);
            packed |= (uint32_t)q << (8 * c);
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
