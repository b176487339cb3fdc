// Standalone probe for the TMA tile-load path (used while bringing up tma.cuh on a B200).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../slam-visual-odometry_b200/csrc/tma.cuh"
#include <cuda_runtime.h>
using namespace b200vo;
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#ifndef BW
#define BW 80
#endif
constexpr int BH = 38;
__device__ __forceinline__ void load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(tma::smem_u32(dst)),
        "l"(map), "r"(tma::smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int x, int y, int z, unsigned char* out) {
    __shared__ __align__(128) unsigned char s[BW * BH];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { tma::mbar_init(&bar, 1); tma::fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) { tma::mbar_arrive_expect_tx(&bar, BW * BH); 
#ifdef USE_2D
        load_2d(s, &tmap, &bar, x, y);
#else
        tma::load_3d(s, &tmap, &bar, x, y, z);
#endif
    }
    tma::mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = s[i];
}
int main() {
    const int W = 1241, H = 376, P = 1280, S = 3; const size_t SB = (size_t)P * H;
    std::vector<unsigned char> h(SB * S);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)((i * 2654435761u) >> 13);
    unsigned char *d, *o; cudaMalloc(&d, h.size() + 256); cudaMalloc(&o, BW * BH);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    printf("entry point: %s q=%d fn=%p\n", cudaGetErrorString(e), (int)q, fnp);
    CUtensorMap m;
    const cuuint64_t dims[3] = {W, H, S}; const cuuint64_t str[2] = {P, SB};
    const cuuint32_t box[3] = {BW, BH, 1}; const cuuint32_t es[3] = {1, 1, 1};
#ifdef USE_2D
    const int RANK = 2;
#else
    const int RANK = 3;
#endif
    CUresult r = ((EncodeTiledFn)fnp)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, RANK, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    const int tests[6][3] = {{64, 0, 1}, {0, 0, 0}, {-16, -3, 0}, {1200, 360, 2}, {16, 27, 1}, {27, 27, 1}};
    for (auto& t : tests) {
        probe<<<1, 128>>>(m, t[0], t[1], t[2], o);
        e = cudaDeviceSynchronize();
        std::vector<unsigned char> g(BW * BH);
        cudaMemcpy(g.data(), o, g.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int yy = 0; yy < BH; ++yy) for (int xx = 0; xx < BW; ++xx) {
            const int gx = t[0] + xx, gy = t[1] + yy;
            const int zz =
#ifdef USE_2D
 0;
#else
 t[2];
#endif
            const unsigned char ref = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[(size_t)zz * SB + (size_t)gy * P + gx] : 0;
            bad += g[yy * BW + xx] != ref;
        }
        printf("box at (%d,%d,%d): %s, mismatches %d\n", t[0], t[1], t[2], cudaGetErrorString(e), bad);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
