// Shared declarations of the b200vo CUDA library: context, level geometry, launch helpers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/b200vo.h"

namespace b200vo {

constexpr int kLevels = B200VO_NLEVELS;
constexpr int kEdge = 31;          // ORB edgeThreshold (reference src/visual_odometry.cpp:68)
constexpr int kHalfPatch = 15;     // patchSize 31
constexpr int kDescBytes = 32;
constexpr int kMaxSolutions = 10;  // real roots of the degree-10 five-point polynomial
// TMA boxes (bytes x rows) of the tile consumers; the inner extent must be a multiple of 16 bytes
constexpr int kFastBoxW = 96, kFastBoxH = 38;      // fast.cu: 70x38 pixel tile + up to 15 bytes of alignment shift
constexpr int kFastBoxWideW = 160;                 // fast.cu: the 134-pixel-wide tile of the large levels
inline bool fast_level_is_wide(int level) { return level <= 4; }
constexpr int kBlurBoxW = 160, kBlurBoxH = 38;     // blur.cu: 128+2*16 x 32+6
constexpr int kResizeBoxW = 176, kResizeBoxH = 24; // pyramid.cu: source footprint of a 128x16 output tile at scale ~1.2
constexpr int kResizeStripRows = 16;                // pyramid.cu, large batches: 8 warp strips x 16 output rows per CTA
constexpr int kResizeStripBoxH = 160;               // source rows of 128 output rows at scale ~1.2 (checked per level at create time)
constexpr int kBriefReach = 19;                     // rotated pattern points stay within |coordinate| <= 18 (max pattern radius 18.38)
constexpr int kOrientBoxW = 48, kOrientBoxH = 31;   // orient_brief.cu: 31-px disc row + up to 15 bytes of alignment shift, per keypoint
constexpr int kBriefBoxW = 64, kBriefBoxH = 39;     // orient_brief.cu: (2*19+1)-px window + up to 15 bytes of alignment shift, per keypoint

struct LevelGeom {
    int w, h, pitch;       // pitch: bytes per row, multiple of 128 (TMA needs 16)
    size_t offset;         // byte offset of the level inside a slot's pyramid buffer
    float scale;           // (float)pow((double)1.2f, l)
    int nfeat;             // per-level budget N_l
    int cand_cap;          // FAST candidate list capacity
    int sel_cap;           // power-of-two capacity of the selection working set
    int kp_cap;            // capacity of the level keypoint list
    int tab_off;           // offset of this level's resize tables (x then y) in d_tab
};

// Small batches (the streaming step: one frame) are launch-latency bound, so one launch covers the tiles
// of several pyramid levels: blockIdx.x runs level by level (largest first), inside a level frame by
// frame, inside a frame tile by tile.  first[l] = index of level l's first CTA (levels outside the launch
// have none), tiles[l] = tiles per frame, tx[l] = tiles per row.  Large batches keep one launch per level
// with the level as a kernel parameter (fast_kernel, blur_kernel<false>): the kernels measured 2-4 %
// faster that way at 500 frames per launch.
constexpr int kMergedLevelsMaxFrames = 16;
constexpr int kFilterFastCap = 4096;               // filter.cu: survivors the compacted sort holds (16 KB of shared memory)
struct LevelTiles {
    int first[kLevels + 1];
    int tiles[kLevels];
    int tx[kLevels];
};
struct TileCoord { int level, frame, bx, by; };
struct LevelMaps { CUtensorMap m[kLevels]; };     // the per-level TMA descriptors of one consumer, as one kernel parameter
__device__ __forceinline__ TileCoord tile_coord(const LevelTiles& lt, int cta) {
    TileCoord t;
    t.level = 0;
#pragma unroll
    for (int l = 1; l < kLevels; ++l) t.level += (cta >= lt.first[l]);
    const int r = cta - lt.first[t.level];
    t.frame = r / lt.tiles[t.level];
    const int tile = r - t.frame * lt.tiles[t.level];
    t.by = tile / lt.tx[t.level];
    t.bx = tile - t.by * lt.tx[t.level];
    return t;
}
// slot (first + k) of the ring, first < slots, k <= slots: one compare instead of an integer division (~20 instructions,
// 2-3 % of the short per-tile kernels in the round-2 profile)
__host__ __device__ __forceinline__ int ring_slot(int first, int k, int slots) { const int s = first + k; return s >= slots ? s - slots : s; }
inline void level_tiles_add(LevelTiles& lt, int l, int tx, int ty, int n) {      // host: append level l (ty == 0: not in this launch)
    lt.tx[l] = tx > 0 ? tx : 1;
    lt.tiles[l] = tx * ty > 0 ? tx * ty : 1;
    lt.first[l + 1] = lt.first[l] + tx * ty * n;
}

// Device-visible description of the frame store (passed by value to kernels).
struct StoreView {
    LevelGeom lv[kLevels];
    uint8_t* pyr;              // [slots][slot_bytes]
    uint8_t* blur;             // [slots][slot_bytes]
    size_t slot_bytes;
    int slots;
    // FAST candidates: per (slot, level)
    uint32_t* cand;            // [slots][cand_total]  packed (y<<20 | x<<8 | score)
    int cand_total;            // sum of cand_cap
    int cand_off[kLevels];
    int* cand_count;           // [slots][kLevels]
    int* fast_hist;            // [slots][kLevels][256]
    // per-level selected keypoints
    uint32_t* lkp_key;         // [slots][lkp_total]  packed (y<<20 | x<<8 | score), raster order
    float* lkp_resp;           // [slots][lkp_total]
    int lkp_total;
    int lkp_off[kLevels];
    int* lkp_count;            // [slots][kLevels]
    // final features
    b200vo_keypoint* kp;       // [slots][kp_cap]
    uint8_t* desc;             // [slots][kp_cap][32]
    int2* kp_mom;              // [slots][kp_cap] intensity-centroid moments (m10, m01), canonical order
    float2* kp_rot;            // [slots][kp_cap] (cos, sin) of the keypoint angle as the descriptor uses them
    int* kp_count;             // [slots]
    int* flags;                // [slots] overflow flags
    int kp_cap;
    int fast_threshold;
};

struct PairView {
    // matching
    uint32_t* nnq;             // [pairs][kp_cap] packed (dist<<16 | idx)
    uint32_t* nnt;             // [pairs][kp_cap]
    uint32_t* nnq2;            // [pairs][kp_cap] second-nearest train of every query (ratio-test mode)
    b200vo_dmatch* good;       // [pairs][kp_cap]
    int* n_good;               // [pairs]
    int* n_cross;              // [pairs]
    double* thr;               // [pairs]
    double* mstat;             // [pairs][4] min, max, median, mean of the unfiltered match distances ([MatchDebug], reference :119-168)
    int* ovf;                  // [pairs] OR of the overflow flags of the pair's two slots
    // geometry (normalised coordinates of the good matches, SoA)
    double* x1x; double* x1y; double* x2x; double* x2y;   // [pairs][kp_cap]
    float4* xf;                // [pairs][kp_cap] the same four coordinates rounded to float (scoring prefilter)
    unsigned long long* best;  // [pairs] packed (count<<32 | ~model)
    double* E;                 // [pairs][9]
    uint8_t* inl;              // [pairs][kp_cap] inlier mask of the best model
    b200vo_pose_record* rec;   // [pairs]
    double* models;            // [pairs][nhyp_cap][10][9] five-point solutions
    int* nsol;                 // [pairs][nhyp_cap]
    int* counts;               // [pairs][nhyp_cap][10] inlier counts (parity hook; may be NULL)
    uint8_t* model_ok;         // [pairs][nhyp_cap][10] 1 where models[...] holds a solution
    int* done;                 // [pairs] adaptive RANSAC: 1 once the stopping rule is met
    int* rounds;               // [pairs] adaptive RANSAC: blocks of 128 hypotheses evaluated
    double* pose_cand;         // [pairs][24] R1(9) R2(9) t(3) from decomposeEssentialMat
    int* pose_cnt;             // [pairs][4] cheirality votes of the four candidates
    int* fb_gid;               // [pairs*nhyp_cap] hypotheses deferred to the exact root-isolation launch (ransac.cu)
    unsigned int* fb_count;    // [1] length of that queue
    int nhyp_cap;
    int kp_cap;
};

}  // namespace b200vo

struct b200vo_ctx {
    b200vo_options opt;
    int device;
    cudaStream_t stream;
    b200vo::StoreView store;
    b200vo::PairView pairs;
    int max_pairs;
    // resize coefficient tables
    int* d_tab;                // per level l>=1: xofs[w], xc1[w], yofs[h], yc1[h]
    // scratch for parity hooks / raw matcher
    void* d_scratch; size_t scratch_bytes;
    void* h_pinned; size_t pinned_bytes;
    void* h_frame_pinned; size_t frame_pinned_bytes;      // b200vo_process_frame: pinned copy of a pageable caller image
    // host uploads: dense H2D copy on a dedicated copy stream into one of two staging buffers, then a
    // device-side repack (or BGR->gray) on the compute stream; events order the two streams
    cudaStream_t copy_stream;
    uint8_t* d_stage[2]; size_t stage_bytes[2];
    cudaEvent_t stage_ready[2], stage_free[2];
    int stage_idx;
    // streaming step (b200vo_process_frame_begin / _end): the features leave on the copy stream behind feat_ready
    cudaEvent_t feat_ready, feat_copied;
    bool frame_open, frame_has_pair;
    // TMA descriptors: one per pyramid level and consumer (the box shape is part of the descriptor)
    CUtensorMap tmap_fast[b200vo::kLevels];
    CUtensorMap tmap_blur[b200vo::kLevels];
    CUtensorMap tmap_blur16[b200vo::kLevels];    // blur_strip_kernel<16>: box 160 x 134
    CUtensorMap tmap_blur24[b200vo::kLevels];    // blur_strip_kernel<24>: box 160 x 198
    CUtensorMap tmap_resize[b200vo::kLevels];    // [l] = level l as the SOURCE of level l+1
    CUtensorMap tmap_resize_strip[b200vo::kLevels];   // the same with the tall box of resize_strip_kernel
    bool resize_strip_ok[b200vo::kLevels];       // [l]: level l's tiles fit the tall box (else resize_tma_kernel)
    bool resize_strip;         // B200VO_RESIZE_STRIP=0 keeps resize_tma_kernel for large batches too
    CUtensorMap tmap_orient[b200vo::kLevels];    // per-keypoint disc box on the unblurred level
    CUtensorMap tmap_brief[b200vo::kLevels];     // per-keypoint pattern box on the BLURRED level
    bool fused_describe;       // describe_kernel (default) or the three round-1 kernels (B200VO_DESCRIBE=0)
    bool use_tma;
    int blur_rows;             // rows per warp strip of the large-batch blur (16 or 24; 0: the two-pass tile kernel; B200VO_BLUR_R)
    int match_variant;         // inner-loop variant of match_kernel (see match.cu)
    int match_csa;             // 0: 8 POPC, 1: 2 CSA + 6 POPC, 2: full Harley-Seal (4 POPC), 3: 3 CSA + 5 POPC (default)
    int small_batch_max;       // batches of at most this many frames / pairs take the small-batch code paths (kMergedLevelsMaxFrames;
                               // B200VO_SMALL_BATCH_MAX=0 sends even single frames through the large-batch kernels: tests)
    bool fill_kernel;          // device_fill32 by kernel (B200VO_FILL_KERNEL=1) instead of cudaMemsetAsync (default; byte-replicated values only)
    bool repack_kernel;        // device-side ingest of gray frames by kernel (default) or cudaMemcpy3DAsync (B200VO_REPACK_KERNEL=0)
    bool pdl;                  // programmatic dependent launch for the current call (small batches only; B200VO_PDL=0 disables)
    bool pdl_enabled;
    int fpg_variant;           // five-point launch shape (B200VO_FPG, ransac.cu)
    int filter_fast_cap;       // filter.cu: survivors the compacted sort may hold (B200VO_FILTER_FAST_CAP, tests set 0)
    int64_t launches;
    bool timing;
    b200vo_stage_times times;
    cudaEvent_t ev[16];
    double K[4];               // fx, fy, cx, cy
    std::vector<void*> allocs;
};

namespace b200vo {

void set_error(const char* fmt, ...);

#define B200VO_CUDA_CHECK(expr)                                                                   \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            b200vo::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return B200VO_ERR_CUDA;                                                               \
        }                                                                                         \
    } while (0)

#define B200VO_LAUNCH_CHECK(ctx)                                                                  \
    do {                                                                                          \
        (ctx)->launches++;                                                                        \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess) {                                                                  \
            b200vo::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return B200VO_ERR_CUDA;                                                               \
        }                                                                                         \
    } while (0)

inline int div_up(int a, int b) { return (a + b - 1) / b; }

// ---- launches ------------------------------------------------------------------------------------------------
// Every kernel of the library starts with pdl_grid_sync(): `griddepcontrol.launch_dependents` lets the NEXT kernel
// of the stream be scheduled while this one still runs, `griddepcontrol.wait` blocks until the PREVIOUS grid has
// completed and its memory is visible -- so nothing is read early.  Both are no-ops unless the launch carries the
// programmatic-stream-serialization attribute, which launch_k sets only for small batches (ctx->pdl): the
// per-frame streaming step is a chain of ~20 dependent launches of a few CTAs each, and the 3-4 us launch latency
// of each then overlaps the tail of its predecessor; large batched launches fill the GPU and keep plain ordering.
__device__ __forceinline__ void pdl_grid_sync() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}

template <typename... KArgs, typename... Args>
inline void launch_k(b200vo_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = ctx->pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);       // the error, if any, is picked up by B200VO_LAUNCH_CHECK
}

// stage launchers (each enqueues on ctx->stream and returns a b200vo_status)
// device fill of 32-bit words on the compute stream: cudaMemsetAsync, or a kernel with B200VO_FILL_KERNEL=1.  (Experiment of
// round 2: were memsets and device-to-device copies served by a copy engine, they would queue behind the upload of the next
// batch in the e2e pipeline.  They are not: e2e 29.7 ms either way, the streaming step 0.325 ms with memsets vs 0.330 with
// kernels -- profiles/README.md.)
int device_fill32(b200vo_ctx* ctx, void* p, uint32_t value, size_t words);
int launch_bgr2gray(b200vo_ctx* ctx, const uint8_t* d_bgr, size_t pitch, size_t frame_stride, int first_slot, int n);
int launch_repack_gray(b200vo_ctx* ctx, const uint8_t* d_src, size_t pitch, size_t frame_stride, int first_slot, int n);
int launch_pyramid(b200vo_ctx* ctx, int first_slot, int n);
int launch_fast(b200vo_ctx* ctx, int first_slot, int n);
int launch_select(b200vo_ctx* ctx, int first_slot, int n);
int launch_orient(b200vo_ctx* ctx, int first_slot, int n);
int launch_describe(b200vo_ctx* ctx, int first_slot, int n);
int launch_blur(b200vo_ctx* ctx, int first_slot, int n);
int launch_brief(b200vo_ctx* ctx, int first_slot, int n);
int launch_match_slots(b200vo_ctx* ctx, int first_slot, int npairs, int slot_b_override);
int launch_match_raw(b200vo_ctx* ctx, const uint8_t* q, int n1, const uint8_t* t, int n2, int npairs,
                     uint32_t* nnq, uint32_t* nnt);
int launch_match_knn2_raw(b200vo_ctx* ctx, const uint8_t* q, int n1, const uint8_t* t, int n2, int npairs,
                          uint32_t* nn1, uint32_t* nn2);
int launch_filter(b200vo_ctx* ctx, int first_slot, int npairs, int slot_b_override);
int launch_ransac(b200vo_ctx* ctx, int npairs, int64_t first_pair_id, const int32_t* d_samples, int nhyp);
int launch_recover_pose(b200vo_ctx* ctx, int npairs);

}  // namespace b200vo
