// extern "C" layer of the b200vo library (include/b200vo.h): context, frame store, entry points.
// Host-side setup mirrors the reference's VisualOdometry constructor
// (src/visual_odometry.cpp:64-101): ORB parameters, pyramid geometry, per-level feature budget,
// camera matrix synthesised from the image size.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "geom_core.cuh"

namespace b200vo {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int upload_brief_pattern();
int launch_score_models(b200vo_ctx* ctx, const double* d_models, int nmodels, float thr2f, int* d_counts);
int launch_recover_pose_masked(b200vo_ctx* ctx, int npairs, uint8_t* d_mask_out);
int measure_popc_rate(b200vo_ctx* ctx, double* popc_per_second);
int init_ransac_kernels();

static inline int cv_round(float v) { return (int)lrintf(v); }
static inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <typename T>
static int dev_alloc(b200vo_ctx* ctx, T** p, size_t count) {
    void* q = nullptr;
    const size_t bytes = std::max<size_t>(count * sizeof(T), 16) + 256;      // tail padding, see pyramid.cu
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return B200VO_ERR_ALLOC; }
    e = cudaMemset(q, 0, bytes);
    if (e != cudaSuccess) { set_error("cudaMemset failed: %s", cudaGetErrorString(e)); return B200VO_ERR_CUDA; }
    ctx->allocs.push_back(q);
    *p = reinterpret_cast<T*>(q);
    return B200VO_OK;
}

// INTER_LINEAR_EXACT taps of one axis: offset and 8-bit weight of the 2nd tap
static void resize_taps(int src, int dst, int* ofs, int* c1) {
    const double inv_scale = (double)dst / (double)src;
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < dst; ++d) {
        volatile double prod = scale * ((double)d + 0.5);     // keep mul and sub separately rounded
        const double f = prod - 0.5;
        const int s = (int)std::floor(f);
        if (s >= 0 && src > 1) {
            if (s < src - 1) { ofs[d] = s; c1[d] = (int)std::nearbyint((f - s) * 256.0); }
            else { ofs[d] = src - 1; c1[d] = 0; }
        } else { ofs[d] = 0; c1[d] = 0; }
    }
}

static int ensure_scratch(b200vo_ctx* ctx, size_t bytes) {
    if (ctx->scratch_bytes >= bytes) return B200VO_OK;
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    ctx->d_scratch = nullptr; ctx->scratch_bytes = 0;
    const size_t want = round_up(bytes, 1 << 20);
    cudaError_t e = cudaMalloc(&ctx->d_scratch, want);
    if (e != cudaSuccess) { set_error("scratch cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e)); return B200VO_ERR_ALLOC; }
    ctx->scratch_bytes = want;
    return B200VO_OK;
}

static int ensure_pinned(b200vo_ctx* ctx, size_t bytes) {
    if (ctx->pinned_bytes >= bytes) return B200VO_OK;
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    ctx->h_pinned = nullptr; ctx->pinned_bytes = 0;
    const size_t want = round_up(bytes, 1 << 16);
    cudaError_t e = cudaMallocHost(&ctx->h_pinned, want);
    if (e != cudaSuccess) { set_error("cudaMallocHost(%zu) failed: %s", want, cudaGetErrorString(e)); return B200VO_ERR_ALLOC; }
    ctx->pinned_bytes = want;
    return B200VO_OK;
}

struct StageClock {
    b200vo_ctx* ctx; int n = 0;
    explicit StageClock(b200vo_ctx* c) : ctx(c) { mark(); }
    void mark() { if (ctx->timing && n < 16) cudaEventRecord(ctx->ev[n++], ctx->stream); }
    float ms(int i) const { float v = 0.f; if (ctx->timing && i + 1 < n) cudaEventElapsedTime(&v, ctx->ev[i], ctx->ev[i + 1]); return v; }
};

// cuTensorMapEncodeTiled through the runtime's driver entry point (no libcuda link dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_level_map(EncodeTiledFn fn, CUtensorMap* out, uint8_t* base, const LevelGeom& g, size_t slot_bytes, int slots,
                            int box_w, int box_h) {
    const cuuint64_t dims[3] = {(cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)slots};
    const cuuint64_t strides[2] = {(cuuint64_t)g.pitch, (cuuint64_t)slot_bytes};       // bytes, multiples of 16
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, base + g.offset, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for level %dx%d box %dx%d", (int)r, g.w, g.h, box_w, box_h); return B200VO_ERR_CUDA; }
    return B200VO_OK;
}

static bool slot_ok(const b200vo_ctx* ctx, int s) { return s >= 0 && s < ctx->store.slots; }

}  // namespace b200vo

using namespace b200vo;

extern "C" {

int b200vo_abi_version(void) { return B200VO_ABI_VERSION; }
const char* b200vo_last_error(void) { return g_err; }

int b200vo_default_options(b200vo_options* o, int width, int height) {
    if (!o || width <= 0 || height <= 0) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    std::memset(o, 0, sizeof(*o));
    o->struct_size = (int32_t)sizeof(*o);
    o->width = width; o->height = height;
    o->nfeatures = 1200;                  // reference src/visual_odometry.cpp:65
    o->fast_threshold = 10;               // :73
    o->max_frames = 2;
    o->ransac_hypotheses = 1024;
    o->match_mode = B200VO_MATCH_CROSSCHECK;
    o->ratio = 0.75f;
    o->ransac_seed = 42;
    o->ransac_threshold_px = 2.0;         // :215
    o->max_hamming = 35.0;                // :147
    o->fx = o->fy = (double)width;        // :90-98
    o->cx = width / 2.0; o->cy = height / 2.0;
    o->min_matches_for_pose = 10;         // :344
    o->min_valid_points = 10;             // :270
    o->min_inliers = 12;                  // :271
    return B200VO_OK;
}

void b200vo_destroy(b200vo_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (void* p : ctx->allocs) cudaFree(p);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    for (int i = 0; i < 2; ++i) {
        if (ctx->d_stage[i]) cudaFree(ctx->d_stage[i]);
        if (ctx->stage_ready[i]) cudaEventDestroy(ctx->stage_ready[i]);
        if (ctx->stage_free[i]) cudaEventDestroy(ctx->stage_free[i]);
    }
    if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->h_frame_pinned) cudaFreeHost(ctx->h_frame_pinned);
    for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
    if (ctx->feat_ready) cudaEventDestroy(ctx->feat_ready);
    if (ctx->feat_copied) cudaEventDestroy(ctx->feat_copied);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int b200vo_create(const b200vo_options* opt, int device, b200vo_ctx** out) {
    if (!opt || !out) { set_error("null argument"); return B200VO_ERR_BAD_ARG; }
    if (opt->struct_size != (int32_t)sizeof(b200vo_options)) { set_error("options struct_size mismatch (ABI)"); return B200VO_ERR_BAD_ARG; }
    if (opt->width < 64 || opt->height < 64 || opt->width > 4095 || opt->height > 4095) {
        set_error("image size %dx%d unsupported (64..4095)", opt->width, opt->height); return B200VO_ERR_BAD_ARG; }
    if (opt->nfeatures < 1 || opt->nfeatures > 30841) { set_error("nfeatures out of range (1..30841: the match filter keeps one 32-bit key per keypoint in shared memory)"); return B200VO_ERR_BAD_ARG; }
    if (opt->max_frames < 2) { set_error("max_frames must be >= 2"); return B200VO_ERR_BAD_ARG; }
    if (opt->ransac_hypotheses < 1) { set_error("ransac_hypotheses must be >= 1"); return B200VO_ERR_BAD_ARG; }
    if (opt->fast_threshold < 1 || opt->fast_threshold > 200) { set_error("fast_threshold out of range"); return B200VO_ERR_BAD_ARG; }
    if (opt->match_mode != B200VO_MATCH_CROSSCHECK && opt->match_mode != B200VO_MATCH_RATIO) { set_error("unknown match_mode %d", opt->match_mode); return B200VO_ERR_BAD_ARG; }
    if (opt->match_mode == B200VO_MATCH_RATIO && !(opt->ratio > 0.f && opt->ratio <= 1.f)) { set_error("ratio must be in (0, 1]"); return B200VO_ERR_BAD_ARG; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error("no CUDA device available (%s); this library has no CPU fallback", cudaGetErrorString(e));
        return B200VO_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return B200VO_ERR_BAD_ARG; }
    B200VO_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    B200VO_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
        return B200VO_ERR_CUDA;
    }

    b200vo_ctx* ctx = new b200vo_ctx();
    ctx->opt = *opt;
    ctx->device = device;
    ctx->stream = nullptr; ctx->d_scratch = nullptr; ctx->scratch_bytes = 0; ctx->h_pinned = nullptr; ctx->pinned_bytes = 0;
    ctx->copy_stream = nullptr; ctx->stage_idx = 0;
    ctx->h_frame_pinned = nullptr; ctx->frame_pinned_bytes = 0;
    for (int i = 0; i < 2; ++i) { ctx->d_stage[i] = nullptr; ctx->stage_bytes[i] = 0; ctx->stage_ready[i] = nullptr; ctx->stage_free[i] = nullptr; }
    ctx->feat_ready = ctx->feat_copied = nullptr; ctx->frame_open = ctx->frame_has_pair = false;
    ctx->launches = 0; ctx->timing = false;
    std::memset(&ctx->times, 0, sizeof(ctx->times));
    for (auto& ev : ctx->ev) ev = nullptr;
    ctx->K[0] = opt->fx > 0 ? opt->fx : (double)opt->width;
    ctx->K[1] = opt->fy > 0 ? opt->fy : (double)opt->width;
    ctx->K[2] = opt->fx > 0 ? opt->cx : opt->width / 2.0;
    ctx->K[3] = opt->fx > 0 ? opt->cy : opt->height / 2.0;
    int rc = B200VO_OK;
#define TRY(x) do { rc = (x); if (rc != B200VO_OK) { b200vo_destroy(ctx); return rc; } } while (0)
#define TRYCUDA(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) { set_error("%s failed: %s", #x, cudaGetErrorString(_e)); b200vo_destroy(ctx); return B200VO_ERR_CUDA; } } while (0)
    TRYCUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    TRYCUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        TRYCUDA(cudaEventCreateWithFlags(&ctx->stage_ready[i], cudaEventDisableTiming));
        TRYCUDA(cudaEventCreateWithFlags(&ctx->stage_free[i], cudaEventDisableTiming));
    }
    for (auto& ev : ctx->ev) TRYCUDA(cudaEventCreate(&ev));
    TRYCUDA(cudaEventCreateWithFlags(&ctx->feat_ready, cudaEventDisableTiming));
    TRYCUDA(cudaEventCreateWithFlags(&ctx->feat_copied, cudaEventDisableTiming));

    // ---- pyramid geometry and per-level budgets
    StoreView& sv = ctx->store;
    std::memset(&sv, 0, sizeof(sv));
    sv.slots = opt->max_frames;
    sv.fast_threshold = opt->fast_threshold;
    const float factor = 1.0f / 1.2f;
    float ndesired = opt->nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)kLevels));
    int sum_feat = 0;
    size_t off = 0;
    int tab_total = 0;
    for (int l = 0; l < kLevels; ++l) {
        LevelGeom& g = sv.lv[l];
        g.scale = (float)std::pow((double)1.2f, (double)l);
        g.w = cv_round((float)opt->width / g.scale);
        g.h = cv_round((float)opt->height / g.scale);
        g.pitch = (int)round_up((size_t)g.w, 128);
        g.offset = off;
        off += (size_t)g.pitch * g.h;
        if (l < kLevels - 1) { g.nfeat = cv_round(ndesired); sum_feat += g.nfeat; ndesired *= factor; }
        else g.nfeat = std::max(opt->nfeatures - sum_feat, 0);
        const long inner = (long)std::max(g.w - 2 * kEdge, 0) * std::max(g.h - 2 * kEdge, 0);
        g.cand_cap = (int)std::max<long>(1024, inner / 4 + 64);   // 3x3 strict NMS bounds the density by 1/4: cannot overflow
        int sel = 2 * g.nfeat + std::max(g.nfeat / 2, 64), p2 = 64;
        while (p2 < sel) p2 <<= 1;
        g.sel_cap = p2;
        g.kp_cap = g.nfeat + std::max(g.nfeat / 8, 32);
        g.tab_off = tab_total;
        tab_total += 2 * (int)(round_up(g.w, 4) + round_up(g.h, 4));
        sv.cand_off[l] = sv.cand_total; sv.cand_total += g.cand_cap;
        sv.lkp_off[l] = sv.lkp_total; sv.lkp_total += g.kp_cap;
        {   // shared memory of the selection kernel: sel_cap (key, response) + the sorted output pairs
            size_t p2k = 64;
            while (p2k < (size_t)g.kp_cap) p2k <<= 1;
            if ((size_t)g.sel_cap * 8 + p2k * 8 > 220 * 1024) { set_error("nfeatures too large for the selection kernel"); b200vo_destroy(ctx); return B200VO_ERR_BAD_ARG; }
        }
    }
    sv.slot_bytes = round_up(off, (size_t)sv.lv[0].pitch);       // frames form a regular 3-D array for batched copies
    sv.kp_cap = (int)round_up((size_t)opt->nfeatures + std::max(opt->nfeatures / 16, 64), 64);
    if (sv.kp_cap > 65535) { set_error("nfeatures too large for 16-bit match indices"); b200vo_destroy(ctx); return B200VO_ERR_BAD_ARG; }
    {   // filter.cu keeps pow2(kp_cap) + kFilterFastCap 32-bit keys in shared memory
        size_t p2 = 1;
        while (p2 < (size_t)sv.kp_cap) p2 <<= 1;
        if ((p2 + kFilterFastCap) * sizeof(uint32_t) > 227 * 1024) {
            set_error("nfeatures %d too large for the shared-memory match filter (at most 30841)", opt->nfeatures); b200vo_destroy(ctx); return B200VO_ERR_BAD_ARG; }
    }

    const size_t S = (size_t)sv.slots;
    TRY(dev_alloc(ctx, &sv.pyr, S * sv.slot_bytes));
    TRY(dev_alloc(ctx, &sv.blur, S * sv.slot_bytes));
    TRY(dev_alloc(ctx, &sv.cand, S * sv.cand_total));
    TRY(dev_alloc(ctx, &sv.cand_count, S * kLevels));
    TRY(dev_alloc(ctx, &sv.fast_hist, S * kLevels * 256));
    TRY(dev_alloc(ctx, &sv.lkp_key, S * sv.lkp_total));
    TRY(dev_alloc(ctx, &sv.lkp_resp, S * sv.lkp_total));
    TRY(dev_alloc(ctx, &sv.lkp_count, S * kLevels));
    TRY(dev_alloc(ctx, &sv.kp, S * sv.kp_cap));
    TRY(dev_alloc(ctx, &sv.desc, S * sv.kp_cap * kDescBytes));
    TRY(dev_alloc(ctx, &sv.kp_mom, S * sv.kp_cap));
    TRY(dev_alloc(ctx, &sv.kp_rot, S * sv.kp_cap));
    TRY(dev_alloc(ctx, &sv.kp_count, S));
    TRY(dev_alloc(ctx, &sv.flags, S));

    // resize tables
    {
        std::vector<int> tab((size_t)tab_total, 0);
        for (int l = 1; l < kLevels; ++l) {
            const LevelGeom& g = sv.lv[l];
            const LevelGeom& s = sv.lv[l - 1];
            const int dwp = (int)round_up(g.w, 4), dhp = (int)round_up(g.h, 4);
            int* xo = tab.data() + g.tab_off; int* xc = xo + dwp; int* yo = xc + dwp; int* yc = yo + dhp;
            resize_taps(s.w, g.w, xo, xc);
            resize_taps(s.h, g.h, yo, yc);
            // the TMA resize kernel stages the source footprint of a 128x16 output tile in a fixed box
            for (int x0 = 0; x0 < g.w; x0 += 128) {
                const int xl = std::min(x0 + 128, g.w) - 1;
                if (std::min(xo[xl] + 1, s.w - 1) - (xo[x0] & ~15) + 1 > kResizeBoxW) { set_error("resize footprint exceeds the TMA box (width)"); b200vo_destroy(ctx); return B200VO_ERR_BAD_ARG; }
            }
            for (int y0 = 0; y0 < g.h; y0 += 16) {
                const int yl = std::min(y0 + 16, g.h) - 1;
                if (std::min(yo[yl] + 1, s.h - 1) - yo[y0] + 1 > kResizeBoxH) { set_error("resize footprint exceeds the TMA box (height)"); b200vo_destroy(ctx); return B200VO_ERR_BAD_ARG; }
            }
            ctx->resize_strip_ok[l] = true;      // resize_strip_kernel: 128 x (8 * kResizeStripRows) output tiles
            for (int y0 = 0; y0 < g.h; y0 += 8 * kResizeStripRows) {
                const int yl = std::min(y0 + 8 * kResizeStripRows, g.h) - 1;
                if (std::min(yo[yl] + 1, s.h - 1) - yo[y0] + 1 > kResizeStripBoxH) ctx->resize_strip_ok[l] = false;
            }
        }
        // resize_tma_kernel loads the taps of a whole 128-column tile as int4 even where the tile overhangs the level
        // (values discarded): keep 128 spare entries behind the last table so that read stays inside the allocation
        TRY(dev_alloc(ctx, &ctx->d_tab, (size_t)tab_total + 256));
        TRYCUDA(cudaMemcpy(ctx->d_tab, tab.data(), tab.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    {   // TMA descriptors
        ctx->use_tma = true;
        if (const char* v = getenv("B200VO_TMA")) ctx->use_tma = (atoi(v) != 0);
        ctx->match_csa = 3;
        ctx->match_variant = 3;
        ctx->blur_rows = 24;
        ctx->resize_strip = true;
        if (const char* v = getenv("B200VO_RESIZE_STRIP")) ctx->resize_strip = (atoi(v) != 0);
        if (const char* v = getenv("B200VO_BLUR_R")) { const int c = atoi(v); if (c == 0 || c == 16 || c == 24) ctx->blur_rows = c; }
        ctx->filter_fast_cap = 1 << 30;
        ctx->fpg_variant = 0;
        ctx->pdl = false; ctx->pdl_enabled = true;
        ctx->repack_kernel = true;
        ctx->fill_kernel = false;
        ctx->small_batch_max = kMergedLevelsMaxFrames;
        if (const char* v = getenv("B200VO_SMALL_BATCH_MAX")) { const int c = atoi(v); if (c >= 0 && c <= kMergedLevelsMaxFrames) ctx->small_batch_max = c; }
        if (const char* v = getenv("B200VO_FILL_KERNEL")) ctx->fill_kernel = (atoi(v) != 0);
        if (const char* v = getenv("B200VO_REPACK_KERNEL")) ctx->repack_kernel = (atoi(v) != 0);
        if (const char* v = getenv("B200VO_PDL")) ctx->pdl_enabled = (atoi(v) != 0);
        ctx->fused_describe = ctx->use_tma;
        if (const char* v = getenv("B200VO_DESCRIBE")) ctx->fused_describe = (atoi(v) != 0);
        if (const char* v = getenv("B200VO_FPG")) ctx->fpg_variant = atoi(v);
        if (const char* v = getenv("B200VO_FILTER_FAST_CAP")) ctx->filter_fast_cap = atoi(v);
        if (const char* v = getenv("B200VO_MATCH_CSA")) { const int c = atoi(v); if (c >= 0 && c <= 3) ctx->match_csa = c; }
        if (const char* v = getenv("B200VO_MATCH_V")) { const int c = atoi(v); if (c >= 0 && c <= 3) ctx->match_variant = c; }
        void* fnp = nullptr;
        cudaDriverEntryPointQueryResult qres;
        TRYCUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qres));
        if (!fnp || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled entry point unavailable"); b200vo_destroy(ctx); return B200VO_ERR_CUDA; }
        for (int l = 0; l < kLevels; ++l) {
            TRY(encode_level_map((EncodeTiledFn)fnp, &ctx->tmap_fast[l], sv.pyr, sv.lv[l], sv.slot_bytes, sv.slots, fast_level_is_wide(l) ? kFastBoxWideW : kFastBoxW, kFastBoxH));
            TRY(encode_level_map((EncodeTiledFn)fnp, &ctx->tmap_blur[l], sv.pyr, sv.lv[l], sv.slot_bytes, sv.slots, kBlurBoxW, kBlurBoxH));
            TRY(encode_level_map((EncodeTiledFn)fnp, &ctx->tmap_blur16[l], sv.pyr, sv.lv[l], sv.slot_bytes, sv.slots, kBlurBoxW, 8 * 16 + 6));
            TRY(encode_level_map((EncodeTiledFn)fnp, &ctx->tmap_blur24[l], sv.pyr, sv.lv[l], sv.slot_bytes, sv.slots, kBlurBoxW, 8 * 24 + 6));
            TRY(encode_level_map((EncodeTiledFn)fnp, &ctx->tmap_resize[l], sv.pyr, sv.lv[l], sv.slot_bytes, sv.slots, kResizeBoxW, kResizeBoxH));
            TRY(encode_level_map((EncodeTiledFn)fnp, &ctx->tmap_resize_strip[l], sv.pyr, sv.lv[l], sv.slot_bytes, sv.slots, kResizeBoxW, kResizeStripBoxH));
            TRY(encode_level_map((EncodeTiledFn)fnp, &ctx->tmap_orient[l], sv.pyr, sv.lv[l], sv.slot_bytes, sv.slots, kOrientBoxW, kOrientBoxH));
            TRY(encode_level_map((EncodeTiledFn)fnp, &ctx->tmap_brief[l], sv.blur, sv.lv[l], sv.slot_bytes, sv.slots, kBriefBoxW, kBriefBoxH));
        }
    }
    TRY(upload_brief_pattern());
    TRY(init_ransac_kernels());

    // ---- pair buffers
    PairView& pv = ctx->pairs;
    std::memset(&pv, 0, sizeof(pv));
    ctx->max_pairs = sv.slots;
    const size_t P = (size_t)ctx->max_pairs, C = (size_t)sv.kp_cap;
    pv.kp_cap = sv.kp_cap;
    pv.nhyp_cap = opt->ransac_hypotheses;
    TRY(dev_alloc(ctx, &pv.nnq, P * C));
    TRY(dev_alloc(ctx, &pv.nnt, P * C));
    TRY(dev_alloc(ctx, &pv.nnq2, P * C));
    TRY(dev_alloc(ctx, &pv.good, P * C));
    TRY(dev_alloc(ctx, &pv.n_good, P));
    TRY(dev_alloc(ctx, &pv.n_cross, P));
    TRY(dev_alloc(ctx, &pv.thr, P));
    TRY(dev_alloc(ctx, &pv.mstat, P * 4));
    TRY(dev_alloc(ctx, &pv.ovf, P));
    TRY(dev_alloc(ctx, &pv.x1x, P * C));
    TRY(dev_alloc(ctx, &pv.x1y, P * C));
    TRY(dev_alloc(ctx, &pv.x2x, P * C));
    TRY(dev_alloc(ctx, &pv.x2y, P * C));
    TRY(dev_alloc(ctx, &pv.xf, P * C));
    TRY(dev_alloc(ctx, &pv.best, P));
    TRY(dev_alloc(ctx, &pv.E, P * 9));
    TRY(dev_alloc(ctx, &pv.inl, P * C));
    TRY(dev_alloc(ctx, &pv.rec, P));
    TRY(dev_alloc(ctx, &pv.models, P * pv.nhyp_cap * 90));
    TRY(dev_alloc(ctx, &pv.nsol, P * pv.nhyp_cap));
    TRY(dev_alloc(ctx, &pv.model_ok, P * pv.nhyp_cap * 10));
    TRY(dev_alloc(ctx, &pv.done, P));
    TRY(dev_alloc(ctx, &pv.rounds, P));
    TRY(dev_alloc(ctx, &pv.pose_cand, P * 24));
    TRY(dev_alloc(ctx, &pv.pose_cnt, P * 4));
    TRY(dev_alloc(ctx, &pv.fb_gid, P * pv.nhyp_cap));
    TRY(dev_alloc(ctx, &pv.fb_count, 4));
    pv.counts = nullptr;
#undef TRY
#undef TRYCUDA
    *out = ctx;
    return B200VO_OK;
}

int b200vo_synchronize(b200vo_ctx* ctx) {
    if (!ctx) { set_error("null ctx"); return B200VO_ERR_BAD_ARG; }
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200VO_OK;
}

int b200vo_level_info(const b200vo_ctx* ctx, int level, int* width, int* height, float* scale, int* nfeatures) {
    if (!ctx || level < 0 || level >= kLevels) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    const LevelGeom& g = ctx->store.lv[level];
    if (width) *width = g.w;
    if (height) *height = g.h;
    if (scale) *scale = g.scale;
    if (nfeatures) *nfeatures = g.nfeat;
    return B200VO_OK;
}

int b200vo_max_keypoints(const b200vo_ctx* ctx) { return ctx ? ctx->store.kp_cap : B200VO_ERR_BAD_ARG; }

static int upload_common(b200vo_ctx* ctx, const void* src, int n, size_t pitch, size_t frame_stride, int channels,
                         int first_slot, cudaMemcpyKind kind) {
    if (!ctx || !src || n <= 0) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    if (channels != 1 && channels != 3) { set_error("channels must be 1 or 3"); return B200VO_ERR_BAD_ARG; }
    const StoreView& sv = ctx->store;
    if (n > sv.slots || !slot_ok(ctx, first_slot)) { set_error("slot range out of bounds"); return B200VO_ERR_BAD_ARG; }
    const LevelGeom& g = sv.lv[0];
    if (pitch < (size_t)g.w * channels || frame_stride < pitch * g.h) { set_error("pitch/frame_stride too small"); return B200VO_ERR_BAD_ARG; }
    B200VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (kind == cudaMemcpyHostToDevice) {
        // Dense H2D copy at full PCIe rate on the copy stream (a pitched 3-D copy of 1241-byte rows
        // from host memory ran at ~10 GB/s), then the device-side repack below on the compute
        // stream.  Two staging buffers: the copy of batch k+1 overlaps the compute of batch k.
        const size_t row = (size_t)g.w * channels, fbytes = row * g.h, need = fbytes * n;
        const int sb = (ctx->stage_idx ^= 1);
        if (ctx->stage_bytes[sb] < need) {
            B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
            B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->copy_stream));
            if (ctx->d_stage[sb]) cudaFree(ctx->d_stage[sb]);
            ctx->d_stage[sb] = nullptr; ctx->stage_bytes[sb] = 0;
            const size_t want = std::max(need, ctx->stage_bytes[sb ^ 1]);      // both buffers grow to the largest batch seen: callers that
            B200VO_CUDA_CHECK(cudaMalloc(&ctx->d_stage[sb], want + 256));      // vary the batch size settle after one pass
            ctx->stage_bytes[sb] = want;
        }
        B200VO_CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->stage_free[sb], 0));
        if (pitch == row && frame_stride == fbytes) {
            B200VO_CUDA_CHECK(cudaMemcpyAsync(ctx->d_stage[sb], src, need, cudaMemcpyDefault, ctx->copy_stream));   // Default: tools/e2e_probe.py feeds a device buffer through the staged path
        } else {
            for (int i = 0; i < n; ++i)
                B200VO_CUDA_CHECK(cudaMemcpy2DAsync(ctx->d_stage[sb] + (size_t)i * fbytes, row,
                                                    static_cast<const uint8_t*>(src) + (size_t)i * frame_stride, pitch, row, g.h,
                                                    cudaMemcpyHostToDevice, ctx->copy_stream));
        }
        B200VO_CUDA_CHECK(cudaEventRecord(ctx->stage_ready[sb], ctx->copy_stream));
        B200VO_CUDA_CHECK(cudaStreamWaitEvent(ctx->stream, ctx->stage_ready[sb], 0));
        const int rc = upload_common(ctx, ctx->d_stage[sb], n, row, fbytes, channels, first_slot, cudaMemcpyDeviceToDevice);
        if (rc) return rc;
        B200VO_CUDA_CHECK(cudaEventRecord(ctx->stage_free[sb], ctx->stream));
        return B200VO_OK;
    }
    if (channels == 1 && ctx->repack_kernel) {
        // handles any pitch / alignment and the ring wrap; reads aligned 32-bit words, never beyond the word of a needed byte
        return launch_repack_gray(ctx, static_cast<const uint8_t*>(src), pitch, frame_stride, first_slot, n);
    }
    if (channels == 1) {
        int done = 0;
        while (done < n) {
            const int slot = (first_slot + done) % sv.slots;
            const int run = std::min(n - done, sv.slots - slot);       // contiguous slots until the ring wraps
            const uint8_t* s = static_cast<const uint8_t*>(src) + (size_t)done * frame_stride;
            if (frame_stride % pitch == 0) {
                cudaMemcpy3DParms p = {};
                p.srcPtr = make_cudaPitchedPtr(const_cast<uint8_t*>(s), pitch, g.w, frame_stride / pitch);
                p.dstPtr = make_cudaPitchedPtr(sv.pyr + (size_t)slot * sv.slot_bytes, g.pitch, g.w, sv.slot_bytes / g.pitch);
                p.extent = make_cudaExtent(g.w, g.h, run);
                p.kind = kind;
                B200VO_CUDA_CHECK(cudaMemcpy3DAsync(&p, ctx->stream));
            } else {
                for (int i = 0; i < run; ++i)
                    B200VO_CUDA_CHECK(cudaMemcpy2DAsync(sv.pyr + (size_t)(slot + i) * sv.slot_bytes, g.pitch,
                                                        s + (size_t)i * frame_stride, pitch, g.w, g.h, kind, ctx->stream));
            }
            done += run;
        }
        return B200VO_OK;
    }
    // BGR already on the device: convert into level 0
    return launch_bgr2gray(ctx, static_cast<const uint8_t*>(src), pitch, frame_stride, first_slot, n);
}

int b200vo_upload_frames(b200vo_ctx* ctx, const uint8_t* host, int n, size_t pitch, size_t frame_stride, int channels,
                         int first_slot) {
    return upload_common(ctx, host, n, pitch, frame_stride, channels, first_slot, cudaMemcpyHostToDevice);
}

int b200vo_upload_frames_device(b200vo_ctx* ctx, const void* dev, int n, size_t pitch, size_t frame_stride, int channels,
                                int first_slot) {
    return upload_common(ctx, dev, n, pitch, frame_stride, channels, first_slot, cudaMemcpyDeviceToDevice);
}

int b200vo_detect(b200vo_ctx* ctx, int first_slot, int n) {
    if (!ctx || n <= 0 || n > ctx->store.slots || !slot_ok(ctx, first_slot)) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    B200VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    int rc;
    // small batches (the per-frame streaming step) are a chain of dependent launches of a few CTAs each:
    // programmatic dependent launch overlaps each launch latency with the predecessor's tail (common.cuh)
    ctx->pdl = ctx->pdl_enabled && !ctx->timing && n <= ctx->small_batch_max;
    struct PdlOff { b200vo_ctx* c; ~PdlOff() { c->pdl = false; } } pdl_off{ctx};
    StageClock clk(ctx);
    if ((rc = launch_pyramid(ctx, first_slot, n))) return rc;  clk.mark();
    if ((rc = launch_fast(ctx, first_slot, n))) return rc;     clk.mark();
    if ((rc = launch_select(ctx, first_slot, n))) return rc;   clk.mark();
    if (ctx->fused_describe) {
        // blur first (it does not depend on the keypoints), then orientation + keypoint record + rBRIEF in one kernel
        if ((rc = launch_blur(ctx, first_slot, n))) return rc;     clk.mark();
        if ((rc = launch_describe(ctx, first_slot, n))) return rc; clk.mark();
    } else {
        if ((rc = launch_orient(ctx, first_slot, n))) return rc;   clk.mark();
        if ((rc = launch_blur(ctx, first_slot, n))) return rc;     clk.mark();
        if ((rc = launch_brief(ctx, first_slot, n))) return rc;    clk.mark();
    }
    if (ctx->timing) {
        B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        ctx->times.pyramid_ms = clk.ms(0); ctx->times.fast_ms = clk.ms(1); ctx->times.select_ms = clk.ms(2);
        if (ctx->fused_describe) {
            ctx->times.blur_ms = clk.ms(3); ctx->times.describe_ms = clk.ms(4); ctx->times.orient_ms = ctx->times.brief_ms = 0.f;
        } else {
            ctx->times.orient_ms = clk.ms(3); ctx->times.blur_ms = clk.ms(4); ctx->times.brief_ms = clk.ms(5); ctx->times.describe_ms = 0.f;
        }
    }
    return B200VO_OK;
}

int b200vo_get_features(b200vo_ctx* ctx, int slot, b200vo_keypoint* kp, uint8_t* desc, int cap, int* n) {
    if (!ctx || !slot_ok(ctx, slot) || !n) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    const StoreView& sv = ctx->store;
    int cnt_flags[2];
    B200VO_CUDA_CHECK(cudaMemcpyAsync(&cnt_flags[0], sv.kp_count + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(&cnt_flags[1], sv.flags + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    *n = cnt_flags[0];
    if (cnt_flags[1]) { set_error("slot %d: internal list overflow (flags=%d)", slot, cnt_flags[1]); return B200VO_ERR_CAPACITY; }
    if ((kp || desc) && cap < cnt_flags[0]) { set_error("capacity %d < %d keypoints", cap, cnt_flags[0]); return B200VO_ERR_CAPACITY; }
    if (kp) B200VO_CUDA_CHECK(cudaMemcpyAsync(kp, sv.kp + (size_t)slot * sv.kp_cap, (size_t)cnt_flags[0] * sizeof(b200vo_keypoint), cudaMemcpyDeviceToHost, ctx->stream));
    if (desc) B200VO_CUDA_CHECK(cudaMemcpyAsync(desc, sv.desc + (size_t)slot * sv.kp_cap * kDescBytes, (size_t)cnt_flags[0] * kDescBytes, cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200VO_OK;
}

int b200vo_set_features(b200vo_ctx* ctx, int slot, const b200vo_keypoint* kp, const uint8_t* desc, int n) {
    if (!ctx || !slot_ok(ctx, slot) || n < 0 || !desc) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    const StoreView& sv = ctx->store;
    if (n > sv.kp_cap) { set_error("n %d exceeds keypoint capacity %d", n, sv.kp_cap); return B200VO_ERR_CAPACITY; }
    if (kp) B200VO_CUDA_CHECK(cudaMemcpyAsync(sv.kp + (size_t)slot * sv.kp_cap, kp, (size_t)n * sizeof(b200vo_keypoint), cudaMemcpyHostToDevice, ctx->stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(sv.desc + (size_t)slot * sv.kp_cap * kDescBytes, desc, (size_t)n * kDescBytes, cudaMemcpyHostToDevice, ctx->stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(sv.kp_count + slot, &n, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    const int zero = 0;
    B200VO_CUDA_CHECK(cudaMemcpyAsync(sv.flags + slot, &zero, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200VO_OK;
}

int b200vo_get_pyramid_level(b200vo_ctx* ctx, int slot, int level, int blurred, uint8_t* out, size_t out_pitch) {
    if (!ctx || !slot_ok(ctx, slot) || level < 0 || level >= kLevels || !out) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    const StoreView& sv = ctx->store;
    const LevelGeom& g = sv.lv[level];
    if (out_pitch < (size_t)g.w) { set_error("out_pitch too small"); return B200VO_ERR_BAD_ARG; }
    const uint8_t* src = (blurred ? sv.blur : sv.pyr) + (size_t)slot * sv.slot_bytes + g.offset;
    B200VO_CUDA_CHECK(cudaMemcpy2DAsync(out, out_pitch, src, g.pitch, g.w, g.h, cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200VO_OK;
}

int b200vo_get_fast_candidates(b200vo_ctx* ctx, int slot, int level, uint32_t* out, int cap, int* n) {
    if (!ctx || !slot_ok(ctx, slot) || level < 0 || level >= kLevels || !n) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    const StoreView& sv = ctx->store;
    int cnt = 0;
    B200VO_CUDA_CHECK(cudaMemcpyAsync(&cnt, sv.cand_count + slot * kLevels + level, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    *n = cnt;
    if (cnt > sv.lv[level].cand_cap) { set_error("FAST candidate list overflow (%d > %d)", cnt, sv.lv[level].cand_cap); return B200VO_ERR_CAPACITY; }
    if (out) {
        if (cap < cnt) { set_error("capacity %d < %d candidates", cap, cnt); return B200VO_ERR_CAPACITY; }
        B200VO_CUDA_CHECK(cudaMemcpyAsync(out, sv.cand + (size_t)slot * sv.cand_total + sv.cand_off[level], (size_t)cnt * 4, cudaMemcpyDeviceToHost, ctx->stream));
        B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    return B200VO_OK;
}

int b200vo_match_descriptors_device(b200vo_ctx* ctx, const void* q_dev, int n1, const void* t_dev, int n2, int npairs,
                                    void* nnq_dev, void* nnt_dev, float* kernel_ms) {
    if (!ctx || !q_dev || !t_dev || !nnq_dev || !nnt_dev || n1 <= 0 || n2 <= 0 || npairs <= 0) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    if (n1 > 65536 || n2 > 65536) { set_error("at most 65536 descriptors per side (16-bit packed indices)"); return B200VO_ERR_BAD_ARG; }
    B200VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (kernel_ms) B200VO_CUDA_CHECK(cudaEventRecord(ctx->ev[14], ctx->stream));
    int rc = launch_match_raw(ctx, static_cast<const uint8_t*>(q_dev), n1, static_cast<const uint8_t*>(t_dev), n2, npairs,
                              static_cast<uint32_t*>(nnq_dev), static_cast<uint32_t*>(nnt_dev));
    if (rc) return rc;
    if (kernel_ms) {
        B200VO_CUDA_CHECK(cudaEventRecord(ctx->ev[15], ctx->stream));
        B200VO_CUDA_CHECK(cudaEventSynchronize(ctx->ev[15]));
        B200VO_CUDA_CHECK(cudaEventElapsedTime(kernel_ms, ctx->ev[14], ctx->ev[15]));
    }
    return B200VO_OK;
}

int b200vo_match_descriptors(b200vo_ctx* ctx, const uint8_t* q, int n1, const uint8_t* t, int n2, int npairs,
                             int32_t* nn_q, int32_t* dist_q, int32_t* nn_t, int32_t* dist_t) {
    if (!ctx || !q || !t || n1 < 0 || n2 < 0 || npairs <= 0) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    if (n1 > 65536 || n2 > 65536) { set_error("at most 65536 descriptors per side (16-bit packed indices)"); return B200VO_ERR_BAD_ARG; }
    if (n1 == 0 || n2 == 0) {      // empty side: no neighbours (cv::BFMatcher returns no matches)
        for (size_t i = 0; i < (size_t)npairs * n1; ++i) { if (nn_q) nn_q[i] = -1; if (dist_q) dist_q[i] = -1; }
        for (size_t i = 0; i < (size_t)npairs * n2; ++i) { if (nn_t) nn_t[i] = -1; if (dist_t) dist_t[i] = -1; }
        return B200VO_OK;
    }
    const size_t stride = (size_t)std::max(n1, n2);
    const size_t qb = (size_t)npairs * n1 * kDescBytes, tb = (size_t)npairs * n2 * kDescBytes, ob = (size_t)npairs * stride * 4;
    int rc = ensure_scratch(ctx, round_up(qb, 256) + round_up(tb, 256) + 2 * round_up(ob, 256));
    if (rc) return rc;
    uint8_t* dq = static_cast<uint8_t*>(ctx->d_scratch);
    uint8_t* dt = dq + round_up(qb, 256);
    uint32_t* dnq = reinterpret_cast<uint32_t*>(dt + round_up(tb, 256));
    uint32_t* dnt = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(dnq) + round_up(ob, 256));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(dq, q, qb, cudaMemcpyHostToDevice, ctx->stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(dt, t, tb, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_match_raw(ctx, dq, n1, dt, n2, npairs, dnq, dnt))) return rc;
    std::vector<uint32_t> hq((size_t)npairs * stride), ht((size_t)npairs * stride);
    B200VO_CUDA_CHECK(cudaMemcpyAsync(hq.data(), dnq, ob, cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(ht.data(), dnt, ob, cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int p = 0; p < npairs; ++p) {
        for (int i = 0; i < n1; ++i) {
            const uint32_t v = hq[(size_t)p * stride + i];
            if (nn_q) nn_q[(size_t)p * n1 + i] = (int32_t)(v & 0xFFFFu);
            if (dist_q) dist_q[(size_t)p * n1 + i] = (int32_t)(v >> 16);
        }
        for (int i = 0; i < n2; ++i) {
            const uint32_t v = ht[(size_t)p * stride + i];
            if (nn_t) nn_t[(size_t)p * n2 + i] = (int32_t)(v & 0xFFFFu);
            if (dist_t) dist_t[(size_t)p * n2 + i] = (int32_t)(v >> 16);
        }
    }
    return B200VO_OK;
}

int b200vo_knn_match_descriptors(b200vo_ctx* ctx, const uint8_t* q, int n1, const uint8_t* t, int n2, int npairs,
                                 int32_t* idx, int32_t* dist) {
    if (!ctx || !q || !t || n1 < 0 || n2 < 0 || npairs <= 0 || !idx || !dist) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    if (n1 > 65536 || n2 > 65536) { set_error("at most 65536 descriptors per side (16-bit packed indices)"); return B200VO_ERR_BAD_ARG; }
    for (size_t i = 0; i < (size_t)npairs * n1 * 2; ++i) { idx[i] = -1; dist[i] = -1; }
    if (n1 == 0 || n2 == 0) return B200VO_OK;
    B200VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    const size_t qb = (size_t)npairs * n1 * kDescBytes, tb = (size_t)npairs * n2 * kDescBytes, ob = (size_t)npairs * n1 * 4;
    int rc = ensure_scratch(ctx, round_up(qb, 256) + round_up(tb, 256) + 2 * round_up(ob, 256));
    if (rc) return rc;
    uint8_t* dq = static_cast<uint8_t*>(ctx->d_scratch);
    uint8_t* dt = dq + round_up(qb, 256);
    uint32_t* d1 = reinterpret_cast<uint32_t*>(dt + round_up(tb, 256));
    uint32_t* d2 = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(d1) + round_up(ob, 256));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(dq, q, qb, cudaMemcpyHostToDevice, ctx->stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(dt, t, tb, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_match_knn2_raw(ctx, dq, n1, dt, n2, npairs, d1, d2))) return rc;
    std::vector<uint32_t> h1((size_t)npairs * n1), h2((size_t)npairs * n1);
    B200VO_CUDA_CHECK(cudaMemcpyAsync(h1.data(), d1, ob, cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(h2.data(), d2, ob, cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < (size_t)npairs * n1; ++i) {
        if (h1[i] != 0xFFFFFFFFu) { idx[2 * i] = (int32_t)(h1[i] & 0xFFFFu); dist[2 * i] = (int32_t)(h1[i] >> 16); }
        if (h2[i] != 0xFFFFFFFFu) { idx[2 * i + 1] = (int32_t)(h2[i] & 0xFFFFu); dist[2 * i + 1] = (int32_t)(h2[i] >> 16); }
    }
    return B200VO_OK;
}

int b200vo_match(b200vo_ctx* ctx, int slot1, int slot2, b200vo_dmatch* out, int cap, int* n_good, int* n_cross,
                 double* threshold) {
    if (!ctx || !slot_ok(ctx, slot1) || !slot_ok(ctx, slot2)) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    B200VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = launch_match_slots(ctx, slot1, 1, slot2))) return rc;
    if ((rc = launch_filter(ctx, slot1, 1, slot2))) return rc;
    int g = 0, c = 0; double thr = 0;
    B200VO_CUDA_CHECK(cudaMemcpyAsync(&g, ctx->pairs.n_good, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(&c, ctx->pairs.n_cross, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(&thr, ctx->pairs.thr, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (n_good) *n_good = g;
    if (n_cross) *n_cross = c;
    if (threshold) *threshold = thr;
    if (out) {
        if (cap < g) { set_error("capacity %d < %d matches", cap, g); return B200VO_ERR_CAPACITY; }
        B200VO_CUDA_CHECK(cudaMemcpyAsync(out, ctx->pairs.good, (size_t)g * sizeof(b200vo_dmatch), cudaMemcpyDeviceToHost, ctx->stream));
        B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    return B200VO_OK;
}

static int run_pairs(b200vo_ctx* ctx, int first_slot, int npairs, int slot_b_override, int64_t first_pair_id,
                     const int32_t* d_samples, int nhyp) {
    int rc;
    ctx->pdl = ctx->pdl_enabled && !ctx->timing && npairs <= ctx->small_batch_max;
    struct PdlOff { b200vo_ctx* c; ~PdlOff() { c->pdl = false; } } pdl_off{ctx};
    StageClock clk(ctx);
    if ((rc = launch_match_slots(ctx, first_slot, npairs, slot_b_override))) return rc;    clk.mark();
    if ((rc = launch_filter(ctx, first_slot, npairs, slot_b_override))) return rc;         clk.mark();
    if ((rc = launch_ransac(ctx, npairs, first_pair_id, d_samples, nhyp))) return rc;      clk.mark();
    if ((rc = launch_recover_pose(ctx, npairs))) return rc;                                clk.mark();
    if (ctx->timing) {
        B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
        ctx->times.match_ms = clk.ms(0); ctx->times.filter_ms = clk.ms(1); ctx->times.ransac_ms = clk.ms(2); ctx->times.pose_ms = clk.ms(3);
        ctx->times.fivept_ms = ctx->times.score_ms = 0.f;
        if (!ctx->opt.ransac_adaptive || d_samples) {
            cudaEventElapsedTime(&ctx->times.fivept_ms, ctx->ev[8], ctx->ev[9]);
            cudaEventElapsedTime(&ctx->times.score_ms, ctx->ev[9], ctx->ev[10]);
        }
    }
    return B200VO_OK;
}

int b200vo_relative_pose(b200vo_ctx* ctx, int slot1, int slot2, int64_t pair_id, const int32_t* samples, int nhyp,
                         b200vo_pose_record* rec, double* essential, uint8_t* inlier_mask, int mask_cap) {
    if (!ctx || !slot_ok(ctx, slot1) || !slot_ok(ctx, slot2) || !rec) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    B200VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    if (nhyp <= 0) nhyp = ctx->opt.ransac_hypotheses;
    if (nhyp > ctx->pairs.nhyp_cap) { set_error("nhyp %d exceeds ransac_hypotheses capacity %d", nhyp, ctx->pairs.nhyp_cap); return B200VO_ERR_CAPACITY; }
    const int32_t* d_samples = nullptr;
    int rc;
    if (samples) {
        if ((rc = ensure_scratch(ctx, (size_t)nhyp * 5 * sizeof(int32_t)))) return rc;
        B200VO_CUDA_CHECK(cudaMemcpyAsync(ctx->d_scratch, samples, (size_t)nhyp * 5 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        d_samples = static_cast<const int32_t*>(ctx->d_scratch);
    }
    if ((rc = run_pairs(ctx, slot1, 1, slot2, pair_id, d_samples, nhyp))) return rc;
    B200VO_CUDA_CHECK(cudaMemcpyAsync(rec, ctx->pairs.rec, sizeof(*rec), cudaMemcpyDeviceToHost, ctx->stream));
    if (essential) B200VO_CUDA_CHECK(cudaMemcpyAsync(essential, ctx->pairs.E, 9 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (inlier_mask) {
        if (mask_cap < rec->n_matches) { set_error("mask capacity %d < %d", mask_cap, rec->n_matches); return B200VO_ERR_CAPACITY; }
        B200VO_CUDA_CHECK(cudaMemcpyAsync(inlier_mask, ctx->pairs.inl, (size_t)rec->n_matches, cudaMemcpyDeviceToHost, ctx->stream));
        B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    }
    return B200VO_OK;
}

// One streaming step of VisualOdometry::process_frame (reference src/visual_odometry.cpp:323-378): upload + detect of the
// new frame, the pair (prev_slot, slot) if prev_slot >= 0, and every result the reference's per-frame code reads --
// features, pose record, E.  Split in two so that the caller's host work on the features (the reference fills
// cv::KeyPoint / descriptor Mat per frame, :103-117) overlaps the pair's kernels:
//   begin: enqueue everything; the features leave on the COPY stream behind an event recorded after the descriptor
//          kernel, so no copy sits between two kernels of the compute stream (a copy node would also end the chain of
//          programmatic dependent launches); returns when the features are on the host
//   end:   wait for the compute stream, hand out the record and E
static int frame_begin_impl(b200vo_ctx* ctx, const uint8_t* host, size_t pitch, int channels, int slot, int prev_slot, int64_t pair_id,
                            b200vo_keypoint* kp, uint8_t* desc, int cap, int* n) {
    if (!ctx || !host || !slot_ok(ctx, slot) || prev_slot >= ctx->store.slots || !n) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    if (ctx->frame_open) { set_error("b200vo_process_frame_begin called twice without b200vo_process_frame_end"); return B200VO_ERR_BAD_ARG; }
    const StoreView& sv = ctx->store;
    const LevelGeom& g0 = sv.lv[0];
    int rc;
    if (channels != 1 && channels != 3) { set_error("channels must be 1 or 3"); return B200VO_ERR_BAD_ARG; }
    if (pitch < (size_t)g0.w * channels) { set_error("pitch too small"); return B200VO_ERR_BAD_ARG; }
    B200VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    // The reference hands over an ordinary (pageable) cv::Mat.  cudaMemcpyAsync from pageable memory goes through the
    // driver's own bounce buffers at a fraction of the PCIe rate and blocks the host meanwhile; the rows are copied into
    // the context's pinned frame buffer instead (private to this entry point, whose previous step ended with a stream
    // synchronisation, so it is free).  A gray frame is laid out there with level 0's 128-byte row pitch, so ONE dense
    // asynchronous copy on the compute stream lands it in the frame store: no staging buffer, no repack kernel, no
    // hop between two streams.  BGR frames and callers whose memory is already pinned keep the staged path.
    cudaPointerAttributes pa;
    const bool pinned = cudaPointerGetAttributes(&pa, host) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    if (!pinned) (void)cudaGetLastError();
    if (!pinned || (channels == 1 && pitch == (size_t)g0.pitch)) {
        const size_t row = (size_t)g0.w * channels;
        const size_t dpitch = (channels == 1) ? (size_t)g0.pitch : row;
        const size_t need = dpitch * (g0.h - 1) + row;               // the last row ends with its last pixel
        const uint8_t* src = host;
        if (!pinned) {
            if (ctx->frame_pinned_bytes < need) {
                if (ctx->h_frame_pinned) cudaFreeHost(ctx->h_frame_pinned);
                ctx->h_frame_pinned = nullptr; ctx->frame_pinned_bytes = 0;
                B200VO_CUDA_CHECK(cudaMallocHost(&ctx->h_frame_pinned, need));
                ctx->frame_pinned_bytes = need;
            }
            uint8_t* dst = static_cast<uint8_t*>(ctx->h_frame_pinned);
            if (pitch == dpitch) std::memcpy(dst, host, need);
            else for (int y = 0; y < g0.h; ++y) std::memcpy(dst + (size_t)y * dpitch, host + (size_t)y * pitch, row);
            src = dst;
        }
        if (channels == 1) {
            B200VO_CUDA_CHECK(cudaMemcpyAsync(sv.pyr + (size_t)slot * sv.slot_bytes + g0.offset, src, need, cudaMemcpyHostToDevice, ctx->stream));
        } else if ((rc = b200vo_upload_frames(ctx, src, 1, dpitch, dpitch * g0.h, channels, slot))) return rc;
    } else if ((rc = b200vo_upload_frames(ctx, host, 1, pitch, pitch * g0.h, channels, slot))) return rc;
    if ((rc = b200vo_detect(ctx, slot, 1))) return rc;
    // results travel through one pinned block: [count, flags | keypoints | descriptors | record | E]
    const size_t kb = (size_t)sv.kp_cap * sizeof(b200vo_keypoint), db = (size_t)sv.kp_cap * kDescBytes;
    const size_t off_kp = 64, off_desc = off_kp + kb, off_rec = round_up(off_desc + db, 64), off_E = off_rec + sizeof(b200vo_pose_record);
    if ((rc = ensure_pinned(ctx, off_E + 9 * sizeof(double)))) return rc;
    uint8_t* hp = static_cast<uint8_t*>(ctx->h_pinned);
    int* cnt_flags = reinterpret_cast<int*>(hp);
    B200VO_CUDA_CHECK(cudaEventRecord(ctx->feat_ready, ctx->stream));
    B200VO_CUDA_CHECK(cudaStreamWaitEvent(ctx->copy_stream, ctx->feat_ready, 0));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(&cnt_flags[0], sv.kp_count + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->copy_stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(&cnt_flags[1], sv.flags + slot, sizeof(int), cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (kp) B200VO_CUDA_CHECK(cudaMemcpyAsync(hp + off_kp, sv.kp + (size_t)slot * sv.kp_cap, kb, cudaMemcpyDeviceToHost, ctx->copy_stream));
    if (desc) B200VO_CUDA_CHECK(cudaMemcpyAsync(hp + off_desc, sv.desc + (size_t)slot * sv.kp_cap * kDescBytes, db, cudaMemcpyDeviceToHost, ctx->copy_stream));
    B200VO_CUDA_CHECK(cudaEventRecord(ctx->feat_copied, ctx->copy_stream));
    ctx->frame_open = true;
    ctx->frame_has_pair = prev_slot >= 0;
    if (prev_slot >= 0) {
        if ((rc = run_pairs(ctx, prev_slot, 1, slot, pair_id, nullptr, ctx->opt.ransac_hypotheses))) return rc;
        B200VO_CUDA_CHECK(cudaMemcpyAsync(hp + off_rec, ctx->pairs.rec, sizeof(b200vo_pose_record), cudaMemcpyDeviceToHost, ctx->stream));
        B200VO_CUDA_CHECK(cudaMemcpyAsync(hp + off_E, ctx->pairs.E, 9 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    B200VO_CUDA_CHECK(cudaEventSynchronize(ctx->feat_copied));
    *n = cnt_flags[0];
    // an overflowed internal list (score ties) is a soft failure here, as everything on the per-frame path of the
    // reference is: the truncation is deterministic and the pair record carries B200VO_ST_OVERFLOW
    if ((kp || desc) && cap < cnt_flags[0]) { set_error("capacity %d < %d keypoints", cap, cnt_flags[0]); return B200VO_ERR_CAPACITY; }
    if (kp) std::memcpy(kp, hp + off_kp, (size_t)cnt_flags[0] * sizeof(b200vo_keypoint));
    if (desc) std::memcpy(desc, hp + off_desc, (size_t)cnt_flags[0] * kDescBytes);
    return B200VO_OK;
}

int b200vo_process_frame_begin(b200vo_ctx* ctx, const uint8_t* host, size_t pitch, int channels, int slot, int prev_slot, int64_t pair_id,
                               b200vo_keypoint* kp, uint8_t* desc, int cap, int* n) {
    const int rc = frame_begin_impl(ctx, host, pitch, channels, slot, prev_slot, pair_id, kp, desc, cap, n);
    if (rc && ctx && ctx->frame_open) {            // a failed step is closed here: nothing stays in flight behind an error
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->copy_stream);
        ctx->frame_open = false;
    }
    return rc;
}

int b200vo_process_frame_end(b200vo_ctx* ctx, b200vo_pose_record* rec, double* essential) {
    if (!ctx || !ctx->frame_open) { set_error("b200vo_process_frame_end without b200vo_process_frame_begin"); return B200VO_ERR_BAD_ARG; }
    ctx->frame_open = false;
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (ctx->frame_has_pair) {
        if (!rec) { set_error("the step ran a pair: rec is required"); return B200VO_ERR_BAD_ARG; }
        const StoreView& sv = ctx->store;
        const size_t kb = (size_t)sv.kp_cap * sizeof(b200vo_keypoint), db = (size_t)sv.kp_cap * kDescBytes;
        const size_t off_rec = round_up(64 + kb + db, 64), off_E = off_rec + sizeof(b200vo_pose_record);
        const uint8_t* hp = static_cast<const uint8_t*>(ctx->h_pinned);
        std::memcpy(rec, hp + off_rec, sizeof(b200vo_pose_record));
        if (essential) std::memcpy(essential, hp + off_E, 9 * sizeof(double));
    }
    return B200VO_OK;
}

// begin + end: the whole step with one call
int b200vo_process_frame(b200vo_ctx* ctx, const uint8_t* host, size_t pitch, int channels, int slot, int prev_slot, int64_t pair_id,
                         b200vo_keypoint* kp, uint8_t* desc, int cap, int* n, b200vo_pose_record* rec, double* essential) {
    if (prev_slot >= 0 && !rec) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    const int rc = b200vo_process_frame_begin(ctx, host, pitch, channels, slot, prev_slot, pair_id, kp, desc, cap, n);
    if (ctx && ctx->frame_open) {
        const int rc2 = b200vo_process_frame_end(ctx, rec, essential);
        return rc ? rc : rc2;
    }
    return rc;
}

int b200vo_process_pairs_device(b200vo_ctx* ctx, int first_slot, int npairs, int64_t first_pair_id, void** records_dev) {
    if (!ctx || !slot_ok(ctx, first_slot) || npairs <= 0 || npairs > ctx->max_pairs - 1) { set_error("bad arguments (npairs must be in 1..max_frames-1)"); return B200VO_ERR_BAD_ARG; }
    B200VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    int rc = run_pairs(ctx, first_slot, npairs, -1, first_pair_id, nullptr, ctx->opt.ransac_hypotheses);
    if (rc) return rc;
    if (records_dev) *records_dev = ctx->pairs.rec;
    return B200VO_OK;
}

int b200vo_process_pairs(b200vo_ctx* ctx, int first_slot, int npairs, int64_t first_pair_id, b200vo_pose_record* out) {
    if (!out) { set_error("null output"); return B200VO_ERR_BAD_ARG; }
    int rc = b200vo_process_pairs_device(ctx, first_slot, npairs, first_pair_id, nullptr);
    if (rc) return rc;
    B200VO_CUDA_CHECK(cudaMemcpyAsync(out, ctx->pairs.rec, (size_t)npairs * sizeof(b200vo_pose_record), cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200VO_OK;
}

// ------------------------------------------------------------------ geometry parity hooks
static int stage_points(b200vo_ctx* ctx, const double* x1, const double* x2, int npts) {
    if (npts < 0 || npts > ctx->pairs.kp_cap) { set_error("npts %d exceeds capacity %d", npts, ctx->pairs.kp_cap); return B200VO_ERR_CAPACITY; }
    std::vector<double> soa((size_t)4 * npts);
    for (int i = 0; i < npts; ++i) {
        soa[i] = x1[2 * i]; soa[(size_t)npts + i] = x1[2 * i + 1];
        soa[(size_t)2 * npts + i] = x2[2 * i]; soa[(size_t)3 * npts + i] = x2[2 * i + 1];
    }
    PairView& pv = ctx->pairs;
    const size_t b = (size_t)npts * sizeof(double);
    B200VO_CUDA_CHECK(cudaMemcpyAsync(pv.x1x, soa.data(), b, cudaMemcpyHostToDevice, ctx->stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(pv.x1y, soa.data() + npts, b, cudaMemcpyHostToDevice, ctx->stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(pv.x2x, soa.data() + 2 * (size_t)npts, b, cudaMemcpyHostToDevice, ctx->stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(pv.x2y, soa.data() + 3 * (size_t)npts, b, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<float> xf((size_t)4 * npts);
    for (int i = 0; i < npts; ++i) {
        xf[4 * (size_t)i] = (float)x1[2 * i]; xf[4 * (size_t)i + 1] = (float)x1[2 * i + 1];
        xf[4 * (size_t)i + 2] = (float)x2[2 * i]; xf[4 * (size_t)i + 3] = (float)x2[2 * i + 1];
    }
    B200VO_CUDA_CHECK(cudaMemcpyAsync(pv.xf, xf.data(), xf.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(pv.n_good, &npts, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));       // soa goes out of scope
    return B200VO_OK;
}

}  // extern "C"

namespace b200vo {
int launch_five_point_only(b200vo_ctx* ctx, const int32_t* d_samples, int nhyp);
}

extern "C" {

int b200vo_five_point(b200vo_ctx* ctx, const double* x1, const double* x2, int npts, const int32_t* samples, int nhyp,
                      double* essentials, int32_t* nsol) {
    if (!ctx || !x1 || !x2 || !samples || !essentials || !nsol || nhyp <= 0 || npts < 5) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    if (nhyp > ctx->pairs.nhyp_cap) { set_error("nhyp %d exceeds capacity %d", nhyp, ctx->pairs.nhyp_cap); return B200VO_ERR_CAPACITY; }
    B200VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = stage_points(ctx, x1, x2, npts))) return rc;
    if ((rc = ensure_scratch(ctx, (size_t)nhyp * 5 * sizeof(int32_t)))) return rc;
    B200VO_CUDA_CHECK(cudaMemcpyAsync(ctx->d_scratch, samples, (size_t)nhyp * 5 * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_five_point_only(ctx, static_cast<const int32_t*>(ctx->d_scratch), nhyp))) return rc;
    std::vector<double> raw((size_t)nhyp * 90);
    std::vector<uint8_t> okv((size_t)nhyp * 10);
    B200VO_CUDA_CHECK(cudaMemcpyAsync(raw.data(), ctx->pairs.models, raw.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(okv.data(), ctx->pairs.model_ok, okv.size(), cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    for (int h = 0; h < nhyp; ++h) {          // solutions of a hypothesis, in root order, without the rejected roots
        int n = 0;
        for (int k = 0; k < 10; ++k)
            if (okv[(size_t)h * 10 + k]) { std::memcpy(essentials + ((size_t)h * 10 + n) * 9, raw.data() + ((size_t)h * 10 + k) * 9, 9 * sizeof(double)); ++n; }
        for (int k = n; k < 10; ++k) std::memset(essentials + ((size_t)h * 10 + k) * 9, 0, 9 * sizeof(double));
        nsol[h] = n;
    }
    return B200VO_OK;
}

int b200vo_score_models(b200vo_ctx* ctx, const double* x1, const double* x2, int npts, const double* essentials,
                        int nmodels, double threshold, int32_t* counts) {
    if (!ctx || !x1 || !x2 || !essentials || !counts || npts <= 0 || nmodels <= 0) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    B200VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = stage_points(ctx, x1, x2, npts))) return rc;
    const size_t mb = round_up((size_t)nmodels * 9 * sizeof(double), 256);
    if ((rc = ensure_scratch(ctx, mb + (size_t)nmodels * sizeof(int)))) return rc;
    uint8_t* base = static_cast<uint8_t*>(ctx->d_scratch);
    double* dm = reinterpret_cast<double*>(base);
    int* dc = reinterpret_cast<int*>(base + mb);
    B200VO_CUDA_CHECK(cudaMemcpyAsync(dm, essentials, (size_t)nmodels * 9 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const float thr2f = (float)(threshold * threshold);
    if ((rc = launch_score_models(ctx, dm, nmodels, thr2f, dc))) return rc;
    B200VO_CUDA_CHECK(cudaMemcpyAsync(counts, dc, (size_t)nmodels * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    return B200VO_OK;
}

int b200vo_recover_pose(b200vo_ctx* ctx, const double* essential, const double* x1, const double* x2, int npts,
                        const uint8_t* mask_in, double* R, double* t, int32_t* n_valid, uint8_t* mask_out) {
    if (!ctx || !essential || !x1 || !x2 || npts <= 0) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    B200VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = stage_points(ctx, x1, x2, npts))) return rc;
    PairView& pv = ctx->pairs;
    std::vector<uint8_t> m((size_t)npts, 1);
    int ninl = npts;
    if (mask_in) { ninl = 0; for (int i = 0; i < npts; ++i) { m[i] = mask_in[i] ? 1 : 0; ninl += m[i]; } }
    b200vo_pose_record r;
    std::memset(&r, 0, sizeof(r));
    r.status = B200VO_ST_MATCHED | B200VO_ST_E_FOUND; r.n_matches = npts; r.n_inliers = ninl;
    B200VO_CUDA_CHECK(cudaMemcpyAsync(pv.inl, m.data(), (size_t)npts, cudaMemcpyHostToDevice, ctx->stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(pv.E, essential, 9 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    B200VO_CUDA_CHECK(cudaMemcpyAsync(pv.rec, &r, sizeof(r), cudaMemcpyHostToDevice, ctx->stream));
    uint8_t* d_mask = nullptr;
    if (mask_out) {
        if ((rc = ensure_scratch(ctx, (size_t)pv.kp_cap))) return rc;
        d_mask = static_cast<uint8_t*>(ctx->d_scratch);
    }
    if ((rc = launch_recover_pose_masked(ctx, 1, d_mask))) return rc;
    B200VO_CUDA_CHECK(cudaMemcpyAsync(&r, pv.rec, sizeof(r), cudaMemcpyDeviceToHost, ctx->stream));
    if (mask_out) B200VO_CUDA_CHECK(cudaMemcpyAsync(mask_out, d_mask, (size_t)npts, cudaMemcpyDeviceToHost, ctx->stream));
    B200VO_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
    if (R) std::memcpy(R, r.R, sizeof(r.R));
    if (t) std::memcpy(t, r.t, sizeof(r.t));
    if (n_valid) *n_valid = r.n_valid;
    return B200VO_OK;
}

int b200vo_draw_samples(uint64_t seed, int64_t pair_id, int nhyp, int m, int32_t* samples) {
    if (!samples || nhyp < 0 || m < 5) { set_error("bad arguments (need m >= 5)"); return B200VO_ERR_BAD_ARG; }
    const uint64_t base = geom::pair_stream(seed, pair_id);
    for (int h = 0; h < nhyp; ++h) {
        int idx[5];
        geom::draw_sample(base, h, m, idx);
        for (int k = 0; k < 5; ++k) samples[h * 5 + k] = idx[k];
    }
    return B200VO_OK;
}

// process_frame's pose update (reference src/visual_odometry.cpp:292-314, :350-360)
int b200vo_chain_pose(const b200vo_pose_record* rec, const double* R_wc, const double* t_wc, double* R_out, double* t_out) {
    if (!rec || !R_wc || !t_wc || !R_out || !t_out) { set_error("null argument"); return B200VO_ERR_BAD_ARG; }
    if (!(rec->status & B200VO_ST_E_FOUND)) {        // fallbacks (a)/(b): pose unchanged
        for (int i = 0; i < 9; ++i) R_out[i] = R_wc[i];
        for (int i = 0; i < 3; ++i) t_out[i] = t_wc[i];
        return B200VO_OK;
    }
    const double scale = (rec->status & B200VO_ST_POSE_OK) ? 0.3 : 0.0;      // :352-353
    double Rinv[9], tinv[3], Rn[9], tn[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rinv[i * 3 + j] = rec->R[j * 3 + i];
    for (int i = 0; i < 3; ++i) tinv[i] = -(Rinv[i * 3] * rec->t[0] + Rinv[i * 3 + 1] * rec->t[1] + Rinv[i * 3 + 2] * rec->t[2]);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Rn[i * 3 + j] = R_wc[i * 3] * Rinv[j] + R_wc[i * 3 + 1] * Rinv[3 + j] + R_wc[i * 3 + 2] * Rinv[6 + j];
        tn[i] = t_wc[i] + scale * (R_wc[i * 3] * tinv[0] + R_wc[i * 3 + 1] * tinv[1] + R_wc[i * 3 + 2] * tinv[2]);
    }
    std::memcpy(R_out, Rn, sizeof(Rn));
    std::memcpy(t_out, tn, sizeof(tn));
    return B200VO_OK;
}

int b200vo_chain_trajectory(const b200vo_pose_record* recs, int n, const double* R0, const double* t0, double* poses_out) {
    if (!recs || n < 0 || !R0 || !t0 || !poses_out) { set_error("bad arguments"); return B200VO_ERR_BAD_ARG; }
    double R[9], t[3];
    std::memcpy(R, R0, sizeof(R));
    std::memcpy(t, t0, sizeof(t));
    for (int i = 0; i < n; ++i) {
        int rc = b200vo_chain_pose(recs + i, R, t, R, t);
        if (rc) return rc;
        std::memcpy(poses_out + (size_t)i * 12, R, sizeof(R));
        std::memcpy(poses_out + (size_t)i * 12 + 9, t, sizeof(t));
    }
    return B200VO_OK;
}

int b200vo_measure_popc_rate(b200vo_ctx* ctx, double* popc_per_second) {
    if (!ctx || !popc_per_second) { set_error("null argument"); return B200VO_ERR_BAD_ARG; }
    B200VO_CUDA_CHECK(cudaSetDevice(ctx->device));
    return measure_popc_rate(ctx, popc_per_second);
}

int b200vo_fundamental_from_essential(const b200vo_ctx* ctx, const double* E, double* F) {
    if (!ctx || !E || !F) { set_error("null argument"); return B200VO_ERR_BAD_ARG; }
    const double fx = ctx->K[0], fy = ctx->K[1], cx = ctx->K[2], cy = ctx->K[3];
    // K^-1 = [[1/fx, 0, -cx/fx], [0, 1/fy, -cy/fy], [0, 0, 1]]
    const double Ki[9] = {1.0 / fx, 0, -cx / fx, 0, 1.0 / fy, -cy / fy, 0, 0, 1};
    double T[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[i * 3 + j] = Ki[0 * 3 + i] * E[0 * 3 + j] + Ki[1 * 3 + i] * E[1 * 3 + j] + Ki[2 * 3 + i] * E[2 * 3 + j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) F[i * 3 + j] = T[i * 3] * Ki[j] + T[i * 3 + 1] * Ki[3 + j] + T[i * 3 + 2] * Ki[6 + j];
    return B200VO_OK;
}

int b200vo_enable_timing(b200vo_ctx* ctx, int on) {
    if (!ctx) { set_error("null ctx"); return B200VO_ERR_BAD_ARG; }
    ctx->timing = on != 0;
    return B200VO_OK;
}

int b200vo_get_stage_times(b200vo_ctx* ctx, b200vo_stage_times* out) {
    if (!ctx || !out) { set_error("null argument"); return B200VO_ERR_BAD_ARG; }
    *out = ctx->times;
    return B200VO_OK;
}

int64_t b200vo_kernel_launches(const b200vo_ctx* ctx) { return ctx ? ctx->launches : -1; }
void* b200vo_stream(b200vo_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

}  // extern "C"
