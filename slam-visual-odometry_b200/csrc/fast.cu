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
#include "tma.cuh"

namespace b200vo {

#ifndef FAST_MINB
#define FAST_MINB 7
#endif

// Tile geometry: 60x30 outputs -> a 64x32 score tile (NMS halo; cell 0 sits on a 4-byte aligned pixel column, so the
// pre-test reads aligned words: the halo is 3 cells on the left, 1 on the right) -> a 70x38 pixel tile (+3 ring).
// The pixel tile is fetched as aligned 32-bit words (row of 20 words; `shift` = misalignment of the
// tile origin), so staging costs ~4 instructions per pixel instead of one byte load each.
// Two tile widths: the wide one (126 outputs) for the large levels -- its survivor / corner lists are
// four and two 256-thread passes long instead of 1.4 and 0.5, so phases 2-4 run on full warps -- the
// narrow one (62) where a wide tile would mostly overhang the level.
template <int CW>
struct FastTile {
    static constexpr int FT_CW = CW, FT_CH = 32;                 // score tile
    static constexpr int FT_W = CW - 4, FT_H = FT_CH - 2;        // outputs: cells 3 .. FT_W + 2 of a row, rows 1 .. FT_H
    static constexpr int FT_SW = (CW + 6 + 15 + 15) / 16 * 16;   // pixel tile row bytes: CW + 6 ring + <= 15 alignment shift, x16
    static constexpr int FT_SWW = FT_SW / 4;
    static constexpr int FT_SH = FT_CH + 6;                      // 38 rows
};
static_assert(FastTile<64>::FT_SW == kFastBoxW && FastTile<128>::FT_SW == kFastBoxWideW && FastTile<64>::FT_SH == kFastBoxH,
              "TMA boxes must match the pixel tiles");

__device__ __forceinline__ bool has_arc9(uint32_t m) {
    uint32_t r = m | (m << 16);
    const uint32_t m2 = r;
    r &= r >> 1;
    r &= r >> 2;
    r &= r >> 4;
    r &= m2 >> 8;
    return (r & 0xFFFFu) != 0;
}

template <int S>
__device__ __forceinline__ void load_ring(const uint8_t* p, int* v) {
    v[0] = p[3 * S];       v[1] = p[3 * S + 1];   v[2] = p[2 * S + 2];   v[3] = p[S + 3];
    v[4] = p[3];           v[5] = p[-S + 3];      v[6] = p[-2 * S + 2];  v[7] = p[-3 * S + 1];
    v[8] = p[-3 * S];      v[9] = p[-3 * S - 1];  v[10] = p[-2 * S - 2]; v[11] = p[-S - 3];
    v[12] = p[-3];         v[13] = p[S - 3];      v[14] = p[2 * S - 2];  v[15] = p[3 * S - 1];
}

// The work is re-packed between phases so that the expensive paths run on dense lane sets:
//   1. every pixel: compass pre-test (a 9-arc always covers two ADJACENT compass pixels 0,4,8,12),
//      four pixels per thread with byte-SIMD arithmetic -> survivors appended to a shared list
//   2. survivors: full 16-pixel segment test -> corners appended to a second list
//   3. corners: corner score (max over 9-arcs of the arc minimum, minus 1) -> score tile
//   4. corners: strict 3x3 non-max suppression + border filter -> candidate list + histogram
template <int CW>
__device__ __forceinline__ void fast_tile(const StoreView& sv, int level, int bx, int by, int frame, int first_slot, const CUtensorMap* tmap, int use_tma) {
    using T = FastTile<CW>;
    constexpr int FT_CW = T::FT_CW, FT_CH = T::FT_CH, FT_W = T::FT_W, FT_H = T::FT_H, FT_SW = T::FT_SW, FT_SWW = T::FT_SWW, FT_SH = T::FT_SH;
    __shared__ __align__(128) uint8_t s_img[FT_SH * FT_SW];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(4) uint8_t s_score[FT_CH * FT_CW];
    __shared__ __align__(4) uint16_t s_surv[FT_CH * FT_CW];
    __shared__ uint16_t s_corn[FT_CH * FT_CW];
    uint32_t* s_list = reinterpret_cast<uint32_t*>(s_surv);      // phase 4's output list (<= cells / 4 entries) reuses the survivor list, dead after phase 2
    __shared__ int s_nsurv, s_ncorn, s_n, s_base;
    __shared__ int s_hist[256];                     // score histogram of this tile, flushed once (one global atomic per non-empty bin)

    const LevelGeom g = sv.lv[level];
    const int slot = ring_slot(first_slot, frame, sv.slots);
    const uint8_t* __restrict__ img = sv.pyr + (size_t)slot * sv.slot_bytes + g.offset;
    const int ox = kEdge + bx * FT_W;       // first output pixel of the tile
    const int oy = kEdge + by * FT_H;
    const int tid = threadIdx.x, lane = tid & 31;
    // score-tile cell (cx, cy) <-> pixel (ox - 3 + cx, oy - 1 + cy): ox - 3 = 28 + bx * FT_W is a multiple of 4
    const int px0 = ox - 6, py0 = oy - 4;           // pixel-tile origin (>= 25: always inside the image)
    // TMA needs a 16-byte aligned inner coordinate, the word-load path a 4-byte aligned one
    const int x0 = use_tma ? (px0 & ~15) : (px0 & ~3), shift = px0 - x0;
    if (tid == 0) { s_nsurv = 0; s_ncorn = 0; s_n = 0; }
    s_hist[tid] = 0;

    if (use_tma) {
        // one thread issues the whole 96x38-byte box; everybody waits on the mbarrier
        if (tid == 0) { tma::mbar_init(&s_bar, 1); tma::fence_mbar_init(); }
        __syncthreads();
        if (tid == 0) {
            tma::mbar_arrive_expect_tx(&s_bar, FT_SH * FT_SW);
            tma::load_3d(s_img, tmap, &s_bar, x0, py0, slot);
        }
        uint32_t* sc32 = reinterpret_cast<uint32_t*>(s_score);
        for (int i = tid; i < FT_CH * FT_CW / 4; i += 256) sc32[i] = 0;
        tma::mbar_wait(&s_bar, 0);
    } else {
        uint32_t* s32 = reinterpret_cast<uint32_t*>(s_img);
        const int wmax = g.pitch / 4 - 1;
        for (int i = tid; i < FT_SH * FT_SWW; i += 256) {
            const int ty = i / FT_SWW, wx = i - ty * FT_SWW;
            const int gy = min(py0 + ty, g.h - 1), gw = min(x0 / 4 + wx, wmax);
            s32[i] = __ldg(reinterpret_cast<const uint32_t*>(img + (size_t)gy * g.pitch) + gw);
        }
        uint32_t* sc32 = reinterpret_cast<uint32_t*>(s_score);
        for (int i = tid; i < FT_CH * FT_CW / 4; i += 256) sc32[i] = 0;
    }
    __syncthreads();

    const int t = sv.fast_threshold;
    // score-tile cell (cx, cy) <-> pixel-tile byte (cx + 3 + shift, cy + 3)
    const uint8_t* s_org = s_img + 3 * FT_SW + 3 + shift;

    // ---- phase 1: compass pre-test on every cell of the score tile, four cells per thread (bytes of a word)
    // A 9-arc always covers two ADJACENT compass pixels (ring 0, 4, 8, 12), so |r_k - c| > t must
    // hold for two adjacent k:  (m0&m4)|(m4&m8)|(m8&m12)|(m12&m0) = (m0|m8)&(m4|m12).  The test ignores
    // the polarity (18 % of the pixels of a textured frame pass instead of 15 % -- phase 2 decides), which
    // lets it run on VABSDIFF4 plus a carry-free byte compare: ~15 instructions per pixel instead of ~45.
    // A thread keeps the pass flags of all its words (16 cells) in one register and appends its survivors
    // after ONE warp scan and ONE shared atomic per warp (round 1: four ballots, four popcounts and an atomic
    // per word -- a fifth of the kernel's instructions).
    {
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(s_img);
        const uint32_t tt = (uint32_t)t;
        const bool t_hi = tt >= 128u;                               // x > t needs bit 7 of x when t >= 128
        const uint32_t kadd = (t_hi ? 255u - tt : 127u - tt) * 0x01010101u;
        const int obyte = 3 + shift;                                // byte column of score-tile cell 0 in a pixel-tile row
        constexpr int ITERS = (FT_CH * FT_CW / 4) / 256;            // words per thread: 4 (wide tile) or 2
        constexpr int ROWS_PER_IT = 256 / (FT_CW / 4);              // tile rows covered by one pass of the CTA
        const int j = tid % (FT_CW / 4), cy0 = tid / (FT_CW / 4);   // this thread's word column and its row in the first pass
        uint32_t flags = 0;                                         // bit 4*it + b: cell (cy0 + it*ROWS_PER_IT, 4*j + b) passed
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const int cy = cy0 + it * ROWS_PER_IT;
            const int bc = obyte + 4 * j;                            // first of the four centres
            const uint32_t* rc = s32 + (cy + 3) * FT_SWW + (bc >> 2);   // bc is a multiple of 4: the four centres are one aligned word
            const uint32_t c = rc[0];
            const uint32_t r4 = __funnelshift_r(c, rc[1], 24), r12 = __funnelshift_r(rc[-1], c, 8);      // the same row, +3 / -3 bytes
            const uint32_t r0 = rc[3 * FT_SWW], r8 = rc[-3 * FT_SWW];
            auto gt_t = [&](uint32_t x) -> uint32_t {                // 0x80 in every byte with x > t
                const uint32_t sm = (x & 0x7F7F7F7Fu) + kadd;        // no carry between bytes
                return (t_hi ? (sm & x) : (sm | x)) & 0x80808080u;
            };
            const uint32_t m0 = gt_t(__vabsdiffu4(r0, c)), m4 = gt_t(__vabsdiffu4(r4, c));
            const uint32_t m8 = gt_t(__vabsdiffu4(r8, c)), m12 = gt_t(__vabsdiffu4(r12, c));
            const uint32_t pass4 = (m0 | m8) & (m4 | m12);           // 0x80 per passing byte
            // gather bits 7, 15, 23, 31 into a nibble: (x >> 7) has them at 0, 8, 16, 24; the multiply moves them to 24..27
            flags |= ((((pass4 >> 7) * 0x00204081u) >> 21) & 0xFu) << (4 * it);
        }
        // exclusive scan of the per-thread counts over the warp, one reservation per warp
        const int cnt = __popc(flags);
        int incl = cnt;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        int base = 0;
        if (lane == 31 && incl) base = atomicAdd(&s_nsurv, incl);
        base = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
        // One predicated store per cell instead of a loop over the set bits: the loop ran as long as the fullest lane of the
        // warp needed (~12 rounds of ~11 instructions with a third of the lanes active: 12 % of the kernel's instructions);
        // unrolled, every lane spends 4 instructions per cell whether it survives or not.
        const int c0 = cy0 * FT_CW + 4 * j;
#pragma unroll
        for (int b = 0; b < 4 * ITERS; ++b) {
            if (flags & (1u << b)) s_surv[base++] = (uint16_t)(c0 + (b >> 2) * ROWS_PER_IT * FT_CW + (b & 3));
        }
    }
    __syncthreads();

    // ---- phase 2: full segment test on the survivors
    const int nsurv = s_nsurv;
    for (int k0 = 0; k0 < nsurv; k0 += 256) {
        const int k = k0 + tid;
        bool corner = false, bright = false;
        int cell = 0;
        if (k < nsurv) {
            cell = s_surv[k];
            const int cy = cell / FT_CW, cx = cell - cy * FT_CW;
            const uint8_t* p = s_org + cy * FT_SW + cx;
            const int c = p[0], hi = c + t, lo = c - t;
            int v[16];
            load_ring<FT_SW>(p, v);
            uint32_t mb = 0, md = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                mb |= (uint32_t)(v[i] > hi) << i;
                md |= (uint32_t)(v[i] < lo) << i;
            }
            bright = has_arc9(mb);
            corner = bright || has_arc9(md);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, corner);
        if (m) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_ncorn, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (corner) s_corn[base + __popc(m & ((1u << lane) - 1))] = (uint16_t)(cell | (bright ? 0x8000 : 0));
        }
    }
    __syncthreads();

    // ---- phase 3: corner score
    const int ncorn = s_ncorn;
    for (int k = tid; k < ncorn; k += 256) {
        const int e = s_corn[k];
        const int cell = e & 0x7FFF;
        const bool bright = (e & 0x8000) != 0;
        const int cy = cell / FT_CW, cx = cell - cy * FT_CW;
        const uint8_t* p = s_org + cy * FT_SW + cx;
        const int c = p[0];
        int v[16];
        load_ring<FT_SW>(p, v);
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
        s_score[cell] = (uint8_t)(best - 1);        // best > t is implied by the mask test
    }
    __syncthreads();

    // ---- phase 4: strict 3x3 NMS on interior corners, border filter, append
    const int xmax = g.w - kEdge, ymax = g.h - kEdge;
    for (int k = tid; k < ncorn; k += 256) {
        const int cell = s_corn[k] & 0x7FFF;
        const int cy = cell / FT_CW, cx = cell - cy * FT_CW;
        if (cx < 3 || cx > FT_W + 2 || cy < 1 || cy > FT_H) continue;   // halo cells only feed their neighbours
        const int gx = ox + cx - 3, gy = oy + cy - 1;
        if (gx >= xmax || gy >= ymax) continue;
        const uint8_t* s = &s_score[cell];
        const int sc = s[0];
        const bool keep = sc > s[-1] && sc > s[1] && sc > s[-FT_CW - 1] && sc > s[-FT_CW] && sc > s[-FT_CW + 1] &&
                          sc > s[FT_CW - 1] && sc > s[FT_CW] && sc > s[FT_CW + 1];
        if (keep) {
            const int j = atomicAdd(&s_n, 1);
            s_list[j] = ((uint32_t)gy << 20) | ((uint32_t)gx << 8) | (uint32_t)sc;
            atomicAdd(&s_hist[sc], 1);
        }
    }
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;
    if (s_hist[tid]) atomicAdd(&sv.fast_hist[((size_t)slot * kLevels + level) * 256 + tid], s_hist[tid]);
    if (tid == 0) s_base = atomicAdd(&sv.cand_count[slot * kLevels + level], n);
    __syncthreads();
    const int base = s_base;
    uint32_t* out = sv.cand + (size_t)slot * sv.cand_total + sv.cand_off[level];
    for (int i = tid; i < n; i += 256)
        if (base + i < g.cand_cap) out[base + i] = s_list[i];
}

// one pyramid level per launch (large batches): the level is a kernel parameter
template <int CW>
__global__ void __launch_bounds__(256, FAST_MINB)
fast_kernel(StoreView sv, int level, int first_slot, const __grid_constant__ CUtensorMap tmap, int use_tma) {
    pdl_grid_sync();
    fast_tile<CW>(sv, level, blockIdx.x, blockIdx.y, blockIdx.z, first_slot, &tmap, use_tma);
}

// several levels per launch (small batches, see LevelTiles)
template <int CW>
__global__ void __launch_bounds__(256, 6)
fast_levels_kernel(StoreView sv, const __grid_constant__ LevelTiles lt, int first_slot, const __grid_constant__ LevelMaps maps, int use_tma) {
    pdl_grid_sync();
    const TileCoord tc = tile_coord(lt, blockIdx.x);
    fast_tile<CW>(sv, tc.level, tc.bx, tc.by, tc.frame, first_slot, &maps.m[tc.level], use_tma);
}

__global__ void clear_detect_state_kernel(StoreView sv, int first_slot, int n) {
    pdl_grid_sync();
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
        launch_k(ctx, clear_detect_state_kernel, dim3(div_up(total, 256)), dim3(256), 0, ctx->store, first_slot, n);
        B200VO_LAUNCH_CHECK(ctx);
    }
    const LevelMaps& maps = *reinterpret_cast<const LevelMaps*>(ctx->tmap_fast);
    const int tma = ctx->use_tma ? 1 : 0;
    LevelTiles wide, narrow;
    wide.first[0] = narrow.first[0] = 0;
    for (int l = 0; l < kLevels; ++l) {
        const LevelGeom& g = ctx->store.lv[l];
        const int iw = g.w - 2 * kEdge, ih = g.h - 2 * kEdge;
        const bool w = fast_level_is_wide(l), any = iw > 0 && ih > 0;
        level_tiles_add(wide, l, any ? div_up(iw, FastTile<128>::FT_W) : 0, (any && w) ? div_up(ih, FastTile<128>::FT_H) : 0, n);
        level_tiles_add(narrow, l, any ? div_up(iw, FastTile<64>::FT_W) : 0, (any && !w) ? div_up(ih, FastTile<64>::FT_H) : 0, n);
    }
    if (n <= ctx->small_batch_max) {
        // two launches: the wide-tile levels (0..4) and the narrow-tile levels (5..7)
        if (wide.first[kLevels] > 0) {
            launch_k(ctx, fast_levels_kernel<128>, dim3(wide.first[kLevels]), dim3(256), 0, ctx->store, wide, first_slot, maps, tma);
            B200VO_LAUNCH_CHECK(ctx);
        }
        if (narrow.first[kLevels] > 0) {
            launch_k(ctx, fast_levels_kernel<64>, dim3(narrow.first[kLevels]), dim3(256), 0, ctx->store, narrow, first_slot, maps, tma);
            B200VO_LAUNCH_CHECK(ctx);
        }
        return B200VO_OK;
    }
    for (int l = 0; l < kLevels; ++l) {
        const LevelGeom& g = ctx->store.lv[l];
        const int iw = g.w - 2 * kEdge, ih = g.h - 2 * kEdge;
        if (iw <= 0 || ih <= 0) continue;
        if (fast_level_is_wide(l)) {
            dim3 grid(div_up(iw, FastTile<128>::FT_W), div_up(ih, FastTile<128>::FT_H), n);
            launch_k(ctx, fast_kernel<128>, dim3(grid), dim3(256), 0, ctx->store, l, first_slot, ctx->tmap_fast[l], tma);
        } else {
            dim3 grid(div_up(iw, FastTile<64>::FT_W), div_up(ih, FastTile<64>::FT_H), n);
            launch_k(ctx, fast_kernel<64>, dim3(grid), dim3(256), 0, ctx->store, l, first_slot, ctx->tmap_fast[l], tma);
        }
        B200VO_LAUNCH_CHECK(ctx);
    }
    return B200VO_OK;
}

}  // namespace b200vo
