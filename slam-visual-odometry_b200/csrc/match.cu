// K9 brute-force Hamming matching with row AND column arg-min in one pass.
//
// What it replaces: cv::BFMatcher(NORM_HAMMING, crossCheck=true)::match
// (reference src/visual_odometry.cpp:87,153), which OpenCV evaluates as three batchDistance
// passes.  Semantics: distance = popcount over 256 bits; nearest train of
// every query and nearest query of every train, LOWEST INDEX wins ties.
//
// Design: this is integer work (XOR + POPC), bound by the POPC pipe, not by HBM and not a dense
// floating-point contraction -- no tensor cores.  Each thread keeps R query descriptors
// (R x 8 words) in registers; a CTA streams a chunk of train descriptors through shared memory
// (every lane reads the same train descriptor: a broadcast, conflict-free).  A distance is used
// twice: folded into the thread's running row minimum, and -- packed as (dist<<16 | row) --
// reduced over the warp with redux.sync.min and merged into a per-column minimum in shared
// memory.  Packing (dist<<16 | index) makes the integer minimum implement "smallest distance,
// then smallest index" under any order of reduction.  Partial results of the (row tile x column
// chunk) grid are merged with atomicMin on the packed words, so the whole distance matrix is
// computed exactly once for both directions.  grid.z batches independent pairs.
#include "common.cuh"
#include <cstdlib>

namespace b200vo {

constexpr int MT_THREADS = 256;
constexpr int MT_MAXCHUNK = 512;

struct MatchArgs {
    const uint8_t* q; const uint8_t* t;   // descriptor bases
    size_t q_stride, t_stride;            // bytes between consecutive pairs (raw mode)
    const int* counts;                    // slot mode: kp_count per slot (NULL in raw mode)
    int n1, n2;                           // raw mode sizes
    int slots, first_slot, slot_b_override, kp_cap;
    uint32_t* nnq; uint32_t* nnt;         // [pairs][out_stride]
    uint32_t* nnq2;                       // [pairs][out_stride] second neighbour (k-NN kernel only)
    int out_stride;
    int chunk;                            // train columns per CTA (<= MT_MAXCHUNK)
};

// 256-bit Hamming distance.  The POPC pipe (16 lanes/clk/SM) is the matcher's bottleneck, LOP3 runs
// at 64 lanes/clk/SM, so part of the population count is moved onto the logic pipe with carry-save
// adders: CSA(a,b,c) = (a^b^c, maj(a,b,c)) turns three words of weight 1 into one of weight 1 and
// one of weight 2 with two LOP3.  Two CSAs + two plain words: 6 POPC instead of 8 per distance.
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
template <int CSA>
__device__ __forceinline__ uint32_t hamming256(const uint32_t* q, const uint4& t0, const uint4& t1) {
    const uint32_t x0 = q[0] ^ t0.x, x1 = q[1] ^ t0.y, x2 = q[2] ^ t0.z, x3 = q[3] ^ t0.w;
    const uint32_t x4 = q[4] ^ t1.x, x5 = q[5] ^ t1.y, x6 = q[6] ^ t1.z, x7 = q[7] ^ t1.w;
    if (CSA == 2) {
    // full Harley-Seal tree: 4 POPC + 22 LOP3
    const uint32_t l1 = lop3_xor3(x0, x1, x2), h1 = lop3_maj(x0, x1, x2);
    const uint32_t l2 = lop3_xor3(x3, x4, x5), h2 = lop3_maj(x3, x4, x5);
    const uint32_t l3 = lop3_xor3(l1, l2, x6), h3 = lop3_maj(l1, l2, x6);
    const uint32_t ones = l3 ^ x7, h4 = l3 & x7;
    const uint32_t tA = lop3_xor3(h1, h2, h3), fA = lop3_maj(h1, h2, h3);
    const uint32_t twos = tA ^ h4, fB = tA & h4;
    const uint32_t fours = fA ^ fB, eights = fA & fB;
    return __popc(ones) + 2 * __popc(twos) + 4 * __popc(fours) + 8 * __popc(eights);
    } else if (CSA == 3) {
    // three carry-save adders: 5 POPC + 14 LOP3 -- POPC pipe (5/16 clk) and ALU pipe about balanced
    const uint32_t l1 = lop3_xor3(x0, x1, x2), h1 = lop3_maj(x0, x1, x2);
    const uint32_t l2 = lop3_xor3(x3, x4, x5), h2 = lop3_maj(x3, x4, x5);
    const uint32_t l3 = lop3_xor3(l1, l2, x6), h3 = lop3_maj(l1, l2, x6);
    return __popc(l3) + __popc(x7) + 2 * (__popc(h1) + __popc(h2) + __popc(h3));
    } else if (CSA == 1) {
    const uint32_t l1 = lop3_xor3(x0, x1, x2), h1 = lop3_maj(x0, x1, x2);
    const uint32_t l2 = lop3_xor3(x3, x4, x5), h2 = lop3_maj(x3, x4, x5);
    return __popc(l1) + __popc(l2) + __popc(x6) + __popc(x7) + 2 * (__popc(h1) + __popc(h2));
    } else {
    return __popc(x0) + __popc(x1) + __popc(x2) + __popc(x3) + __popc(x4) + __popc(x5) + __popc(x6) + __popc(x7);
    }
}

// query / train descriptor arrays and their sizes of pair `pair`: ring slots of the frame store, or
// back-to-back raw arrays
__device__ __forceinline__ void resolve_pair(const MatchArgs& a, int pair, const uint8_t*& q, const uint8_t*& t, int& n1, int& n2) {
    if (a.counts) {
        const int sa = (a.first_slot + pair) % a.slots;
        const int sb = (a.slot_b_override >= 0) ? a.slot_b_override : (a.first_slot + pair + 1) % a.slots;
        q = a.q + (size_t)sa * a.kp_cap * kDescBytes;
        t = a.t + (size_t)sb * a.kp_cap * kDescBytes;
        n1 = a.counts[sa]; n2 = a.counts[sb];
    } else {
        q = a.q + (size_t)pair * a.q_stride;
        t = a.t + (size_t)pair * a.t_stride;
        n1 = a.n1; n2 = a.n2;
    }
}

// V selects the inner-loop variant (B200VO_MATCH_V, A/B knob):
//   0  round-1 loop: after the warp reduction lane 0 merges every column minimum with a shared atomicMin
//   1  lane (j & 31) keeps the reduced minimum of column j in a register and the warp merges 32 columns with ONE shared
//      atomicMin per lane: the compiler expands a shared atomic at a warp-uniform address into a second CREDUX plus a
//      leader election (~10 instructions per column); packed words are formed with + instead of | so the shift-and-add
//      is one IMAD on the FMA pipe
//   2  as 1, and the last of the R rows takes the full Harley-Seal tree (4 POPC + 22 LOP3): 19 instead of 20 POPC per
//      column of four rows, which balances the POPC pipe (8 clk per warp instruction) against the ALU pipe (2 clk)
//   3  as 1 with __launch_bounds__(256, 4): 64 registers instead of 76, four CTAs (32 warps) per SM instead of three -- the
//      default (2.802 -> 2.774 ms per 250 pairs; unrolling the column loop by 4 instead of 2, and column chunks of 500
//      instead of 512 so that all CTAs are equal, measured no gain)
template <int R, int CSA, int V>
__global__ void __launch_bounds__(MT_THREADS, V == 3 ? 4 : 0)
match_kernel(MatchArgs a) {
    pdl_grid_sync();
    __shared__ __align__(16) uint32_t s_t[MT_MAXCHUNK * 8];
    __shared__ uint32_t s_col[MT_MAXCHUNK];

    const int pair = blockIdx.z;
    const uint8_t* q; const uint8_t* t;
    int n1, n2;
    resolve_pair(a, pair, q, t, n1, n2);
    const int row0 = blockIdx.y * (MT_THREADS * R);
    const int col0 = blockIdx.x * a.chunk;
    if (row0 >= n1 || col0 >= n2) return;
    const int ncols = min(a.chunk, n2 - col0);
    const int tid = threadIdx.x;

    // stage the train chunk (coalesced 16-B loads) and reset the column minima
    {
        const uint4* src = reinterpret_cast<const uint4*>(t + (size_t)col0 * kDescBytes);
        uint4* dst = reinterpret_cast<uint4*>(s_t);
        for (int i = tid; i < ncols * 2; i += MT_THREADS) dst[i] = __ldg(src + i);
        for (int i = tid; i < ncols; i += MT_THREADS) s_col[i] = 0xFFFFFFFFu;
    }

    uint32_t qd[R][8];
    uint32_t rowpack[R];       // row index, or a huge distance offset for padding rows
    uint32_t best[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = row0 + r * MT_THREADS + tid;
        best[r] = 0xFFFFFFFFu;
        if (row < n1) {
            const uint4* p = reinterpret_cast<const uint4*>(q + (size_t)row * kDescBytes);
            const uint4 lo = __ldg(p), hi = __ldg(p + 1);
            qd[r][0] = lo.x; qd[r][1] = lo.y; qd[r][2] = lo.z; qd[r][3] = lo.w;
            qd[r][4] = hi.x; qd[r][5] = hi.y; qd[r][6] = hi.z; qd[r][7] = hi.w;
            rowpack[r] = (uint32_t)row;
        } else {
#pragma unroll
            for (int w = 0; w < 8; ++w) qd[r][w] = 0;
            rowpack[r] = 0x80000000u;      // loses every minimum
        }
    }
    __syncthreads();

    const int lane = tid & 31;
    if (V == 0) {
#pragma unroll 2
        for (int j = 0; j < ncols; ++j) {
            const uint4 t0 = *reinterpret_cast<const uint4*>(&s_t[j * 8]);
            const uint4 t1 = *reinterpret_cast<const uint4*>(&s_t[j * 8 + 4]);
            uint32_t colmin = 0xFFFFFFFFu;
            const uint32_t colidx = (uint32_t)(col0 + j);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t d = hamming256<CSA>(qd[r], t0, t1);
                best[r] = min(best[r], (d << 16) | colidx);
                colmin = min(colmin, (d << 16) + rowpack[r]);
            }
            colmin = __reduce_min_sync(0xffffffffu, colmin);
            if (lane == 0) atomicMin(&s_col[j], colmin);
        }
    } else {
        uint32_t mycol = 0xFFFFFFFFu;          // reduced minimum of column (j & ~31) + lane
        for (int j0 = 0; j0 < ncols; j0 += 32) {
            const int nj = min(32, ncols - j0);
#pragma unroll 2
            for (int jj = 0; jj < nj; ++jj) {
                const int j = j0 + jj;
                const uint4 t0 = *reinterpret_cast<const uint4*>(&s_t[j * 8]);
                const uint4 t1 = *reinterpret_cast<const uint4*>(&s_t[j * 8 + 4]);
                uint32_t colmin = 0xFFFFFFFFu;
                const uint32_t colidx = (uint32_t)(col0 + j);
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t d = (V == 2 && R >= 4 && r == R - 1) ? hamming256<2>(qd[r], t0, t1) : hamming256<CSA>(qd[r], t0, t1);
                    best[r] = min(best[r], (d << 16) + colidx);
                    colmin = min(colmin, (d << 16) + rowpack[r]);
                }
                colmin = __reduce_min_sync(0xffffffffu, colmin);
                if (jj == lane) mycol = colmin;
            }
            if (lane < nj) atomicMin(&s_col[j0 + lane], mycol);
        }
    }

#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = row0 + r * MT_THREADS + tid;
        if (row < n1) atomicMin(&a.nnq[(size_t)pair * a.out_stride + row], best[r]);
    }
    __syncthreads();
    for (int i = tid; i < ncols; i += MT_THREADS) atomicMin(&a.nnt[(size_t)pair * a.out_stride + col0 + i], s_col[i]);
}

// K9b: the two nearest trains of every query -- cv::BFMatcher(NORM_HAMMING, false)::knnMatch(k = 2),
// the matcher behind a Lowe ratio test (north star; the reference itself only cross-checks,
// src/visual_odometry.cpp:87).  OpenCV orders neighbours by (distance, trainIdx) (pinned in
// tests/test_oracle_match.py), which is again the integer order of the packed words, so
//   second = min(second, max(best, v));  best = min(best, v)
// keeps the exact top two in three integer min/max.  A CTA owns its query rows for ALL train
// columns (it streams the chunks itself), so the top-2 never has to be merged across CTAs, and the
// column minima of the cross-check kernel (redux + shared atomics) are not needed at all.
template <int R>
__global__ void __launch_bounds__(MT_THREADS)
match_knn2_kernel(MatchArgs a) {
    pdl_grid_sync();
    __shared__ __align__(16) uint32_t s_t[MT_MAXCHUNK * 8];
    const int pair = blockIdx.z;
    const uint8_t* q; const uint8_t* t;
    int n1, n2;
    resolve_pair(a, pair, q, t, n1, n2);
    const int row0 = blockIdx.y * (MT_THREADS * R);
    if (row0 >= n1) return;
    const int tid = threadIdx.x;
    uint32_t qd[R][8], best[R], second[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = row0 + r * MT_THREADS + tid;
        best[r] = second[r] = 0xFFFFFFFFu;
#pragma unroll
        for (int w = 0; w < 8; ++w) qd[r][w] = 0;
        if (row < n1) {
            const uint4* p = reinterpret_cast<const uint4*>(q + (size_t)row * kDescBytes);
            const uint4 lo = __ldg(p), hi = __ldg(p + 1);
            qd[r][0] = lo.x; qd[r][1] = lo.y; qd[r][2] = lo.z; qd[r][3] = lo.w;
            qd[r][4] = hi.x; qd[r][5] = hi.y; qd[r][6] = hi.z; qd[r][7] = hi.w;
        }
    }
    for (int col0 = 0; col0 < n2; col0 += MT_MAXCHUNK) {
        const int ncols = min(MT_MAXCHUNK, n2 - col0);
        __syncthreads();                       // the previous chunk has been consumed
        {
            const uint4* src = reinterpret_cast<const uint4*>(t + (size_t)col0 * kDescBytes);
            uint4* dst = reinterpret_cast<uint4*>(s_t);
            for (int i = tid; i < ncols * 2; i += MT_THREADS) dst[i] = __ldg(src + i);
        }
        __syncthreads();
#pragma unroll 2
        for (int j = 0; j < ncols; ++j) {
            const uint4 t0 = *reinterpret_cast<const uint4*>(&s_t[j * 8]);
            const uint4 t1 = *reinterpret_cast<const uint4*>(&s_t[j * 8 + 4]);
            const uint32_t colidx = (uint32_t)(col0 + j);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const uint32_t v = (hamming256<3>(qd[r], t0, t1) << 16) | colidx;
                second[r] = min(second[r], max(best[r], v));
                best[r] = min(best[r], v);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = row0 + r * MT_THREADS + tid;
        if (row < n1) {
            a.nnq[(size_t)pair * a.out_stride + row] = best[r];
            a.nnq2[(size_t)pair * a.out_stride + row] = second[r];
        }
    }
}

static int run_match_knn2(b200vo_ctx* ctx, MatchArgs a, int n1max, int n1_expected, int npairs) {
    if (npairs <= 0 || n1max <= 0) return B200VO_OK;
    // rows per thread: as many as still leave >= 2 CTAs per SM (the CTA count is rows x pairs only)
    int R = 4;
    while (R > 1 && (long)div_up(n1_expected, MT_THREADS * R) * npairs < 2 * 148) R >>= 1;
    dim3 grid(1, div_up(n1max, MT_THREADS * R), npairs);
    switch (R) {
        case 4: launch_k(ctx, match_knn2_kernel<4>, dim3(grid), dim3(MT_THREADS), 0, a); break;
        case 2: launch_k(ctx, match_knn2_kernel<2>, dim3(grid), dim3(MT_THREADS), 0, a); break;
        default: launch_k(ctx, match_knn2_kernel<1>, dim3(grid), dim3(MT_THREADS), 0, a); break;
    }
    B200VO_LAUNCH_CHECK(ctx);
    return B200VO_OK;
}

static int pick_rows_per_thread(int n1) {
    // least padding wins; prefer more rows per thread (fewer shared-memory reads per distance)
    int bestR = 1;
    long bestpad = -1;
    const int cand[3] = {4, 2, 1};
    for (int k = 0; k < 3; ++k) {
        const int tile = MT_THREADS * cand[k];
        const long padded = (long)div_up(n1, tile) * tile;
        if (bestpad < 0 || padded < bestpad) { bestpad = padded; bestR = cand[k]; }
    }
    return bestR;
}

static int run_match(b200vo_ctx* ctx, MatchArgs a, int n1max, int n2max, int n1_expected, int npairs) {
    if (npairs <= 0 || n1max <= 0 || n2max <= 0) return B200VO_OK;
    int R = pick_rows_per_thread(n1_expected);
    if (const char* v = getenv("B200VO_MATCH_R")) { const int f = atoi(v); if (f == 1 || f == 2 || f == 3 || f == 4 || f == 6 || f == 8) R = f; }
    const int row_tiles = div_up(n1max, MT_THREADS * R);
    // enough CTAs to fill 148 SMs a few times over when the batch is small
    int chunk = MT_MAXCHUNK;
    while (chunk > 64 && (long)row_tiles * div_up(n2max, chunk) * npairs < 4 * 148) chunk >>= 1;
    a.chunk = chunk;
    dim3 grid(div_up(n2max, chunk), row_tiles, npairs);
    const int csa = ctx->match_csa;
    const int variant = ctx->match_variant;
#define LAUNCH_MATCH_V(RR, VV) \
    do { if (csa == 3) launch_k(ctx, match_kernel<RR, 3, VV>, dim3(grid), dim3(MT_THREADS), 0, a); \
         else if (csa == 2) launch_k(ctx, match_kernel<RR, 2, VV>, dim3(grid), dim3(MT_THREADS), 0, a); \
         else if (csa == 1) launch_k(ctx, match_kernel<RR, 1, VV>, dim3(grid), dim3(MT_THREADS), 0, a); \
         else launch_k(ctx, match_kernel<RR, 0, VV>, dim3(grid), dim3(MT_THREADS), 0, a); } while (0)
#define LAUNCH_MATCH(RR) \
    do { if (variant == 2) LAUNCH_MATCH_V(RR, 2); else if (variant == 1) LAUNCH_MATCH_V(RR, 1); else if (variant == 3) LAUNCH_MATCH_V(RR, 3); else LAUNCH_MATCH_V(RR, 0); } while (0)
    switch (R) {
        case 8: LAUNCH_MATCH(8); break;
        case 6: LAUNCH_MATCH(6); break;
        case 4: LAUNCH_MATCH(4); break;
        case 3: LAUNCH_MATCH(3); break;
        case 2: LAUNCH_MATCH(2); break;
        default: LAUNCH_MATCH(1); break;
    }
#undef LAUNCH_MATCH_V
#undef LAUNCH_MATCH
    B200VO_LAUNCH_CHECK(ctx);
    return B200VO_OK;
}

int launch_match_slots(b200vo_ctx* ctx, int first_slot, int npairs, int slot_b_override) {
    const StoreView& sv = ctx->store;
    if (ctx->opt.match_mode == B200VO_MATCH_RATIO) {
        MatchArgs a{};
        a.q = sv.desc; a.t = sv.desc; a.counts = sv.kp_count;
        a.slots = sv.slots; a.first_slot = first_slot; a.slot_b_override = slot_b_override; a.kp_cap = sv.kp_cap;
        a.nnq = ctx->pairs.nnq; a.nnq2 = ctx->pairs.nnq2; a.out_stride = sv.kp_cap;
        return run_match_knn2(ctx, a, sv.kp_cap, ctx->opt.nfeatures, npairs);
    }
    { int rc; if ((rc = device_fill32(ctx, ctx->pairs.nnq, 0xFFFFFFFFu, (size_t)npairs * sv.kp_cap))) return rc;
      if ((rc = device_fill32(ctx, ctx->pairs.nnt, 0xFFFFFFFFu, (size_t)npairs * sv.kp_cap))) return rc; }
    MatchArgs a{};
    a.q = sv.desc; a.t = sv.desc; a.counts = sv.kp_count;
    a.slots = sv.slots; a.first_slot = first_slot; a.slot_b_override = slot_b_override; a.kp_cap = sv.kp_cap;
    a.nnq = ctx->pairs.nnq; a.nnt = ctx->pairs.nnt; a.out_stride = sv.kp_cap;
    // the grid is sized for the capacity; CTAs beyond the actual counts exit immediately
    return run_match(ctx, a, sv.kp_cap, sv.kp_cap, ctx->opt.nfeatures, npairs);
}

int launch_match_raw(b200vo_ctx* ctx, const uint8_t* q, int n1, const uint8_t* t, int n2, int npairs,
                     uint32_t* nnq, uint32_t* nnt) {
    const int stride = max(n1, n2);
    { int rc; if ((rc = device_fill32(ctx, nnq, 0xFFFFFFFFu, (size_t)npairs * stride))) return rc;
      if ((rc = device_fill32(ctx, nnt, 0xFFFFFFFFu, (size_t)npairs * stride))) return rc; }
    MatchArgs a{};
    a.q = q; a.t = t; a.q_stride = (size_t)n1 * kDescBytes; a.t_stride = (size_t)n2 * kDescBytes;
    a.counts = nullptr; a.n1 = n1; a.n2 = n2;
    a.nnq = nnq; a.nnt = nnt; a.out_stride = stride;
    return run_match(ctx, a, n1, n2, n1, npairs);
}

int launch_match_knn2_raw(b200vo_ctx* ctx, const uint8_t* q, int n1, const uint8_t* t, int n2, int npairs,
                          uint32_t* nn1, uint32_t* nn2) {
    MatchArgs a{};
    a.q = q; a.t = t; a.q_stride = (size_t)n1 * kDescBytes; a.t_stride = (size_t)n2 * kDescBytes;
    a.counts = nullptr; a.n1 = n1; a.n2 = n2;
    a.nnq = nn1; a.nnq2 = nn2; a.out_stride = n1;
    return run_match_knn2(ctx, a, n1, n1, npairs);
}

// ------------------------------------------------------------------ POPC-pipe microbenchmark
// Same instruction mix as the matcher's inner loop (LOP3 xor, POPC, IADD), 8 independent chains
// per thread, enough CTAs to fill every SM; reports sustained popc32/s.
__global__ void __launch_bounds__(256)
popc_rate_kernel(uint32_t* out, int iters) {
    pdl_grid_sync();
    uint32_t a[8], acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { a[k] = threadIdx.x * 2654435761u + k * 40503u + blockIdx.x; acc[k] = 0; }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += __popc(a[k] ^ (uint32_t)i);
    }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k];
    if (s == 0xFFFFFFFFu) out[0] = s;       // never true; keeps the loop alive
}

int measure_popc_rate(b200vo_ctx* ctx, double* popc_per_second) {
    uint32_t* d = nullptr;
    B200VO_CUDA_CHECK(cudaMalloc(&d, 256));
    const int blocks = 148 * 8, iters = 1 << 15;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        B200VO_CUDA_CHECK(cudaEventRecord(ctx->ev[14], ctx->stream));
        launch_k(ctx, popc_rate_kernel, dim3(blocks), dim3(256), 0, d, iters);
        B200VO_LAUNCH_CHECK(ctx);
        B200VO_CUDA_CHECK(cudaEventRecord(ctx->ev[15], ctx->stream));
        B200VO_CUDA_CHECK(cudaEventSynchronize(ctx->ev[15]));
        float ms = 0.f;
        B200VO_CUDA_CHECK(cudaEventElapsedTime(&ms, ctx->ev[14], ctx->ev[15]));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaFree(d);
    *popc_per_second = (double)blocks * 256.0 * iters * 8.0 / (best * 1e-3);
    return B200VO_OK;
}

}  // namespace b200vo
