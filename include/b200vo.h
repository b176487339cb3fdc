/*
 * b200vo.h -- C ABI of the B200-native visual-odometry front end.
 *
 * This is the drop-in boundary for the per-frame-pair hot path of evya1/slam-visual-odometry:
 * everything VisualOdometry::process_frame obtains from OpenCV
 * (reference src/visual_odometry.cpp:323-378) is computed by hand-written sm_100a kernels behind
 * these entry points.  Plain pointers and sizes only; the caller owns every host buffer and
 * passes capacities; the library owns all device memory inside the context.  Every function
 * returns 0 on success or a negative b200vo_status; nothing throws across this boundary.
 * Reference-style soft failures (too few matches, no essential matrix, failed cheirality) are
 * status BITS in the pose record, not errors.  There is no CPU fallback: if no CUDA device or
 * kernel image is available b200vo_create fails with B200VO_ERR_CUDA.
 *
 * A context is NOT thread-safe: one context per GPU, one host thread per context; all work is
 * ordered on the context's stream.
 *
 * Entry point                       replaces (reference file:line)
 * --------------------------------  ------------------------------------------------------------
 * b200vo_default_options/create     VisualOdometry ctor: ORB::create, BFMatcher::create, K
 *                                     src/visual_odometry.cpp:64-101
 * b200vo_upload_frames              Frame(id, image) + cvtColor(BGR2GRAY)
 *                                     src/frame.h:32-34, src/visual_odometry.cpp:105-111
 * b200vo_detect                     orb_detector_->detectAndCompute   src/visual_odometry.cpp:113
 * b200vo_get_features               frame.keypoints / frame.descriptors   src/frame.h:22-23
 * b200vo_match_descriptors          matcher_->match (cross-check NN)  src/visual_odometry.cpp:153
 * b200vo_knn_match_descriptors      [ext] BFMatcher::knnMatch(k=2) for the ratio-test mode (same call site)
 * b200vo_match                      get_good_matches_of_features      src/visual_odometry.cpp:145-180
 * b200vo_relative_pose              estimate_relative_pose: findEssentialMat + recoverPose
 *                                     src/visual_odometry.cpp:182-281
 * b200vo_process_frame              one whole streaming step of process_frame    src/visual_odometry.cpp:323-378
 * b200vo_process_frame_begin/_end   the same step in two halves (host work overlaps) src/visual_odometry.cpp:103-117,323-378
 * b200vo_process_pairs              the batched detect-free remainder of process_frame for a
 *                                     contiguous range of frame pairs  src/visual_odometry.cpp:340-366
 * b200vo_chain_pose                 invert_relative_... + compose_next_...  src/visual_odometry.cpp:292-314,350-360
 * b200vo_fundamental_from_essential F = K^-T E K^-1                   src/visual_odometry.cpp:225-246
 * b200vo_group_process_sequence     [ext] the frame loop of main() over several GPUs   src/main.cpp:128-194
 */
#ifndef B200VO_H
#define B200VO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VO_ABI_VERSION 3
#define B200VO_NLEVELS 8            /* reference src/visual_odometry.cpp:67 */

typedef enum {
    B200VO_OK = 0,
    B200VO_ERR_BAD_ARG = -1,
    B200VO_ERR_CAPACITY = -2,       /* a caller buffer or an internal fixed-capacity list is too small */
    B200VO_ERR_CUDA = -3,
    B200VO_ERR_NOT_READY = -4,      /* reserved; a slot without features is a soft failure (0 matches, status 0), as in the reference */
    B200VO_ERR_ALLOC = -5
} b200vo_status;

/* pose-record status bits (soft failures of the reference, src/visual_odometry.cpp:340-366) */
#define B200VO_ST_MATCHED 1         /* >= 10 good matches: pose estimation attempted (:344-345) */
#define B200VO_ST_E_FOUND 2         /* >= 8 matches and a model was found: R,t valid (:189,:218) */
#define B200VO_ST_POSE_OK 4         /* valid_points >= 10 and inliers >= 12 -> scale 0.3 (:270-277,:352) */
#define B200VO_ST_OVERFLOW 256      /* an internal fixed-capacity list of one of the two frames overflowed (only possible
                                     * through score ties); its feature set was truncated deterministically, so the
                                     * pair's results are reproducible but not the reference's */
#define B200VO_ST_ROUNDS_SHIFT 16    /* adaptive mode: bits 16..23 = number of 128-hypothesis blocks evaluated */

#define B200VO_MATCH_CROSSCHECK 0   /* reference behaviour: cross-check + adaptive threshold */
#define B200VO_MATCH_RATIO 1        /* [ext] Lowe ratio test on the 2 nearest trains (no cross-check) replaces the
                                     * matcher_->match call (:153); the sort + adaptive threshold (:155-177) still
                                     * follow.  n_cross then counts the matches that passed the ratio test. */

typedef struct b200vo_ctx b200vo_ctx;

typedef struct b200vo_options {
    int32_t struct_size;            /* sizeof(b200vo_options), set by b200vo_default_options */
    int32_t width, height;          /* level-0 image size */
    int32_t nfeatures;              /* ORB nfeatures; reference default 1200 (:65); 1..30841 */
    int32_t fast_threshold;         /* 10 (:73) */
    int32_t max_frames;             /* slots in the device-resident frame store (ring) */
    int32_t ransac_hypotheses;      /* five-point hypotheses per pair; default 1024 (>= OpenCV maxIters 1000) */
    int32_t match_mode;             /* B200VO_MATCH_* */
    float ratio;                    /* ratio-test threshold for B200VO_MATCH_RATIO (0.75) */
    int32_t ransac_adaptive;        /* 0 (default): evaluate all ransac_hypotheses; 1: OpenCV's adaptive stopping rule
                                     * (confidence 0.999, reference :215), checked after every block of 128 hypotheses */
    uint64_t ransac_seed;           /* hypothesis samples are f(seed, pair id, hypothesis index) */
    double ransac_threshold_px;     /* 2.0 (:215) */
    double max_hamming;             /* 35.0 (:147) */
    double fx, fy, cx, cy;          /* intrinsics; default fx=fy=width, cx=w/2, cy=h/2 (:90-98) */
    int32_t min_matches_for_pose;   /* 10 (:344) */
    int32_t min_valid_points;       /* 10 (:270) */
    int32_t min_inliers;            /* 12 (:271) */
    int32_t reserved1;
} b200vo_options;

/* Field-compatible with cv::KeyPoint's data members (pt.x, pt.y, size, angle, response, octave,
 * class_id) followed by the integer level coordinates the kernels work in. */
typedef struct b200vo_keypoint {
    float x, y;                     /* level-0 pixel coordinates: (lx,ly) * scale_octave */
    float size;                     /* 31 * scale_octave */
    float angle;                    /* degrees, [0,360) */
    float response;                 /* Harris response */
    int32_t octave;
    int32_t class_id;               /* -1 */
    int16_t lx, ly;                 /* integer coordinates in pyramid level `octave` */
} b200vo_keypoint;

/* Same layout as cv::DMatch. */
typedef struct b200vo_dmatch {
    int32_t queryIdx, trainIdx, imgIdx;
    float distance;
} b200vo_dmatch;

/* 160-byte relative-pose record: the wire format gathered across GPUs.  (R,t) is OpenCV's
 * recoverPose convention: x_c2 = R x_c1 + t, |t| = 1.  The tail carries what the reference prints per
 * pair: the [MatchDebug] statistics of the unfiltered match list (src/visual_odometry.cpp:119-143,
 * 160-168) and the mean algebraic epipolar residual over the RANSAC inliers (:248-263). */
typedef struct b200vo_pose_record {
    double R[9];                    /* row-major */
    double t[3];
    int32_t n_matches;              /* good matches after the adaptive threshold */
    int32_t n_inliers;              /* RANSAC inliers of the selected model (before recoverPose) */
    int32_t n_valid;                /* recoverPose cheirality count */
    int32_t status;                 /* B200VO_ST_* bits */
    int32_t n_cross;                /* cross-checked matches before the threshold */
    int32_t best_hyp;               /* index of the winning hypothesis, -1 if none */
    int32_t best_sol;               /* which of its solutions */
    float threshold;                /* adaptive Hamming threshold used */
    double dist_mean;               /* mean distance of the n_cross matches (get_mean_dist_ham, :130-134); 0 if none */
    double epi_residual;            /* mean |x2^T F x1| over the inliers of findEssentialMat, pixel coordinates, F = K^-T E K^-1
                                     * with E as returned here (|E|_F = 1; the residual is algebraic, so it scales with E); 0 when no E (:248-263) */
    float dist_min, dist_max;       /* matches.front() / matches.back() after the sort (:160-161) */
    float dist_median;              /* element size/2 of the sorted list (get_median_dist, :136-143) */
    int32_t reserved;
} b200vo_pose_record;

typedef struct b200vo_stage_times {
    float upload_ms, gray_ms, pyramid_ms, fast_ms, select_ms, orient_ms, blur_ms, brief_ms;
    float match_ms, filter_ms, ransac_ms, pose_ms, download_ms;
    /* kernels inside ransac_ms (fixed-H mode: one launch each): five-point solver, Sampson scoring */
    float fivept_ms, score_ms;
    /* fused orientation + keypoint record + rBRIEF kernel (then orient_ms = brief_ms = 0) */
    float describe_ms;
} b200vo_stage_times;

int b200vo_abi_version(void);
/* message for the most recent failure on this thread (also valid when create failed) */
const char* b200vo_last_error(void);

int b200vo_default_options(b200vo_options* opt, int width, int height);
int b200vo_create(const b200vo_options* opt, int device, b200vo_ctx** out);
void b200vo_destroy(b200vo_ctx* ctx);
int b200vo_synchronize(b200vo_ctx* ctx);

/* geometry of the pyramid and capacities, for sizing caller buffers */
int b200vo_level_info(const b200vo_ctx* ctx, int level, int* width, int* height, float* scale, int* nfeatures);
int b200vo_max_keypoints(const b200vo_ctx* ctx);          /* per-frame capacity (nfeatures + slack for ties) */

/* Copy n frames from host memory into slots first_slot.. (mod max_frames).  channels 1 (gray) or
 * 3 (BGR, converted on the device).  pitch = bytes between rows; frame_stride = bytes between
 * frames.  Asynchronous on the context stream when `host` is pinned memory. */
int b200vo_upload_frames(b200vo_ctx* ctx, const uint8_t* host, int n, size_t pitch, size_t frame_stride,
                         int channels, int first_slot);
/* Same, from device memory (used when the frames are already resident in HBM). */
int b200vo_upload_frames_device(b200vo_ctx* ctx, const void* dev, int n, size_t pitch, size_t frame_stride,
                                int channels, int first_slot);

/* Pyramid + FAST + selection + Harris + orientation + blur + rBRIEF for n consecutive slots. */
int b200vo_detect(b200vo_ctx* ctx, int first_slot, int n);

/* Features of one slot in canonical (octave, y, x) order.  desc may be NULL. */
int b200vo_get_features(b200vo_ctx* ctx, int slot, b200vo_keypoint* kp, uint8_t* desc, int cap, int* n);
/* Debug / parity: copy one pyramid level (blurred != 0: the 7x7-blurred copy) to host. */
int b200vo_get_pyramid_level(b200vo_ctx* ctx, int slot, int level, int blurred, uint8_t* out, size_t out_pitch);
/* Debug / parity: FAST corners after NMS + border filter of one level, packed (y<<20 | x<<8 | score), unordered. */
int b200vo_get_fast_candidates(b200vo_ctx* ctx, int slot, int level, uint32_t* out, int cap, int* n);
/* Replace the features of a slot (parity tests feed identical descriptors to both sides). */
int b200vo_set_features(b200vo_ctx* ctx, int slot, const b200vo_keypoint* kp, const uint8_t* desc, int n);

/* Raw brute-force Hamming NN on caller descriptors (n x 32 bytes each, host memory), batched
 * over npairs independent problems laid out back to back.  nn_q[i] = index of the nearest train
 * of query i, dist_q its distance; nn_t/dist_t the same for trains over queries.  Lowest index
 * wins ties.  Any output pointer may be NULL. */
int b200vo_match_descriptors(b200vo_ctx* ctx, const uint8_t* q, int n1, const uint8_t* t, int n2, int npairs,
                             int32_t* nn_q, int32_t* dist_q, int32_t* nn_t, int32_t* dist_t);
/* Device-resident variant used by the matcher stress benchmark; returns elapsed kernel ms. */
int b200vo_match_descriptors_device(b200vo_ctx* ctx, const void* q_dev, int n1, const void* t_dev, int n2, int npairs,
                                    void* nnq_packed_dev, void* nnt_packed_dev, float* kernel_ms);

/* [ext] cv::BFMatcher(NORM_HAMMING, false)::knnMatch(k = 2) on caller descriptors: the two nearest trains
 * of every query ordered by (distance, trainIdx), as OpenCV orders them.  idx/dist: [npairs][n1][2],
 * -1 where a neighbour does not exist (n2 < 2).  This is the matcher behind B200VO_MATCH_RATIO; the
 * reference itself only cross-checks (src/visual_odometry.cpp:87). */
int b200vo_knn_match_descriptors(b200vo_ctx* ctx, const uint8_t* q, int n1, const uint8_t* t, int n2, int npairs,
                                 int32_t* idx, int32_t* dist);

/* get_good_matches_of_features between two slots: cross-check (or ratio test), sort by
 * (distance, queryIdx), adaptive threshold.  out may be NULL. */
int b200vo_match(b200vo_ctx* ctx, int slot1, int slot2, b200vo_dmatch* out, int cap, int* n_good, int* n_cross,
                 double* threshold);

/* estimate_relative_pose for one pair.  samples: nhyp x 5 indices into the good-match list
 * (NULL: drawn on the device from (seed, pair_id)).  essential (9 doubles, may be NULL) gets E. */
int b200vo_relative_pose(b200vo_ctx* ctx, int slot1, int slot2, int64_t pair_id, const int32_t* samples, int nhyp,
                         b200vo_pose_record* rec, double* essential, uint8_t* inlier_mask, int mask_cap);

/* One streaming step, i.e. what VisualOdometry::process_frame does per frame (src/visual_odometry.cpp:323-378):
 * upload + detect the frame into `slot`, run the pair (prev_slot, slot) when prev_slot >= 0, and return the
 * frame's features, the pose record and E -- with ONE host synchronisation (the separate calls above
 * synchronise four times).  kp / desc / essential may be NULL; rec is required when prev_slot >= 0. */
int b200vo_process_frame(b200vo_ctx* ctx, const uint8_t* host, size_t pitch, int channels, int slot, int prev_slot, int64_t pair_id,
                         b200vo_keypoint* kp, uint8_t* desc, int cap, int* n, b200vo_pose_record* rec, double* essential);
/* The same step in two halves, so the caller's per-frame host work on the features (the reference fills cv::KeyPoint and
 * the descriptor Mat, src/visual_odometry.cpp:103-117) overlaps the pair's kernels: _begin enqueues the whole step and
 * returns as soon as the frame's features are on the host; _end waits for the pair and hands out the record and E
 * (rec is required when the step had prev_slot >= 0).  One step at a time per context. [ext] */
int b200vo_process_frame_begin(b200vo_ctx* ctx, const uint8_t* host, size_t pitch, int channels, int slot, int prev_slot, int64_t pair_id,
                               b200vo_keypoint* kp, uint8_t* desc, int cap, int* n);
int b200vo_process_frame_end(b200vo_ctx* ctx, b200vo_pose_record* rec, double* essential);

/* Batched: pairs (first_slot+i, first_slot+i+1), i < npairs, with pair ids first_pair_id+i.
 * Features of all touched slots must exist (b200vo_detect).  Records are copied to `out`. */
int b200vo_process_pairs(b200vo_ctx* ctx, int first_slot, int npairs, int64_t first_pair_id, b200vo_pose_record* out);
/* Same but leaves the records in device memory (for the NCCL gather); returns the device pointer. */
int b200vo_process_pairs_device(b200vo_ctx* ctx, int first_slot, int npairs, int64_t first_pair_id, void** records_dev);

/* Parity hooks for the geometry kernels (normalised double coordinates, host memory). */
int b200vo_five_point(b200vo_ctx* ctx, const double* x1, const double* x2, int npts, const int32_t* samples, int nhyp,
                      double* essentials /* nhyp*10*9 */, int32_t* nsol /* nhyp */);
int b200vo_score_models(b200vo_ctx* ctx, const double* x1, const double* x2, int npts, const double* essentials,
                        int nmodels, double threshold, int32_t* counts);
int b200vo_recover_pose(b200vo_ctx* ctx, const double* essential, const double* x1, const double* x2, int npts,
                        const uint8_t* mask_in, double* R, double* t, int32_t* n_valid, uint8_t* mask_out);
int b200vo_draw_samples(uint64_t seed, int64_t pair_id, int nhyp, int m, int32_t* samples /* nhyp*5 */);

/* Host-side helpers (tiny 3x3 algebra; the reference does these on the host too). */
int b200vo_chain_pose(const b200vo_pose_record* rec, const double* R_wc, const double* t_wc, double* R_out, double* t_out);
int b200vo_fundamental_from_essential(const b200vo_ctx* ctx, const double* essential, double* fundamental);

/* process_frame's accumulation over a whole range of records (reference src/visual_odometry.cpp:340-374):
 * poses_out[i] = pose after applying recs[0..i] to (R0, t0); 12 doubles each (R_wc row-major, t_wc). */
int b200vo_chain_trajectory(const b200vo_pose_record* recs, int n, const double* R0, const double* t0, double* poses_out);

/* ---- [ext] device groups: ONE sequence over the GPUs of one node ------------------------------------------
 * The reference walks a sequence strictly in order on one thread (src/main.cpp:128-194).  Frame pairs are
 * independent until the poses are chained, so a group partitions the pair range into contiguous blocks, one
 * per device, runs every block on its own context (one host thread each), all-gathers the 160-byte records
 * with NCCL (ncclCommInitAll + ncclAllGather over NVLink; libnccl.so.2 is bound at first use) and hands them to
 * the caller in pair order for b200vo_chain_trajectory.  Pair results are byte-identical to the one-device
 * run: nothing in a pair's computation depends on device, slot or batch.
 * A device may be listed k times in adjacent entries (the same k for every device), e.g. {0, 0} or {0, 0, 1, 1}: it
 * then carries k contexts, each with its own CUDA stream, host thread and contiguous sub-range, so the latency-bound
 * tail of one context's launch batch overlaps the other's throughput-bound kernels (about 3 % at the headline
 * configuration with k = 2); NCCL runs over the distinct devices. */
typedef struct b200vo_group b200vo_group;
int b200vo_group_create(const b200vo_options* opt, const int* devices, int ndev, b200vo_group** out);
void b200vo_group_destroy(b200vo_group* g);
int b200vo_group_size(const b200vo_group* g);
b200vo_ctx* b200vo_group_ctx(b200vo_group* g, int i);                 /* the i-th context (owned by the group) */
/* Contiguous pair range [lo, hi) of `rank` among `world` (the first ranks take the remainder). */
int b200vo_partition_pairs(int total_pairs, int world, int rank, int* lo, int* hi);
/* nframes host frames (layout as b200vo_upload_frames) -> nframes-1 records in pair order.  opt.max_frames of the
 * group must be >= batch_frames + 1. */
int b200vo_group_process_sequence(b200vo_group* g, const uint8_t* frames, int nframes, size_t pitch, size_t frame_stride,
                                  int channels, int batch_frames, int64_t first_pair_id, b200vo_pose_record* out);
/* The exchange step alone: context i contributes counts[i] records from its part of the send buffer (filled by
 * b200vo_group_process_sequence); out receives sum(counts) records in rank order. */
int b200vo_group_allgather_records(b200vo_group* g, const int* counts, b200vo_pose_record* out);
int b200vo_nccl_version(void);                                          /* 0 when NCCL could not be loaded */

/* Microbenchmark of the XOR+POPC+ADD instruction mix the matcher is bound by: sustained 32-bit
 * population counts per second on this device (the matcher's roofline denominator). */
int b200vo_measure_popc_rate(b200vo_ctx* ctx, double* popc_per_second);

/* Per-stage CUDA-event timings of the most recent detect / process_pairs call (ms). */
int b200vo_enable_timing(b200vo_ctx* ctx, int on);
int b200vo_get_stage_times(b200vo_ctx* ctx, b200vo_stage_times* out);
/* Number of kernels this library has launched on the context since creation. */
int64_t b200vo_kernel_launches(const b200vo_ctx* ctx);
/* cudaStream_t of the context (so a caller can record events on the launching stream). */
void* b200vo_stream(b200vo_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* B200VO_H */
