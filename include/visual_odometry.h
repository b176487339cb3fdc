// VisualOdometry: the reference's per-frame VO API (src/visual_odometry.h:31-66) on top of the
// B200 kernels.  Same class name, constructor, process_frame / get_trajectory /
// get_trajectory_poses / has_last_F / last_F, same fallback semantics
// (src/visual_odometry.cpp:323-378).  Extensions that the reference lacks are marked [ext]:
// an options overload (the reference hard-codes nfeatures = 1200) and a batched
// process_sequence (a per-frame streaming call cannot be batched or partitioned).
#pragma once
#include <cstdint>
#include <mutex>
#include <vector>

#include "frame.h"
#include "vo_cv_compat.h"

struct b200vo_ctx;
struct b200vo_group;

class VisualOdometry {
public:
    struct Options {                       // [ext]
        int nfeatures = 1200;              // reference src/visual_odometry.cpp:65
        int ransac_hypotheses = 1024;
        uint64_t ransac_seed = 42;
        int device = 0;
        std::vector<int> devices;          // [ext] more than one entry: process_sequence partitions the pair range over these GPUs
                                           // (contiguous blocks, one host thread per GPU, NCCL all-gather of the pose records);
                                           // process_frame always runs on the first.  A device may be repeated in adjacent
                                           // entries ({0, 0} or {0, 0, 1, 1}): that many contexts / CUDA streams on it, whose
                                           // launch batches overlap (b200vo.h, device groups)
        int batch_frames = 64;             // frames per launch batch in process_sequence
        bool ratio_test = false;           // false: cross-check (reference :87); true: Lowe ratio test on 2-NN
        float ratio = 0.75f;
    };

    VisualOdometry(int image_width, int image_height);
    VisualOdometry(int image_width, int image_height, const Options& options);   // [ext]
    ~VisualOdometry();
    VisualOdometry(const VisualOdometry&) = delete;
    VisualOdometry& operator=(const VisualOdometry&) = delete;

    // Detects features into frame.keypoints/descriptors, writes frame.pose (T_wc), appends to the
    // trajectory and MOVES the frame into the internal previous frame (as the reference does).
    // Returns the display image; this headless build draws nothing (the reference overlays the keypoints for its
    // GUI) and returns the input image itself -- a reference-counted handle, as every cv::Mat copy is.
    cv::Mat process_frame(Frame& frame);

    // [ext] Batched path: all frames are uploaded, detected and matched in launch batches; the
    // relative poses are chained on the host exactly as process_frame would.  Appends to the
    // trajectory; returns the poses of the processed frames.
    std::vector<Pose> process_sequence(const std::vector<cv::Mat>& images);

    std::vector<cv::Mat> get_trajectory();
    std::vector<Pose> get_trajectory_poses();
    bool has_last_F() const;
    cv::Matx33d last_F() const;

    const cv::Mat& camera_matrix() const { return camera_matrix_; }   // [ext]
    Pose current_pose() const;                                         // [ext] T_wc of the most recent frame (no copy of the whole trajectory)
    size_t trajectory_size() const { return trajectory_poses_.size(); }   // [ext]

    // [ext] What the reference prints to stdout for the most recent pair -- the [MatchDebug] line
    // (src/visual_odometry.cpp:119-129,160-169), the inlier / valid-point counts (:265-279) and the mean
    // |x2^T F x1| over the inliers (:248-263) -- as values instead of log lines.
    struct PairDiagnostics {
        int status = 0;                 // B200VO_ST_* bits (matched / E found / pose ok / overflow)
        int cross_matches = 0, good_matches = 0, inliers = 0, valid_points = 0;
        double dist_min = 0, dist_max = 0, dist_mean = 0, dist_median = 0, threshold = 0;
        double epipolar_residual = 0;
    };
    PairDiagnostics last_pair() const { return last_pair_; }

private:
    void init(int w, int h, const Options& o);
    void push_pose(const Pose& p);

    Options options_;
    int width_ = 0, height_ = 0;
    b200vo_ctx* ctx_ = nullptr;         // context of the first device (owned by group_)
    b200vo_group* group_ = nullptr;
    cv::Mat camera_matrix_;
    bool has_last_F_ = false;
    PairDiagnostics last_pair_{};
    cv::Matx33d last_F_{};
    std::vector<Pose> trajectory_poses_;
    std::vector<cv::Mat> trajectory_positions_;
    std::mutex trajectory_mutex_;
    Frame previous_frame_;
    int previous_slot_ = 0;
    int64_t frame_counter_ = 0;
    bool initialized_ = false;
    std::vector<struct b200vo_keypoint> kp_buf_;      // reused per frame (no allocation in process_frame)
    std::vector<uint8_t> desc_buf_;
};
