// Frame: one time step of the VO pipeline (image + features + pose).  Same public fields as the
// reference's src/frame.h:19-28; move-only like the reference (its user-declared move operations
// suppress copying, src/frame.h:36-57); the image is cloned on construction (src/frame.h:32-34).
#pragma once
#include <utility>
#include <vector>

#include "pose.h"
#include "vo_cv_compat.h"

class Frame {
public:
    int id = -1;
    std::vector<cv::KeyPoint> keypoints;   // pt = level-0 pixel coordinates (0-based)
    cv::Mat descriptors;                   // N x 32, CV_8U, row i belongs to keypoints[i]
    Pose pose;                             // estimated T_wc of this frame
    cv::Mat image;
    double timestamp = 0.0;
    bool processed = false;

    Frame() = default;
    Frame(int frame_id, const cv::Mat& img) : id(frame_id), image(img.clone()) {}
    Frame(Frame&&) noexcept = default;
    Frame& operator=(Frame&&) noexcept = default;
    Frame(const Frame&) = delete;
    Frame& operator=(const Frame&) = delete;
};
