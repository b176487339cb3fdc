// Host side of the VO pipeline: the reference's VisualOdometry::process_frame orchestration
// (reference src/visual_odometry.cpp:323-378) over the b200vo C ABI.  Everything the reference
// delegates to OpenCV runs on the GPU; what stays here is what the reference also does on the
// host: the fallback logic, F = K^-T E K^-1 bookkeeping, pose chaining and trajectory storage.
#include "visual_odometry.h"

#include <algorithm>
#include <stdexcept>
#include <string>

#include "b200vo.h"

namespace {
void check(int rc, const char* what) {
    if (rc != B200VO_OK) throw std::runtime_error(std::string(what) + ": " + b200vo_last_error());
}

Pose chained(const Pose& prev, const b200vo_pose_record& rec) {
    double R[9], t[3], Ro[9], to[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = prev.R_wc.at<double>(i, j);
        t[i] = prev.t_wc.at<double>(i, 0);
    }
    check(b200vo_chain_pose(&rec, R, t, Ro, to), "chain_pose");
    cv::Mat Rm(3, 3, CV_64F), tm(3, 1, CV_64F);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Rm.at<double>(i, j) = Ro[i * 3 + j];
        tm.at<double>(i, 0) = to[i];
    }
    return Pose(Rm, tm);
}

VisualOdometry::PairDiagnostics diagnostics_of(const b200vo_pose_record& r) {
    VisualOdometry::PairDiagnostics d;
    d.status = r.status; d.cross_matches = r.n_cross; d.good_matches = r.n_matches; d.inliers = r.n_inliers; d.valid_points = r.n_valid;
    d.dist_min = r.dist_min; d.dist_max = r.dist_max; d.dist_mean = r.dist_mean; d.dist_median = r.dist_median; d.threshold = r.threshold;
    d.epipolar_residual = r.epi_residual;
    return d;
}

// rows of an image as a tightly described u8 buffer
struct ImageView { const uint8_t* data; size_t pitch; int channels; };
ImageView view_of(const cv::Mat& img, int w, int h) {
    if (img.empty() || img.cols != w || img.rows != h) throw std::runtime_error("image size does not match the VisualOdometry instance");
    const int ch = img.channels();
    if (img.depth() != CV_8U || (ch != 1 && ch != 3)) throw std::runtime_error("image must be 8-bit with 1 or 3 channels");
    return ImageView{img.data, (size_t)img.step, ch};
}
}  // namespace

VisualOdometry::VisualOdometry(int image_width, int image_height) { init(image_width, image_height, Options{}); }
VisualOdometry::VisualOdometry(int image_width, int image_height, const Options& options) { init(image_width, image_height, options); }

void VisualOdometry::init(int w, int h, const Options& o) {
    options_ = o;
    width_ = w; height_ = h;
    b200vo_options opt;
    check(b200vo_default_options(&opt, w, h), "default_options");
    opt.nfeatures = o.nfeatures;
    opt.ransac_hypotheses = o.ransac_hypotheses;
    opt.ransac_seed = o.ransac_seed;
    opt.match_mode = o.ratio_test ? B200VO_MATCH_RATIO : B200VO_MATCH_CROSSCHECK;
    opt.ratio = o.ratio;
    opt.max_frames = std::max(2, o.batch_frames + 1);
    std::vector<int> devs = o.devices.empty() ? std::vector<int>{o.device} : o.devices;
    check(b200vo_group_create(&opt, devs.data(), (int)devs.size(), &group_), "group_create");
    ctx_ = b200vo_group_ctx(group_, 0);
    // intrinsics synthesised from the image size (reference src/visual_odometry.cpp:90-98)
    camera_matrix_ = cv::Mat::eye(3, 3, CV_64F);
    camera_matrix_.at<double>(0, 0) = opt.fx; camera_matrix_.at<double>(1, 1) = opt.fy;
    camera_matrix_.at<double>(0, 2) = opt.cx; camera_matrix_.at<double>(1, 2) = opt.cy;
}

VisualOdometry::~VisualOdometry() { b200vo_group_destroy(group_); }

void VisualOdometry::push_pose(const Pose& p) {
    std::lock_guard<std::mutex> lock(trajectory_mutex_);
    trajectory_positions_.push_back(p.C_w().clone());
    trajectory_poses_.emplace_back(p.R_wc.clone(), p.t_wc.clone());
}

cv::Mat VisualOdometry::process_frame(Frame& frame) {
    const int slots = std::max(2, options_.batch_frames + 1);
    const int cur = initialized_ ? (previous_slot_ + 1) % slots : 0;
    const ImageView v = view_of(frame.image, width_, height_);
    // one call, one synchronisation: upload + detect + (if there is a previous frame) match / RANSAC / pose
    const int cap = b200vo_max_keypoints(ctx_);
    kp_buf_.resize((size_t)cap);
    desc_buf_.resize((size_t)cap * 32);
    int n = 0;
    b200vo_pose_record rec;
    double E[9];
    // first half: returns when the features are on the host; the pair's kernels are still running
    check(b200vo_process_frame_begin(ctx_, v.data, v.pitch, v.channels, cur, initialized_ ? previous_slot_ : -1, frame_counter_ - 1,
                                     kp_buf_.data(), desc_buf_.data(), cap, &n), "process_frame_begin");
    const b200vo_keypoint* kps = kp_buf_.data();
    // detect_features(): fill frame.keypoints / frame.descriptors (reference :103-117) -- overlaps matching and RANSAC
    frame.keypoints.resize((size_t)n);
    for (int i = 0; i < n; ++i) {
        cv::KeyPoint& k = frame.keypoints[(size_t)i];
        k.pt.x = kps[i].x; k.pt.y = kps[i].y; k.size = kps[i].size; k.angle = kps[i].angle;
        k.response = kps[i].response; k.octave = kps[i].octave; k.class_id = kps[i].class_id;
    }
    cv::Mat d(n, 32, CV_8UC1);
    for (int i = 0; i < n; ++i) std::copy(desc_buf_.data() + (size_t)i * 32, desc_buf_.data() + (size_t)(i + 1) * 32, d.data + (size_t)i * d.step);
    frame.descriptors = d;
    // second half: the pose record and E of the pair (previous frame, this frame)
    check(b200vo_process_frame_end(ctx_, &rec, E), "process_frame_end");
    frame.processed = true;
    const cv::Mat display_image = frame.image;      // shares the pixels; nothing is drawn in the headless build

    if (!initialized_) {
        frame.pose = Pose();
        push_pose(frame.pose);
        previous_frame_ = std::move(frame);
        previous_slot_ = cur;
        frame_counter_ = 1;
        initialized_ = true;
        return display_image;
    }

    // Default to the last known pose; the record's status bits encode the reference's fallbacks.
    frame.pose = Pose(previous_frame_.pose.R_wc, previous_frame_.pose.t_wc);
    last_pair_ = diagnostics_of(rec);
    if (rec.status & B200VO_ST_E_FOUND) {
        double F[9];
        check(b200vo_fundamental_from_essential(ctx_, E, F), "fundamental");
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) last_F_(i, j) = F[i * 3 + j];
        has_last_F_ = true;
        frame.pose = chained(previous_frame_.pose, rec);
    }
    push_pose(frame.pose);
    previous_frame_ = std::move(frame);
    previous_slot_ = cur;
    ++frame_counter_;
    return display_image;
}

std::vector<Pose> VisualOdometry::process_sequence(const std::vector<cv::Mat>& images) {
    std::vector<Pose> out;
    const int slots = std::max(2, options_.batch_frames + 1);
    const int B = slots - 1;
    size_t next = 0;
    if (!initialized_ && !images.empty()) {
        const ImageView v = view_of(images[0], width_, height_);
        check(b200vo_upload_frames(ctx_, v.data, 1, v.pitch, v.pitch * height_, v.channels, 0), "upload_frames");
        check(b200vo_detect(ctx_, 0, 1), "detect");
        Frame f(0, cv::Mat());
        f.pose = Pose();
        push_pose(f.pose);
        out.emplace_back(f.pose.R_wc, f.pose.t_wc);
        previous_frame_ = std::move(f);
        previous_slot_ = 0;
        frame_counter_ = 1;
        initialized_ = true;
        next = 1;
    }
    if (b200vo_group_size(group_) > 1 && images.size() - next >= 2 * (size_t)b200vo_group_size(group_)) {
        // Several GPUs: the remaining frames (plus the previous frame, which is pair 0's first image) form one
        // sequence that the group partitions by contiguous pair ranges; the records come back in pair order and
        // are chained here exactly as below.  The previous frame's pixels are needed again (another device may
        // own the first block), so this path requires it in `images`: it is taken from images[next - 1] when the
        // whole sequence goes through one call, which is how main() would use it.
        if (next == 0) throw std::runtime_error("process_sequence on several GPUs must start a sequence (call it once with all frames)");
        const size_t first_img = next - 1, nimg = images.size() - first_img;
        const ImageView v0 = view_of(images[first_img], width_, height_);
        const size_t fbytes = (size_t)width_ * v0.channels * height_;
        std::vector<uint8_t> packed(nimg * fbytes);
        for (size_t i = 0; i < nimg; ++i) {
            const ImageView v = view_of(images[first_img + i], width_, height_);
            if (v.channels != v0.channels) throw std::runtime_error("all images of a sequence must have the same channel count");
            for (int y = 0; y < height_; ++y) std::copy(v.data + (size_t)y * v.pitch, v.data + (size_t)y * v.pitch + (size_t)width_ * v.channels,
                                                        packed.data() + i * fbytes + (size_t)y * width_ * v.channels);
        }
        std::vector<b200vo_pose_record> recs(nimg - 1);
        check(b200vo_group_process_sequence(group_, packed.data(), (int)nimg, (size_t)width_ * v0.channels, fbytes, v0.channels, B,
                                            frame_counter_ - 1, recs.data()), "group_process_sequence");
        Pose cur(previous_frame_.pose.R_wc, previous_frame_.pose.t_wc);
        for (size_t i = 0; i < recs.size(); ++i) {
            if (recs[i].status & B200VO_ST_E_FOUND) cur = chained(cur, recs[i]);
            push_pose(cur);
            out.emplace_back(cur.R_wc, cur.t_wc);
        }
        if (!recs.empty()) last_pair_ = diagnostics_of(recs.back());
        previous_frame_.pose = Pose(cur.R_wc, cur.t_wc);
        previous_frame_.id = (int)(frame_counter_ + (int64_t)recs.size() - 1);
        frame_counter_ += (int64_t)recs.size();
        // the last frame's features live on the last device; a following process_frame / process_sequence call
        // re-detects it on the first device
        {
            const ImageView v = view_of(images.back(), width_, height_);
            check(b200vo_upload_frames(ctx_, v.data, 1, v.pitch, v.pitch * height_, v.channels, 0), "upload_frames");
            check(b200vo_detect(ctx_, 0, 1), "detect");
            check(b200vo_synchronize(ctx_), "synchronize");
            previous_slot_ = 0;
        }
        return out;
    }
    std::vector<b200vo_pose_record> recs((size_t)B);
    while (next < images.size()) {
        const int n = (int)std::min<size_t>((size_t)B, images.size() - next);
        const int first = (previous_slot_ + 1) % slots;
        for (int i = 0; i < n; ++i) {
            const ImageView v = view_of(images[next + i], width_, height_);
            check(b200vo_upload_frames(ctx_, v.data, 1, v.pitch, v.pitch * height_, v.channels, (first + i) % slots), "upload_frames");
        }
        check(b200vo_detect(ctx_, first, n), "detect");
        check(b200vo_process_pairs(ctx_, previous_slot_, n, frame_counter_ - 1, recs.data()), "process_pairs");
        Pose cur(previous_frame_.pose.R_wc, previous_frame_.pose.t_wc);
        for (int i = 0; i < n; ++i) {
            if (recs[i].status & B200VO_ST_E_FOUND) cur = chained(cur, recs[i]);
            push_pose(cur);
            out.emplace_back(cur.R_wc, cur.t_wc);
        }
        if (n > 0) last_pair_ = diagnostics_of(recs[(size_t)n - 1]);
        previous_frame_.pose = Pose(cur.R_wc, cur.t_wc);
        previous_frame_.id = (int)(frame_counter_ + n - 1);
        previous_slot_ = (first + n - 1) % slots;
        frame_counter_ += n;
        next += (size_t)n;
    }
    return out;
}

std::vector<cv::Mat> VisualOdometry::get_trajectory() {
    std::lock_guard<std::mutex> lock(trajectory_mutex_);
    return trajectory_positions_;
}

std::vector<Pose> VisualOdometry::get_trajectory_poses() {
    std::lock_guard<std::mutex> lock(trajectory_mutex_);
    std::vector<Pose> copy;
    copy.reserve(trajectory_poses_.size());
    for (const Pose& p : trajectory_poses_) copy.emplace_back(p.R_wc, p.t_wc);
    return copy;
}

Pose VisualOdometry::current_pose() const {
    if (trajectory_poses_.empty()) return Pose();
    const Pose& p = trajectory_poses_.back();
    return Pose(p.R_wc, p.t_wc);
}

bool VisualOdometry::has_last_F() const { return has_last_F_; }
cv::Matx33d VisualOdometry::last_F() const { return last_F_; }
