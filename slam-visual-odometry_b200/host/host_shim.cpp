// Flat C wrappers around the C++ VisualOdometry class so the Python parity tests can drive the
// reference-shaped API (construct, process_frame per frame, read the trajectory) through ctypes.
#include <cstring>
#include <exception>
#include <string>
#include <vector>

#include "sequence_io.h"
#include "visual_odometry.h"

namespace { thread_local std::string g_msg; }

extern "C" {

const char* vo_host_last_error() { return g_msg.c_str(); }

void* vo_host_create(int width, int height, int nfeatures, int ransac_hypotheses, unsigned long long seed, int device,
                     int batch_frames) {
    try {
        VisualOdometry::Options o;
        if (nfeatures > 0) o.nfeatures = nfeatures;
        if (ransac_hypotheses > 0) o.ransac_hypotheses = ransac_hypotheses;
        o.ransac_seed = seed; o.device = device;
        if (batch_frames > 0) o.batch_frames = batch_frames;
        return new VisualOdometry(width, height, o);
    } catch (const std::exception& e) { g_msg = e.what(); return nullptr; }
}

// the same over several devices (Options::devices): process_sequence partitions the pair range over them
void* vo_host_create_multi(int width, int height, int nfeatures, int ransac_hypotheses, unsigned long long seed, const int* devices,
                           int ndev, int batch_frames) {
    try {
        VisualOdometry::Options o;
        if (nfeatures > 0) o.nfeatures = nfeatures;
        if (ransac_hypotheses > 0) o.ransac_hypotheses = ransac_hypotheses;
        o.ransac_seed = seed;
        o.devices.assign(devices, devices + ndev);
        if (batch_frames > 0) o.batch_frames = batch_frames;
        return new VisualOdometry(width, height, o);
    } catch (const std::exception& e) { g_msg = e.what(); return nullptr; }
}

void vo_host_destroy(void* h) { delete static_cast<VisualOdometry*>(h); }

// pose_out: 12 doubles (R_wc row-major, then t_wc); returns number of keypoints or <0 on error
int vo_host_process_frame(void* h, int frame_id, const unsigned char* img, int width, int height, int channels,
                          double* pose_out, int* moved_from) {
    try {
        VisualOdometry* vo = static_cast<VisualOdometry*>(h);
        cv::Mat wrap(height, width, channels == 3 ? CV_8UC3 : CV_8UC1, const_cast<unsigned char*>(img));
        Frame f(frame_id, wrap);
        const cv::Mat disp = vo->process_frame(f);
        // the reference moves the frame into previous_frame_: the caller's keypoints are gone
        if (moved_from) *moved_from = f.keypoints.empty() ? 1 : 0;
        const Pose p = vo->current_pose();
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) pose_out[i * 3 + j] = p.R_wc.at<double>(i, j);
            pose_out[9 + i] = p.t_wc.at<double>(i, 0);
        }
        return disp.empty() ? -2 : (int)vo->trajectory_size();
    } catch (const std::exception& e) { g_msg = e.what(); return -1; }
}

// frames: n tightly packed images; poses_out: n*12 doubles
int vo_host_process_sequence(void* h, const unsigned char* frames, int n, int width, int height, int channels,
                             double* poses_out) {
    try {
        VisualOdometry* vo = static_cast<VisualOdometry*>(h);
        std::vector<cv::Mat> imgs;
        const size_t fb = (size_t)width * height * channels;
        for (int i = 0; i < n; ++i)
            imgs.emplace_back(height, width, channels == 3 ? CV_8UC3 : CV_8UC1, const_cast<unsigned char*>(frames + (size_t)i * fb));
        const std::vector<Pose> ps = vo->process_sequence(imgs);
        for (size_t k = 0; k < ps.size(); ++k)
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) poses_out[k * 12 + i * 3 + j] = ps[k].R_wc.at<double>(i, j);
                poses_out[k * 12 + 9 + i] = ps[k].t_wc.at<double>(i, 0);
            }
        return (int)ps.size();
    } catch (const std::exception& e) { g_msg = e.what(); return -1; }
}

// out[11]: status, cross, good, inliers, valid, dist min / max / mean / median, threshold, epipolar residual
void vo_host_last_pair(void* h, double* out) {
    const VisualOdometry::PairDiagnostics d = static_cast<VisualOdometry*>(h)->last_pair();
    const double v[11] = {(double)d.status, (double)d.cross_matches, (double)d.good_matches, (double)d.inliers, (double)d.valid_points,
                          d.dist_min, d.dist_max, d.dist_mean, d.dist_median, d.threshold, d.epipolar_residual};
    for (int i = 0; i < 11; ++i) out[i] = v[i];
}

int vo_host_trajectory_size(void* h) { return (int)static_cast<VisualOdometry*>(h)->get_trajectory().size(); }

int vo_host_last_F(void* h, double* F) {
    VisualOdometry* vo = static_cast<VisualOdometry*>(h);
    if (!vo->has_last_F()) return 0;
    const cv::Matx33d f = vo->last_F();
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) F[i * 3 + j] = f(i, j);
    return 1;
}

// ---- sequence I/O (include/sequence_io.h): listing, timestamps, Netpbm round trip
int vo_host_list_images(const char* dir, char* out, int cap) {
    try {
        std::string joined;
        const std::vector<std::string> p = load_image_paths(dir);
        for (const std::string& s : p) { joined += s; joined += '\n'; }
        if ((int)joined.size() + 1 > cap) return -2;
        std::memcpy(out, joined.c_str(), joined.size() + 1);
        return (int)p.size();
    } catch (const std::exception& e) { g_msg = e.what(); return -1; }
}

double vo_host_timestamp(const char* path, int index) { return timestamp_from_path(path, (size_t)index); }

// returns channels (1 / 3), 0 on failure; data = rows*cols*channels bytes, BGR for colour
int vo_host_read_pnm(const char* path, unsigned char* data, int cap, int* rows, int* cols) {
    const cv::Mat m = read_pnm(path);
    if (m.empty()) return 0;
    const int ch = m.channels();
    if (cap < m.rows * m.cols * ch) return -2;
    for (int y = 0; y < m.rows; ++y) std::memcpy(data + (size_t)y * m.cols * ch, m.ptr<uint8_t>(y), (size_t)m.cols * ch);
    *rows = m.rows; *cols = m.cols;
    return ch;
}

// poses12: n rows [R_wc row-major | t_wc]; stamps: n doubles or NULL
int vo_host_save_tum(const char* path, const double* poses12, int n, const double* stamps) {
    try {
        std::vector<Pose> poses;
        std::vector<double> ts;
        for (int i = 0; i < n; ++i) {
            cv::Mat R(3, 3, CV_64F), t(3, 1, CV_64F);
            for (int r = 0; r < 3; ++r) {
                for (int c = 0; c < 3; ++c) R.at<double>(r, c) = poses12[i * 12 + r * 3 + c];
                t.at<double>(r, 0) = poses12[i * 12 + 9 + r];
            }
            poses.emplace_back(R, t);
            if (stamps) ts.push_back(stamps[i]);
        }
        return save_trajectory_tum(path, ts, poses) ? 1 : 0;
    } catch (const std::exception& e) { g_msg = e.what(); return -1; }
}

// est12 / gt12: n rows [R_wc row-major | t_wc]; out[5]: frames, ATE rmse, scale, RPE rotation, RPE translation direction
int vo_host_trajectory_report(const double* est12, const double* gt12, int n, double* out) {
    try {
        auto mk = [n](const double* p) {
            std::vector<Pose> v;
            for (int i = 0; i < n; ++i) {
                cv::Mat R(3, 3, CV_64F), t(3, 1, CV_64F);
                for (int r = 0; r < 3; ++r) {
                    for (int c = 0; c < 3; ++c) R.at<double>(r, c) = p[i * 12 + r * 3 + c];
                    t.at<double>(r, 0) = p[i * 12 + 9 + r];
                }
                v.emplace_back(R, t);
            }
            return v;
        };
        TrajectoryReport rep;
        if (!trajectory_report(mk(est12), mk(gt12), &rep)) return 0;
        out[0] = rep.frames; out[1] = rep.ate_rmse_aligned; out[2] = rep.alignment_scale; out[3] = rep.rpe_rot_deg_mean; out[4] = rep.rpe_trans_dir_deg_mean;
        return 1;
    } catch (const std::exception& e) { g_msg = e.what(); return -1; }
}

// poses12_out: up to cap rows [R_wc | t_wc]; returns the number of poses in the file or < 0
int vo_host_load_tum(const char* path, double* poses12_out, double* stamps_out, int cap) {
    try {
        std::vector<Pose> poses;
        std::vector<double> ts;
        if (!load_trajectory_tum(path, &ts, &poses)) return -1;
        for (int i = 0; i < (int)poses.size() && i < cap; ++i) {
            for (int r = 0; r < 3; ++r) {
                for (int c = 0; c < 3; ++c) poses12_out[i * 12 + r * 3 + c] = poses[(size_t)i].R_wc.at<double>(r, c);
                poses12_out[i * 12 + 9 + r] = poses[(size_t)i].t_wc.at<double>(r, 0);
            }
            if (stamps_out) stamps_out[i] = ts[(size_t)i];
        }
        return (int)poses.size();
    } catch (const std::exception& e) { g_msg = e.what(); return -2; }
}

// the same through read_image (PNG or Netpbm by content)
int vo_host_read_image(const char* path, unsigned char* data, int cap, int* rows, int* cols) {
    const cv::Mat m = read_image(path);
    if (m.empty()) return 0;
    const int ch = m.channels();
    if (cap < m.rows * m.cols * ch) return -2;
    for (int y = 0; y < m.rows; ++y) std::memcpy(data + (size_t)y * m.cols * ch, m.ptr<uint8_t>(y), (size_t)m.cols * ch);
    *rows = m.rows; *cols = m.cols;
    return ch;
}

int vo_host_write_pnm(const char* path, const unsigned char* data, int rows, int cols, int channels) {
    cv::Mat m(rows, cols, channels == 3 ? CV_8UC3 : CV_8UC1, const_cast<unsigned char*>(data));
    return write_pnm(path, m) ? 1 : 0;
}

}  // extern "C"
