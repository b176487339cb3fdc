// Image-sequence input for the headless build: what the reference's main() does around
// process_frame (src/main.cpp:26-49 load_image_paths, :136 cv::imread, :146-151 timestamp from the
// file stem).  OpenCV's image codecs are not available in this build, so the reader brings its own decoders: PNG (every
// colour type and bit depth, non-interlaced, on zlib) -- the format of the reference's datasets -- and the binary Netpbm
// formats (P5 gray, P6 RGB -> BGR like cv::imread).  JPEG is not decoded here; with real OpenCV installed use cv::imread
// instead and everything downstream is unchanged.
#pragma once
#include <string>
#include <vector>

#include "vo_cv_compat.h"

// png / jpg / jpeg as in the reference, plus pgm / ppm; regular files only; lexicographic order
std::vector<std::string> load_image_paths(const std::string& dataset_path);
// the file stem parsed as a double, or `index` when it is not a number (src/main.cpp:146-151)
double timestamp_from_path(const std::string& path, size_t index);
// P5 -> CV_8UC1, P6 -> CV_8UC3 in BGR order; empty Mat on failure (like cv::imread)
cv::Mat read_pnm(const std::string& path);
bool write_pnm(const std::string& path, const cv::Mat& image);
// PNG -> CV_8UC1 (gray, gray + alpha) or CV_8UC3 BGR (RGB, RGBA, palette); 16-bit samples reduced to their high byte; empty on failure
cv::Mat read_png(const std::string& path);
// PNG or Netpbm, decided by the file's first bytes; empty Mat for anything else (jpeg: no codec in this build)
cv::Mat read_image(const std::string& path);
// "timestamp tx ty tz r00 r01 ... r22" per line (T_wc)
bool save_trajectory_text(const std::string& path, const std::vector<double>& timestamps, const std::vector<class Pose>& poses);
// "timestamp tx ty tz qx qy qz qw" per line (TUM format, T_wc): what trajectory evaluation tools read
bool save_trajectory_tum(const std::string& path, const std::vector<double>& timestamps, const std::vector<class Pose>& poses);
bool load_trajectory_tum(const std::string& path, std::vector<double>* timestamps, std::vector<class Pose>* poses);
// Accuracy of an estimated trajectory against a reference one (positions after a similarity alignment, since the reference
// steps with a fixed scale of 0.3; relative rotation error and angle between the translation directions per step): the
// numbers b200vo/trajectory.py::accuracy_report and bench.py's result_check give
struct TrajectoryReport { int frames; double ate_rmse_aligned, alignment_scale, rpe_rot_deg_mean, rpe_trans_dir_deg_mean; };
bool trajectory_report(const std::vector<class Pose>& estimated, const std::vector<class Pose>& reference, TrajectoryReport* out);
