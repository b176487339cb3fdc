// slam_vo_headless: the reference's slam_vo_run (src/main.cpp:51-220) without its Pangolin / highgui
// front end -- find the dataset directory, list the images in order, and run
//     image -> Frame(i, image) -> timestamp from the file stem -> VisualOdometry::process_frame
// once per frame exactly as src/main.cpp:128-155 does; then write the trajectory as text instead of
// screenshots.  `--batched` uses the process_sequence extension (same poses, launch batches).
#include <cctype>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "sequence_io.h"
#include "visual_odometry.h"

int main(int argc, char** argv) {
    std::string dir, out = "trajectory.txt", tum, gt;
    bool batched = false;
    int nfeatures = 1200;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--batched")) batched = true;
        else if (!std::strcmp(argv[i], "--features") && i + 1 < argc) nfeatures = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--out") && i + 1 < argc) out = argv[++i];
        else if (!std::strcmp(argv[i], "--tum") && i + 1 < argc) tum = argv[++i];
        else if (!std::strcmp(argv[i], "--gt") && i + 1 < argc) gt = argv[++i];
        else dir = argv[i];
    }
    if (dir.empty()) { std::fprintf(stderr, "usage: %s <image directory> [--batched] [--features N] [--out file] [--tum file] [--gt tum-file]\n", argv[0]); return 2; }
    std::vector<std::string> paths;
    try { paths = load_image_paths(dir); } catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
    if (paths.empty()) { std::fprintf(stderr, "No images found in dataset directory: %s\n", dir.c_str()); return 1; }      // main.cpp:92-95
    // load_image_paths lists what the reference lists (png / jpg / jpeg, src/main.cpp:26-49) plus Netpbm; this build decodes
    // png and Netpbm, so a directory it cannot read completely is refused up front instead of silently pairing other frames
    for (const std::string& p : paths) {
        const size_t dot = p.find_last_of('.');
        std::string ext = dot == std::string::npos ? "" : p.substr(dot + 1);
        for (char& c : ext) c = (char)std::tolower((unsigned char)c);
        if (ext != "png" && ext != "pgm" && ext != "ppm" && ext != "pnm") {
            std::fprintf(stderr, "%s: this build decodes png and binary Netpbm (P5/P6) only -- convert the sequence, or build with OpenCV and use cv::imread\n", p.c_str());
            return 1;
        }
    }
    const cv::Mat first = read_image(paths[0]);
    if (first.empty()) { std::fprintf(stderr, "Failed to load first image: %s (this build decodes png and P5/P6 only)\n", paths[0].c_str()); return 1; }
    try {
        VisualOdometry::Options opt;
        opt.nfeatures = nfeatures;
        VisualOdometry vo(first.cols, first.rows, opt);
        std::vector<double> stamps;
        if (batched) {
            std::vector<cv::Mat> images;
            for (size_t i = 0; i < paths.size(); ++i) {
                cv::Mat im = read_image(paths[i]);
                if (im.empty()) { std::fprintf(stderr, "Failed to load image: %s\n", paths[i].c_str()); continue; }         // main.cpp:137-140
                images.push_back(im);
                stamps.push_back(timestamp_from_path(paths[i], i));
            }
            vo.process_sequence(images);
        } else {
            for (size_t i = 0; i < paths.size(); ++i) {
                cv::Mat im = read_image(paths[i]);
                if (im.empty()) { std::fprintf(stderr, "Failed to load image: %s\n", paths[i].c_str()); continue; }
                Frame frame((int)i, im);
                frame.timestamp = timestamp_from_path(paths[i], i);
                stamps.push_back(frame.timestamp);
                vo.process_frame(frame);
            }
        }
        const std::vector<Pose> traj = vo.get_trajectory_poses();
        if (!save_trajectory_text(out, stamps, traj)) { std::fprintf(stderr, "cannot write %s\n", out.c_str()); return 1; }
        if (!tum.empty() && !save_trajectory_tum(tum, stamps, traj)) { std::fprintf(stderr, "cannot write %s\n", tum.c_str()); return 1; }
        std::printf("%zu frames, trajectory written to %s\n", traj.size(), out.c_str());
        if (!gt.empty()) {       // accuracy against a ground-truth trajectory in TUM format (frame i <-> line i)
            std::vector<Pose> ref;
            TrajectoryReport rep;
            if (!load_trajectory_tum(gt, nullptr, &ref) || !trajectory_report(traj, ref, &rep)) { std::fprintf(stderr, "cannot evaluate against %s\n", gt.c_str()); return 1; }
            std::printf("vs %s: %d frames, ATE rmse (similarity-aligned) %.6f, scale %.6f, RPE rotation %.4f deg, translation direction %.4f deg\n",
                        gt.c_str(), rep.frames, rep.ate_rmse_aligned, rep.alignment_scale, rep.rpe_rot_deg_mean, rep.rpe_trans_dir_deg_mean);
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
