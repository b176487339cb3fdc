// See include/sequence_io.h.  Reference counterpart: src/main.cpp:26-49, :136, :146-151.
#include "sequence_io.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <filesystem>
#include <cstdlib>

#include "pose.h"

namespace fs = std::filesystem;

std::vector<std::string> load_image_paths(const std::string& dataset_path) {
    std::vector<std::string> paths;
    for (const auto& entry : fs::directory_iterator(dataset_path)) {
        if (!entry.is_regular_file()) continue;
        std::string ext = entry.path().extension().string();
        std::transform(ext.begin(), ext.end(), ext.begin(), [](unsigned char c) { return (char)std::tolower(c); });
        if (ext == ".png" || ext == ".jpg" || ext == ".jpeg" || ext == ".pgm" || ext == ".ppm") paths.push_back(entry.path().string());
    }
    std::sort(paths.begin(), paths.end());
    return paths;
}

double timestamp_from_path(const std::string& path, size_t index) {
    const std::string stem = fs::path(path).stem().string();
    try {
        return std::stod(stem);
    } catch (...) {
        return static_cast<double>(index);
    }
}

// C stdio on purpose: the library may be loaded into processes that carry another C++ runtime
// (Python extensions), and iostream formatting depends on locale state that is per runtime copy.
static bool next_token(FILE* in, std::string& tok) {
    tok.clear();
    int c;
    while ((c = std::fgetc(in)) != EOF) {
        if (c == '#') { while ((c = std::fgetc(in)) != EOF && c != '\n') {} continue; }
        if (std::isspace(c)) { if (!tok.empty()) return true; continue; }
        tok.push_back((char)c);
    }
    return !tok.empty();
}

cv::Mat read_pnm(const std::string& path) {
    FILE* in = std::fopen(path.c_str(), "rb");
    if (!in) return cv::Mat();
    struct Closer { FILE* f; ~Closer() { std::fclose(f); } } closer{in};
    std::string magic, sw, sh, smax;
    if (!next_token(in, magic) || (magic != "P5" && magic != "P6")) return cv::Mat();
    if (!next_token(in, sw) || !next_token(in, sh) || !next_token(in, smax)) return cv::Mat();
    const int w = std::atoi(sw.c_str()), h = std::atoi(sh.c_str()), mx = std::atoi(smax.c_str());
    if (w <= 0 || h <= 0 || mx != 255) return cv::Mat();        // next_token consumed the single whitespace after maxval
    const int ch = magic == "P6" ? 3 : 1;
    cv::Mat img(h, w, ch == 3 ? CV_8UC3 : CV_8UC1);
    for (int y = 0; y < h; ++y) {
        uint8_t* p = img.ptr<uint8_t>(y);
        if (std::fread(p, 1, (size_t)w * ch, in) != (size_t)w * ch) return cv::Mat();
        if (ch == 3)                                              // RGB on disk -> BGR in memory (cv::imread's order)
            for (int x = 0; x < w; ++x) std::swap(p[3 * x], p[3 * x + 2]);
    }
    return img;
}

bool write_pnm(const std::string& path, const cv::Mat& image) {
    if (image.empty() || image.depth() != CV_8U || (image.channels() != 1 && image.channels() != 3)) return false;
    FILE* out = std::fopen(path.c_str(), "wb");
    if (!out) return false;
    const int ch = image.channels();
    std::fprintf(out, "%s\n%d %d\n255\n", ch == 3 ? "P6" : "P5", image.cols, image.rows);
    std::vector<uint8_t> row((size_t)image.cols * ch);
    bool ok = true;
    for (int y = 0; y < image.rows && ok; ++y) {
        const uint8_t* p = image.ptr<uint8_t>(y);
        for (int x = 0; x < image.cols; ++x)
            for (int c = 0; c < ch; ++c) row[(size_t)x * ch + c] = p[(size_t)x * ch + (ch == 3 ? 2 - c : c)];
        ok = std::fwrite(row.data(), 1, row.size(), out) == row.size();
    }
    return std::fclose(out) == 0 && ok;
}

bool save_trajectory_text(const std::string& path, const std::vector<double>& timestamps, const std::vector<Pose>& poses) {
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return false;
    std::fprintf(f, "# timestamp tx ty tz r00 r01 r02 r10 r11 r12 r20 r21 r22   (T_wc)\n");
    for (size_t i = 0; i < poses.size(); ++i) {
        const Pose& p = poses[i];
        std::fprintf(f, "%.9f %.9f %.9f %.9f", i < timestamps.size() ? timestamps[i] : (double)i, p.t_wc.at<double>(0, 0),
                     p.t_wc.at<double>(1, 0), p.t_wc.at<double>(2, 0));
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) std::fprintf(f, " %.12f", p.R_wc.at<double>(r, c));
        std::fprintf(f, "\n");
    }
    std::fclose(f);
    return true;
}
