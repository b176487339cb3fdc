// See include/sequence_io.h.  Reference counterpart: src/main.cpp:26-49, :136, :146-151.
#include "sequence_io.h"

#include <algorithm>
#include <cctype>
#include <cstdio>
#include <filesystem>
#include <cstdlib>
#include <cmath>
#include <cstring>

#include <zlib.h>

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

// ---- PNG (the reference's datasets are png files read with cv::imread, src/main.cpp:136) -------------------------------
// A plain decoder on zlib: every colour type and bit depth of the PNG specification, no interlacing.  The result is what
// the reference's pipeline gets out of cv::imread + cvtColor: 8 bits per channel, alpha dropped, 16-bit samples reduced to
// their high byte (libpng's strip_16, which OpenCV uses), gray kept as one channel (cv::imread would replicate it into
// three equal channels that BGR2GRAY maps back to the same value), colour as BGR.
static uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

cv::Mat read_png(const std::string& path) {
    FILE* in = std::fopen(path.c_str(), "rb");
    if (!in) return cv::Mat();
    struct Closer { FILE* f; ~Closer() { std::fclose(f); } } closer{in};
    uint8_t sig[8];
    static const uint8_t kSig[8] = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1A, '\n'};
    if (std::fread(sig, 1, 8, in) != 8 || std::memcmp(sig, kSig, 8) != 0) return cv::Mat();
    uint32_t w = 0, h = 0;
    int depth = 0, ctype = -1;
    std::vector<uint8_t> idat, plte;
    bool have_ihdr = false, done = false;
    while (!done) {
        uint8_t hdr[8];
        if (std::fread(hdr, 1, 8, in) != 8) return cv::Mat();
        const uint32_t len = be32(hdr);
        if (len > (1u << 30)) return cv::Mat();
        std::vector<uint8_t> data(len);
        uint8_t crcb[4];
        if ((len && std::fread(data.data(), 1, len, in) != len) || std::fread(crcb, 1, 4, in) != 4) return cv::Mat();
        uLong crc = crc32(0L, hdr + 4, 4);
        if (len) crc = crc32(crc, data.data(), len);
        if ((uint32_t)crc != be32(crcb)) return cv::Mat();
        if (std::memcmp(hdr + 4, "IHDR", 4) == 0) {
            if (len != 13) return cv::Mat();
            w = be32(data.data()); h = be32(data.data() + 4);
            depth = data[8]; ctype = data[9];
            if (data[10] != 0 || data[11] != 0 || data[12] != 0) return cv::Mat();     // compression, filter method, interlace: only the plain ones
            have_ihdr = true;
        } else if (std::memcmp(hdr + 4, "PLTE", 4) == 0) {
            plte = data;
        } else if (std::memcmp(hdr + 4, "IDAT", 4) == 0) {
            idat.insert(idat.end(), data.begin(), data.end());
        } else if (std::memcmp(hdr + 4, "IEND", 4) == 0) {
            done = true;
        }
    }
    if (!have_ihdr || w == 0 || h == 0 || w > 65535 || h > 65535) return cv::Mat();
    int samples;                                                  // samples per pixel
    switch (ctype) {
        case 0: samples = 1; break;                               // gray
        case 2: samples = 3; break;                               // RGB
        case 3: samples = 1; break;                               // palette index
        case 4: samples = 2; break;                               // gray + alpha
        case 6: samples = 4; break;                               // RGBA
        default: return cv::Mat();
    }
    const bool depth_ok = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                          (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                          ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
    if (!depth_ok || (ctype == 3 && plte.size() < 3)) return cv::Mat();
    const size_t row_bytes = ((size_t)w * samples * depth + 7) / 8;
    const int bpp = std::max(1, samples * depth / 8);             // filter distance in bytes
    std::vector<uint8_t> raw((row_bytes + 1) * h);
    uLongf raw_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &raw_len, idat.data(), (uLong)idat.size()) != Z_OK || raw_len != raw.size()) return cv::Mat();
    // undo the per-row filters in place (row r starts at r * (row_bytes + 1) with its filter type byte)
    std::vector<uint8_t> zero(row_bytes, 0);
    for (uint32_t y = 0; y < h; ++y) {
        uint8_t* cur = raw.data() + (size_t)y * (row_bytes + 1);
        const int ft = cur[0];
        uint8_t* x = cur + 1;
        const uint8_t* up = y ? raw.data() + (size_t)(y - 1) * (row_bytes + 1) + 1 : zero.data();
        for (size_t i = 0; i < row_bytes; ++i) {
            const int a = i >= (size_t)bpp ? x[i - bpp] : 0, b = up[i], c = i >= (size_t)bpp ? up[i - bpp] : 0;
            int pred;
            switch (ft) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: {
                    const int pp = a + b - c, pa = std::abs(pp - a), pb = std::abs(pp - b), pc = std::abs(pp - c);
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: return cv::Mat();
            }
            x[i] = (uint8_t)(x[i] + pred);
        }
    }
    const bool colour = (ctype == 2 || ctype == 6 || ctype == 3);
    cv::Mat img((int)h, (int)w, colour ? CV_8UC3 : CV_8UC1);
    const int maxv = (1 << depth) - 1;
    for (uint32_t y = 0; y < h; ++y) {
        const uint8_t* x = raw.data() + (size_t)y * (row_bytes + 1) + 1;
        uint8_t* o = img.ptr<uint8_t>((int)y);
        for (uint32_t px = 0; px < w; ++px) {
            auto sample = [&](int k) -> int {                     // k-th sample of pixel px as an 8-bit value (16 bit: high byte)
                if (depth == 8) return x[(size_t)px * samples + k];
                if (depth == 16) return x[((size_t)px * samples + k) * 2];
                const size_t bit = (size_t)px * depth;            // 1 / 2 / 4 bits: one sample per pixel
                return (x[bit >> 3] >> (8 - depth - (bit & 7))) & maxv;
            };
            if (ctype == 3) {
                const size_t idx = (size_t)sample(0) * 3;
                if (idx + 2 >= plte.size()) { o[3 * px] = o[3 * px + 1] = o[3 * px + 2] = 0; continue; }      // index beyond the palette
                o[3 * px] = plte[idx + 2]; o[3 * px + 1] = plte[idx + 1]; o[3 * px + 2] = plte[idx];
            } else if (colour) {
                o[3 * px] = (uint8_t)sample(2); o[3 * px + 1] = (uint8_t)sample(1); o[3 * px + 2] = (uint8_t)sample(0);
            } else {
                const int v = sample(0);
                o[px] = (uint8_t)(depth < 8 ? v * 255 / maxv : v);
            }
        }
    }
    return img;
}

// by content: PNG signature or Netpbm magic (cv::imread decides by content too)
cv::Mat read_image(const std::string& path) {
    FILE* in = std::fopen(path.c_str(), "rb");
    if (!in) return cv::Mat();
    uint8_t m[2] = {0, 0};
    const size_t got = std::fread(m, 1, 2, in);
    std::fclose(in);
    if (got != 2) return cv::Mat();
    if (m[0] == 0x89 && m[1] == 'P') return read_png(path);
    if (m[0] == 'P' && (m[1] == '5' || m[1] == '6')) return read_pnm(path);
    return cv::Mat();                                             // jpeg and the rest: no codec in this build
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

// TUM trajectory format, one line per frame: "timestamp tx ty tz qx qy qz qw" of T_wc (camera -> world, as Pose) -- what
// evaluation tools (evo, the TUM / KITTI scripts) read; the reference itself only shows the trajectory in its viewer
// (src/trajectory_viewer.cpp).  Same quaternion branch rule as b200vo/trajectory.py::save_tum.
bool save_trajectory_tum(const std::string& path, const std::vector<double>& timestamps, const std::vector<Pose>& poses) {
    FILE* f = std::fopen(path.c_str(), "w");
    if (!f) return false;
    std::fprintf(f, "# timestamp tx ty tz qx qy qz qw   (T_wc: camera -> world, as Pose in include/pose.h)\n");
    for (size_t n = 0; n < poses.size(); ++n) {
        const Pose& p = poses[n];
        double R[3][3];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) R[r][c] = p.R_wc.at<double>(r, c);
        double q[4];
        const double tr = R[0][0] + R[1][1] + R[2][2];
        if (tr > 0) {
            const double s = std::sqrt(tr + 1.0) * 2;
            q[0] = (R[2][1] - R[1][2]) / s; q[1] = (R[0][2] - R[2][0]) / s; q[2] = (R[1][0] - R[0][1]) / s; q[3] = 0.25 * s;
        } else {
            int i = 0;
            if (R[1][1] > R[i][i]) i = 1;
            if (R[2][2] > R[i][i]) i = 2;
            const int j = (i + 1) % 3, k = (i + 2) % 3;
            const double s = std::sqrt(1.0 + R[i][i] - R[j][j] - R[k][k]) * 2;
            q[i] = 0.25 * s; q[j] = (R[j][i] + R[i][j]) / s; q[k] = (R[k][i] + R[i][k]) / s; q[3] = (R[k][j] - R[j][k]) / s;
        }
        const double nrm = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
        std::fprintf(f, "%.9f %.9f %.9f %.9f %.9f %.9f %.9f %.9f\n", n < timestamps.size() ? timestamps[n] : (double)n,
                     p.t_wc.at<double>(0, 0), p.t_wc.at<double>(1, 0), p.t_wc.at<double>(2, 0), q[0] / nrm, q[1] / nrm, q[2] / nrm, q[3] / nrm);
    }
    return std::fclose(f) == 0;
}

// ---- accuracy report against a reference trajectory (b200vo/trajectory.py::accuracy_report in C++) ----------------------
// The reference applies a FIXED translation scale of 0.3 per successful step (src/visual_odometry.cpp:352), so positions are
// compared after a similarity alignment (Horn's closed form: the rotation is the dominant eigenvector of a symmetric 4x4
// matrix built from the cross-covariance, found here by Jacobi rotations), and relative poses by rotation angle and by
// the angle between the translation DIRECTIONS.
namespace {
void jacobi_eigen4(double A[4][4], double V[4][4]) {
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = 0.0;
        for (int i = 0; i < 4; ++i) for (int j = i + 1; j < 4; ++j) off += A[i][j] * A[i][j];
        if (off < 1e-30) break;
        for (int p = 0; p < 4; ++p)
            for (int q = p + 1; q < 4; ++q) {
                if (std::fabs(A[p][q]) < 1e-300) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 4; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - sn * akq; A[k][q] = sn * akp + c * akq; }
                for (int k = 0; k < 4; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - sn * aqk; A[q][k] = sn * apk + c * aqk; }
                for (int k = 0; k < 4; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - sn * vkq; V[k][q] = sn * vkp + c * vkq; }
            }
    }
}
double angle_deg(double c) { return std::acos(std::max(-1.0, std::min(1.0, c))) * 180.0 / 3.14159265358979323846; }
}  // namespace

bool trajectory_report(const std::vector<Pose>& est, const std::vector<Pose>& gt, TrajectoryReport* out) {
    const size_t n = std::min(est.size(), gt.size());
    if (!out || n < 2) return false;
    std::vector<double> pe(3 * n), pg(3 * n);
    double me[3] = {0, 0, 0}, mg[3] = {0, 0, 0};
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            pe[3 * i + k] = est[i].t_wc.at<double>(k, 0); pg[3 * i + k] = gt[i].t_wc.at<double>(k, 0);
            me[k] += pe[3 * i + k] / (double)n; mg[k] += pg[3 * i + k] / (double)n;
        }
    // cross-covariance S = sum (est - me)(gt - mg)^T, Horn's 4x4 matrix, rotation = eigenvector of the largest eigenvalue
    double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, var = 0.0;
    for (size_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            const double ea = pe[3 * i + a] - me[a];
            var += ea * ea;
            for (int b = 0; b < 3; ++b) S[a][b] += ea * (pg[3 * i + b] - mg[b]);
        }
    double N[4][4] = {{S[0][0] + S[1][1] + S[2][2], S[1][2] - S[2][1], S[2][0] - S[0][2], S[0][1] - S[1][0]},
                      {S[1][2] - S[2][1], S[0][0] - S[1][1] - S[2][2], S[0][1] + S[1][0], S[2][0] + S[0][2]},
                      {S[2][0] - S[0][2], S[0][1] + S[1][0], -S[0][0] + S[1][1] - S[2][2], S[1][2] + S[2][1]},
                      {S[0][1] - S[1][0], S[2][0] + S[0][2], S[1][2] + S[2][1], -S[0][0] - S[1][1] + S[2][2]}};
    double V[4][4];
    jacobi_eigen4(N, V);
    int best = 0;
    for (int k = 1; k < 4; ++k) if (N[k][k] > N[best][best]) best = k;
    const double w = V[0][best], x = V[1][best], y = V[2][best], z = V[3][best];
    const double R[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)},
                            {2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)},
                            {2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)}};
    // scale that minimises the residual for this rotation: sum (gt_c . R est_c) / sum |est_c|^2
    double num = 0.0;
    for (size_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            double re = 0.0;
            for (int b = 0; b < 3; ++b) re += R[a][b] * (pe[3 * i + b] - me[b]);
            num += re * (pg[3 * i + a] - mg[a]);
        }
    const double sc = var > 0 ? num / var : 1.0;
    double sq = 0.0;
    for (size_t i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) {
            double re = 0.0;
            for (int b = 0; b < 3; ++b) re += R[a][b] * (pe[3 * i + b] - me[b]);
            const double d = (pg[3 * i + a] - mg[a]) - sc * re;
            sq += d * d;
        }
    out->frames = (int)n;
    out->ate_rmse_aligned = std::sqrt(sq / (double)n);
    out->alignment_scale = sc;
    // relative pose errors between consecutive frames
    double rot = 0.0, dir = 0.0;
    int ndir = 0;
    auto Rm = [](const Pose& p, int r, int c) { return p.R_wc.at<double>(r, c); };
    for (size_t i = 0; i + 1 < n; ++i) {
        // trace(dRe^T dRg) with dR = R_i^T R_{i+1}
        double dRe[3][3], dRg[3][3];
        for (int a = 0; a < 3; ++a)
            for (int b = 0; b < 3; ++b) {
                dRe[a][b] = dRg[a][b] = 0.0;
                for (int k = 0; k < 3; ++k) { dRe[a][b] += Rm(est[i], k, a) * Rm(est[i + 1], k, b); dRg[a][b] += Rm(gt[i], k, a) * Rm(gt[i + 1], k, b); }
            }
        double tr = 0.0;
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) tr += dRe[a][b] * dRg[a][b];
        rot += angle_deg((tr - 1.0) / 2.0);
        double te[3] = {0, 0, 0}, tg[3] = {0, 0, 0};
        for (int a = 0; a < 3; ++a)
            for (int k = 0; k < 3; ++k) {
                te[a] += Rm(est[i], k, a) * (pe[3 * (i + 1) + k] - pe[3 * i + k]);
                tg[a] += Rm(gt[i], k, a) * (pg[3 * (i + 1) + k] - pg[3 * i + k]);
            }
        const double ne = std::sqrt(te[0] * te[0] + te[1] * te[1] + te[2] * te[2]), ng = std::sqrt(tg[0] * tg[0] + tg[1] * tg[1] + tg[2] * tg[2]);
        if (ne > 1e-12 && ng > 1e-12) { dir += angle_deg((te[0] * tg[0] + te[1] * tg[1] + te[2] * tg[2]) / (ne * ng)); ++ndir; }
    }
    out->rpe_rot_deg_mean = rot / (double)(n - 1);
    out->rpe_trans_dir_deg_mean = ndir ? dir / ndir : 0.0;
    return true;
}

// the inverse of save_trajectory_tum: lines "timestamp tx ty tz qx qy qz qw", '#' comments and blank lines skipped
bool load_trajectory_tum(const std::string& path, std::vector<double>* timestamps, std::vector<Pose>* poses) {
    FILE* f = std::fopen(path.c_str(), "r");
    if (!f || !poses) { if (f) std::fclose(f); return false; }
    char line[512];
    bool ok = true;
    while (std::fgets(line, sizeof(line), f)) {
        const char* p = line;
        while (*p == ' ' || *p == '\t') ++p;
        if (*p == '#' || *p == '\n' || *p == '\r' || *p == 0) continue;
        double v[8];
        if (std::sscanf(p, "%lf %lf %lf %lf %lf %lf %lf %lf", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]) != 8) { ok = false; break; }
        const double nq = std::sqrt(v[4] * v[4] + v[5] * v[5] + v[6] * v[6] + v[7] * v[7]);
        if (!(nq > 0)) { ok = false; break; }
        const double x = v[4] / nq, y = v[5] / nq, z = v[6] / nq, w = v[7] / nq;
        cv::Mat R(3, 3, CV_64F), t(3, 1, CV_64F);
        const double r[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                             2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                             2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
        for (int i = 0; i < 9; ++i) R.at<double>(i / 3, i % 3) = r[i];
        for (int i = 0; i < 3; ++i) t.at<double>(i, 0) = v[1 + i];
        poses->emplace_back(R, t);
        if (timestamps) timestamps->push_back(v[0]);
    }
    std::fclose(f);
    return ok;
}
