// Minimal stand-ins for the OpenCV value types that appear in the reference's public headers
// (cv::Mat, cv::KeyPoint, cv::DMatch, cv::Matx33d, cv::Point2f -- reference src/frame.h:21-27,
// src/pose.h:31-77, src/visual_odometry.h:33-66).  When real OpenCV headers are installed they are
// used instead, so the API is source-compatible with the reference; this container has no OpenCV
// C++ development package, hence the fallback.  Only what the VO API needs
// is provided: small dense double matrices, u8 images, and the keypoint / match records.
#pragma once

#if defined(B200VO_USE_OPENCV) || (defined(__has_include) && __has_include(<opencv2/core.hpp>) && !defined(B200VO_NO_OPENCV))
#include <opencv2/core.hpp>
#define B200VO_HAVE_OPENCV 1
#else
#define B200VO_HAVE_OPENCV 0

#include <cassert>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_Assert(expr) do { if (!(expr)) throw std::runtime_error("CV_Assert failed: " #expr); } while (0)

namespace cv {

struct Point2f { float x = 0, y = 0; Point2f() = default; Point2f(float a, float b) : x(a), y(b) {} };

struct KeyPoint {
    Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
};

struct DMatch {
    int queryIdx = -1, trainIdx = -1, imgIdx = -1; float distance = 0;
};

class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    uint8_t* data = nullptr;

    Mat() = default;
    Mat(int r, int c, int type) { create(r, c, type); }
    // wraps caller memory (no ownership), like cv::Mat(rows, cols, type, ptr, step)
    Mat(int r, int c, int type, void* ptr, size_t stp = 0) : rows(r), cols(c), type_(type) {
        step = stp ? stp : (size_t)c * elemSize();
        data = static_cast<uint8_t*>(ptr);
    }
    void create(int r, int c, int type) {
        rows = r; cols = c; type_ = type;
        step = (size_t)c * elemSize();
        buf_ = std::make_shared<std::vector<uint8_t>>((size_t)r * step, 0);
        data = buf_->data();
    }
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat eye(int r, int c, int type) {
        Mat m(r, c, type);
        CV_Assert(type == CV_64F);
        for (int i = 0; i < (r < c ? r : c); ++i) m.at<double>(i, i) = 1.0;
        return m;
    }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    size_t elemSize() const { const int d = depth(); return (size_t)channels() * (d == CV_64F ? 8 : d == CV_32F ? 4 : 1); }
    size_t total() const { return (size_t)rows * cols; }
    template <typename T> T& at(int r, int c) { return *reinterpret_cast<T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> const T& at(int r, int c) const { return *reinterpret_cast<const T*>(data + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <typename T> T& at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
    template <typename T> const T& at(int i) const { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
    template <typename T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
    template <typename T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
    Mat clone() const {
        Mat m;
        if (empty()) return m;
        m.rows = rows; m.cols = cols; m.type_ = type_;
        m.step = (size_t)cols * elemSize();
        if (step == m.step) {                       // contiguous source: one copy, no zero fill first
            m.buf_ = std::make_shared<std::vector<uint8_t>>(data, data + (size_t)rows * step);
            m.data = m.buf_->data();
            return m;
        }
        m.create(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat& dst) const { dst = clone(); }
    void convertTo(Mat& dst, int type) const {
        CV_Assert(channels() == 1);
        Mat m(rows, cols, type);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                const double v = depth() == CV_64F ? at<double>(r, c) : depth() == CV_32F ? (double)at<float>(r, c) : (double)at<uint8_t>(r, c);
                if (type == CV_64F) m.at<double>(r, c) = v;
                else if (type == CV_32F) m.at<float>(r, c) = (float)v;
                else m.at<uint8_t>(r, c) = (uint8_t)v;
            }
        dst = m;
    }
    Mat t() const {
        CV_Assert(type_ == CV_64F);
        Mat m(cols, rows, CV_64F);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) m.at<double>(c, r) = at<double>(r, c);
        return m;
    }

private:
    int type_ = 0;
    std::shared_ptr<std::vector<uint8_t>> buf_;
};

inline Mat operator*(const Mat& a, const Mat& b) {
    CV_Assert(a.type() == CV_64F && b.type() == CV_64F && a.cols == b.rows);
    Mat m(a.rows, b.cols, CV_64F);
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < b.cols; ++j) {
            double s = 0;
            for (int k = 0; k < a.cols; ++k) s += a.at<double>(i, k) * b.at<double>(k, j);
            m.at<double>(i, j) = s;
        }
    return m;
}
inline Mat operator*(double s, const Mat& a) {
    Mat m = a.clone();
    for (int i = 0; i < m.rows; ++i)
        for (int j = 0; j < m.cols; ++j) m.at<double>(i, j) *= s;
    return m;
}
inline Mat operator+(const Mat& a, const Mat& b) {
    CV_Assert(a.rows == b.rows && a.cols == b.cols);
    Mat m = a.clone();
    for (int i = 0; i < m.rows; ++i)
        for (int j = 0; j < m.cols; ++j) m.at<double>(i, j) += b.at<double>(i, j);
    return m;
}
inline Mat operator-(const Mat& a) { return -1.0 * a; }

template <typename T, int M, int N>
struct Matx {
    T val[M * N] = {};
    T& operator()(int i, int j) { return val[i * N + j]; }
    const T& operator()(int i, int j) const { return val[i * N + j]; }
};
using Matx33d = Matx<double, 3, 3>;

}  // namespace cv
#endif
