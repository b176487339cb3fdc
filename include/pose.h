// Pose: rigid camera pose stored camera->world (T_wc).  Same public members and methods as the
// reference's src/pose.h:26-78 (R_wc, t_wc, set_R_t, T_wc, R_cw, t_cw, T_cw, C_w).
#pragma once
#include "vo_cv_compat.h"

class Pose {
public:
    cv::Mat R_wc;   // 3x3 CV_64F:  x_w = R_wc x_c + t_wc
    cv::Mat t_wc;   // 3x1 CV_64F:  camera centre in world coordinates

    Pose() : R_wc(cv::Mat::eye(3, 3, CV_64F)), t_wc(cv::Mat::zeros(3, 1, CV_64F)) {}
    Pose(const cv::Mat& R, const cv::Mat& t) { set_R_t(R, t); }

    // Accepts any depth for R (3x3) and t (3x1 or 1x3); stores doubles, t as a column.
    void set_R_t(const cv::Mat& R, const cv::Mat& t) {
        CV_Assert(R.rows == 3 && R.cols == 3);
        CV_Assert((t.rows == 3 && t.cols == 1) || (t.rows == 1 && t.cols == 3));
        R.convertTo(R_wc, CV_64F);
        cv::Mat col(3, 1, CV_64F);
        cv::Mat t64;
        t.convertTo(t64, CV_64F);
        for (int i = 0; i < 3; ++i) col.at<double>(i, 0) = (t.rows == 3) ? t64.at<double>(i, 0) : t64.at<double>(0, i);
        t_wc = col;
    }

    cv::Mat T_wc() const { return assemble(R_wc, t_wc); }
    cv::Mat R_cw() const { return R_wc.t(); }
    cv::Mat t_cw() const { return -(R_wc.t() * t_wc); }
    cv::Mat T_cw() const { return assemble(R_cw(), t_cw()); }
    const cv::Mat& C_w() const { return t_wc; }

private:
    static cv::Mat assemble(const cv::Mat& R, const cv::Mat& t) {
        cv::Mat T = cv::Mat::eye(4, 4, CV_64F);
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) T.at<double>(i, j) = R.at<double>(i, j);
            T.at<double>(i, 3) = t.at<double>(i, 0);
        }
        return T;
    }
};
