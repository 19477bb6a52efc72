// compat.hpp - the two third-party types in motcpp's public signature, Eigen::MatrixXf and cv::Mat.
// When the real headers are available they are used; otherwise (this image ships neither) tiny
// API-surface stand-ins with the SAME memory layout (column-major float) are defined, so the
// binding below compiles and is testable anywhere.  Reference: include/motcpp/tracker.hpp:67-69.
#pragma once
#include <cstddef>
#include <vector>

#if __has_include(<Eigen/Dense>) && !defined(MOTB200_NO_EIGEN)
#include <Eigen/Dense>
#else
namespace Eigen {
class MatrixXf {
public:
    MatrixXf() : r_(0), c_(0) {}
    MatrixXf(long r, long c) : r_(r), c_(c), d_(static_cast<size_t>(r * c), 0.0f) {}
    long rows() const { return r_; }
    long cols() const { return c_; }
    long size() const { return r_ * c_; }
    float& operator()(long i, long j) { return d_[static_cast<size_t>(j * r_ + i)]; }      // column-major
    float operator()(long i, long j) const { return d_[static_cast<size_t>(j * r_ + i)]; }
    const float* data() const { return d_.data(); }
    float* data() { return d_.data(); }
private:
    long r_, c_;
    std::vector<float> d_;
};
}  // namespace Eigen
#endif

#if __has_include(<opencv2/core.hpp>) && !defined(MOTB200_NO_OPENCV)
#include <opencv2/core.hpp>
#else
namespace cv {
// the hot path only asks a frame for empty()/rows/cols (src/tracker.cpp:114,166-171)
struct Mat {
    int rows = 0, cols = 0;
    Mat() = default;
    Mat(int r, int c) : rows(r), cols(c) {}
    bool empty() const { return rows == 0 || cols == 0; }
};
}  // namespace cv
#endif
