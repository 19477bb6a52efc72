// TEST INFRASTRUCTURE - stand-in for the OpenCV headers, written for this repository so that the
// reference's own sources compile in place into oracle/_ref/libref_core*.so (oracle/Makefile).
// On the association hot path the reference touches OpenCV only through the `const cv::Mat& img`
// parameter of BaseTracker::update (empty() / rows / cols, src/tracker.cpp:110-124,166-172); camera-
// motion compensation, ReID inference and rotated-box IoU are outside the path and
// their OpenCV calls are declared here only so that the headers parse - calling them throws.
#pragma once
#include <cstdint>
#include <string>
#include <stdexcept>
#include <vector>

typedef unsigned char uchar;

namespace cv {

[[noreturn]] inline void shim_fail(const char* what) { throw std::logic_error(std::string("ref_shim OpenCV: ") + what); }

enum { CV_8UC1 = 0, CV_8UC3 = 16, CV_32F = 5, CV_32FC1 = 5, CV_32FC3 = 21, CV_64F = 6 };
enum { MOTION_TRANSLATION = 0, MOTION_EUCLIDEAN = 1, MOTION_AFFINE = 2, MOTION_HOMOGRAPHY = 3 };
enum { COLOR_BGR2GRAY = 6, COLOR_BGR2RGB = 4 };
enum { INTER_LINEAR = 1 };
enum { INTERSECT_NONE = 0, INTERSECT_PARTIAL = 1, INTERSECT_FULL = 2 };

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
using Point = Point_<int>;
using Point2f = Point_<float>;
using Point2d = Point_<double>;

template <class T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
using Size = Size_<int>;
using Size2f = Size_<float>;

template <class T> struct Rect_ {
    T x, y, width, height;
    Rect_() : x(0), y(0), width(0), height(0) {}
    Rect_(T x_, T y_, T w, T h) : x(x_), y(y_), width(w), height(h) {}
};
using Rect = Rect_<int>;

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
    double operator[](int i) const { return val[i]; }
};

struct TermCriteria {
    enum { COUNT = 1, MAX_ITER = 1, EPS = 2 };
    int type, maxCount;
    double epsilon;
    TermCriteria() : type(0), maxCount(0), epsilon(0) {}
    TermCriteria(int t, int c, double e) : type(t), maxCount(c), epsilon(e) {}
};

struct RotatedRect {
    Point2f center;
    Size2f size;
    float angle;
    RotatedRect() : angle(0) {}
    RotatedRect(const Point2f& c, const Size2f& s, float a) : center(c), size(s), angle(a) {}
};

// header-only image handle: shape only, no pixels
class Mat {
public:
    int rows, cols;
    Mat() : rows(0), cols(0), type_(0) {}
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type) {}
    Mat(int r, int c, int type, const Scalar&) : rows(r), cols(c), type_(type) {}
    Mat(Size s, int type) : rows(s.height), cols(s.width), type_(type) {}
    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat eye(int r, int c, int type) { return Mat(r, c, type); }
    bool empty() const { return rows == 0 || cols == 0; }
    Mat clone() const { return *this; }
    int channels() const { return (type_ >> 3) + 1; }
    int type() const { return type_; }
    Size size() const { return Size(cols, rows); }
    template <class T> T& at(int, int) { shim_fail("Mat::at - the stand-in holds no pixels"); }
    template <class T> const T& at(int, int) const { shim_fail("Mat::at - the stand-in holds no pixels"); }
    void copyTo(Mat& o) const { o = *this; }
    void convertTo(Mat& o, int type, double = 1.0, double = 0.0) const { o = *this; o.type_ = type; }
    Mat operator()(const Rect&) const { shim_fail("Mat ROI - the stand-in holds no pixels"); }
private:
    int type_;
};
using InputArray = const Mat&;
using OutputArray = Mat&;

inline int rotatedRectangleIntersection(const RotatedRect&, const RotatedRect&, std::vector<Point2f>&) {
    shim_fail("rotatedRectangleIntersection (oriented boxes are outside the hot path)");
}
inline double contourArea(const std::vector<Point2f>&) { shim_fail("contourArea"); }
inline void cvtColor(const Mat&, Mat&, int) { shim_fail("cvtColor"); }
inline void resize(const Mat&, Mat&, Size, double = 0, double = 0, int = INTER_LINEAR) { shim_fail("resize"); }
inline double findTransformECC(const Mat&, const Mat&, Mat&, int, TermCriteria, const Mat& = Mat(), int = 5) { shim_fail("findTransformECC"); }

}  // namespace cv
