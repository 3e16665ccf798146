// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Minimal stand-in for the slice of the OpenCV C++ API that the reference
// (better-flow) touches, so that the reference's OWN, UNMODIFIED sources under
// /root/reference can be compiled into oracle/_ref/ in an image that ships no
// OpenCV C++ headers.  On the motion-compensation hot path the reference uses
// cv:: purely as a 2-D float container and a 2-vector (accel_lib.h:148-175,
// 522-543; object_model.cpp:14-33,111-120; event.h:100-105): those pieces are
// real here (Mat, Point_).  Every drawing / GUI / filtering entry point is a
// no-op stub that exists only so the visualisation code still parses and links.
//
// Nothing in here is copied from OpenCV; the container is a plain ref-counted
// row-major buffer.
#pragma once

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <climits>
#include <cmath>
#include <memory>
#include <ostream>
#include <string>
#include <vector>

#define CV_MAJOR_VERSION 3

typedef unsigned char uchar;
typedef unsigned short ushort;

// type codes: depth in the low 3 bits, (channels-1) above (same convention as
// the real library so that CV_32FC1 == CV_32F).
#define CV_8U 0
#define CV_16S 3
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_RGB(r, g, b) cv::Scalar((b), (g), (r), 0)
#define CV_FOURCC(a, b, c, d) 0
#define CV_GRAY2RGB 8
#define CV_HSV2BGR 54
#define CV_PUSH_BUTTON 0

namespace cv {

enum {
    WINDOW_NORMAL = 0, WINDOW_AUTOSIZE = 1, NORM_MINMAX = 32, LINE_AA = 16,
    FONT_HERSHEY_DUPLEX = 2, COLOR_HSV2BGR = 54, THRESH_BINARY = 0,
    ADAPTIVE_THRESH_GAUSSIAN_C = 1, BORDER_DEFAULT = 4, BORDER_CONSTANT = 0
};

template <class T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <class U> Point_(const Point_<U> &o) : x(T(o.x)), y(T(o.y)) {}
    double cross(const Point_ &o) const { return double(x) * o.y - double(y) * o.x; }
    double ddot(const Point_ &o) const { return double(x) * o.x + double(y) * o.y; }
};
template <class T> inline Point_<T> operator+(const Point_<T> &a, const Point_<T> &b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <class T> inline Point_<T> operator-(const Point_<T> &a, const Point_<T> &b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <class T> inline Point_<T> operator-(const Point_<T> &a) { return Point_<T>(-a.x, -a.y); }
template <class T> inline Point_<T> operator*(const Point_<T> &a, double k) { return Point_<T>(T(a.x * k), T(a.y * k)); }
template <class T> inline Point_<T> operator*(double k, const Point_<T> &a) { return Point_<T>(T(a.x * k), T(a.y * k)); }
template <class T> inline Point_<T> operator/(const Point_<T> &a, double k) { return Point_<T>(T(a.x / k), T(a.y / k)); }
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

struct Size {
    int width, height;
    Size() : width(0), height(0) {}
    Size(int w, int h) : width(w), height(h) {}
    bool operator==(const Size &o) const { return width == o.width && height == o.height; }
    bool operator!=(const Size &o) const { return !(*this == o); }
};
inline std::ostream &operator<<(std::ostream &os, const Size &s) { return os << "[" << s.width << " x " << s.height << "]"; }

struct Rect {
    int x, y, width, height;
    Rect() : x(0), y(0), width(0), height(0) {}
    Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
    double &operator[](int i) { return val[i]; }
    const double &operator[](int i) const { return val[i]; }
};

template <class T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; ++i) val[i] = T(0); }
    Vec(T a, T b, T c) { static_assert(N == 3, "3-vector ctor"); val[0] = a; val[1] = b; val[2] = c; }
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
};
typedef Vec<unsigned char, 3> Vec3b;
typedef Vec<float, 3> Vec3f;

// Ref-counted, row-major, zero-initialised buffer with OpenCV's (rows, cols)
// convention.  Copies are shallow, as in the real library.
class Mat {
public:
    int rows, cols;
    unsigned char *data;

    Mat() : rows(0), cols(0), data(nullptr), type_(0), esz_(1) {}
    Mat(int r, int c, int type) { create(r, c, type, nullptr); }
    Mat(int r, int c, int type, const Scalar &s) { create(r, c, type, &s); }
    Mat(Size sz, int type) { create(sz.height, sz.width, type, nullptr); }
    Mat(Size sz, int type, const Scalar &s) { create(sz.height, sz.width, type, &s); }

    static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
    static Mat zeros(Size sz, int type) { return Mat(sz, type); }

    Size size() const { return Size(cols, rows); }
    int type() const { return type_; }
    int channels() const { return (type_ >> 3) + 1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t total() const { return size_t(rows) * size_t(cols); }
    size_t elemSize() const { return esz_; }

    template <class T> T &at(int r, int c) { return *reinterpret_cast<T *>(data + (size_t(r) * cols + c) * esz_); }
    template <class T> const T &at(int r, int c) const { return *reinterpret_cast<const T *>(data + (size_t(r) * cols + c) * esz_); }
    template <class T> T &at(int i) { return *reinterpret_cast<T *>(data + size_t(i) * esz_); }
    template <class T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + size_t(r) * cols * esz_); }
    template <class T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + size_t(r) * cols * esz_); }

    Mat clone() const {
        Mat m(rows, cols, type_);
        if (data) std::memcpy(m.data, data, total() * esz_);
        return m;
    }
    void copyTo(Mat &dst) const { dst = clone(); }
    void convertTo(Mat &dst, int, double = 1, double = 0) const { dst = clone(); }
    Mat operator()(const Rect &) const { return *this; }
    Mat &operator=(const Scalar &) { return *this; }
    Mat &setTo(const Scalar &) { return *this; }

private:
    void create(int r, int c, int type, const Scalar *fill) {
        rows = r; cols = c; type_ = type;
        const int depth = type & 7;
        const size_t bytes = (depth == CV_8U) ? 1 : (depth == CV_16S) ? 2 : 4;
        esz_ = bytes * size_t(channels());
        const size_t n = size_t(r > 0 ? r : 0) * size_t(c > 0 ? c : 0) * esz_;
        store_.reset(new unsigned char[n ? n : 1](), std::default_delete<unsigned char[]>());
        data = store_.get();
        if (fill) {
            const int cn = channels();
            for (size_t i = 0; i < size_t(rows) * cols; ++i)
                for (int k = 0; k < cn; ++k) {
                    if (depth == CV_8U) data[i * esz_ + k] = (unsigned char)fill->val[k];
                    else if (depth == CV_32F) reinterpret_cast<float *>(data + i * esz_)[k] = float(fill->val[k]);
                }
        }
    }
    int type_;
    size_t esz_;
    std::shared_ptr<unsigned char> store_;
};

template <class T> class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(const Mat &m) : Mat(m) {}
};

// Arithmetic on whole images is never reached by OptimizerRolling::run();
// these exist so the visualisation helpers compile.
inline Mat operator-(const Mat &a) { return a; }
inline Mat operator-(const Mat &a, const Mat &) { return a; }
inline Mat operator+(const Mat &a, const Mat &) { return a; }
inline Mat operator*(const Mat &a, double) { return a; }
inline Mat operator*(double, const Mat &a) { return a; }
inline Mat operator/(const Mat &a, double) { return a; }
inline Mat operator/(const Mat &a, const Mat &) { return a; }
inline Mat operator+(const Mat &a, const Scalar &) { return a; }
inline Mat operator-(const Mat &a, const Scalar &) { return a; }
inline Mat &operator+=(Mat &a, const Mat &) { return a; }
inline Mat &operator-=(Mat &a, const Mat &) { return a; }
inline Mat &operator*=(Mat &a, double) { return a; }
inline Mat &operator/=(Mat &a, double) { return a; }
inline Mat &operator+=(Mat &a, const Scalar &) { return a; }
inline Mat abs(const Mat &a) { return a; }

// ---- no-op stubs (GUI, drawing, filtering, codecs) ------------------------
inline void addWeighted(const Mat &, double, const Mat &, double, double, Mat &, int = -1) {}
inline void normalize(const Mat &, Mat &, double = 1, double = 0, int = 0, int = -1) {}
inline void convertScaleAbs(const Mat &, Mat &, double = 1, double = 0) {}
inline void hconcat(const Mat &, const Mat &, Mat &) {}
inline void vconcat(const Mat &, const Mat &, Mat &) {}
inline void transpose(const Mat &, Mat &) {}
inline void cvtColor(const Mat &, Mat &, int, int = 0) {}
inline void resize(const Mat &, Mat &, Size, double = 0, double = 0, int = 1) {}
// The one filtering call that carries arithmetic on a path we check: OptimizerLocal::iteration_step
// blurs its CV_8UC1 event-count image with GaussianBlur(img, img, Size(scale, scale), 0, 0)
// (optimizer_sampler.cpp:147-149).  OpenCV's 8-bit path for ksize 3 / 5 with sigma <= 0 uses the
// binomial kernels [1 2 1]/4 and [1 4 6 4 1]/16 in fixed point, i.e. the exactly computed weighted
// sum rounded half up, with BORDER_REFLECT_101 (the default).  tests/test_oracle_local.py checks
// this stand-in against fixtures produced by the real cv2.GaussianBlur (oracle/make_golden_local.py).
inline unsigned long long &bf_shim_blur_calls() {
    static unsigned long long n = 0;
    return n;
}
inline void GaussianBlur(const Mat &src, Mat &dst, Size ksize, double, double = 0, int = BORDER_DEFAULT) {
    ++bf_shim_blur_calls();
    if (src.type() != CV_8UC1 || ksize.width != ksize.height || (ksize.width != 1 && ksize.width != 3 && ksize.width != 5)) {
        std::fprintf(stderr, "cv shim: GaussianBlur supports CV_8UC1 with ksize 1, 3 or 5 only\n");
        std::abort();
    }
    const int k = ksize.width, r = k / 2, R = src.rows, C = src.cols;
    static const int w3[3] = {1, 2, 1}, w5[5] = {1, 4, 6, 4, 1}, w1[1] = {1};
    const int *w = k == 1 ? w1 : k == 3 ? w3 : w5;
    const int shift = k == 1 ? 0 : k == 3 ? 4 : 8;
    auto refl = [](int i, int n) { if (n == 1) return 0; while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i; return i; };
    std::vector<int> tmp(size_t(R) * C);
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j) {
            int s = 0;
            for (int d = -r; d <= r; ++d) s += w[d + r] * int(src.data[size_t(i) * C + refl(j + d, C)]);
            tmp[size_t(i) * C + j] = s;
        }
    Mat out(R, C, CV_8UC1);
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j) {
            int s = 0;
            for (int d = -r; d <= r; ++d) s += w[d + r] * tmp[size_t(refl(i + d, R)) * C + j];
            out.data[size_t(i) * C + j] = (unsigned char)((s + ((1 << shift) >> 1)) >> shift);
        }
    dst = out;
}
inline void Sobel(const Mat &, Mat &, int, int, int, int = 3, double = 1, double = 0, int = BORDER_DEFAULT) {}
inline void Scharr(const Mat &, Mat &, int, int, int, double = 1, double = 0, int = BORDER_DEFAULT) {}
inline void Laplacian(const Mat &, Mat &, int, int = 1, double = 1, double = 0, int = BORDER_DEFAULT) {}
inline void adaptiveThreshold(const Mat &, Mat &, double, int, int, int, double) {}
inline void line(Mat &, Point, Point, const Scalar &, int = 1, int = 8, int = 0) {}
inline void arrowedLine(Mat &, Point, Point, const Scalar &, int = 1, int = 8, int = 0, double = 0.1) {}
inline void putText(Mat &, const std::string &, Point, int, double, Scalar, int = 1, int = 8, bool = false) {}
inline void namedWindow(const std::string &, int = WINDOW_AUTOSIZE) {}
inline void imshow(const std::string &, const Mat &) {}
inline int waitKey(int = 0) { return 32; }
typedef void (*TrackbarCallback)(int, void *);
typedef void (*ButtonCallback)(int, void *);
inline int createTrackbar(const std::string &, const std::string &, int *, int, TrackbarCallback = nullptr, void * = nullptr) { return 0; }
inline void setTrackbarPos(const std::string &, const std::string &, int) {}
inline void displayStatusBar(const std::string &, const std::string &, int = 0) {}
inline int createButton(const std::string &, ButtonCallback, void * = nullptr, int = 0, bool = false) { return 0; }
inline bool imwrite(const std::string &, const Mat &) { return true; }

class VideoWriter {
public:
    VideoWriter() {}
    VideoWriter(const std::string &, int, double, Size, bool = true) {}
    static int fourcc(char, char, char, char) { return 0; }
    bool isOpened() const { return true; }
    VideoWriter &operator<<(const Mat &) { return *this; }
};

}  // namespace cv
