// better_flow/image.h -- minimal row-major float image.  Stands in for the cv::Mat objects that the
// reference's public API hands around (AccelLib::get_time_img returns one, ObjectModel::update and
// AccelLib::Sobel take them); only the members the hot path used are provided.
#ifndef BF_IMAGE_H
#define BF_IMAGE_H

#include <cstddef>
#include <vector>

template <class T> class Image2D {
public:
    int rows, cols;

    Image2D() : rows(0), cols(0) {}
    Image2D(int r, int c) : rows(r), cols(c), buf_(size_t(r > 0 ? r : 0) * size_t(c > 0 ? c : 0), T(0)) {}

    static Image2D zeros(int r, int c) { return Image2D(r, c); }

    bool empty() const { return buf_.empty(); }
    size_t total() const { return buf_.size(); }
    T *data() { return buf_.data(); }
    const T *data() const { return buf_.data(); }
    T *ptr(int r) { return buf_.data() + size_t(r) * cols; }
    const T *ptr(int r) const { return buf_.data() + size_t(r) * cols; }
    T &at(int r, int c) { return buf_[size_t(r) * cols + c]; }
    const T &at(int r, int c) const { return buf_[size_t(r) * cols + c]; }

private:
    std::vector<T> buf_;
};

typedef Image2D<float> ImageF;

#endif  // BF_IMAGE_H
