// better_flow/datastructures.h -- slice containers with the reference's semantics
// (reference: better_flow_core/include/better_flow/datastructures.h).
//
// CircularArray<DType, SZ, SPAN>: ring buffer holding at most SZ elements no older than SPAN
// (same units as DType::operator-) relative to the newest one.  The behaviours callers can observe
// are kept, including the quirks (datastructures.h:31-76):
//   * the first push lands in physical slot 1, not 0;
//   * eviction of too-old elements is lazy (done by size()/begin()/end());
//   * iteration runs from the NEWEST element to the oldest;
//   * when the buffer is full, end() stops one element early: a range-for visits SZ-1 elements.
// The capacity/span are template parameters as in the reference; a run-time sized core
// (RingCore) sits underneath so the CLI can change them with flags.
#ifndef BF_DATASTRUCTURES_H
#define BF_DATASTRUCTURES_H

#include <cassert>
#include <climits>
#include <cstddef>
#include <utility>
#include <vector>

template <class DType> class RingCore {
public:
    typedef DType value_type;

    RingCore(size_t capacity, long long span)
        : cap_(capacity), span_(span), store_(capacity), count_(0), head_(0), span_ok_(true), newest_(nullptr) {}

    size_t capacity() const { return cap_; }
    long long span() const { return span_; }

    size_t size() {
        trim();
        return count_;
    }

    void push_back(DType &d) {
        span_ok_ = false;
        if (count_ < cap_) ++count_;
        head_ = (head_ + 1 >= cap_) ? 0 : head_ + 1;
        store_[head_] = d;
        newest_ = &store_[head_];
    }

    // push_back that copies only what DType::copy_header_to copies (for an Event: coordinates, times, noise / valid
    // flags -- the first 26 of its 152 bytes); the slot's other fields keep their old contents.  Used by DVS_flow when
    // the slices are cut on the device and nothing on the host ever reads the per-event state (set_device_ring).
    void push_back_header(const DType &d) {
        span_ok_ = false;
        if (count_ < cap_) ++count_;
        head_ = (head_ + 1 >= cap_) ? 0 : head_ + 1;
        d.copy_header_to(store_[head_]);
        newest_ = &store_[head_];
    }

    // idx counts back from the newest element
    DType &operator[](size_t idx) {
        assert(idx < count_);
        return store_[(head_ + cap_ - idx % cap_) % cap_];
    }

    class iterator {
    public:
        DType &operator*() { return core_->store_[pos_]; }
        DType *operator->() { return &core_->store_[pos_]; }
        iterator &operator++() {
            pos_ = (pos_ == 0) ? core_->cap_ - 1 : pos_ - 1;
            return *this;
        }
        bool operator!=(const iterator &o) const { return pos_ != o.pos_; }
        bool operator==(const iterator &o) const { return pos_ == o.pos_; }

    private:
        friend class RingCore;
        iterator(RingCore *c, size_t p) : core_(c), pos_(p) {}
        RingCore *core_;
        size_t pos_;
    };

    iterator begin() {
        trim();
        return iterator(this, head_);
    }

    iterator end() {
        trim();
        // one-past-the-oldest in walking order; when full this aliases the oldest element itself,
        // which is why a full buffer yields capacity-1 elements (reference quirk, see header)
        const long long full = (count_ >= cap_) ? 1 : 0;
        const long long pos = ((full - ((long long)count_ - (long long)head_)) % (long long)cap_ + (long long)cap_) % (long long)cap_;
        return iterator(this, size_t(pos));
    }

private:
    void trim() {
        if (span_ok_) return;
        span_ok_ = true;
        long long tail = ((1 - ((long long)count_ - (long long)head_)) % (long long)cap_ + (long long)cap_) % (long long)cap_;
        size_t dropped = 0;
        while ((long long)(*newest_ - store_[size_t(tail)]) > span_) {
            ++dropped;
            tail = (tail + 1 >= (long long)cap_) ? 0 : tail + 1;
        }
        count_ -= dropped;
    }

    size_t cap_;
    long long span_;
    std::vector<DType> store_;
    size_t count_, head_;
    bool span_ok_;
    DType *newest_;
};

template <class DType, size_t SZ, long long SPAN> class CircularArray final : public RingCore<DType> {
public:
    CircularArray() : RingCore<DType>(SZ, SPAN) {}
    CircularArray(size_t capacity, long long span) : RingCore<DType>(capacity, span) {}
};

// A simple linear event cloud with no structure (datastructures.h:119-168)
template <class DType> class LinearEventCloudTemplate {
public:
    int x_min, y_min, x_max, y_max;

    LinearEventCloudTemplate() : x_min(INT_MAX), y_min(INT_MAX), x_max(INT_MIN), y_max(INT_MIN) {}
    explicit LinearEventCloudTemplate(std::vector<DType> &v) : LinearEventCloudTemplate() {
        items_.reserve(v.size());
        for (auto &d : v) push_back(d);
    }

    void push_back(const DType &d) {
        grow_box((int)d.get_x(), (int)d.get_y());
        items_.push_back(d);
    }
    void reserve(size_t n) { items_.reserve(n); }
    void swap(LinearEventCloudTemplate &o) {
        items_.swap(o.items_);
        std::swap(x_min, o.x_min); std::swap(y_min, o.y_min); std::swap(x_max, o.x_max); std::swap(y_max, o.y_max);
    }
    DType &operator[](size_t i) {
        assert(i < items_.size());
        return items_[i];
    }
    size_t size() { return items_.size(); }
    auto begin() { return items_.begin(); }
    auto end() { return items_.end(); }

private:
    void grow_box(int x, int y) {
        if (x > x_max) x_max = x;
        if (y > y_max) y_max = y;
        if (x < x_min) x_min = x;
        if (y < y_min) y_min = y;
    }
    std::vector<DType> items_;
};

// Same interface, but referring to events that live elsewhere (datastructures.h:172-259)
template <class DType> class LinearEventPtrsTemplate {
public:
    int x_min, y_min, x_max, y_max;

    LinearEventPtrsTemplate() : x_min(INT_MAX), y_min(INT_MAX), x_max(INT_MIN), y_max(INT_MIN) {}

    void push_back(DType *d) {
        const int x = (int)d->get_x(), y = (int)d->get_y();
        if (x > x_max) x_max = x;
        if (y > y_max) y_max = y;
        if (x < x_min) x_min = x;
        if (y < y_min) y_min = y;
        refs_.push_back(d);
    }
    void push_back(DType &d) { push_back(&d); }
    void reserve(size_t n) { refs_.reserve(n); }
    DType &operator[](size_t i) {
        assert(i < refs_.size());
        return *refs_[i];
    }
    size_t size() { return refs_.size(); }

    class iterator {
    public:
        explicit iterator(typename std::vector<DType *>::iterator it) : it_(it) {}
        DType &operator*() { return **it_; }
        DType *operator->() { return *it_; }
        iterator &operator++() {
            ++it_;
            return *this;
        }
        bool operator!=(const iterator &o) const { return it_ != o.it_; }
        bool operator==(const iterator &o) const { return it_ == o.it_; }

    private:
        typename std::vector<DType *>::iterator it_;
    };
    iterator begin() { return iterator(refs_.begin()); }
    iterator end() { return iterator(refs_.end()); }

private:
    std::vector<DType *> refs_;
};

#endif  // BF_DATASTRUCTURES_H
