// TEST INFRASTRUCTURE ONLY.  Stand-in for tbb::blocked_range<T>: the reference uses
// only begin()/end() of an integer range (accel_lib.h:168,528).
#pragma once
namespace tbb {
template <class T> class blocked_range {
public:
    blocked_range(T b, T e) : b_(b), e_(e) {}
    T begin() const { return b_; }
    T end() const { return e_; }
private:
    T b_, e_;
};
}  // namespace tbb
