// TEST INFRASTRUCTURE ONLY.  Stand-in for tbb::parallel_for(range, body): splits the
// range into one contiguous chunk per thread (OpenMP when compiled with -fopenmp,
// serial otherwise).  The reference's two hot-path uses are independent row loops
// (accel_lib.h:168-175, 528-542), so results are bit-identical for any thread count.
// Thread count: BF_ORACLE_THREADS env var, default = all host threads.
#pragma once
#include <tbb/blocked_range.h>
#include <cstdlib>
#ifdef _OPENMP
#include <omp.h>
#endif
namespace tbb {
// Number of parallel_for calls so far.  OptimizerRolling::iteration_step issues exactly two
// (normalise accel_lib.h:168, Scharr accel_lib.h:528), which lets the oracle driver recover
// run()'s local `itercount` (optimizer_rolling.h:60) without touching the reference.
inline unsigned long long &bf_shim_call_count() {
    static unsigned long long c = 0;
    return c;
}
inline int bf_shim_threads() {
    static int n = -1;
    if (n < 0) {
        const char *s = std::getenv("BF_ORACLE_THREADS");
        n = s ? std::atoi(s) : 0;
#ifdef _OPENMP
        if (n <= 0) n = omp_get_max_threads();
#else
        n = 1;
#endif
        if (n < 1) n = 1;
    }
    return n;
}
template <class Range, class Body> inline void parallel_for(const Range &r, const Body &body) {
    const long b = r.begin(), e = r.end();
    const long n = e - b;
    const int nt = bf_shim_threads();
    ++bf_shim_call_count();
    if (n <= 0) return;
    if (nt <= 1 || n < 4L * nt) { body(r); return; }
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nt)
#endif
    for (int k = 0; k < nt; ++k) {
        const long lo = b + n * k / nt, hi = b + n * (k + 1) / nt;
        if (hi > lo) body(Range((decltype(r.begin()))lo, (decltype(r.begin()))hi));
    }
}
}  // namespace tbb
