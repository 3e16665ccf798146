// Host build of the text event reader (better_flow/event_file.h: TextEventReader) next to the
// reference's way of parsing the same file (`ifstream >> double >> uint >> uint >> bool`,
// bf_motion_compensator.cpp:190-202), behind a tiny C interface for the CPU test-suite.
#include <better_flow/event_file.h>

#include <chrono>
#include <iomanip>

extern "C" {

// Both return the number of records parsed (up to cap) and the seconds spent.
long long rd_fast(const char *path, long long cap, double *t, unsigned *x, unsigned *y, unsigned char *p, double *secs) {
    const auto t0 = std::chrono::steady_clock::now();
    TextEventReader in(path);
    long long n = 0;
    double tv; uint xv, yv; bool pv;
    while (n < cap && in.next(tv, xv, yv, pv)) { t[n] = tv; x[n] = xv; y[n] = yv; p[n] = pv; ++n; }
    *secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return n;
}

long long rd_prefetch(const char *path, long long cap, double *t, unsigned *x, unsigned *y, unsigned char *p, double *secs) {
    const auto t0 = std::chrono::steady_clock::now();
    PrefetchingTextEventReader in(path, 1 << 12);
    long long n = 0;
    double tv; uint xv, yv; bool pv;
    while (n < cap && in.next(tv, xv, yv, pv)) { t[n] = tv; x[n] = xv; y[n] = yv; p[n] = pv; ++n; }
    *secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return n;
}

long long rd_iostream(const char *path, long long cap, double *t, unsigned *x, unsigned *y, unsigned char *p, double *secs) {
    const auto t0 = std::chrono::steady_clock::now();
    std::ifstream in(path, std::ifstream::in);
    long long n = 0;
    double tv; uint xv, yv; bool pv;
    while (n < cap && (in >> tv >> xv >> yv >> pv)) { t[n] = tv; x[n] = xv; y[n] = yv; p[n] = pv; ++n; }
    *secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return n;
}

// EventFile::from_file end to end: timestamps (ns) and coordinates of the resulting cloud.
long long rd_from_file(const char *path, long long cap, unsigned long long *ts, unsigned *fr_x, unsigned *fr_y) {
    LinearEventCloud ec;
    std::streambuf *old = std::cout.rdbuf(nullptr);   // silence the progress prints
    EventFile::from_file(&ec, path);
    std::cout.rdbuf(old);
    long long n = 0;
    for (auto &e : ec) {
        if (n >= cap) break;
        ts[n] = e.timestamp; fr_x[n] = e.fr_x; fr_y[n] = e.fr_y; ++n;
    }
    return n;
}

// EventFile::to_file_uv next to the reference's way of writing the same records (`ofstream << fixed << setprecision(9)`,
// event_file.h:265-289); both return the seconds spent.
static LinearEventCloud wr_cloud(long long n, const unsigned long long *ts, const unsigned *fr_x, const unsigned *fr_y, const double *bu, const double *bv) {
    LinearEventCloud ec;
    ec.reserve((size_t)n);
    for (long long i = 0; i < n; ++i) {
        Event e(fr_x[i], fr_y[i], ts[i]);
        e.best_u = bu[i]; e.best_v = bv[i];
        ec.push_back(e);
    }
    return ec;
}
double wr_fast(const char *path, long long n, const unsigned long long *ts, const unsigned *fr_x, const unsigned *fr_y, const double *bu, const double *bv) {
    LinearEventCloud ec = wr_cloud(n, ts, fr_x, fr_y, bu, bv);
    std::streambuf *old = std::cout.rdbuf(nullptr);
    const auto t0 = std::chrono::steady_clock::now();
    EventFile::to_file_uv(&ec, path);
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout.rdbuf(old);
    return s;
}
double wr_iostream(const char *path, long long n, const unsigned long long *ts, const unsigned *fr_x, const unsigned *fr_y, const double *bu, const double *bv) {
    LinearEventCloud ec = wr_cloud(n, ts, fr_x, fr_y, bu, bv);
    const auto t0 = std::chrono::steady_clock::now();
    std::ofstream out(path, std::ofstream::out);
    out << std::fixed << std::setprecision(9);
    for (auto &e : ec)
        out << double(e.timestamp) / 1000000000 << " " << e.fr_y << " " << e.fr_x << " " << 1 << " " << e.best_v << " " << e.best_u << "\n";
    out.close();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
}
