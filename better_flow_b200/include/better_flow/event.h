// better_flow/event.h -- the host-side Event record (reference: better_flow_core/include/better_flow/event.h).
// Same public fields and methods.  The device never sees this 150-byte object: the optimiser packs
// fr_x / fr_y / t into 8-byte bf_event records (include/bf_cuda.h) and writes pr_x, pr_y, nx, ny back
// after run().  The small per-event formulas below are needed on the host for set-up and for the
// post-processing the reference does outside the hot loop (compute_uv for -o output).
#ifndef BF_EVENT_H
#define BF_EVENT_H

#include <better_flow/common.h>

class Cluster;

class Event {
public:
    uint fr_x, fr_y;     // fr_x = sensor row, fr_y = sensor column (bf_motion_compensator.cpp:200)
    sll t;               // local time, ns, relative to the slice start (set_local_time)
    ull timestamp;       // absolute time, ns
    bool noise;
    bool valid;

    double pr_x, pr_y;   // warped position
    double nx, ny, nz;   // direction vector
    double u, v;         // flow, px/s

    double best_u, best_v, max_score;
    double best_pr_x, best_pr_y;

    bool visited;
    Cluster *cl;
    int cl_id;

    Event()
        : fr_x(UINT_MAX), fr_y(UINT_MAX), t(sll(ULLONG_MAX)), timestamp(LLONG_MAX), noise(true), valid(false),
          pr_x(NAN), pr_y(NAN), nx(0), ny(0), nz(NZ), u(0), v(0), best_u(0), best_v(0), max_score(0),
          best_pr_x(NAN), best_pr_y(NAN), visited(false), cl(nullptr), cl_id(-1) {}

    Event(uint x_, uint y_, ull t_)
        : fr_x(x_), fr_y(y_), t(sll(t_)), timestamp(t_), noise(false), valid(false), pr_x(x_), pr_y(y_), nx(0), ny(0),
          nz(NZ), u(0), v(0), best_u(0), best_v(0), max_score(0), best_pr_x(x_), best_pr_y(y_), visited(false),
          cl(nullptr), cl_id(-1) {}

    // time difference in ns (event.h:36-38)
    sll operator-(const Event &rhs) const { return sll(timestamp) - sll(rhs.timestamp); }

    // same pixel and closer than 0.1 ms (event.h:40-45; the asymmetric use of `t` in the second branch
    // is the reference's and is kept)
    bool operator==(const Event &rhs) const {
        if (fr_x != rhs.fr_x || fr_y != rhs.fr_y) return false;
        const ull dt = (timestamp >= rhs.timestamp) ? timestamp - rhs.timestamp : rhs.timestamp - ull(t);
        return dt < 100000;
    }
    bool operator!=(const Event &rhs) const { return !(*this == rhs); }

    // what identifies an event (everything else is per-slice state that Event::reset / the optimiser rewrite)
    void copy_header_to(Event &o) const {
        o.fr_x = fr_x; o.fr_y = fr_y; o.t = t; o.timestamp = timestamp; o.noise = noise; o.valid = valid;
    }

    uint get_x() const { return fr_x; }
    uint get_y() const { return fr_y; }

    void reset() {                       // event.h:54-59
        pr_x = fr_x;
        pr_y = fr_y;
        nx = ny = 0;
        u = v = 0;
    }

    void set_local_time(ull t0) {        // event.h:61-63
        t = (timestamp > t0) ? sll(timestamp - t0) : -sll(t0 - timestamp);
    }

    void project(double nx_, double ny_, double nz_ = NZ) {
        nx = nx_; ny = ny_; nz = nz_;
        warp_from_n();
    }
    void project_dn(double dnx, double dny) {
        nx += dnx; ny += dny;
        warp_from_n();
    }

    // 4-parameter warp about (cx, cy), restarted from the current warped position (event.h:99-110)
    void project_4param_reinit(double dnx, double dny, double cx, double cy, double div, double crl) {
        const double rx = pr_x - cx, ry = pr_y - cy;
        const double c = std::cos(crl), s = std::sin(crl);
        const double qx = c * rx - s * ry, qy = s * rx + c * ry;
        nx = ((-qx) * div + (qx - rx)) + dnx;
        ny = ((-qy) * div + (qy - ry)) + dny;
        warp_from_n();
    }

    void apply_score(double score) {
        if (score > max_score) assume_score(score);
    }
    void assume_score(double score) {    // event.h:123-129
        max_score = score;
        best_u = u; best_v = v;
        best_pr_x = pr_x; best_pr_y = pr_y;
    }

    // px/s <-> direction-vector units: 1000000000/(T_DIVIDER*10000) is an INTEGER expression (= 100000)
    double n_from_u(double vel) const { return vel * (nz / (1000000000 / (T_DIVIDER * 10000))); }

    void compute_uv() {                  // event.h:135-142
        const double len = std::hypot(nx, ny);
        const double speed = len / (nz / (1000000000 / (T_DIVIDER * 10000)));
        u = (len == 0) ? 0 : speed * nx / len;
        v = (len == 0) ? 0 : speed * ny / len;
        best_u = u;
        best_v = v;
    }

private:
    // event.h:164-168: f32 slope, f32 product with the f32 time, then f64 divide and subtract
    void warp_from_n() {
        const float kx = float(nx) / nz, ky = float(ny) / nz;
        pr_x = float(fr_x) - kx * float(t) / 10000.0;
        pr_y = float(fr_y) - ky * float(t) / 10000.0;
    }
};

#endif  // BF_EVENT_H
