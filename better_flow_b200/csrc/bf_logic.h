// bf_logic.h -- scalar logic shared by host and device (geometry of a slice's image, the
// packed-accumulator configuration, and the gradient-descent control flow of
// OptimizerRolling::run).  Everything here is plain arithmetic so that the same code can be
// unit-tested on the CPU and executed by one thread per CTA on the GPU.
//
// Reference paths are relative to /root/reference/better_flow_core/.
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/bf_cuda.h"

#if defined(__CUDACC__)
#define BF_HD __host__ __device__ __forceinline__
#else
#define BF_HD inline
#endif

// ---- geometry: OptimizerRolling::set_cloud + set_scale (optimizer_rolling.h:248-283) ---------
struct BfGeom {
    int x_min, x_max, y_min, y_max;  // bbox over fr_x (rows), fr_y (cols)
    int w, h;                        // metric_wsizex / metric_wsizey = scale * extent
    int rows, cols;                  // scale_img_x / scale_img_y   = w + scale, h + scale
    int x_sh, y_sh;                  // shifts as get_time_img receives them: truncated to int (accel_lib.h:211)
    double x_shift, y_shift;         // untruncated (used for the centre, optimizer_rolling.h:330-331)
    int scale, half;
};

BF_HD void bf_make_geom(BfGeom &g, int x_min, int x_max, int y_min, int y_max, int scale) {
    g.x_min = x_min; g.x_max = x_max; g.y_min = y_min; g.y_max = y_max;
    g.scale = scale; g.half = scale / 2;
    g.w = scale * (x_max - x_min);
    g.h = scale * (y_max - y_min);
    g.rows = g.w + scale;
    g.cols = g.h + scale;
    // integer /2 on the extent and on scale, exactly as written in the reference (:279-282)
    g.x_shift = -(double)((x_max - x_min) / 2 + x_min) * (double)scale + (double)g.w / 2.0 + scale / 2;
    g.y_shift = -(double)((y_max - y_min) / 2 + y_min) * (double)scale + (double)g.h / 2.0 + scale / 2;
    g.x_sh = (int)g.x_shift;
    g.y_sh = (int)g.y_shift;
}

// run()'s first guard (optimizer_rolling.h:49): "window too small"
BF_HD bool bf_guard_tiny(const BfGeom &g, int res_x, int res_y) {
    return (g.rows < g.scale * res_x / 15) && (g.cols < g.scale * res_y / 15);
}

// ---- packed per-pixel accumulator --------------------------------------------------------
// One 64-bit word per pixel: [ count : cnt_bits | sum of (t - t_min) >> q : 64 - cnt_bits ].
// A single 64-bit integer atomic add per event accumulates both, exactly and in any order.
// Width budget: a pixel (or any box of pixels) receives at most n events, each with
// (t - t_min) < 2^t_bits, so the sum needs t_bits + cnt_bits bits and the count cnt_bits.
// If t_bits + 2*cnt_bits > 64 the times are right-shifted by q (never at BASELINE sizes).
struct BfPack {
    int cnt_shift;           // 64 - cnt_bits
    int q;                   // time quantisation shift (0 = exact)
    int32_t t_min;           // offset subtracted from every t before packing (0 when all t >= 0 fit as they are)
    int fast;                // 1: q == 0, t_min == 0 and every box sum < 2^52 (device fast unpack path)
    unsigned long long sum_mask;
};

// Every event of a compact batch inside the res_x x res_y sensor?  (Host-side precondition check of
// bf_batch_add_packed / bf_batch_add_staged: the device sizes a slice's images from the bounding box of its
// events, so a coordinate beyond the context's sensor would index past the image allocation.)
BF_HD bool bf_events_in_sensor(const bf_event *ev, long long n, int res_x, int res_y) {
    unsigned bad = 0;
    for (long long i = 0; i < n; ++i)
        bad |= (unsigned)((int)ev[i].fr_x >= res_x) | (unsigned)((int)(ev[i].fr_y & 0x7fffu) >= res_y);
    return bad == 0;
}

BF_HD int bf_bits(unsigned long long v) {
    int b = 0;
    while (v) { ++b; v >>= 1; }
    return b;
}

BF_HD void bf_make_pack(BfPack &p, int n, int32_t t_min, int32_t t_max) {
    const int cnt_bits = bf_bits((unsigned long long)(n > 0 ? n : 1));
    p.cnt_shift = 64 - cnt_bits;
    p.sum_mask = (1ull << p.cnt_shift) - 1ull;
    // Local times are normally >= 0 (t = timestamp - slice start): then they are packed as they are
    // and the unpack needs no offset correction.  Only when that would not fit (or t < 0 occurs:
    // events older than the slice start of an overflowed buffer, dvs_flow.h:187-190) the times are
    // re-based to t_min.
    if (t_min >= 0 && bf_bits((unsigned long long)t_max) + 2 * cnt_bits <= 64) {
        p.q = 0;
        p.t_min = 0;
        p.fast = (bf_bits((unsigned long long)t_max) + cnt_bits <= 52) ? 1 : 0;
        return;
    }
    const unsigned long long span = (unsigned long long)((long long)t_max - (long long)t_min);
    const int t_bits = bf_bits(span);
    int q = t_bits + 2 * cnt_bits - 64;
    if (q < 0) q = 0;
    p.q = q;
    p.t_min = t_min;
    p.fast = 0;
}

BF_HD unsigned long long bf_pack_value(const BfPack &p, int32_t t) {
    const unsigned long long dt = (unsigned long long)((long long)t - (long long)p.t_min);
    return (1ull << p.cnt_shift) + (dt >> p.q);
}

// Mean timestamp of a pixel from its (box-summed) packed word, in seconds, as f32.
// The sum of t_ns is exact; it is rounded to f32 once and divided in f32 like the reference's
// normalise loop (accel_lib.h:171-172).  (Reference: f32 running sum, rounded at every add.)
BF_HD float bf_unpack_avg(const BfPack &p, unsigned long long v) {
    const unsigned long long cnt = v >> p.cnt_shift;
    if (cnt == 0) return 0.0f;
    long long sum = (long long)((v & p.sum_mask) << p.q);
    if (p.t_min != 0) sum += (long long)cnt * (long long)p.t_min;
    // sum / 1e9 as multiply-by-reciprocal + one fma correction step: equal to the correctly rounded
    // quotient for every integer dividend tried (2e9 random |a| < 2^53, tools/ in DESIGN.md) and
    // ~10x cheaper than an IEEE fp64 divide on the GPU.
    const double a = (double)sum;
    const double q0 = a * 1e-9;
    const double qd = fma(fma(-1000000000.0, q0, a), 1e-9, q0);
    const float s = (float)qd;
    return s / (float)cnt;
}

// ---- reduction result of one image pass ------------------------------------------------------
// Sums over pixels whose mean timestamp is > 1e-6 (object_model.cpp:17-33,111-120).  Moments are
// taken about a fixed origin (i0, j0) so that the centre of mass does not have to be known first:
//   rot = [ S(i-i0)gy - S(j-j0)gx - (cx-i0) Sgy + (cy-j0) Sgx ] / cnt
//   div = [ S(i-i0)gx + S(j-j0)gy - (cx-i0) Sgx - (cy-j0) Sgy ] / cnt
struct BfSums {
    double cnt, si, sj;           // exact integers held in doubles
    double sgx, sgy;
    double sigx, sjgx, sigy, sjgy;
};
#define BF_NSUMS 9

// ObjectModel::center_of_mass + compute (object_model.cpp:103-126, 4-39): fills cx, cy (image
// units), dx, dy, rot, div, cnt of `m`.
BF_HD void bf_sums_to_model(bf_model &m, const BfSums &s, int i0, int j0) {
    const double cnt = s.cnt;
    m.cnt = (uint32_t)cnt;
    m.cx = s.si / cnt;   // cnt == 0 -> NaN, as the NDEBUG reference
    m.cy = s.sj / cnt;
    const double ox = m.cx - (double)i0, oy = m.cy - (double)j0;
    double rot = s.sigy - s.sjgx;
    rot = rot - ox * s.sgy;
    rot = rot + oy * s.sgx;
    double div = s.sigx + s.sjgy;
    div = div - ox * s.sgx;
    div = div - oy * s.sgy;
    m.rot = rot / cnt;
    m.div = div / cnt;
    m.dx = s.sgx / cnt;
    m.dy = s.sgy / cnt;
}

// ---- warp parameters ---------------------------------------------------------------------------
// Arguments of Event::project_4param_reinit (event.h:99-110) with cos/sin of crl evaluated once.
struct BfProj {
    double dnx, dny, cx, cy, div, c, s;
};

BF_HD void bf_make_proj(BfProj &p, double dnx, double dny, double cx, double cy, double div, double crl) {
    p.dnx = dnx; p.dny = dny; p.cx = cx; p.cy = cy; p.div = div;
    p.c = cos(crl);
    p.s = sin(crl);
}

// ---- OptimizerRolling::run control flow (optimizer_rolling.h:48-125) ----------------------------
struct BfOpt {
    bf_model m;
    float x_div, y_div, rot_div, div_div;      // f32, as declared (:36)
    float old_dx, old_dy, old_rot, old_div;    // f32 copies (:86-89)
    int iters;                                 // itercount (:60)
    int rc;
};

BF_HD void bf_opt_init(BfOpt &o, const bf_model *init) {
    if (init) o.m = *init;
    else {
        o.m.cx = o.m.cy = o.m.dx = o.m.dy = o.m.rot = o.m.div = 0;
        o.m.cnt = 0; o.m.pad_ = 0;
        o.m.total_dx = o.m.total_dy = o.m.total_rot = o.m.total_div = 0;
    }
    o.x_div = o.y_div = 1.0f;       // :61-63
    o.rot_div = 10000;
    o.div_div = 10000;
    o.old_dx = o.old_dy = o.old_rot = o.old_div = 0;
    o.iters = 0;
    o.rc = BF_RC_OK;
}

// Consumes the reduction of one iteration_step's image (optimizer_rolling.h:327-346) and then
// advances run()'s loop to the point just before the next iteration_step.  Returns true when
// another step must be executed (with warp parameters `next`), false when run() is over; in both
// cases `next` holds the re-projection that iteration_step performs last (:340-344).
BF_HD bool bf_opt_advance(BfOpt &o, const BfGeom &g, const BfSums &s, int i0, int j0,
                          int max_iter, int iter_cap, BfProj &next) {
    // fast_model -> ObjectModel::update
    bf_sums_to_model(o.m, s, i0, j0);
    // update_accumulators(rot_divider, div_divider, x_divider, y_divider) (object_model.h:48-53)
    o.m.total_rot += o.m.rot / o.rot_div;
    o.m.total_div += o.m.div / o.div_div;
    o.m.total_dx += o.m.dx / o.x_div;
    o.m.total_dy += o.m.dy / o.y_div;
    // centre back to sensor units with the untruncated shifts (:330-331, :345-346)
    const double cx = (o.m.cx - g.x_shift) / g.scale;
    const double cy = (o.m.cy - g.y_shift) / g.scale;
    o.m.cx = cx;
    o.m.cy = cy;
    bf_make_proj(next, -o.m.total_dx, -o.m.total_dy, cx, cy, o.m.total_div, -o.m.total_rot);
    o.iters += 1;

    if (!(s.cnt > 0)) { o.rc = BF_RC_DEGENERATE; return false; }
    if (o.iters > 1) {
        // tail of the while body (:94-101)
        if (max_iter > 0 && o.iters > max_iter) return false;
        if (o.m.dx * o.old_dx < 0) o.x_div *= 2;
        if (o.m.dy * o.old_dy < 0) o.y_div *= 2;
        if (o.m.rot * o.old_rot < 0) o.rot_div *= 2;
        if (o.m.div * o.old_div < 0) o.div_div *= 2;
    }
    // while condition (:76-79)
    if (!(o.x_div < 32 * 10 || o.y_div < 32 * 10 || o.rot_div < 32 * 1000 || o.div_div < 32 * 1000))
        return false;
    // convergence test (:81-84)
    if (fabs(o.m.dx / o.x_div) < 1e-5 && fabs(o.m.dy / o.y_div) < 1e-5 &&
        fabs(o.m.rot / o.rot_div) < 1e-4 && fabs(o.m.div / o.div_div) < 1e-1)
        return false;
    if (o.iters >= iter_cap) { o.rc = BF_RC_ITER_CAP; return false; }
    o.old_dx = (float)o.m.dx;      // :86-89
    o.old_dy = (float)o.m.dy;
    o.old_rot = (float)o.m.rot;
    o.old_div = (float)o.m.div;
    return true;
}
