// better_flow/optimizer_rolling.h -- per-slice optimiser (reference:
// better_flow_core/include/better_flow/optimizer_rolling.h).  Same class name and public methods
// (run, set_maxiter, set_time, set_cloud, set_scale, get_model, set_model, get_time_img).
//
// The difference is where the work happens: the reference's run() loops on the host over
// iteration_step() (time image -> Scharr -> reductions -> accumulators -> re-projection,
// optimizer_rolling.h:48-125,305-347); here run() hands the slice to the CUDA library once and the
// whole gradient descent -- including the divider / sign-flip / convergence control flow -- runs in
// one persistent kernel.  Events are mutated in place afterwards (pr_x, pr_y, nx, ny, noise, best_*),
// as callers of the reference expect.  The interactive manual() mode and the debug image getters
// are GUI-only and not provided.
#ifndef BF_OPTIMIZER_ROLLING_H
#define BF_OPTIMIZER_ROLLING_H

#include <better_flow/accel_lib.h>
#include <better_flow/common.h>
#include <better_flow/event.h>
#include <better_flow/object_model.h>

template <class T> class OptimizerRolling {
protected:
    AccelLib accel;
    T *events;
    int scale;
    int metric_wsizex, metric_wsizey;
    int max_itercount;
    int scale_img_x, scale_img_y;
    double x_shift, y_shift;
    int x_min, y_min, x_max, y_max;
    ull current_time;
    ObjectModel model;
    float x_divider, y_divider, rot_divider, div_divider;
    bool warm_start_;       // set_model() was called: the device applies it before the first step
    int itercount_;
    int last_rc_;

public:
    OptimizerRolling()
        : events(nullptr), scale(0), metric_wsizex(0), metric_wsizey(0), max_itercount(-1), scale_img_x(0), scale_img_y(0),
          x_shift(0), y_shift(0), x_min(0), y_min(0), x_max(0), y_max(0), current_time(0), x_divider(1), y_divider(1),
          rot_divider(10000), div_divider(10000), warm_start_(false), itercount_(0), last_rc_(0) {}

    // optimizer_rolling.h:48-125.  Returns 0 when the slice was optimised, 1 when it was skipped.
    int run() {
        const int n = (int)events->size();
        fx_.resize(n); fy_.resize(n); t_.resize(n); nz_.resize(n);
        px_.resize(n); py_.resize(n); nx_.resize(n); ny_.resize(n);
        int i = 0;
        for (auto &e : *events) {
            if (e.t > INT32_MAX || e.t < INT32_MIN) {
                std::cerr << "OptimizerRolling: local time of an event exceeds +-2.1 s; shorten the slice" << std::endl;
                std::exit(1);
            }
            fx_[i] = (uint16_t)e.fr_x; fy_[i] = (uint16_t)e.fr_y; t_[i] = (int32_t)e.t; nz_[i] = e.noise ? 1 : 0;
            ++i;
        }
        bf_ctx *ctx = CudaDriver::context(n, 1, scale);
        bf_model init = model.to_pod();
        bf_slice_result res;
        const int rc = bf_minimize(ctx, fx_.data(), fy_.data(), t_.data(), nz_.data(), n, scale, max_itercount,
                                   warm_start_ ? &init : nullptr, &res, px_.data(), py_.data(), nx_.data(), ny_.data());
        if (rc < 0) {
            std::cerr << "bf_minimize failed: " << bf_last_error() << std::endl;
            std::exit(1);
        }
        last_rc_ = rc;
        itercount_ = res.iters;
        model.from_pod(res.model);
        x_divider = res.dividers[0]; y_divider = res.dividers[1]; rot_divider = res.dividers[2]; div_divider = res.dividers[3];
        i = 0;
        const bool all_noise = (res.flags & BF_FLAG_ALL_NOISE) != 0;   // tiny window (:49-55)
        for (auto &e : *events) {
            e.pr_x = px_[i]; e.pr_y = py_[i]; e.nx = nx_[i]; e.ny = ny_[i];
            if (all_noise) e.noise = true;
            ++i;
        }
        if (rc == BF_RC_SKIPPED) return 1;
        if (rc == BF_RC_DEGENERATE)
            std::cerr << "OptimizerRolling: empty time image (the reference would not terminate here)" << std::endl;
        for (auto &e : *events) e.assume_score(0);   // :121-122
        return 0;
    }

    void set_maxiter(int val) { max_itercount = val; }

    // optimizer_rolling.h:241-245
    void set_time(ull t_) {
        current_time = t_;
        for (auto &e : *events) e.set_local_time(current_time);
    }

    // optimizer_rolling.h:248-270: bounding box (minima start at RES_X / RES_Y, maxima at 0), reset
    void set_cloud(T *events_, int sc_) {
        events = events_;
        scale = sc_;
        x_min = RES_X; y_min = RES_Y;
        x_max = 0; y_max = 0;
        for (auto &e : *events) {
            if ((int)e.fr_x > x_max) x_max = e.fr_x;
            if ((int)e.fr_y > y_max) y_max = e.fr_y;
            if ((int)e.fr_x < x_min) x_min = e.fr_x;
            if ((int)e.fr_y < y_min) y_min = e.fr_y;
            e.reset();
        }
        metric_wsizex = sc_ * (x_max - x_min);
        metric_wsizey = sc_ * (y_max - y_min);
        set_scale(scale);
        accel.init_gpu(events, metric_wsizex + scale, metric_wsizey + scale);
    }

    // optimizer_rolling.h:272-283 (integer halving of the extent and of the scale is intentional)
    void set_scale(int sc_) {
        scale = sc_;
        assert(scale % 2 != 0);
        scale_img_x = metric_wsizex + scale;
        scale_img_y = metric_wsizey + scale;
        x_shift = -double((x_max - x_min) / 2 + x_min) * double(scale) + double(metric_wsizex) / 2.0 + scale / 2;
        y_shift = -double((y_max - y_min) / 2 + y_min) * double(scale) + double(metric_wsizey) / 2.0 + scale / 2;
    }

    ObjectModel get_model() { return model; }

    // optimizer_rolling.h:289-299.  The warm-start re-projection is applied on the device at the start
    // of run() (also when run() then bails out on a guard), not here.
    void set_model(ObjectModel m) {
        model = m;
        warm_start_ = true;
    }

    // Mean-timestamp image of the events' current warped positions (debug helper, :301,351-357)
    ImageF get_time_img() {
        return accel.get_time_img(events, metric_wsizex, metric_wsizey, scale, (int)x_shift, (int)y_shift);
    }

    // extensions
    int iterations() const { return itercount_; }
    int last_rc() const { return last_rc_; }
    double get_x_shift() const { return x_shift; }
    double get_y_shift() const { return y_shift; }

    auto begin() { return events->begin(); }
    auto end() { return events->end(); }

private:
    std::vector<uint16_t> fx_, fy_;
    std::vector<int32_t> t_;
    std::vector<uint8_t> nz_;
    std::vector<double> px_, py_, nx_, ny_;
};

#endif  // BF_OPTIMIZER_ROLLING_H
