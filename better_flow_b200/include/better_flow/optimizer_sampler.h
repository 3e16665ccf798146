// better_flow/optimizer_sampler.h -- OptimizerLocal, the contrast-driven optimiser (reference:
// better_flow_core/include/better_flow/optimizer_sampler.h, src/optimizer_sampler.cpp).  Same class
// name, constructors and public methods (run, get_nx, get_ny).
//
// The reference's run() evaluates iteration_step() on the host once per trial position: project
// every event with the global (nx, ny), splat a saturating 8-bit event-count image, Gaussian-blur it
// and score it by the mean of its non-zero pixels (optimizer_sampler.cpp:120-153, 192-204), moving nx
// and ny alternately with step halving (:20-23, 90-117).  Here run() hands the cloud to the CUDA
// library once; the whole descent runs inside the persistent kernel (bf_local_minimize) and the
// events' warped positions are written back afterwards.  manual() is a GUI mode and is not provided.
#ifndef BF_OPTIMIZER_SAMPLER_H
#define BF_OPTIMIZER_SAMPLER_H

#include <better_flow/common.h>
#include <better_flow/event.h>
#include <better_flow/opencl_driver.h>

class OptimizerLocal {
protected:
    LinearEventCloud *events;
    Event event_c;
    int scale;
    int metric_wsizex, metric_wsizey;
    int scale_img_x, scale_img_y;
    double nx, ny;
    double last_score, dscore;
    double dnx, dny, dn_th;
    int steps_;
    bool explicit_window_;

public:
    // optimizer_sampler.h:35-39 (explicit centre event and window size)
    OptimizerLocal(LinearEventCloud *events_, Event &e_, int sc_, int wsz_)
        : events(events_), event_c(e_), scale(sc_), metric_wsizex(sc_ * wsz_), metric_wsizey(sc_ * wsz_), nx(0), ny(0),
          last_score(0), dscore(0), dnx(0.01), dny(0.01), dn_th(0), steps_(0), explicit_window_(true) {
        update_fields();
    }

    // optimizer_sampler.h:41-56: window = bounding box of the cloud, centre event at its middle, t = 0
    OptimizerLocal(LinearEventCloud *events_, int sc_)
        : events(events_), scale(sc_), nx(0), ny(0), last_score(0), dscore(0), dnx(0.01), dny(0.01), dn_th(0), steps_(0),
          explicit_window_(false) {
        const int x_min = events->x_min, y_min = events->y_min;
        const int x_max = events->x_max, y_max = events->y_max;
        metric_wsizex = sc_ * (x_max - x_min);
        metric_wsizey = sc_ * (y_max - y_min);
        event_c = Event((x_max - x_min) / 2 + x_min, (y_max - y_min) / 2 + y_min, 0);
        update_fields();
    }

    // optimizer_sampler.cpp:4-38.  Returns 0, or 1 when the window is too small (:9-13).
    int run() {
        if (explicit_window_) {
            // (the per-event windowed variant is only reachable from the reference's unreleased clustering code)
            std::cerr << "OptimizerLocal: the explicit-window constructor has no CUDA path; use OptimizerLocal(cloud, scale)" << std::endl;
            std::exit(1);
        }
        const int n = (int)events->size();
        std::vector<uint16_t> fx((size_t)n), fy((size_t)n);
        std::vector<int32_t> t((size_t)n);
        int i = 0;
        for (auto &e : *events) {
            if (e.t > INT32_MAX || e.t < INT32_MIN) {
                std::cerr << "OptimizerLocal: an event's t exceeds +-2.1 s; set local times first" << std::endl;
                std::exit(1);
            }
            fx[(size_t)i] = (uint16_t)e.fr_x; fy[(size_t)i] = (uint16_t)e.fr_y; t[(size_t)i] = (int32_t)e.t;
            ++i;
        }
        bf_ctx *ctx = CudaDriver::context(n, 1, scale);
        bf_slice_result res;
        const int rc = bf_local_minimize(ctx, fx.data(), fy.data(), t.data(), n, scale, &res);
        if (rc < 0) {
            std::cerr << "bf_local_minimize failed: " << bf_last_error() << std::endl;
            std::exit(1);
        }
        nx = res.model.total_dx; ny = res.model.total_dy;
        last_score = res.model.dx; dnx = res.model.dy; dny = res.model.rot; dn_th = res.model.div;
        steps_ = res.iters;
        if (rc == BF_RC_SKIPPED) return 1;
        // the events are left as the last iteration_step projected them: compute_new_ny's (nx, ny)
        for (auto &e : *events) e.project(nx, ny);
        event_c.project(nx, ny);
        return 0;
    }

    double get_nx() { return nx; }
    double get_ny() { return ny; }

    // extensions
    double get_score() const { return last_score; }
    int steps() const { return steps_; }

private:
    void update_fields() {   // optimizer_sampler.cpp:207-214
        assert(scale % 2 != 0);
        scale_img_x = metric_wsizex + scale;
        scale_img_y = metric_wsizey + scale;
    }

protected:
    auto begin() { return events->begin(); }
    auto end() { return events->end(); }
};

#endif  // BF_OPTIMIZER_SAMPLER_H
