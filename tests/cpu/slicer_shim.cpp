// Host build of the slice manager (better_flow/dvs_flow.h: DVS_flow + CircularArray) in its queue-only mode
// (stm disabled, a batch that never fills: every recompute() snapshots its slice, nothing is minimised, the
// CUDA library is linked but never called), behind a tiny C interface for the CPU test-suite.  What a slice
// IS -- which events, in which order, with which local times -- is decided here, before any kernel runs.
#include <better_flow/dvs_flow.h>

namespace {

template <class Flow> struct FlowProbe : Flow {
    using Flow::Flow;
    size_t queued() const { return this->pending_.size(); }
    const std::vector<bf_event> &packed(size_t k) const { return this->pending_[k].packed; }
    ull first_ts(size_t k) const { return this->pending_[k].log.ts_first; }
    ull last_ts(size_t k) const { return this->pending_[k].log.ts_last; }
    ull start(size_t k) const { return this->pending_[k].slice_start; }
};

// info:  3 per slice = {events consumed, buffer size, buffer time diff}   (what oracle/ref_driver.cpp records)
// slice: 6 per slice = {n iterated, newest timestamp, oldest timestamp, slice start, sum of local t, sum of (k+1) * fr_x[k]}
template <class Flow>
int run(Flow &est, int n, const uint32_t *fr_x, const uint32_t *fr_y, const uint64_t *ts, int flush, int max_slices,
        long long *info, long long *slice) {
    est.set_stm_disable(true);
    est.set_batch(1 << 30);
    est.set_quiet(true);
    int ns = 0;
    auto record = [&](long long consumed) {
        if (ns < max_slices) {
            info[3 * ns + 0] = consumed;
            info[3 * ns + 1] = est.get_buf_size();
            info[3 * ns + 2] = est.get_buf_time_diff();
            const size_t k = est.queued() - 1;
            const std::vector<bf_event> &p = est.packed(k);
            long long st = 0, sx = 0;
            for (size_t i = 0; i < p.size(); ++i) { st += p[i].t_ns; sx += (long long)(i + 1) * p[i].fr_x; }
            slice[6 * ns + 0] = (long long)p.size();
            slice[6 * ns + 1] = (long long)est.first_ts(k);
            slice[6 * ns + 2] = (long long)est.last_ts(k);
            slice[6 * ns + 3] = (long long)est.start(k);
            slice[6 * ns + 4] = st;
            slice[6 * ns + 5] = sx;
        }
        ++ns;
    };
    for (int i = 0; i < n; ++i) {
        Event e(fr_x[i], fr_y[i], ts[i]);
        if (est.add_event(e)) record(i + 1);
    }
    if (flush) { est.recompute(); record(n); }
    return ns;
}

}  // namespace

extern "C" int sl_stream(int config, int n, const uint32_t *fr_x, const uint32_t *fr_y, const uint64_t *ts,
                         unsigned long long ev_refresh, unsigned long long time_refresh_ns, int flush, int max_slices,
                         long long capacity, long long span_ns, long long *info, long long *slice) {
    bf::set_sensor(180, 240);
    if (config == 0) {          // bf_motion_compensator.cpp:6-7,135
        FlowProbe<DVS_flow<50000, FROM_SEC(0.2)>> est(ev_refresh, time_refresh_ns);
        return run(est, n, fr_x, fr_y, ts, flush, max_slices, info, slice);
    }
    if (config == 1) {          // ros_nodes_src/bf_visualizer.cpp:30-34
        FlowProbe<DVS_flow<30000, FROM_MS(70)>> est(ev_refresh, time_refresh_ns);
        return run(est, n, fr_x, fr_y, ts, flush, max_slices, info, slice);
    }
    if (config == 2) {          // run-time sized buffer (CLI --max-events / --slice-time)
        FlowProbe<DVS_flow<50000, FROM_SEC(0.2)>> est(ev_refresh, time_refresh_ns, 0, (size_t)capacity, span_ns);
        return run(est, n, fr_x, fr_y, ts, flush, max_slices, info, slice);
    }
    return -1;
}

// ---- the per-event formulas of the host mirror (better_flow/event.h), for comparison with the oracle ----
// Event::project_4param_reinit + apply_project (event.h:99-110,164-168), restarted from the given pr.
extern "C" void ev_project_4param(int n, const uint32_t *fr_x, const uint32_t *fr_y, const long long *t_local, double *pr_x,
                                  double *pr_y, double *nx, double *ny, double dnx, double dny, double cx, double cy,
                                  double div, double crl) {
    for (int i = 0; i < n; ++i) {
        Event e(fr_x[i], fr_y[i], 0);
        e.t = t_local[i];
        e.pr_x = pr_x[i]; e.pr_y = pr_y[i];
        e.project_4param_reinit(dnx, dny, cx, cy, div, crl);
        pr_x[i] = e.pr_x; pr_y[i] = e.pr_y; nx[i] = e.nx; ny[i] = e.ny;
    }
}

// Event::compute_uv (event.h:135-142)
extern "C" void ev_compute_uv(int n, const double *nx, const double *ny, double *u, double *v) {
    for (int i = 0; i < n; ++i) {
        Event e(0, 0, 0);
        e.nx = nx[i]; e.ny = ny[i];
        e.compute_uv();
        u[i] = e.u; v[i] = e.v;
    }
}

// Event::set_local_time (event.h:61-63) and Event::project (event.h:144-149 via apply_project)
extern "C" void ev_project(int n, const uint32_t *fr_x, const uint32_t *fr_y, const unsigned long long *ts,
                           unsigned long long t0, double nx, double ny, long long *t_local, double *pr_x, double *pr_y) {
    for (int i = 0; i < n; ++i) {
        Event e(fr_x[i], fr_y[i], ts[i]);
        e.set_local_time(t0);
        e.project(nx, ny);
        t_local[i] = e.t; pr_x[i] = e.pr_x; pr_y[i] = e.pr_y;
    }
}
