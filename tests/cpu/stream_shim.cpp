// Whole-stream runs of the HOST mirror (better_flow/dvs_flow.h with OptimizerRolling / OptimizerLocal) linked
// against the oracle-backed test double of the C ABI (mock_bf_cuda.cpp) instead of the CUDA library: everything
// the drop-in does on the host -- triggers, ring buffer, slice hand-over, warm start via last_model, noise
// marking, per-event write-back and compute_uv, batching of independent slices -- with an exact back end, so the
// per-slice models must equal the compiled reference's DVS_flow bit for bit.  CPU test-suite only.
#include <better_flow/dvs_flow.h>

namespace {

constexpr int UV = 14;

template <class Flow> struct FlowProbe : Flow {
    using Flow::Flow;
    size_t remembered() const { return this->motion_memory.size(); }
    const ObjectModel &remembered_model(size_t k) const { return this->motion_memory[k].model; }
    size_t remembered_size(size_t k) const { return this->motion_memory[k].size; }
};

void model_out(const ObjectModel &m, double *a) {
    const bf_model p = m.to_pod();
    a[0] = p.cx; a[1] = p.cy; a[2] = p.dx; a[3] = p.dy; a[4] = p.rot; a[5] = p.div; a[6] = (double)p.cnt;
    a[7] = p.total_dx; a[8] = p.total_dy; a[9] = p.total_rot; a[10] = p.total_div;
}

// models: 11 doubles per slice; info: 3 per slice = {events consumed, buffer size, buffer time diff} (unbatched runs);
// uv: UV doubles per slice = weighted sums of the per-event fields in the buffer right after the slice (what -o / the viewers read)
template <class Flow>
int run(Flow &est, int n, const uint32_t *fr_x, const uint32_t *fr_y, const uint64_t *ts, int scale, int max_iter, int stm_disable,
        int flush, int batch, int local, int lazy, int max_slices, double *models, long long *info, double *uv) {
    est.set_lazy_events(lazy != 0);
    est.set_device_ring(lazy >= 2);          // lazy == 2: the slice ring lives behind the C ABI (bf_ring_*)
                                             // lazy == 3: ... and is switched off at n / 3 and on again at 2 n / 3
                                             // lazy == 4: ... and another user of the pooled context re-creates it at n / 2
                                             //            (the ring and its staging buffer are gone from one event to the next)
    est.set_scale(scale);
    est.set_max_iter(max_iter);
    est.set_stm_disable(stm_disable != 0);
    est.set_quiet(true);
    if (batch > 1) est.set_batch(batch);
    if (local) est.set_optimizer_local(true);
    int ns = 0;
    auto record = [&](long long consumed) {
        if (ns < max_slices && batch <= 1) {
            model_out(est.get_last_model(), models + 11 * ns);
            info[3 * ns + 0] = consumed;
            info[3 * ns + 1] = est.get_buf_size();
            info[3 * ns + 2] = est.get_buf_time_diff();
            // position-weighted sums of every per-event field the slice leaves behind in the buffer
            if (uv == nullptr) { ++ns; return; }
            double c[UV] = {0};
            double w = 1.0;
            for (auto &e : est.ev_buffer) {
                c[0] += w * e.u; c[1] += w * e.v; c[2] += w * e.pr_x; c[3] += w * e.pr_y; c[4] += w * e.nx; c[5] += w * e.ny;
                c[6] += w * e.best_u; c[7] += w * e.best_v; c[8] += w * e.best_pr_x; c[9] += w * e.best_pr_y;
                c[10] += w * e.max_score; c[11] += w * (double)e.t; c[12] += e.noise ? w : 0.0; c[13] += 1.0;
                w += 1e-3;
            }
            for (int k = 0; k < UV; ++k) uv[UV * ns + k] = c[k];
        }
        ++ns;
    };
    for (int i = 0; i < n; ++i) {
        if (lazy == 3 && i == n / 3) est.set_device_ring(false);
        if (lazy == 3 && i == 2 * (n / 3)) est.set_device_ring(true);
        if (lazy == 4 && i == n / 2) {
            est.get_last_model();                                  // (settle the outstanding tickets first, as such a user must)
            CudaDriver::context(4000000, 3, scale);                // does not fit the pooled context: destroyed and re-created
        }
        Event e(fr_x[i], fr_y[i], ts[i]);
        if (est.add_event(e)) record(i + 1);
    }
    if (flush) { est.recompute(); record(n); }
    if (batch > 1) {
        est.flush();
        for (size_t k = 0; k < est.remembered() && (int)k < max_slices; ++k) {
            model_out(est.remembered_model(k), models + 11 * k);
            info[3 * k + 1] = (long long)est.remembered_size(k);
        }
    }
    return ns;
}

}  // namespace

extern "C" int st_stream(int config, int n, const uint32_t *fr_x, const uint32_t *fr_y, const uint64_t *ts,
                         unsigned long long ev_refresh, unsigned long long time_refresh_ns, int scale, int max_iter, int stm_disable,
                         int flush, int batch, int local, int lazy, int max_slices, double *models, long long *info, double *uv) {
    bf::set_sensor(180, 240);
    if (config == 0) {
        FlowProbe<DVS_flow<50000, FROM_SEC(0.2)>> est(ev_refresh, time_refresh_ns);
        return run(est, n, fr_x, fr_y, ts, scale, max_iter, stm_disable, flush, batch, local, lazy, max_slices, models, info, uv);
    }
    if (config == 1) {
        FlowProbe<DVS_flow<30000, FROM_MS(70)>> est(ev_refresh, time_refresh_ns);
        return run(est, n, fr_x, fr_y, ts, scale, max_iter, stm_disable, flush, batch, local, lazy, max_slices, models, info, uv);
    }
    return -1;
}
