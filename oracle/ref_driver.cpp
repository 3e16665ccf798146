// TEST INFRASTRUCTURE ONLY -- never linked into, loaded by, or called from the product.
//
// C-ABI driver around the reference's OWN, UNMODIFIED classes (compiled from
// /root/reference by oracle/Makefile into oracle/_ref/libbf_ref_<rows>x<cols>.so).
// It exposes, for ctypes:
//   * bf_ref_minimize   -- OptimizerRolling<LinearEventPtrs>: set_cloud -> set_time ->
//                          set_maxiter -> [set_model] -> run -> get_model, exactly the
//                          sequence DVS_flow::recompute performs (dvs_flow.h:210-224);
//   * bf_ref_time_img   -- AccelLib::get_time_img_cpu (accel_lib.h:147-178);
//   * bf_ref_model      -- ObjectModel::update(Mat)  (object_model.h:31-34) and
//                          AccelLib::Sobel_cpu       (accel_lib.h:513-543);
//   * bf_ref_project    -- Event::project_4param_reinit (event.h:99-110);
//   * bf_ref_stream     -- DVS_flow<50000,200ms> / <30000,70ms> fed event by event,
//                          returning the per-slice models (dvs_flow.h:163-252);
//   * bf_ref_local      -- OptimizerLocal(LinearEventCloud*, scale)::run(), the contrast-driven
//                          (nx, ny) coordinate descent (optimizer_sampler.cpp:4-38); its
//                          cv::GaussianBlur is the shim's restatement (see oracle/shim);
//   * bf_ref_blur       -- that GaussianBlur stand-in alone, for the check against real cv2.
// The only thing the driver adds is plumbing: it builds Event objects from SoA
// arrays, reads results out of public/protected members through a derived class,
// and recovers run()'s iteration count from the TBB stand-in's call counter.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <unistd.h>
#include <fcntl.h>

#include <better_flow/common.h>
#include <better_flow/event.h>
#include <better_flow/object_model.h>
#include <better_flow/accel_lib.h>
#include <better_flow/optimizer_rolling.h>
#include <better_flow/optimizer_sampler.h>
#include <better_flow/dvs_flow.h>

namespace {

// The reference prints to std::cout from recompute() (O(slices^2), dvs_flow.h:245-252);
// silence fd 1 for the duration of a call.
struct StdoutSilencer {
    int saved;
    StdoutSilencer() {
        std::cout.flush();
        fflush(stdout);
        saved = dup(1);
        int nul = open("/dev/null", O_WRONLY);
        dup2(nul, 1);
        close(nul);
    }
    ~StdoutSilencer() {
        std::cout.flush();
        fflush(stdout);
        dup2(saved, 1);
        close(saved);
    }
};

struct Probe : public OptimizerRolling<LinearEventPtrs> {
    void setup_out(int *ints, double *dbls) {
        ints[0] = x_min; ints[1] = x_max; ints[2] = y_min; ints[3] = y_max;
        ints[4] = metric_wsizex; ints[5] = metric_wsizey;
        ints[6] = scale_img_x; ints[7] = scale_img_y;
        dbls[0] = x_shift; dbls[1] = y_shift;
    }
    void dividers_out(float *d) {
        d[0] = x_divider; d[1] = y_divider; d[2] = rot_divider; d[3] = div_divider;
    }
};

void model_to_array(const ObjectModel &m, double *o) {
    o[0] = m.cx; o[1] = m.cy; o[2] = m.dx; o[3] = m.dy; o[4] = m.rot; o[5] = m.div;
    o[6] = double(m.cnt);
    o[7] = m.total_dx; o[8] = m.total_dy; o[9] = m.total_rot; o[10] = m.total_div;
}

ObjectModel model_from_array(const double *a) {
    ObjectModel m;
    m.cx = a[0]; m.cy = a[1]; m.dx = a[2]; m.dy = a[3]; m.rot = a[4]; m.div = a[5];
    m.cnt = uint(a[6]);
    m.total_dx = a[7]; m.total_dy = a[8]; m.total_rot = a[9]; m.total_div = a[10];
    return m;
}

struct LocalProbe : public OptimizerLocal {
    using OptimizerLocal::OptimizerLocal;
    void state_out(double *o) {
        o[0] = nx; o[1] = ny; o[2] = last_score; o[3] = dnx; o[4] = dny; o[5] = dn_th;
        o[6] = metric_wsizex; o[7] = metric_wsizey; o[8] = scale_img_x; o[9] = scale_img_y;
    }
    const cv::Mat &image() const { return project_img; }
};

template <class Flow> struct FlowProbe : public Flow {
    using Flow::Flow;
    ObjectModel model() { return this->last_model; }
};

template <class Flow>
int run_stream(int n, const uint32_t *fr_x, const uint32_t *fr_y, const uint64_t *ts,
               unsigned long long ev_refresh, unsigned long long time_refresh_ns,
               int scale, int max_iter, int stm_disable, int flush,
               int max_slices, double *models, long long *slice_info) {
    FlowProbe<Flow> est(ev_refresh, time_refresh_ns);
    est.set_scale(scale);
    est.set_max_iter(max_iter);
    est.set_stm_disable(stm_disable != 0);
    int ns = 0;
    for (int i = 0; i < n; ++i) {
        Event e(fr_x[i], fr_y[i], ts[i]);
        bool processed = est.add_event(e);
        if (processed) {
            if (ns < max_slices) {
                model_to_array(est.model(), models + 11 * ns);
                slice_info[3 * ns + 0] = i + 1;
                slice_info[3 * ns + 1] = est.get_buf_size();
                slice_info[3 * ns + 2] = est.get_buf_time_diff();
            }
            ++ns;
        }
    }
    if (flush) {
        est.recompute();
        if (ns < max_slices) {
            model_to_array(est.model(), models + 11 * ns);
            slice_info[3 * ns + 0] = n;
            slice_info[3 * ns + 1] = est.get_buf_size();
            slice_info[3 * ns + 2] = est.get_buf_time_diff();
        }
        ++ns;
    }
    return ns;
}

}  // namespace

extern "C" {

// Compiled-in sensor size (common.h:39-40) and the thread count the TBB stand-in uses.
void bf_ref_info(int *res_x, int *res_y, int *threads) {
    *res_x = RES_X;
    *res_y = RES_Y;
    *threads = tbb::bf_shim_threads();
}

// One slice through OptimizerRolling.  Events are given in the order the optimiser
// iterates them (DVS_flow passes newest -> oldest, dvs_flow.h:196-198).
//   ts[i]        absolute timestamp (ns);  slice_start: passed to set_time
//   noise        nullable, per-event initial `noise` flag
//   init_model   nullable (== --stm-disable); 11 doubles cx,cy,dx,dy,rot,div,cnt,total_dx,total_dy,total_rot,total_div
//   out_setup_i  8 ints: x_min,x_max,y_min,y_max,metric_wsizex,metric_wsizey,scale_img_x,scale_img_y
//   out_setup_d  2 doubles: x_shift,y_shift;  out_div: 4 final dividers
//   out_pr       nullable, 4*n doubles: pr_x[n], pr_y[n], nx[n], ny[n] after run()
// Returns run()'s return value (0 optimised, 1 skipped).
int bf_ref_minimize(int n, const uint32_t *fr_x, const uint32_t *fr_y, const uint64_t *ts,
                    const uint8_t *noise, uint64_t slice_start, int scale, int max_iter,
                    const double *init_model, double *out_model, int *out_iters,
                    int *out_setup_i, double *out_setup_d, float *out_div,
                    double *out_pr, uint8_t *out_noise, double *out_seconds) {
    StdoutSilencer quiet;
    std::vector<Event> store;
    store.reserve(n);
    for (int i = 0; i < n; ++i) {
        store.emplace_back(fr_x[i], fr_y[i], ts[i]);
        if (noise) store.back().noise = noise[i] != 0;
    }
    LinearEventPtrs ptrs;
    for (auto &e : store) ptrs.push_back(&e);

    Probe opt;
    opt.set_cloud(&ptrs, scale);
    opt.set_time(slice_start);
    opt.set_maxiter(max_iter);
    if (init_model) opt.set_model(model_from_array(init_model));

    const unsigned long long c0 = tbb::bf_shim_call_count();
    auto t0 = std::chrono::steady_clock::now();
    int rc = opt.run();
    auto t1 = std::chrono::steady_clock::now();
    const unsigned long long c1 = tbb::bf_shim_call_count();

    if (out_seconds) *out_seconds = std::chrono::duration<double>(t1 - t0).count();
    if (out_iters) *out_iters = int((c1 - c0) / 2);
    if (out_model) model_to_array(opt.get_model(), out_model);
    if (out_setup_i && out_setup_d) opt.setup_out(out_setup_i, out_setup_d);
    if (out_div) opt.dividers_out(out_div);
    if (out_pr) {
        for (int i = 0; i < n; ++i) {
            out_pr[i] = store[i].pr_x;
            out_pr[n + i] = store[i].pr_y;
            out_pr[2 * n + i] = store[i].nx;
            out_pr[3 * n + i] = store[i].ny;
        }
    }
    if (out_noise)
        for (int i = 0; i < n; ++i) out_noise[i] = store[i].noise ? 1 : 0;
    return rc;
}

// AccelLib::get_time_img_cpu on events with explicit warped positions and local times.
// out must hold (w+scale)*(h+scale) floats (row-major, rows = w+scale).
void bf_ref_time_img(int n, const double *pr_x, const double *pr_y, const int64_t *t_local,
                     const uint8_t *noise, int w, int h, int scale, int x_sh, int y_sh, float *out) {
    std::vector<Event> store(n, Event(0, 0, 0));
    for (int i = 0; i < n; ++i) {
        store[i].pr_x = pr_x[i];
        store[i].pr_y = pr_y[i];
        store[i].t = t_local[i];
        store[i].noise = noise ? (noise[i] != 0) : false;
    }
    LinearEventPtrs ptrs;
    for (auto &e : store) ptrs.push_back(&e);
    cv::Mat img = AccelLib::get_time_img_cpu<LinearEventPtrs>(&ptrs, w, h, scale, x_sh, y_sh);
    std::memcpy(out, img.data, size_t(img.rows) * img.cols * sizeof(float));
}

// ObjectModel::update(Mat) (center_of_mass + compute) on a given time image; optionally
// also returns the two Scharr images from AccelLib::Sobel_cpu.
// out7: cx, cy, dx, dy, rot, div, cnt
void bf_ref_model(int rows, int cols, const float *img, double *out7, float *gx, float *gy) {
    cv::Mat m(rows, cols, CV_32FC1);
    std::memcpy(m.data, img, size_t(rows) * cols * sizeof(float));
    ObjectModel model;
    model.update(m);
    out7[0] = model.cx; out7[1] = model.cy; out7[2] = model.dx; out7[3] = model.dy;
    out7[4] = model.rot; out7[5] = model.div; out7[6] = double(model.cnt);
    if (gx && gy) {
        cv::Mat grad_x, grad_y;
        AccelLib::Sobel_cpu(m, grad_x, grad_y);
        std::memcpy(gx, grad_x.data, size_t(rows) * cols * sizeof(float));
        std::memcpy(gy, grad_y.data, size_t(rows) * cols * sizeof(float));
    }
}

// Event::project_4param_reinit over n events (in place on pr_x/pr_y; nx/ny out).
void bf_ref_project(int n, const uint32_t *fr_x, const uint32_t *fr_y, const int64_t *t_local,
                    double *pr_x, double *pr_y, double *nx, double *ny,
                    double dnx, double dny, double cx, double cy, double div, double crl) {
    for (int i = 0; i < n; ++i) {
        Event e(fr_x[i], fr_y[i], 0);
        e.t = t_local[i];
        e.pr_x = pr_x[i];
        e.pr_y = pr_y[i];
        e.project_4param_reinit(dnx, dny, cx, cy, div, crl);
        pr_x[i] = e.pr_x; pr_y[i] = e.pr_y; nx[i] = e.nx; ny[i] = e.ny;
    }
}

// Event::compute_uv (event.h:135-142): u,v from nx,ny.
void bf_ref_compute_uv(int n, const double *nx, const double *ny, double *u, double *v) {
    for (int i = 0; i < n; ++i) {
        Event e(0, 0, 0);
        e.nx = nx[i]; e.ny = ny[i];
        e.compute_uv();
        u[i] = e.u; v[i] = e.v;
    }
}

// Whole-stream run through DVS_flow::add_event (triggers, ring buffer, warm start).
// config 0: DVS_flow<50000, 200 ms> (bf_motion_compensator.cpp:6-7,135)
// config 1: DVS_flow<30000,  70 ms> (ros_nodes_src/bf_visualizer.cpp:30-34)
// models: 11 doubles per slice; slice_info: 3 int64 per slice = {events consumed, buffer size, buffer time diff}
// Returns number of slices computed (may exceed max_slices; only the first max_slices are stored).
// OptimizerLocal on one cloud.  t_local[i] becomes Event::t (the class never calls set_local_time;
// the caller's t is what Event::project uses, event.h:164-168).
//   out10   nx, ny, last_score, dnx, dny, dn_th, metric_wsizex, metric_wsizey, scale_img_x, scale_img_y
//   out_img nullable: the CV_8UC1 project_img of the LAST iteration_step, scale_img_x * scale_img_y bytes
//   out_pr  nullable: 2*n doubles pr_x[n], pr_y[n] after run()
// Returns run()'s value (0 ok, 1 window too small); *out_steps = number of iteration_steps (blur calls; -1 for scale 1).
int bf_ref_local(int n, const uint32_t *fr_x, const uint32_t *fr_y, const int64_t *t_local, int scale,
                 double *out10, int *out_steps, uint8_t *out_img, double *out_pr, double *out_seconds) {
    StdoutSilencer quiet;
    LinearEventCloud cloud;
    for (int i = 0; i < n; ++i) {
        Event e(fr_x[i], fr_y[i], 0);
        e.t = t_local[i];
        cloud.push_back(e);
    }
    LocalProbe opt(&cloud, scale);
    const unsigned long long b0 = cv::bf_shim_blur_calls();
    auto t0 = std::chrono::steady_clock::now();
    const int rc = opt.run();
    auto t1 = std::chrono::steady_clock::now();
    if (out_seconds) *out_seconds = std::chrono::duration<double>(t1 - t0).count();
    if (out_steps) *out_steps = scale > 1 ? int(cv::bf_shim_blur_calls() - b0) : -1;
    if (out10) opt.state_out(out10);
    if (out_img) std::memcpy(out_img, opt.image().data, size_t(opt.image().rows) * opt.image().cols);
    if (out_pr)
        for (int i = 0; i < n; ++i) { out_pr[i] = cloud[i].pr_x; out_pr[n + i] = cloud[i].pr_y; }
    return rc;
}

// The shim's cv::GaussianBlur on a CV_8UC1 image (in place), for validation against cv2.
void bf_ref_blur(int rows, int cols, int ksize, uint8_t *img) {
    cv::Mat m(rows, cols, CV_8UC1);
    std::memcpy(m.data, img, size_t(rows) * cols);
    cv::GaussianBlur(m, m, cv::Size(ksize, ksize), 0, 0);
    std::memcpy(img, m.data, size_t(rows) * cols);
}

int bf_ref_stream(int config, int n, const uint32_t *fr_x, const uint32_t *fr_y, const uint64_t *ts,
                  unsigned long long ev_refresh, unsigned long long time_refresh_ns,
                  int scale, int max_iter, int stm_disable, int flush,
                  int max_slices, double *models, long long *slice_info) {
    StdoutSilencer quiet;
    if (config == 0)
        return run_stream<DVS_flow<50000, FROM_SEC(0.2)>>(n, fr_x, fr_y, ts, ev_refresh, time_refresh_ns, scale,
                                                           max_iter, stm_disable, flush, max_slices, models, slice_info);
    if (config == 1)
        return run_stream<DVS_flow<30000, FROM_MS(70)>>(n, fr_x, fr_y, ts, ev_refresh, time_refresh_ns, scale,
                                                         max_iter, stm_disable, flush, max_slices, models, slice_info);
    return -1;
}

}  // extern "C"
