// TEST DOUBLE of the C ABI (include/bf_cuda.h) for the CPU test-suite: the entry points the host mirror calls,
// computed by the ORACLE (oracle/bf_oracle.c, the C restatement pinned bit for bit on the compiled reference).
// It exists so that the HOST side of the drop-in -- DVS_flow / OptimizerRolling / OptimizerLocal plumbing: slice
// hand-over, warm start through last_model, noise marking, per-event write-back, compute_uv, batching -- can be
// checked against the reference's own DVS_flow without a GPU.  Never shipped, never linked by the product.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <algorithm>
#include <vector>

#include "../../include/bf_cuda.h"

extern "C" {
// oracle/bf_oracle.c
typedef struct { int x_min, x_max, y_min, y_max, wsize_x, wsize_y, img_rows, img_cols; double x_shift, y_shift; } bfo_setup;
int bfo_minimize(int n, const uint16_t *fr_x, const uint16_t *fr_y, const int64_t *t, uint8_t *noise, int res_x, int res_y,
                 int scale, int max_iter, const double *init_model, int accum_mode, double *out_model, int *out_iters,
                 bfo_setup *out_setup, float *out_div, double *pr_out);
int bfo_local_minimize(int n, const uint16_t *fr_x, const uint16_t *fr_y, const int64_t *t, int res_x, int res_y, int scale,
                       double *out10, int *out_steps, uint8_t *out_img, double *out_pr);
}

struct MockSlice {
    std::vector<uint16_t> fx, fy;
    std::vector<int64_t> t;
    std::vector<uint8_t> noise;
    int scale, max_iter;
    bool has_init;
    bf_model init;
    bf_slice_result res;
    std::vector<double> pr;     // pr_x | pr_y | nx | ny
};
struct bf_ring;
struct bf_ctx {
    int rows, cols;
    std::vector<MockSlice> batch;
    long long launches = 0;
    std::vector<bf_ring *> rings;         // destroyed with the context, as in the library
};
static std::string g_err;
static long long g_minimize_calls = 0, g_batch_runs = 0;

static void model_to_array(const bf_model &m, double *a) {
    a[0] = m.cx; a[1] = m.cy; a[2] = m.dx; a[3] = m.dy; a[4] = m.rot; a[5] = m.div; a[6] = (double)m.cnt;
    a[7] = m.total_dx; a[8] = m.total_dy; a[9] = m.total_rot; a[10] = m.total_div;
}
static void array_to_model(const double *a, bf_model &m) {
    memset(&m, 0, sizeof m);
    m.cx = a[0]; m.cy = a[1]; m.dx = a[2]; m.dy = a[3]; m.rot = a[4]; m.div = a[5]; m.cnt = (uint32_t)a[6];
    m.total_dx = a[7]; m.total_dy = a[8]; m.total_rot = a[9]; m.total_div = a[10];
}

static int g_null = 0;   // 1: skip the compute (host-overhead timing of the mirror)

static void run_slice(bf_ctx *c, MockSlice &s) {
    const int n = (int)s.fx.size();
    if (g_null) {
        memset(&s.res, 0, sizeof s.res);
        s.res.n_events = n;
        s.pr.assign(4 * (size_t)(n > 0 ? n : 1), 0.25);
        return;
    }
    double init[11], out[11];
    if (s.has_init) model_to_array(s.init, init);
    std::vector<uint8_t> nz = s.noise;
    bool any_clear = false;
    for (uint8_t v : nz) any_clear |= v == 0;
    int iters = 0;
    bfo_setup su;
    float div[4];
    s.pr.assign(4 * (size_t)(n > 0 ? n : 1), 0.0);
    const int rc = bfo_minimize(n, s.fx.data(), s.fy.data(), s.t.data(), nz.data(), c->rows, c->cols, s.scale, s.max_iter,
                                s.has_init ? init : nullptr, /*accum_mode=*/0, out, &iters, &su, div, s.pr.data());
    memset(&s.res, 0, sizeof s.res);
    array_to_model(out, s.res.model);
    s.res.rc = rc; s.res.iters = iters;
    memcpy(s.res.dividers, div, sizeof div);
    s.res.x_min = su.x_min; s.res.x_max = su.x_max; s.res.y_min = su.y_min; s.res.y_max = su.y_max;
    s.res.img_rows = su.img_rows; s.res.img_cols = su.img_cols; s.res.x_shift = su.x_shift; s.res.y_shift = su.y_shift;
    s.res.n_events = n;
    bool all_set = n > 0;
    for (uint8_t v : nz) all_set &= v != 0;
    s.res.flags = (rc == 1 && all_set && any_clear) ? BF_FLAG_ALL_NOISE : 0u;   // the tiny-window guard marked every event
}

extern "C" {

int bf_cuda_init(int) {
    if (const char *e = std::getenv("BF_MOCK_NULL")) g_null = atoi(e);   // host-overhead timing of the tool: no compute
    return BF_OK;
}
int bf_device_count(void) { return 1; }
const char *bf_last_error(void) { return g_err.c_str(); }
const char *bf_version(void) { return "mock (oracle back end, tests only)"; }
long long bf_mock_minimize_calls(void) { return g_minimize_calls; }
long long bf_mock_batch_runs(void) { return g_batch_runs; }
void bf_mock_set_null(int v) { g_null = v; }

bf_ctx *bf_ctx_create(int rows, int cols, int, long long, int) {
    bf_ctx *c = new bf_ctx;
    c->rows = rows; c->cols = cols;
    return c;
}
void bf_ring_destroy(bf_ring *r);
void bf_ctx_destroy(bf_ctx *c) {
    if (!c) return;
    while (!c->rings.empty()) bf_ring_destroy(c->rings.back());   // (each removes itself from the list)
    delete c;
}
int bf_ctx_set_option(bf_ctx *, const char *, long long) { return BF_OK; }
long long bf_ctx_get_option(bf_ctx *, const char *) { return 0; }

int bf_batch_reset(bf_ctx *c) { c->batch.clear(); return BF_OK; }

int bf_batch_add_packed(bf_ctx *c, const bf_event *ev, int n, int scale, int max_iter, const bf_model *init) {
    MockSlice s;
    s.fx.resize(n); s.fy.resize(n); s.t.resize(n); s.noise.resize(n);
    for (int i = 0; i < n; ++i) {
        s.fx[i] = ev[i].fr_x; s.fy[i] = (uint16_t)(ev[i].fr_y & 0x7fffu); s.t[i] = ev[i].t_ns;
        s.noise[i] = (ev[i].fr_y & BF_EVENT_NOISE) ? 1 : 0;
        if (s.fx[i] >= c->rows || s.fy[i] >= c->cols) { g_err = "event outside the sensor"; return BF_ERR_ARG; }
    }
    s.scale = scale; s.max_iter = max_iter; s.has_init = init != nullptr;
    if (init) s.init = *init;
    c->batch.push_back(std::move(s));
    return (int)c->batch.size() - 1;
}

int bf_batch_run(bf_ctx *c, int) {
    for (MockSlice &s : c->batch) run_slice(c, s);
    c->launches += 1;
    g_batch_runs += 1;
    return BF_OK;
}
int bf_batch_size(bf_ctx *c) { return (int)c->batch.size(); }
int bf_batch_result(bf_ctx *c, int slot, bf_slice_result *out) {
    if (slot < 0 || slot >= (int)c->batch.size()) { g_err = "bad slot"; return BF_ERR_ARG; }
    *out = c->batch[slot].res;
    return BF_OK;
}
int bf_batch_events(bf_ctx *c, int slot, double *pr_x, double *pr_y, double *nx, double *ny) {
    if (slot < 0 || slot >= (int)c->batch.size()) { g_err = "bad slot"; return BF_ERR_ARG; }
    const MockSlice &s = c->batch[slot];
    const size_t n = s.fx.size();
    if (pr_x) memcpy(pr_x, s.pr.data(), n * sizeof(double));
    if (pr_y) memcpy(pr_y, s.pr.data() + n, n * sizeof(double));
    if (nx) memcpy(nx, s.pr.data() + 2 * n, n * sizeof(double));
    if (ny) memcpy(ny, s.pr.data() + 3 * n, n * sizeof(double));
    return BF_OK;
}

int bf_minimize(bf_ctx *c, const uint16_t *fr_x, const uint16_t *fr_y, const int32_t *t_ns, const uint8_t *noise, int n,
                int scale, int max_iter, const bf_model *init, bf_slice_result *out, double *pr_x, double *pr_y, double *nx,
                double *ny) {
    MockSlice s;
    s.fx.assign(fr_x, fr_x + n); s.fy.assign(fr_y, fr_y + n);
    s.t.resize(n); s.noise.resize(n);
    for (int i = 0; i < n; ++i) { s.t[i] = t_ns[i]; s.noise[i] = noise ? noise[i] : 0; }
    s.scale = scale; s.max_iter = max_iter; s.has_init = init != nullptr;
    if (init) s.init = *init;
    run_slice(c, s);
    g_minimize_calls += 1;
    if (out) *out = s.res;
    const size_t m = (size_t)n;
    if (pr_x) memcpy(pr_x, s.pr.data(), m * sizeof(double));
    if (pr_y) memcpy(pr_y, s.pr.data() + m, m * sizeof(double));
    if (nx) memcpy(nx, s.pr.data() + 2 * m, m * sizeof(double));
    if (ny) memcpy(ny, s.pr.data() + 3 * m, m * sizeof(double));
    return s.res.rc;
}

int bf_local_minimize(bf_ctx *c, const uint16_t *fr_x, const uint16_t *fr_y, const int32_t *t_ns, int n, int scale,
                      bf_slice_result *out) {
    std::vector<int64_t> t(t_ns, t_ns + n);
    double o[10];
    int steps = 0;
    const int rc = bfo_local_minimize(n, fr_x, fr_y, t.data(), c->rows, c->cols, scale, o, &steps, nullptr, nullptr);
    bf_slice_result r;
    memset(&r, 0, sizeof r);
    r.model.total_dx = o[0]; r.model.total_dy = o[1]; r.model.dx = o[2]; r.model.dy = o[3]; r.model.rot = o[4]; r.model.div = o[5];
    r.rc = rc; r.iters = steps; r.n_events = n;
    if (out) *out = r;
    return rc;
}

// ---- device-resident slice ring, restated on the host with the oracle as the minimiser ----
}  // extern "C"
struct bf_ring {
    bf_ctx *c;
    long long cap;
    int max_pending;
    std::vector<bf_ring_event> ring;      // event g of the stream at g % cap
    long long pushed = 0;
    std::vector<bf_slice_result> res;     // by ticket % max_pending
    int next_ticket = 0;
    bool have_prev = false;
    bf_slice_result prev;
    long long prev_lo = 0, prev_hi = 0;
    std::vector<bf_ring_event> stage;     // bf_ring_reserve / bf_ring_commit
    int res_n = 0;
};
static long long g_ring_pushes = 0, g_ring_slices = 0;
extern "C" {
long long bf_mock_ring_pushed_events(void) { return g_ring_pushes; }
long long bf_mock_ring_slices(void) { return g_ring_slices; }

bf_ring *bf_ring_create(bf_ctx *c, long long capacity, int max_pending) {
    if (!c || capacity <= 0 || max_pending < 2) { g_err = "bf_ring_create: bad arguments"; return nullptr; }
    bf_ring *r = new bf_ring;
    r->c = c; r->cap = capacity; r->max_pending = max_pending;
    r->ring.resize((size_t)capacity);
    r->res.resize((size_t)max_pending);
    memset(&r->prev, 0, sizeof r->prev);
    c->rings.push_back(r);
    return r;
}
void bf_ring_destroy(bf_ring *r) {
    if (!r) return;
    auto &v = r->c->rings;
    v.erase(std::remove(v.begin(), v.end(), r), v.end());
    delete r;
}
long long bf_ring_pushed(bf_ring *r) { return r ? r->pushed : 0; }
int bf_ring_push(bf_ring *r, const bf_ring_event *ev, int n) {
    if (g_null) { r->pushed += n; g_ring_pushes += n; return BF_OK; }   // host-overhead timing
    for (int i = 0; i < n; ++i) {
        if ((int)ev[i].fr_x >= r->c->rows || (int)(ev[i].fr_y & 0x7fffu) >= r->c->cols) { g_err = "event outside the sensor"; return BF_ERR_ARG; }
        r->ring[(size_t)(r->pushed % r->cap)] = ev[i];
        r->pushed += 1;
    }
    g_ring_pushes += n;
    return BF_OK;
}
int bf_ring_reserve(bf_ring *r, int n, bf_ring_event **where) {
    if (!r || !where || n <= 0) { g_err = "bf_ring_reserve: bad arguments"; return BF_ERR_ARG; }
    if (r->stage.size() < (size_t)n) r->stage.resize((size_t)n);
    r->res_n = n;
    *where = r->stage.data();
    return BF_OK;
}
int bf_ring_commit(bf_ring *r, int n) {
    if (!r || n < 0 || n > r->res_n) { g_err = "bf_ring_commit: n exceeds the reservation"; return BF_ERR_ARG; }
    r->res_n = 0;
    return bf_ring_push(r, r->stage.data(), n);
}
int bf_ring_slice(bf_ring *r, int n, uint64_t slice_start, int scale, int max_iter, int chain) {
    if (n < 0 || n > r->cap || n > r->pushed) { g_err = "bf_ring_slice: n exceeds the ring's content"; return BF_ERR_ARG; }
    if (g_null) {                          // host-overhead timing: no gather, no compute
        const int ticket = r->next_ticket++;
        memset(&r->res[(size_t)(ticket % r->max_pending)], 0, sizeof(bf_slice_result));
        r->res[(size_t)(ticket % r->max_pending)].n_events = n;
        g_ring_slices += 1;
        return ticket;
    }
    MockSlice s;
    s.fx.resize(n); s.fy.resize(n); s.t.resize(n); s.noise.resize(n);
    const bool prev_noise = r->have_prev && (r->prev.flags & BF_FLAG_ALL_NOISE);
    for (int i = 0; i < n; ++i) {
        const long long g = r->pushed - 1 - i;
        bf_ring_event &e = r->ring[(size_t)(g % r->cap)];
        if (prev_noise && g >= r->prev_lo && g < r->prev_hi) e.fr_y |= BF_EVENT_NOISE;
        s.fx[i] = e.fr_x; s.fy[i] = (uint16_t)(e.fr_y & 0x7fffu); s.noise[i] = (e.fr_y & BF_EVENT_NOISE) ? 1 : 0;
        s.t[i] = (int64_t)(e.timestamp - slice_start);
    }
    s.scale = scale; s.max_iter = max_iter; s.has_init = chain != 0;
    if (chain) s.init = r->prev.model;     // (all-zero for the first slice: last_model of a fresh DVS_flow)
    run_slice(r->c, s);
    const int ticket = r->next_ticket++;
    r->res[(size_t)(ticket % r->max_pending)] = s.res;
    r->prev = s.res; r->have_prev = true; r->prev_lo = r->pushed - n; r->prev_hi = r->pushed;
    g_ring_slices += 1;
    return ticket;
}
int bf_ring_result(bf_ring *r, int ticket, bf_slice_result *out) {
    if (ticket < 0 || ticket >= r->next_ticket || ticket < r->next_ticket - r->max_pending) { g_err = "bad ticket"; return BF_ERR_ARG; }
    *out = r->res[(size_t)(ticket % r->max_pending)];
    return BF_OK;
}
int bf_ring_seed(bf_ring *r, const bf_model *m) {
    if (!r || !m) { g_err = "bf_ring_seed: bad arguments"; return BF_ERR_ARG; }
    memset(&r->prev, 0, sizeof r->prev);
    r->prev.model = *m;
    r->have_prev = false;                  // (no noise marks carried over)
    return BF_OK;
}
int bf_ring_sync(bf_ring *) { return BF_OK; }

// entry points the mirror references but these tests never reach
static int unavailable(const char *what) { g_err = std::string(what) + ": not part of the CPU test double"; return BF_ERR_STATE; }
int bf_time_img(bf_ctx *, int, const double *, const double *, const int32_t *, const uint8_t *, int, int, int, int, int, float *) { return unavailable("bf_time_img"); }
int bf_projection_img(bf_ctx *, int, const double *, const double *, const uint8_t *, int, uint8_t *, double *) { return unavailable("bf_projection_img"); }
int bf_color_time_img(bf_ctx *, int, const double *, const double *, const int32_t *, const uint8_t *, int, uint8_t *) { return unavailable("bf_color_time_img"); }
int bf_project(bf_ctx *, int, const uint16_t *, const uint16_t *, const int32_t *, double *, double *, double *, double *, double, double, double, double, double, double) { return unavailable("bf_project"); }
int bf_model_from_image(bf_ctx *, int, int, const float *, double *, float *, float *) { return unavailable("bf_model_from_image"); }
int bf_multi_owner(int slice, int n_devices, int block) { return (n_devices <= 0 || block <= 0 || slice < 0) ? -1 : (slice / block) % n_devices; }
bf_multi *bf_multi_create(int, const int *, int, int, int, long long, int) { unavailable("bf_multi_create"); return nullptr; }
void bf_multi_destroy(bf_multi *) {}
int bf_multi_reset(bf_multi *) { return unavailable("bf_multi"); }
int bf_multi_set_option(bf_multi *, const char *, long long) { return unavailable("bf_multi"); }
int bf_multi_add_packed(bf_multi *, const bf_event *, int, int, int) { return unavailable("bf_multi"); }
int bf_multi_run(bf_multi *, int) { return unavailable("bf_multi"); }
int bf_multi_sync(bf_multi *) { return unavailable("bf_multi"); }
int bf_multi_result(bf_multi *, int, bf_slice_result *) { return unavailable("bf_multi"); }
int bf_multi_locate(bf_multi *, int, bf_ctx **, int *, int *) { return unavailable("bf_multi"); }

}  // extern "C"
