// better_flow/dvs_flow.h -- slice manager (reference: better_flow_core/include/better_flow/dvs_flow.h).
// Same class template, constructor and public methods: events are pushed one by one, a slice is
// (re)computed when enough new events or enough time have accumulated, the model of the previous
// slice warm-starts the next one unless stm is disabled.
//
// Additions over the reference, all opt-in:
//   * run-time buffer capacity / time span (the template arguments stay the defaults),
//   * set_batch(n): with stm disabled slices are independent, so n of them are queued and minimised
//     by ONE persistent-kernel launch (bf_batch_*),
//   * set_gpus(n): the queued batch is dealt to n devices (bf_multi_*: one launch per device, one
//     NCCL all-gather of the per-slice flow records),
//   * set_optimizer_local(): minimise every slice with OptimizerLocal (contrast-driven nx, ny descent)
//     instead of OptimizerRolling; the slice's model then carries total_dx = -nx, total_dy = -ny,
//   * set_flow_out(stream): one machine-readable line per slice,
//   * set_quiet(): suppress the reference's per-slice dump of every past model,
//   * set_device_ring(): keep the slice ring on the device (bf_ring_*): only new events are uploaded, a slice is an
//     index range, and the warm-start chain is stream-ordered device work -- the host enqueues slices and reads the
//     models back later (at once when the per-slice dump is printed).  Used when no per-event state is wanted
//     (set_lazy_events, no accumulation).  The host then only has to know WHICH events the window holds (triggers,
//     eviction, slice start, a rebuild of the device ring): they are kept as 16-byte records in a ring of the same
//     capacity / span / quirks (`hdr_buffer_`) and `ev_buffer` stays empty; noise marks of the tiny-window guard live
//     on the device only.  Switching the mode with events in the window moves them across.
//   * set_generate_pictures / set_generate_video (the reference's --img / --video, dvs_flow.h:256-335): after every
//     slice a 2 x 2 montage is written -- EventFile::projection_img of the events as recorded and as warped (computed
//     on the device: bf_projection_img), next to EventFile::color_time_img of the same events (bf_color_time_img).
//     Without OpenCV there is no JPEG / AVI encoder and no text rendering: pictures are binary PPM files
//     (frame_N.ppm) with the overlay text of the reference's frames in a sidecar frame_N.txt, the video is an
//     uncompressed YUV4MPEG2 stream (4:4:4), playable by ffplay / mpv.
// The interactive mode is a GUI feature and is accepted but ignored.
#ifndef BF_DVS_FLOW_H
#define BF_DVS_FLOW_H

#include <algorithm>
#include <chrono>
#include <deque>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>

#include <better_flow/common.h>
#include <better_flow/event.h>
#include <better_flow/event_file.h>
#include <better_flow/optimizer_rolling.h>
#include <better_flow/optimizer_sampler.h>

template <size_t MAX_SZ, sll SPAN> class DVS_flow {
public:
    // Buffer for incoming events (aka 'slice')
    CircularArray<Event, MAX_SZ, SPAN> ev_buffer;

protected:
    ull on_ev_change, on_time_change;        // triggers
    sll time_diff, event_diff;               // time passed / new events since the last slice
    ull last_slice_time, current_slice_time;
    ObjectModel last_model;                  // starting point of the next minimisation
    bool accumulate;
    std::vector<LinearEventCloudTemplate<Event>> accumulated;
    bool manual_mode;
    int max_iter;
    int scale;
    bool stm_disable;

    struct SliceLog {                        // what the reference prints per remembered slice (:245-252)
        ObjectModel model;
        size_t size;
        ull ts_first, ts_last;
    };
    std::vector<SliceLog> motion_memory;

    struct Pending {                         // a slice queued for a batched launch
        std::vector<bf_event> packed;
        SliceLog log;
        LinearEventCloudTemplate<Event> copy; // only when accumulating
        ull slice_start;
    };
    std::vector<Pending> pending_;
    std::vector<uint16_t> fx_, fy_;          // SoA staging of the slice handed to the back end (reused across slices)
    std::vector<int32_t> t_;
    std::vector<uint8_t> nz_;
    std::vector<double> px_, py_, nx_, ny_;
    bool lazy_events_ = false;
    int batch_;
    int gpus_;
    bool local_;
    bool quiet_;
    std::ostream *flow_out_;
    ull slices_done_;
    ull events_done_;
    ull iters_done_;
    bool unsorted_ = false;                  // a timestamp decreased somewhere in the input (see get_accumulated)
    ull events_added_ = 0;                   // add_event calls (reserve hint for get_accumulated)
    bool generate_pictures_ = false, generate_video_ = false;
    std::string img_prefix_ = "./", video_name_ = "out.avi";
    int video_fps_ = 30;
    ull frame_count_ = 0;
    std::unique_ptr<std::ofstream> video_out_;
    void dump_frame(size_t n_slice);
    // device-resident ring (set_device_ring)
    bool device_ring_ = false;
    bool ring_on_ = false;                   // device_ring_active(), re-evaluated by every setter it depends on
    // The window as the device ring's own 16-byte records (noise = bit 15 of fr_y, as on the device).
    struct RingHeader : bf_ring_event {
        sll operator-(const RingHeader &rhs) const { return sll(timestamp) - sll(rhs.timestamp); }   // Event::operator-
        void copy_header_to(RingHeader &o) const { o = *this; }
    };
    static RingHeader header_of(const Event &e) {
        RingHeader r;
        r.fr_x = (uint16_t)e.fr_x; r.fr_y = (uint16_t)(e.fr_y | (e.noise ? BF_EVENT_NOISE : 0u)); r.reserved = 0; r.timestamp = e.timestamp;
        return r;
    }
    static_assert(sizeof(RingHeader) == sizeof(bf_ring_event), "RingHeader adds no data");
    RingCore<RingHeader> hdr_buffer_;
    void update_ring_on() {
        const bool on = device_ring_ && lazy_events_ && !accumulate && !local_ && gpus_ == 1 && !(batch_ > 1 && stm_disable);
        if (on != ring_on_) migrate_window(on);
        ring_on_ = on;
    }
    void migrate_window(bool to_headers);
    // the window, whichever ring holds it
    size_t win_size() { return ring_on_ ? hdr_buffer_.size() : ev_buffer.size(); }
    ull win_timestamp(size_t idx) { return ring_on_ ? hdr_buffer_[idx].timestamp : ev_buffer[idx].timestamp; }
    bf_ring *ring_ = nullptr;
    unsigned long ring_gen_ = 0;             // CudaDriver::generation() the ring was created under
    int ring_pending_ = 64;
    std::vector<RingHeader> ring_new_;       // scratch for a rebuild of the device ring
    // New events go straight into the device ring's pinned staging buffer (bf_ring_reserve / bf_ring_commit): `stage_`
    // is the open reservation, valid while the pooled context is the one the ring was created under.
    bf_ring_event *stage_ = nullptr;
    int stage_fill_ = 0, stage_room_ = 0;
    const unsigned long *ctx_gen_ = CudaDriver::generation_ptr();
    static constexpr int kStageChunk = 32768;
    __attribute__((noinline)) void stage_slow(const RingHeader &r);
    void stage_drop() { stage_ = nullptr; stage_fill_ = stage_room_ = 0; }
    struct Deferred { SliceLog log; int ticket; };
    std::deque<Deferred> deferred_;          // slices enqueued on the device whose models have not been read back yet
    double t_push_ = 0, t_slice_ = 0, t_resolve_ = 0;   // host seconds spent in bf_ring_push / bf_ring_slice / bf_ring_result
    double t_recompute_ = 0;                            // ... and in recompute() as a whole (device-ring mode)
public:
    // (BF_TIMING diagnostics of the tool) host time spent inside the ring's three entry points
    void ring_host_seconds(double &push, double &slice, double &resolve) const { push = t_push_; slice = t_slice_; resolve = t_resolve_; }
    double recompute_host_seconds() const { return t_recompute_; }
protected:

public:
    DVS_flow(ull on_ev_change_, ull on_time_change_, ull start_time = 0)
        : on_ev_change(on_ev_change_), on_time_change(on_time_change_), time_diff(0), event_diff(0),
          last_slice_time(start_time), current_slice_time(start_time), accumulate(false), manual_mode(false), max_iter(-1),
          scale(3), stm_disable(false), batch_(1), gpus_(1), local_(false), quiet_(false), flow_out_(nullptr), slices_done_(0), events_done_(0),
          iters_done_(0), hdr_buffer_(MAX_SZ, SPAN) {}

    // run-time sized variant (CLI flags --max-events / --slice-time)
    DVS_flow(ull on_ev_change_, ull on_time_change_, ull start_time, size_t capacity, sll span)
        : ev_buffer(capacity, span), on_ev_change(on_ev_change_), on_time_change(on_time_change_), time_diff(0), event_diff(0),
          last_slice_time(start_time), current_slice_time(start_time), accumulate(false), manual_mode(false), max_iter(-1),
          scale(3), stm_disable(false), batch_(1), gpus_(1), local_(false), quiet_(false), flow_out_(nullptr), slices_done_(0), events_done_(0),
          iters_done_(0), hdr_buffer_(capacity, span) {}

    ~DVS_flow() {
        // the device ring belongs to the pooled context and would otherwise live until that is destroyed
        if (ring_ && ring_gen_ == CudaDriver::generation()) bf_ring_destroy(ring_);
    }

    // The per-event part is small and inlined into the caller's loop; the slice itself is not.
    __attribute__((always_inline)) inline bool add_event(Event &ev) {
        if (ring_on_) {
            // the slices are cut on the device: 16 bytes per event instead of the 152-byte record
            const RingHeader r = header_of(ev);
            hdr_buffer_.push_back_header(r);
            if (stage_fill_ < stage_room_ && *ctx_gen_ == ring_gen_) stage_[stage_fill_++] = r;
            else stage_slow(r);
        } else {
            ev_buffer.push_back(ev);
        }
        events_added_++;
        event_diff++;
        if (ev.timestamp < current_slice_time) unsorted_ = true;
        current_slice_time = ev.timestamp;
        time_diff = current_slice_time - last_slice_time;   // time only increases
        if ((event_diff < (sll)on_ev_change) && (time_diff < (sll)on_time_change)) return false;
        recompute();
        return true;
    }
    __attribute__((noinline)) void recompute();
    void flush();   // minimise whatever is still queued (batch mode)

    void set_accumulate(bool val = true) { accumulate = val; update_ring_on(); }
    LinearEventCloudTemplate<Event> get_accumulated();
    void set_manual_mode(bool val = true) {
        manual_mode = val;
        if (val) std::cerr << "interactive mode is a GUI feature of the reference and is ignored" << std::endl;
    }
    void set_max_iter(int val = -1) { max_iter = val; }
    void set_scale(int val = 3) { scale = val; }
    void set_generate_video(bool val = true, std::string fname = "out.avi", int fps = 30) {
        generate_video_ = val; video_name_ = fname; video_fps_ = fps;
        if (val) std::cerr << "video output: uncompressed YUV4MPEG2 (4:4:4) stream in '" << fname << "' (no AVI encoder without OpenCV)" << std::endl;
    }
    void set_generate_pictures(bool val = true, std::string prefix = "./") {
        generate_pictures_ = val; img_prefix_ = prefix;
        if (val) std::cerr << "picture output: " << prefix << "/frame_N.ppm + frame_N.txt (no JPEG encoder without OpenCV)" << std::endl;
    }
    void set_stm_disable(bool val = true) { stm_disable = val; update_ring_on(); }

    sll get_buf_size() { return win_size(); }
    sll get_time_diff() { return time_diff; }
    sll get_buf_time_diff() { return current_slice_time - slice_start_time(); }

    // extensions
    void set_batch(int n) { batch_ = n < 1 ? 1 : n; update_ring_on(); }
    void set_gpus(int n) { gpus_ = n < 1 ? 1 : n; update_ring_on(); }
    void set_optimizer_local(bool v = true) { local_ = v; update_ring_on(); }
    void set_quiet(bool q = true) { quiet_ = q; }
    // Do not read the per-event state (pr, nx/ny, u/v) back after a slice unless it is being accumulated: the
    // buffer's events then keep what Event::reset left.  For callers that only want the per-slice models.
    void set_lazy_events(bool v = true) { lazy_events_ = v; update_ring_on(); }
    void set_flow_out(std::ostream *os) { flow_out_ = os; }
    void set_device_ring(bool v = true) { device_ring_ = v; update_ring_on(); }
    void prepare();   // optional: create the device-side objects now instead of inside the first slice
    bool device_ring_active() const { return ring_on_; }
    ObjectModel get_last_model() { resolve_deferred(); return last_model; }
    ull slices_done() { resolve_deferred(); return slices_done_; }
    ull events_done() { resolve_deferred(); return events_done_; }
    ull iterations_done() { resolve_deferred(); return iters_done_; }

protected:
    // dvs_flow.h:186-193: the oldest timestamp when the buffer overflowed, else now - SPAN
    ull slice_start_time() {
        if (win_size() == ev_buffer.capacity()) return win_timestamp(ev_buffer.capacity() - 1);
        const ull span = (ull)ev_buffer.span();
        return (current_slice_time > span) ? current_slice_time - span : 0;
    }

    void log_slice(const SliceLog &l, int iters, int rc) {
        motion_memory.push_back(l);
        last_model = l.model;
        slices_done_ += 1;
        events_done_ += l.size;
        iters_done_ += (ull)iters;
        if (!quiet_) {
            // the reference dumps every remembered slice after each recompute (dvs_flow.h:245-252)
            std::cout << "\n\n------------------------\n";
            for (auto &s : motion_memory) {
                std::cout << s.model << "\n";
                std::cout << s.size << "\t" << s.ts_first << "\t" << s.ts_last << "\n";
            }
        }
        if (flow_out_) {
            std::ostream &o = *flow_out_;
            const auto old = o.precision(17);
            o << (slices_done_ - 1) << " " << l.size << " " << iters << " " << rc << " " << l.model.total_dx << " " << l.model.total_dy
              << " " << l.model.total_rot << " " << l.model.total_div << " " << l.model.cx << " " << l.model.cy << " " << l.model.dx
              << " " << l.model.dy << " " << l.model.rot << " " << l.model.div << " " << l.model.cnt << "\n";
            o.precision(old);
        }
    }

    void run_pending();
    void resolve_deferred();
    void ring_slice(const SliceLog &log, ull start);
    bool ensure_ring();
};

template <size_t MAX_SZ, sll SPAN> void DVS_flow<MAX_SZ, SPAN>::migrate_window(bool to_headers) {
    // oldest -> newest, so that the other ring ends up with the same visible content (size, order, full-buffer quirk)
    resolve_deferred();   // models of slices already enqueued on the device are logged before the mode changes
    if (ring_ && ring_gen_ == CudaDriver::generation()) bf_ring_destroy(ring_);   // (its content would be stale when the mode comes back)
    ring_ = nullptr;
    stage_drop();
    if (to_headers) {
        for (long int i = (long int)ev_buffer.size() - 1; i >= 0; i--) hdr_buffer_.push_back_header(header_of(ev_buffer[i]));
        ev_buffer = CircularArray<Event, MAX_SZ, SPAN>(ev_buffer.capacity(), ev_buffer.span());
    } else {
        for (long int i = (long int)hdr_buffer_.size() - 1; i >= 0; i--) {
            const RingHeader &h = hdr_buffer_[i];
            Event e(h.fr_x, h.fr_y & 0x7fffu, h.timestamp);
            e.noise = (h.fr_y & BF_EVENT_NOISE) != 0;
            ev_buffer.push_back(e);
        }
        hdr_buffer_ = RingCore<RingHeader>(hdr_buffer_.capacity(), hdr_buffer_.span());
    }
}

template <size_t MAX_SZ, sll SPAN> void DVS_flow<MAX_SZ, SPAN>::recompute() {
    const ull start = slice_start_time();

    // The slice = what a range-for over the buffer visits: newest -> oldest, and SZ-1 elements when the buffer is
    // full (dvs_flow.h:196-198 builds a LinearEventPtrs of exactly these).
    const size_t held = win_size();
    SliceLog log;
    log.size = held - ((held == ev_buffer.capacity() && held > 0) ? 1 : 0);
    log.ts_first = log.size ? win_timestamp(0) : 0;
    log.ts_last = log.size ? win_timestamp(log.size - 1) : 0;

    LinearEventPtrs e_ptrs;
    if (local_) {
        e_ptrs.reserve(log.size);
        for (auto &e : ev_buffer) e_ptrs.push_back(&e);
        assert(e_ptrs.size() == log.size);
    }

    if (local_) {
        // OptimizerLocal works on a LinearEventCloud of its own and never warm-starts
        LinearEventCloud cloud;
        cloud.reserve(log.size);
        for (auto &e : e_ptrs) {
            e.reset();
            e.set_local_time(start);
            cloud.push_back(e);
        }
        OptimizerLocal optimizer(&cloud, scale);
        const int rc = optimizer.run();
        log.model = ObjectModel();
        log.model.total_dx = -optimizer.get_nx();   // same sign convention as OptimizerRolling's totals:
        log.model.total_dy = -optimizer.get_ny();   // events are projected with n = -total (optimizer_rolling.h:340-344)
        log.model.dx = optimizer.get_score();
        size_t k = 0;
        for (auto &e : e_ptrs) {
            Event &c = cloud[k++];
            e.pr_x = c.pr_x; e.pr_y = c.pr_y; e.nx = c.nx; e.ny = c.ny;
            e.compute_uv();
            if (rc == 0) e.assume_score(0);
        }
        log_slice(log, optimizer.steps(), rc);
        if (accumulate) {
            LinearEventCloudTemplate<Event> cur;
            cur.reserve(ev_buffer.size());
            for (long int i = (long int)ev_buffer.size() - 1; i >= 0; i--) cur.push_back(ev_buffer[i]);
            accumulated.push_back(std::move(cur));
        }
    } else if (device_ring_active()) {
        const auto tr0 = std::chrono::steady_clock::now();
        ring_slice(log, start);
        t_recompute_ += std::chrono::duration<double>(std::chrono::steady_clock::now() - tr0).count();
    } else if (batch_ > 1 && stm_disable) {
        // independent slice: snapshot it and minimise later together with its neighbours
        Pending p;
        p.slice_start = start;
        p.packed.resize(log.size);
        size_t k = 0;
        // bounding box as set_cloud computes it (optimizer_rolling.h:252-260: minima start at RES_X / RES_Y, maxima at 0)
        uint x_min = (uint)RES_X, y_min = (uint)RES_Y, x_max = 0, y_max = 0;
        for (auto &e : ev_buffer) {
            e.reset();
            e.set_local_time(start);
            if (e.t > INT32_MAX || e.t < INT32_MIN) {
                std::cerr << "DVS_flow: local time of an event exceeds +-2.1 s; shorten the slice" << std::endl;
                std::exit(1);
            }
            bf_event &b = p.packed[k++];
            b.fr_x = (uint16_t)e.fr_x;
            b.fr_y = (uint16_t)(e.fr_y | (e.noise ? BF_EVENT_NOISE : 0u));
            b.t_ns = (int32_t)e.t;
            x_min = std::min(x_min, e.fr_x); x_max = std::max(x_max, e.fr_x);
            y_min = std::min(y_min, e.fr_y); y_max = std::max(y_max, e.fr_y);
        }
        assert(k == log.size);
        // run()'s tiny-window guard marks the slice's events as noise IN THE BUFFER (optimizer_rolling.h:49-55), and
        // later overlapping slices skip them.  The batch is minimised later, so the same test (bf_guard_tiny in the
        // back end, which will skip the slice) is evaluated here, before the next slice is snapshotted.
        {
            const int rows_img = scale * (int)(x_max - x_min) + scale, cols_img = scale * (int)(y_max - y_min) + scale;
            if ((rows_img < scale * (int)RES_X / 15) && (cols_img < scale * (int)RES_Y / 15))
                for (auto &e : ev_buffer) e.noise = true;
        }
        p.log = log;
        if (accumulate) {
            p.copy.reserve(ev_buffer.size());
            for (long int i = (long int)ev_buffer.size() - 1; i >= 0; i--) p.copy.push_back(ev_buffer[i]);
        }
        pending_.push_back(std::move(p));
        if ((int)pending_.size() >= batch_) run_pending();
    } else {
        // The reference's sequence (dvs_flow.h:210-235)
        //     OptimizerRolling opt; opt.set_cloud(&e_ptrs, scale); opt.set_time(start); opt.set_maxiter(max_iter);
        //     if (!stm_disable) opt.set_model(last_model); opt.run(); for (e : ev_buffer) e.compute_uv();
        // touches every 152-byte Event seven times (reset, local time, gather, write-back, assume_score,
        // compute_uv, plus the pointer vector) -- on the host that costs more than the kernel does.  The same
        // per-event operations are done here in two passes, in the same order per event; the OptimizerRolling
        // class itself is unchanged for callers that drive it directly.
        // (tests/test_host_stream_cpu.py holds both to the reference's DVS_flow and to each other.)
        const int n = (int)log.size;
        fx_.resize(n); fy_.resize(n); t_.resize(n); nz_.resize(n);
        int i = 0;
        for (auto &e : ev_buffer) {
            e.reset();                                        // set_cloud (optimizer_rolling.h:262)
            e.set_local_time(start);                          // set_time (optimizer_rolling.h:241-245)
            if (e.t > INT32_MAX || e.t < INT32_MIN) {
                std::cerr << "DVS_flow: local time of an event exceeds +-2.1 s; shorten the slice" << std::endl;
                std::exit(1);
            }
            fx_[i] = (uint16_t)e.fr_x; fy_[i] = (uint16_t)e.fr_y; t_[i] = (int32_t)e.t; nz_[i] = e.noise ? 1 : 0;
            ++i;
        }
        assert(i == n);
        // per-event state back to the host only when something reads it: the reference always fills it, so the
        // default is to do so; set_lazy_events() (the CLI without -o) skips the read-back and the second pass
        const bool want_events = accumulate || !lazy_events_;
        if (want_events) { px_.resize(n); py_.resize(n); nx_.resize(n); ny_.resize(n); }
        bf_ctx *ctx = CudaDriver::context(n, 1, scale);
        const bf_model init = last_model.to_pod();            // set_model(last_model), dvs_flow.h:218-219
        bf_slice_result res;
        const int rc = bf_minimize(ctx, fx_.data(), fy_.data(), t_.data(), nz_.data(), n, scale, max_iter,
                                   stm_disable ? nullptr : &init, &res, want_events ? px_.data() : nullptr,
                                   want_events ? py_.data() : nullptr, want_events ? nx_.data() : nullptr,
                                   want_events ? ny_.data() : nullptr);
        if (rc < 0) {
            std::cerr << "bf_minimize failed: " << bf_last_error() << std::endl;
            std::exit(1);
        }
        if (rc == BF_RC_DEGENERATE)
            std::cerr << "OptimizerRolling: empty time image (the reference would not terminate here)" << std::endl;
        log.model.from_pod(res.model);
        const bool all_noise = (res.flags & BF_FLAG_ALL_NOISE) != 0;   // tiny window (optimizer_rolling.h:49-55)
        if (want_events) {
            i = 0;
            for (auto &e : ev_buffer) {
                e.pr_x = px_[i]; e.pr_y = py_[i]; e.nx = nx_[i]; e.ny = ny_[i];
                if (all_noise) e.noise = true;
                if (rc != BF_RC_SKIPPED) e.assume_score(0);   // optimizer_rolling.h:121-122
                e.compute_uv();                               // dvs_flow.h:234-235
                ++i;
            }
        } else if (all_noise) {
            for (auto &e : ev_buffer) e.noise = true;         // (the flag outlives the slice: later slices skip these events)
        }
        log_slice(log, res.iters, rc);
        if ((generate_pictures_ || generate_video_) && want_events) dump_frame(log.size);   // dvs_flow.h:256-335
        if (accumulate) {                                     // dvs_flow.h:341-346: oldest -> newest copy
            LinearEventCloudTemplate<Event> cur;
            cur.reserve(ev_buffer.size());
            for (long int i = (long int)ev_buffer.size() - 1; i >= 0; i--) cur.push_back(ev_buffer[i]);
            accumulated.push_back(std::move(cur));
        }
    }
    event_diff = 0;
    last_slice_time = current_slice_time;
}

template <size_t MAX_SZ, sll SPAN> void DVS_flow<MAX_SZ, SPAN>::run_pending() {
    if (pending_.empty()) return;
    auto check = [](int rc, const char *what) {
        if (rc < 0) {
            std::cerr << what << " failed: " << bf_last_error() << std::endl;
            std::exit(1);
        }
    };
    const int n_pending = (int)pending_.size();
    bf_ctx *ctx = nullptr;
    bf_multi *multi = nullptr;
    if (gpus_ > 1) {
        // capacity per device = the largest share the block-cyclic deal can produce
        const int block = 4;
        std::vector<long long> ev((size_t)gpus_, 0);
        std::vector<int> sl((size_t)gpus_, 0);
        for (int k = 0; k < n_pending; ++k) {
            const int o = bf_multi_owner(k, gpus_, block);
            ev[(size_t)o] += (long long)pending_[(size_t)k].packed.size();
            sl[(size_t)o] += 1;
        }
        multi = CudaDriver::multi(gpus_, *std::max_element(ev.begin(), ev.end()), *std::max_element(sl.begin(), sl.end()), scale);
        check(bf_multi_reset(multi), "bf_multi_reset");
        check(bf_multi_set_option(multi, "block", block), "bf_multi_set_option");
        for (auto &p : pending_) check(bf_multi_add_packed(multi, p.packed.data(), (int)p.packed.size(), scale, max_iter), "bf_multi_add_packed");
        check(bf_multi_run(multi, accumulate ? 1 : 0), "bf_multi_run");
        check(bf_multi_sync(multi), "bf_multi_sync");
    } else {
        long long total = 0;
        for (auto &p : pending_) total += (long long)p.packed.size();
        ctx = CudaDriver::context(total, n_pending, scale);
        check(bf_batch_reset(ctx), "bf_batch_reset");
        for (auto &p : pending_) check(bf_batch_add_packed(ctx, p.packed.data(), (int)p.packed.size(), scale, max_iter, nullptr), "bf_batch_add_packed");
        check(bf_batch_run(ctx, accumulate ? 1 : 0), "bf_batch_run");
    }
    std::vector<double> nx, ny, px, py;
    for (size_t k = 0; k < pending_.size(); ++k) {
        Pending &p = pending_[k];
        bf_slice_result r;
        bf_ctx *ev_ctx = ctx;
        int ev_slot = (int)k;
        if (multi) {
            check(bf_multi_result(multi, (int)k, &r), "bf_multi_result");
            check(bf_multi_locate(multi, (int)k, &ev_ctx, &ev_slot, nullptr), "bf_multi_locate");
        } else {
            check(bf_batch_result(ctx, (int)k, &r), "bf_batch_result");
        }
        p.log.model.from_pod(r.model);
        log_slice(p.log, r.iters, r.rc);
        if (accumulate) {
            // the snapshot copy is oldest -> newest and includes the element the iterator skips when the
            // buffer is full; the packed slice is newest -> oldest.  Map by walking backwards.
            const size_t n = p.packed.size();
            nx.resize(n); ny.resize(n); px.resize(n); py.resize(n);
            check(bf_batch_events(ev_ctx, ev_slot, px.data(), py.data(), nx.data(), ny.data()), "bf_batch_events");
            const size_t m = p.copy.size();
            for (size_t i = 0; i < m; ++i) {
                Event &e = p.copy[i];
                const size_t from_newest = m - 1 - i;
                if (from_newest < n) {
                    e.pr_x = px[from_newest]; e.pr_y = py[from_newest]; e.nx = nx[from_newest]; e.ny = ny[from_newest];
                    e.set_local_time(p.slice_start);
                    if (r.flags & BF_FLAG_ALL_NOISE) e.noise = true;
                    e.compute_uv();
                    if (r.rc != BF_RC_SKIPPED) e.assume_score(0);
                }
            }
            accumulated.push_back(std::move(p.copy));
        }
    }
    pending_.clear();
}

// dvs_flow.h:256-335: the frame the reference writes after every slice when pictures / video are requested.
//   top    : projection_img(ev_buffer, 3, show_final = true)  | mean-timestamp image of the events as recorded
//   bottom : projection_img(ev_buffer, 3, show_final = false) | mean-timestamp image of the warped events
template <size_t MAX_SZ, sll SPAN> void DVS_flow<MAX_SZ, SPAN>::dump_frame(size_t n_slice) {
    const int S = 3;
    const int rows = (int)RES_X * S, cols = (int)RES_Y * S;
    const int n = (int)n_slice;
    std::vector<double> px((size_t)n), py((size_t)n), fx((size_t)n), fy((size_t)n);
    std::vector<int32_t> tl((size_t)n);
    std::vector<uint8_t> nz((size_t)n);
    int i = 0;
    for (auto &e : ev_buffer) {
        px[(size_t)i] = e.pr_x; py[(size_t)i] = e.pr_y; fx[(size_t)i] = (double)e.fr_x; fy[(size_t)i] = (double)e.fr_y;
        tl[(size_t)i] = (int32_t)e.t; nz[(size_t)i] = e.noise ? 1 : 0;
        ++i;
    }
    bf_ctx *ctx = CudaDriver::context(n, 1, std::max(scale, S));
    std::vector<uint8_t> pr_t((size_t)rows * cols), pr_f((size_t)rows * cols);
    const int crow = rows + S, ccol = cols + S;                 // color_time_img spans (S * RES + S) pixels per axis
    std::vector<uint8_t> col_t((size_t)crow * ccol * 3), col_f((size_t)crow * ccol * 3);
    auto check = [](int rc, const char *what) {
        if (rc < 0) { std::cerr << what << " failed: " << bf_last_error() << std::endl; std::exit(1); }
    };
    check(bf_projection_img(ctx, n, fx.data(), fy.data(), nz.data(), S, pr_t.data(), nullptr), "bf_projection_img");
    check(bf_projection_img(ctx, n, px.data(), py.data(), nz.data(), S, pr_f.data(), nullptr), "bf_projection_img");
    check(bf_color_time_img(ctx, n, fx.data(), fy.data(), tl.data(), nz.data(), S, col_t.data()), "bf_color_time_img");
    check(bf_color_time_img(ctx, n, px.data(), py.data(), tl.data(), nz.data(), S, col_f.data()), "bf_color_time_img");
    // 2 x 2 montage, 3 channels (R G B for the PPM / planar for the video): the grey projection images replicated
    // (cv::cvtColor GRAY2RGB), the colour images cropped to the frame (the reference resizes 543 x 723 -> 540 x 720)
    std::vector<uint8_t> frame((size_t)4 * rows * cols * 3);
    auto put = [&](int r0, int c0, const uint8_t *grey, const uint8_t *bgr, int bgr_cols) {
        for (int r = 0; r < rows; ++r)
            for (int cc = 0; cc < cols; ++cc) {
                uint8_t *o = &frame[((size_t)(r0 + r) * 2 * cols + (c0 + cc)) * 3];
                if (grey) { o[0] = o[1] = o[2] = grey[(size_t)r * cols + cc]; }
                else { const uint8_t *q = bgr + ((size_t)r * bgr_cols + cc) * 3; o[0] = q[2]; o[1] = q[1]; o[2] = q[0]; }
            }
    };
    put(0, 0, pr_t.data(), nullptr, 0);
    put(0, cols, nullptr, col_t.data(), ccol);
    put(rows, 0, pr_f.data(), nullptr, 0);
    put(rows, cols, nullptr, col_f.data(), ccol);
    if (generate_pictures_) {
        const std::string base = img_prefix_ + "/frame_" + std::to_string(frame_count_);
        std::ofstream ppm(base + ".ppm", std::ofstream::binary);
        ppm << "P6\n" << 2 * cols << " " << 2 * rows << "\n255\n";
        ppm.write(reinterpret_cast<const char *>(frame.data()), (std::streamsize)frame.size());
        // the text the reference renders into the frame (dvs_flow.h:274-318)
        std::ofstream txt(base + ".txt");
        txt << "timestamp: " << double(current_slice_time) / 1000000000.0 << "\n"
            << "%realtime: " << double(on_time_change) / double(time_diff) << "\n"
            << "Time diff (new): " << double(time_diff) / 1000000000.0 << "\n"
            << "Events: " << ev_buffer.size() << "\n"
            << "New events: " << event_diff << "\n"
            << "Model:\n" << last_model << "\n";
        frame_count_++;
    }
    if (generate_video_) {
        if (!video_out_) {
            video_out_.reset(new std::ofstream(video_name_, std::ofstream::binary));
            if (!*video_out_) std::cout << "Could not open the output video for write" << std::endl;
            *video_out_ << "YUV4MPEG2 W" << 2 * cols << " H" << 2 * rows << " F" << video_fps_ << ":1 Ip A1:1 C444\n";
        }
        // planar Y Cb Cr (BT.601 full range) of the RGB frame
        const size_t np_ = (size_t)4 * rows * cols;
        std::vector<uint8_t> yuv(3 * np_);
        for (size_t k = 0; k < np_; ++k) {
            const double R = frame[3 * k], G = frame[3 * k + 1], B = frame[3 * k + 2];
            yuv[k] = (uint8_t)std::min(255.0, std::max(0.0, 0.299 * R + 0.587 * G + 0.114 * B + 0.5));
            yuv[np_ + k] = (uint8_t)std::min(255.0, std::max(0.0, -0.168736 * R - 0.331264 * G + 0.5 * B + 128.5));
            yuv[2 * np_ + k] = (uint8_t)std::min(255.0, std::max(0.0, 0.5 * R - 0.418688 * G - 0.081312 * B + 128.5));
        }
        *video_out_ << "FRAME\n";
        video_out_->write(reinterpret_cast<const char *>(yuv.data()), (std::streamsize)yuv.size());
    }
}

template <size_t MAX_SZ, sll SPAN> void DVS_flow<MAX_SZ, SPAN>::flush() {
    run_pending();
    resolve_deferred();
}

// The open reservation is full, or there is none (no ring yet, or the pooled context -- and with it the ring and its
// staging buffer -- was re-created by another user).  Without a ring nothing is lost: the next slice rebuilds the
// device ring from the window (`hdr_buffer_`).
template <size_t MAX_SZ, sll SPAN> void DVS_flow<MAX_SZ, SPAN>::stage_slow(const RingHeader &r) {
    if (!ring_ || ring_gen_ != CudaDriver::generation()) {
        stage_drop();
        return;
    }
    auto check = [](int rc, const char *what) {
        if (rc < 0) {
            std::cerr << what << " failed: " << bf_last_error() << std::endl;
            std::exit(1);
        }
    };
    check(bf_ring_commit(ring_, stage_fill_), "bf_ring_commit");
    check(bf_ring_reserve(ring_, kStageChunk, &stage_), "bf_ring_reserve");
    stage_room_ = kStageChunk;
    stage_[0] = r;
    stage_fill_ = 1;
}

// Makes sure the device ring exists under the current pooled context; returns true if it had to be (re)built -- it then
// holds the whole window, staged events included.
template <size_t MAX_SZ, sll SPAN> bool DVS_flow<MAX_SZ, SPAN>::ensure_ring() {
    auto check = [](int rc, const char *what) {
        if (rc < 0) {
            std::cerr << what << " failed: " << bf_last_error() << std::endl;
            std::exit(1);
        }
    };
    const long long cap = (long long)ev_buffer.capacity();
    // a request that re-creates the pooled context destroys the ring: read the outstanding models back first
    if (ring_ && ring_gen_ == CudaDriver::generation() && !CudaDriver::fits(cap + 64, 1, scale)) resolve_deferred();
    bf_ctx *ctx = CudaDriver::context(cap + 64, 1, scale);
    if (ring_ && ring_gen_ == CudaDriver::generation()) return false;
    // (a context re-created for more capacity took its rings with it)
    stage_drop();
    ring_ = bf_ring_create(ctx, cap, ring_pending_);
    if (!ring_) check(-1, "bf_ring_create");
    ring_gen_ = CudaDriver::generation();
    // a ring created in the middle of a stream continues the warm-start chain from the host's last model
    const bf_model seed = last_model.to_pod();
    check(bf_ring_seed(ring_, &seed), "bf_ring_seed");
    // everything the window holds, oldest -> newest
    ring_new_.clear();
    for (long int i = (long int)hdr_buffer_.size() - 1; i >= 0; i--) ring_new_.push_back(hdr_buffer_[i]);
    if (!ring_new_.empty()) check(bf_ring_push(ring_, ring_new_.data(), (int)ring_new_.size()), "bf_ring_push");
    ring_new_.clear();
    return true;
}

// Start-up work the first slice would otherwise pay for (pooled context, device ring and its pinned staging buffer).
template <size_t MAX_SZ, sll SPAN> void DVS_flow<MAX_SZ, SPAN>::prepare() {
    if (!ring_on_) {
        if (gpus_ > 1 || local_) return;   // (the multi-GPU front and OptimizerLocal size their contexts themselves)
        const int nb = (batch_ > 1 && stm_disable) ? batch_ : 1;
        CudaDriver::context(((long long)ev_buffer.capacity() + 64) * nb, nb, scale);
        return;
    }
    if (ensure_ring()) {
        if (bf_ring_reserve(ring_, kStageChunk, &stage_) < 0) {
            std::cerr << "bf_ring_reserve failed: " << bf_last_error() << std::endl;
            std::exit(1);
        }
        stage_fill_ = 0; stage_room_ = kStageChunk;
    }
}

// Default mode on the device-resident ring (include/bf_cuda.h: bf_ring_*).  Per slice the host uploads the events
// that arrived since the last slice and enqueues "newest n events, local time relative to `start`, warm-started from
// the previous slice's model ON THE DEVICE"; nothing here waits for the GPU unless the per-slice dump is wanted.
template <size_t MAX_SZ, sll SPAN> void DVS_flow<MAX_SZ, SPAN>::ring_slice(const SliceLog &log, ull start) {
    auto check = [](int rc, const char *what) {
        if (rc < 0) {
            std::cerr << what << " failed: " << bf_last_error() << std::endl;
            std::exit(1);
        }
    };
    const auto tp0 = std::chrono::steady_clock::now();
    if (!ensure_ring()) check(bf_ring_commit(ring_, stage_fill_), "bf_ring_commit");   // the events that arrived since the last slice
    check(bf_ring_reserve(ring_, kStageChunk, &stage_), "bf_ring_reserve");
    stage_fill_ = 0; stage_room_ = kStageChunk;
    const auto tp1 = std::chrono::steady_clock::now();
    t_push_ += std::chrono::duration<double>(tp1 - tp0).count();
    if (log.size > 0) {
        const sll t_new = (sll)(log.ts_first - start), t_old = (sll)(log.ts_last - start);
        if (t_new > INT32_MAX || t_old > INT32_MAX || t_new < INT32_MIN || t_old < INT32_MIN) {
            std::cerr << "DVS_flow: local time of an event exceeds +-2.1 s; shorten the slice" << std::endl;
            std::exit(1);
        }
    }
    const auto ts0 = std::chrono::steady_clock::now();
    const int ticket = bf_ring_slice(ring_, (int)log.size, start, scale, max_iter, stm_disable ? 0 : 1);
    check(ticket, "bf_ring_slice");
    t_slice_ += std::chrono::duration<double>(std::chrono::steady_clock::now() - ts0).count();
    deferred_.push_back(Deferred{log, ticket});
    // the reference dumps every remembered slice after each recompute: that needs the model now
    if (!quiet_ || (int)deferred_.size() >= ring_pending_ - 1) resolve_deferred();
}

template <size_t MAX_SZ, sll SPAN> void DVS_flow<MAX_SZ, SPAN>::resolve_deferred() {
    if (!deferred_.empty() && ring_gen_ != CudaDriver::generation()) {
        // another user of the pooled context re-created it (and with it the ring) while slices were outstanding
        std::cerr << "DVS_flow: the CUDA context was re-created with " << deferred_.size() << " slices outstanding; their models are lost" << std::endl;
        std::exit(1);
    }
    while (!deferred_.empty()) {
        Deferred d = deferred_.front();
        deferred_.pop_front();
        bf_slice_result res;
        const auto tr0 = std::chrono::steady_clock::now();
        const int rrc = bf_ring_result(ring_, d.ticket, &res);
        t_resolve_ += std::chrono::duration<double>(std::chrono::steady_clock::now() - tr0).count();
        if (rrc < 0) {
            std::cerr << "bf_ring_result failed: " << bf_last_error() << std::endl;
            std::exit(1);
        }
        if (res.rc == BF_RC_DEGENERATE)
            std::cerr << "OptimizerRolling: empty time image (the reference would not terminate here)" << std::endl;
        d.log.model.from_pod(res.model);
        log_slice(d.log, res.iters, res.rc);
    }
}

// dvs_flow.h:350-389: concatenate the remembered slices, dropping from LATER slices every event that an
// earlier slice already contains (same pixel, not newer, closer than 0.1 ms).  The reference does
// this with nested linear scans; an index by pixel gives the same set in near-linear time.
template <size_t MAX_SZ, sll SPAN> LinearEventCloudTemplate<Event> DVS_flow<MAX_SZ, SPAN>::get_accumulated() {
    flush();
    LinearEventCloudTemplate<Event> ret;
    std::cout << "Aggregating events into one cloud...\n";
    if (unsorted_) {
        // The indexed scan below equals the reference's nested scan only for non-decreasing timestamps (it stops at
        // the first LATER BUFFER that starts after e, the reference only leaves the inner loop at the first newer
        // event and still visits the buffers after it).  Out-of-order input takes the reference's literal scan.
        for (ull i = 0; i < accumulated.size(); ++i) {
            std::cout << "\tBuffer: " << i << "\n";
            for (auto &e : accumulated[i]) {
                if (e.t == -1) continue;
                for (ull j = i + 1; j < accumulated.size(); ++j) {
                    for (auto &o : accumulated[j]) {
                        if (o - e > 0) break;
                        if (o.t == -1) continue;
                        if (e != o) continue;
                        o.t = -1;
                    }
                }
                ret.push_back(e);
            }
        }
        std::cout << "FInal buffer contains " << ret.size() << " events." << std::endl;
        return ret;
    }
    {
        // every event that was added survives in at most one copy, plus the few that several buffers hold more than
        // 0.1 ms apart from any same-pixel neighbour cannot exceed what the buffers hold: reserve once instead of doubling
        ull held = 0;
        for (auto &buf : accumulated) held += buf.size();
        ret.reserve((size_t)std::min<ull>(held, events_added_ + events_added_ / 64 + 1024));
    }
    // Per-buffer index "pixel -> its events, oldest first" (counting sort by pixel), built when a buffer is first
    // scanned and dropped once no earlier buffer can reach it any more: a few buffers are alive at a time.  The scan
    // works on the index alone -- timestamps in pixel order and one "erased" byte per event (the reference's o.t = -1
    // mark, which only ever decides whether a copy is emitted) -- so the 152-byte copies of the LATER buffers are not
    // touched at all; each copy is read once, in order, when its own buffer is emitted.
    struct PixelIndex {
        std::vector<uint32_t> first, pos;    // first[pixel] .. first[pixel + 1] into pos / ts
        std::vector<ull> ts;                 // timestamp of event pos[q]
        std::vector<uint8_t> dead;           // by position in the buffer
    };
    uint rows = 1, cols = 1;
    for (auto &buf : accumulated)
        for (auto &e : buf) { rows = std::max(rows, e.fr_x + 1); cols = std::max(cols, e.fr_y + 1); }
    std::vector<std::unique_ptr<PixelIndex>> index(accumulated.size());
    std::vector<uint32_t> fill;
    auto index_of = [&](size_t j) -> PixelIndex & {
        if (!index[j]) {
            auto &buf = accumulated[j];
            std::unique_ptr<PixelIndex> ix(new PixelIndex);
            ix->first.assign((size_t)rows * cols + 1, 0u);
            for (size_t k = 0; k < buf.size(); ++k) ix->first[(size_t)buf[k].fr_x * cols + buf[k].fr_y + 1] += 1;
            for (size_t p = 1; p < ix->first.size(); ++p) ix->first[p] += ix->first[p - 1];
            ix->pos.resize(buf.size());
            ix->ts.resize(buf.size());
            ix->dead.assign(buf.size(), 0);
            fill.assign(ix->first.begin(), ix->first.end() - 1);
            for (size_t k = 0; k < buf.size(); ++k) {
                const uint32_t q = fill[(size_t)buf[k].fr_x * cols + buf[k].fr_y]++;
                ix->pos[q] = (uint32_t)k;
                ix->ts[q] = buf[k].timestamp;
            }
            index[j] = std::move(ix);
        }
        return *index[j];
    };
    std::vector<ull> starts(accumulated.size(), 0);      // timestamp of every buffer's first (oldest) copy
    for (size_t j = 0; j < accumulated.size(); ++j)
        if (accumulated[j].size() > 0) starts[j] = accumulated[j][0].timestamp;
    for (ull i = 0; i < accumulated.size(); ++i) {
        std::cout << "\tBuffer: " << i << "\n";
        auto &buf = accumulated[i];
        const uint8_t *dead_i = index[i] ? index[i]->dead.data() : nullptr;   // (nobody looked into this buffer: nothing erased)
        for (size_t k = 0; k < buf.size(); ++k) {
            Event &e = buf[k];
            if (e.t == -1 || (dead_i && dead_i[k])) continue;
            for (ull j = i + 1; j < accumulated.size(); ++j) {
                // buffers are oldest -> newest copies of a ring whose oldest timestamp never decreases: once a
                // later buffer STARTS after e, neither it nor any buffer after it holds a candidate (o - e <= 0)
                if (accumulated[j].size() == 0) continue;
                if (sll(starts[j]) - sll(e.timestamp) > 0) break;
                PixelIndex &ix = index_of(j);
                const size_t pixel = (size_t)e.fr_x * cols + e.fr_y;
                for (uint32_t q = ix.first[pixel]; q < ix.first[pixel + 1]; ++q) {
                    const ull o_ts = ix.ts[q];
                    if (sll(o_ts) - sll(e.timestamp) > 0) continue;      // newer than e: the reference's scan has stopped by then
                    // e != o (event.h:40-45) with the pixel equal and o not newer: closer than 0.1 ms or not
                    if (e.timestamp - o_ts >= 100000) continue;
                    // (a copy whose local time is the reference's mark value -1 is skipped by its scan as well)
                    if (ix.dead[ix.pos[q]] || accumulated[j][ix.pos[q]].t == -1) continue;
                    ix.dead[ix.pos[q]] = 1;
                }
            }
            ret.push_back(e);
        }
        index[i].reset();
        LinearEventCloudTemplate<Event>().swap(buf);   // (done with this copy: later buffers only look forward)
    }
    std::cout << "FInal buffer contains " << ret.size() << " events." << std::endl;
    return ret;
}

#endif  // BF_DVS_FLOW_H
