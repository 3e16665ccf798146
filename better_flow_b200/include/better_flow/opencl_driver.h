// better_flow/opencl_driver.h -- back-end selection.  The reference's OpenCLDriver is a static
// singleton that picks an OpenCL device and JIT-builds gpu_impl.cl (reference:
// better_flow_core/src/opencl_driver.cpp:14-60).  Here the same role is played by the CUDA library
// behind include/bf_cuda.h; the class keeps its name and its two public members so that
// `OpenCLDriver::init()` / `OpenCLDriver::enabled` in caller code keep compiling.
// There is no CPU path: if no CUDA device can be opened, init() reports and exits, exactly as the
// reference exits when its kernel file cannot be built (opencl_driver.cpp:43,50).
#ifndef BF_OPENCL_DRIVER_H
#define BF_OPENCL_DRIVER_H

#include <chrono>

#include <better_flow/common.h>
#include <bf_cuda.h>

class CudaDriver {
public:
    static bool &enabled_ref() {
        static bool e = false;
        return e;
    }
    static int &device_ref() {
        static int d = 0;
        return d;
    }

    static void init(int device = 0) {
        if (bf_cuda_init(device) != BF_OK) {
            std::cerr << "CUDA back-end unavailable: " << bf_last_error() << std::endl;
            std::exit(1);
        }
        device_ref() = device;
        enabled_ref() = true;
    }

    // Would context(need...) return the existing context?  (Callers that hold objects tied to it -- a bf_ring with
    // outstanding tickets -- settle them before a request that re-creates the context.)
    static bool fits(long long need_events, int need_slices, int need_scale) {
        State &s = state();
        return s.ctx && s.rows == RES_X && s.cols == RES_Y && need_events <= s.events && need_slices <= s.slices && need_scale <= s.scale;
    }

    // One pooled context per process (the reference re-allocates device buffers for every slice,
    // accel_lib.h:71-145 via dvs_flow.h:210).  Re-created only if a caller needs more capacity.
    static bf_ctx *context(long long need_events, int need_slices, int need_scale) {
        if (!enabled_ref()) init(device_ref());
        State &s = state();
        const bool fits = s.ctx && s.rows == RES_X && s.cols == RES_Y && need_events <= s.events &&
                          need_slices <= s.slices && need_scale <= s.scale;
        if (!fits) {
            const auto t0 = std::chrono::steady_clock::now();
            if (s.ctx) bf_ctx_destroy(s.ctx);          // (destroys the context's rings with it)
            s.generation += 1;
            s.rows = RES_X; s.cols = RES_Y;
            s.events = std::max<long long>(need_events + need_events / 4, 1 << 16);
            s.slices = std::max(need_slices, 64);
            s.scale = std::max(need_scale, s.scale);
            s.ctx = bf_ctx_create(s.rows, s.cols, s.scale, s.events, s.slices);
            if (!s.ctx) {
                std::cerr << "bf_ctx_create failed: " << bf_last_error() << std::endl;
                std::exit(1);
            }
            // development knobs for same-box A/B runs of the tool (tools/cli_ring.py)
            if (const char *v = std::getenv("BF_RING_CLUSTER")) bf_ctx_set_option(s.ctx, "ring_cluster", atoi(v));
            if (const char *v = std::getenv("BF_CLUSTER")) bf_ctx_set_option(s.ctx, "cluster", atoi(v));
            if (std::getenv("BF_TIMING"))
                std::cerr << "[timing] context (re)created for " << s.events << " events / " << s.slices << " slices in "
                          << std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() << " s" << std::endl;
        }
        return s.ctx;
    }

    // Pooled multi-GPU front (bf_multi_*): `gpus` devices starting at the selected one.
    static bf_multi *multi(int gpus, long long need_events_per_dev, int need_slices_per_dev, int need_scale) {
        if (!enabled_ref()) init(device_ref());
        State &s = state();
        const bool fits = s.multi && s.m_gpus == gpus && s.m_rows == RES_X && s.m_cols == RES_Y &&
                          need_events_per_dev <= s.m_events && need_slices_per_dev <= s.m_slices && need_scale <= s.m_scale;
        if (!fits) {
            if (s.multi) bf_multi_destroy(s.multi);
            s.m_gpus = gpus; s.m_rows = RES_X; s.m_cols = RES_Y;
            s.m_events = std::max<long long>(need_events_per_dev + need_events_per_dev / 4, 1 << 16);
            s.m_slices = std::max(need_slices_per_dev, 16);
            s.m_scale = std::max(need_scale, s.m_scale);
            std::vector<int> ids;
            for (int i = 0; i < gpus; ++i) ids.push_back(device_ref() + i);
            s.multi = bf_multi_create(gpus, ids.data(), s.m_rows, s.m_cols, s.m_scale, s.m_events, s.m_slices);
            if (!s.multi) {
                std::cerr << "bf_multi_create failed: " << bf_last_error() << std::endl;
                std::exit(1);
            }
        }
        return s.multi;
    }

    // bumped whenever the pooled context is (re)created: objects tied to the old context (bf_ring) are gone then
    static unsigned long generation() { return state().generation; }
    static const unsigned long *generation_ptr() { return &state().generation; }   // (for per-event checks)

    static void shutdown() {
        State &s = state();
        if (s.ctx) bf_ctx_destroy(s.ctx);
        s.ctx = nullptr;
        s.generation += 1;
        if (s.multi) bf_multi_destroy(s.multi);
        s.multi = nullptr;
    }

private:
    struct State {
        bf_ctx *ctx = nullptr;
        unsigned long generation = 0;
        int rows = 0, cols = 0, slices = 0, scale = 3;
        long long events = 0;
        bf_multi *multi = nullptr;
        int m_gpus = 0, m_rows = 0, m_cols = 0, m_slices = 0, m_scale = 3;
        long long m_events = 0;
    };
    static State &state() {
        static State s;
        return s;
    }
};

// Drop-in spelling used by the reference's callers (bf_motion_compensator.cpp:132-133, accel_lib.h:46).
class OpenCLDriver : public CudaDriver {
public:
    static bool enabled_get() { return enabled_ref(); }
};

#endif  // BF_OPENCL_DRIVER_H
