// bf_multi.cpp -- slice-sharded multi-GPU front of the C ABI (include/bf_cuda.h: bf_multi_*), SURVEY 8(e).
//
// Time slices are independent once the warm start is disabled (reference --stm-disable semantics,
// dvs_flow.h:218-219), so N GPUs are used the obvious way: blocks of consecutive slices are dealt
// round-robin to the devices, every device minimises its share with ITS OWN persistent launch (no
// data-path collective), and the fixed-size per-slice flow records are exchanged with ONE NCCL
// all-gather per batch (160 B x slices: latency-bound, the NVLink fabric is irrelevant for it).
// One host process drives all devices (ncclCommInitAll); everything between "add" and "sync" is
// asynchronous, so the N launches and the event uploads run concurrently.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2): the library has no link-time dependency on
// it, single-GPU users never load it, and inside a PyTorch process the copy torch already loaded is
// the one that gets used.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/bf_cuda.h"

// ---- the handful of NCCL entry points we need ----------------------------------------------------
typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;      // ncclSuccess == 0
enum { NCCL_CHAR = 0 };        // ncclInt8 / ncclChar
struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
extern "C" void bf_set_error_(const char *msg);   // bf_cuda.cu: stores the text bf_last_error() returns

static int mfail(int code, const char *fmt, const char *a = "", const char *b = "") {
    char buf[512];
    snprintf(buf, sizeof buf, fmt, a, b);
    bf_set_error_(buf);
    return code;
}

static int load_nccl() {
    if (g_nccl.lib) return BF_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *n : names)
        if ((h = dlopen(n, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
    if (!h) return mfail(BF_ERR_CUDA, "multi-GPU needs NCCL: dlopen(libnccl.so.2) failed: %s", dlerror());
    NcclApi a;
    a.lib = h;
    a.CommInitAll = (decltype(a.CommInitAll))dlsym(h, "ncclCommInitAll");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
    a.GroupStart = (decltype(a.GroupStart))dlsym(h, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))dlsym(h, "ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!a.CommInitAll || !a.CommDestroy || !a.AllGather || !a.GroupStart || !a.GroupEnd || !a.GetErrorString)
        return mfail(BF_ERR_CUDA, "libnccl lacks a required symbol");
    g_nccl = a;
    return BF_OK;
}

#define MCU(call)                                                                         \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) return mfail(BF_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)
#define MNCCL(call)                                                                       \
    do {                                                                                  \
        ncclResult_t r_ = (call);                                                         \
        if (r_ != 0) return mfail(BF_ERR_CUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r_)); \
    } while (0)

struct Dev {
    int device = 0;
    bf_ctx *ctx = nullptr;
    cudaStream_t stream = nullptr;
    ncclComm_t comm = nullptr;
    unsigned char *d_gather = nullptr;     // [n_dev][cap] result records of every device
    int n_local = 0;                       // slices of the current batch on this device
};

struct bf_multi {
    std::vector<Dev> devs;
    int block = 4;                         // slices per block of the block-cyclic deal
    int max_slices_per_dev = 0;
    int n_slices = 0;
    std::vector<int> owner, slot;          // global slice -> device index / slot in that device's batch
    unsigned char *h_gather = nullptr;     // pinned copy of device 0's gather buffer
    int cap = 0;                           // records per device in the last gather
    bool ran = false;
};

extern "C" {

// Block-cyclic owner of global slice `k` (same rule as better_flow_b200/shard.py:partition).
int bf_multi_owner(int slice, int n_devices, int block) {
    if (n_devices <= 0 || block <= 0 || slice < 0) return -1;
    return (slice / block) % n_devices;
}

void bf_multi_destroy(bf_multi *m) {
    if (!m) return;
    for (Dev &d : m->devs) {
        cudaSetDevice(d.device);
        if (d.stream) cudaStreamSynchronize(d.stream);
        if (d.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(d.comm);
        if (d.ctx) bf_ctx_destroy(d.ctx);
        if (d.d_gather) cudaFree(d.d_gather);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    if (m->h_gather) cudaFreeHost(m->h_gather);
    delete m;
}

bf_multi *bf_multi_create(int n_devices, const int *devices, int sensor_rows, int sensor_cols, int max_scale,
                          long long max_events_per_device, int max_slices_per_device) {
    if (n_devices <= 0 || max_slices_per_device <= 0) {
        mfail(BF_ERR_ARG, "bf_multi_create: bad arguments");
        return nullptr;
    }
    const int have = bf_device_count();
    if (have < n_devices) {
        char a[32], b[32];
        snprintf(a, sizeof a, "%d", n_devices);
        snprintf(b, sizeof b, "%d", have);
        mfail(BF_ERR_CUDA, "bf_multi_create: %s devices requested, %s visible (no CPU fallback)", a, b);
        return nullptr;
    }
    if (n_devices > 1 && load_nccl() != BF_OK) return nullptr;
    bf_multi *m = new bf_multi();
    m->max_slices_per_dev = max_slices_per_device;
    m->devs.resize((size_t)n_devices);
    std::vector<int> ids((size_t)n_devices);
    for (int i = 0; i < n_devices; ++i) ids[(size_t)i] = devices ? devices[i] : i;
    const size_t gbytes = (size_t)n_devices * (size_t)max_slices_per_device * sizeof(bf_slice_result);
    for (int i = 0; i < n_devices; ++i) {
        Dev &d = m->devs[(size_t)i];
        d.device = ids[(size_t)i];
        if (bf_cuda_init(d.device) != BF_OK) { bf_multi_destroy(m); return nullptr; }
        d.ctx = bf_ctx_create(sensor_rows, sensor_cols, max_scale, max_events_per_device, max_slices_per_device);
        if (!d.ctx) { bf_multi_destroy(m); return nullptr; }
        if (cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaMalloc(&d.d_gather, gbytes) != cudaSuccess) {
            mfail(BF_ERR_CUDA, "bf_multi_create: stream / gather buffer allocation failed");
            bf_multi_destroy(m);
            return nullptr;
        }
        bf_ctx_set_stream(d.ctx, d.stream);   // launches and the collective are ordered on one stream per device
    }
    if (cudaMallocHost(&m->h_gather, gbytes) != cudaSuccess) {
        mfail(BF_ERR_CUDA, "bf_multi_create: pinned gather buffer allocation failed");
        bf_multi_destroy(m);
        return nullptr;
    }
    if (n_devices > 1) {
        std::vector<ncclComm_t> comms((size_t)n_devices);
        const ncclResult_t r = g_nccl.CommInitAll(comms.data(), n_devices, ids.data());
        if (r != 0) {
            mfail(BF_ERR_CUDA, "ncclCommInitAll failed: %s", g_nccl.GetErrorString(r));
            bf_multi_destroy(m);
            return nullptr;
        }
        for (int i = 0; i < n_devices; ++i) m->devs[(size_t)i].comm = comms[(size_t)i];
    }
    bf_cuda_init(ids[0]);
    return m;
}

int bf_multi_device_count(bf_multi *m) { return m ? (int)m->devs.size() : 0; }

int bf_multi_set_option(bf_multi *m, const char *key, long long value) {
    if (!m || !key) return mfail(BF_ERR_ARG, "null argument");
    if (!strcmp(key, "block")) {
        const int b = (int)std::max(1LL, value);
        if (m->n_slices && b != m->block) return mfail(BF_ERR_STATE, "the block size cannot change inside a batch");
        m->block = b;
        return BF_OK;
    }
    for (Dev &d : m->devs) {
        const int rc = bf_ctx_set_option(d.ctx, key, value);
        if (rc != BF_OK) return rc;
    }
    return BF_OK;
}

int bf_multi_reset(bf_multi *m) {
    if (!m) return mfail(BF_ERR_ARG, "null argument");
    for (Dev &d : m->devs) {
        bf_batch_reset(d.ctx);
        d.n_local = 0;
    }
    m->n_slices = 0;
    m->owner.clear();
    m->slot.clear();
    m->ran = false;
    return BF_OK;
}

// Next slice of the batch (global index = order of the calls): goes to device (index / block) % N.
int bf_multi_add_packed(bf_multi *m, const bf_event *events, int n, int scale, int max_iter) {
    if (!m) return mfail(BF_ERR_ARG, "null argument");
    const int k = m->n_slices;
    const int o = bf_multi_owner(k, (int)m->devs.size(), m->block);
    Dev &d = m->devs[(size_t)o];
    const int slot = bf_batch_add_packed(d.ctx, events, n, scale, max_iter, nullptr);   // stm-disabled by construction
    if (slot < 0) return slot;
    d.n_local = slot + 1;
    m->owner.push_back(o);
    m->slot.push_back(slot);
    m->n_slices = k + 1;
    m->ran = false;
    return k;
}

// Upload + one persistent launch per device + ONE all-gather of the result records + D2H on device 0.
// Asynchronous; bf_multi_sync waits.
int bf_multi_run(bf_multi *m, int want_events) {
    if (!m) return mfail(BF_ERR_ARG, "null argument");
    const int nd = (int)m->devs.size();
    int cap = 0;
    for (Dev &d : m->devs) cap = std::max(cap, d.n_local);
    m->cap = cap;
    if (cap == 0) { m->ran = true; return BF_OK; }
    for (Dev &d : m->devs) {
        if (d.n_local == 0) continue;
        int rc;
        if ((rc = bf_batch_upload(d.ctx)) != BF_OK) return rc;
        if ((rc = bf_batch_launch(d.ctx, want_events)) != BF_OK) return rc;
    }
    const size_t chunk = (size_t)cap * sizeof(bf_slice_result);
    if (nd > 1) {
        // equal-size contributions: devices with fewer slices send padding that nobody reads
        MNCCL(g_nccl.GroupStart());
        for (Dev &d : m->devs) {
            void *src = nullptr;
            long long bytes = 0;
            bf_batch_results_device(d.ctx, &src, &bytes);
            const ncclResult_t r = g_nccl.AllGather(src, d.d_gather, chunk, NCCL_CHAR, d.comm, d.stream);
            if (r != 0) {
                g_nccl.GroupEnd();
                return mfail(BF_ERR_CUDA, "ncclAllGather failed: %s", g_nccl.GetErrorString(r));
            }
        }
        MNCCL(g_nccl.GroupEnd());
        Dev &d0 = m->devs[0];
        MCU(cudaSetDevice(d0.device));
        MCU(cudaMemcpyAsync(m->h_gather, d0.d_gather, chunk * (size_t)nd, cudaMemcpyDeviceToHost, d0.stream));
    } else {
        Dev &d0 = m->devs[0];
        void *src = nullptr;
        long long bytes = 0;
        bf_batch_results_device(d0.ctx, &src, &bytes);
        MCU(cudaSetDevice(d0.device));
        MCU(cudaMemcpyAsync(m->h_gather, src, chunk, cudaMemcpyDeviceToHost, d0.stream));
    }
    m->ran = true;
    return BF_OK;
}

int bf_multi_sync(bf_multi *m) {
    if (!m) return mfail(BF_ERR_ARG, "null argument");
    for (Dev &d : m->devs) {
        MCU(cudaSetDevice(d.device));
        MCU(cudaStreamSynchronize(d.stream));
    }
    return BF_OK;
}

int bf_multi_size(bf_multi *m) { return m ? m->n_slices : 0; }

// Result of global slice `slice`, read from the gathered records (valid after bf_multi_sync).
int bf_multi_result(bf_multi *m, int slice, bf_slice_result *out) {
    if (!m || !out || slice < 0 || slice >= m->n_slices) return mfail(BF_ERR_ARG, "bf_multi_result: bad slice");
    if (!m->ran) return mfail(BF_ERR_STATE, "bf_multi_result before bf_multi_run");
    const size_t o = (size_t)m->owner[(size_t)slice], s = (size_t)m->slot[(size_t)slice];
    memcpy(out, m->h_gather + (o * (size_t)m->cap + s) * sizeof(bf_slice_result), sizeof(bf_slice_result));
    return BF_OK;
}

// Where a slice lives: for per-event read-back through bf_batch_events(ctx, slot, ...).
int bf_multi_locate(bf_multi *m, int slice, bf_ctx **ctx, int *slot, int *device) {
    if (!m || slice < 0 || slice >= m->n_slices) return mfail(BF_ERR_ARG, "bf_multi_locate: bad slice");
    const Dev &d = m->devs[(size_t)m->owner[(size_t)slice]];
    if (ctx) *ctx = d.ctx;
    if (slot) *slot = m->slot[(size_t)slice];
    if (device) *device = d.device;
    return BF_OK;
}

long long bf_multi_launch_count(bf_multi *m) {
    long long n = 0;
    if (m) for (Dev &d : m->devs) n += bf_ctx_launch_count(d.ctx);
    return n;
}

}  // extern "C"
