/* bf_cuda.h -- C ABI of the B200-native motion-compensation backend.
 *
 * This is the drop-in boundary for better-flow's per-time-slice hot path.  In the reference the
 * accelerator seam is class AccelLib (better_flow_core/include/better_flow/accel_lib.h:14-616),
 * owned by value by OptimizerRolling (optimizer_rolling.h:19) and selected at run time by
 * OpenCLDriver::enabled (accel_lib.h:41,46; bf_motion_compensator.cpp:132-133).  Every entry
 * point below names the reference interface it replaces.  All paths in comments are relative
 * to /root/reference/better_flow_core/.
 *
 * Conventions: plain pointers and sizes only; the caller owns every host array; the library
 * owns all device memory behind the opaque bf_ctx; one CUDA stream per context; a context is
 * not thread-safe, distinct contexts are.  Return values: >= 0 success (see each function),
 * < 0 error (bf_last_error() gives the text).  There is NO CPU fallback: without a CUDA device
 * every compute entry point fails with BF_ERR_CUDA.
 *
 * Axis convention is the reference's: fr_x = sensor ROW (file "y"), fr_y = sensor COLUMN
 * (file "x") (bf_motion_compensator.cpp:200); image row index = "x" (accel_lib.h:148).
 */
#ifndef BF_CUDA_H
#define BF_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BF_OK 0
#define BF_ERR_ARG (-1)      /* bad argument / capacity exceeded */
#define BF_ERR_CUDA (-2)     /* CUDA runtime error or no device */
#define BF_ERR_STATE (-3)    /* call out of sequence */

/* run() outcomes stored in bf_slice_result.rc */
#define BF_RC_OK 0           /* optimised                      (optimizer_rolling.h:124) */
#define BF_RC_SKIPPED 1      /* guard hit: tiny window or < 1000 events (optimizer_rolling.h:49-58) */
#define BF_RC_ITER_CAP 2     /* safety cap on iterations reached (the reference has none) */
#define BF_RC_DEGENERATE 3   /* no pixel above the 1e-6 occupancy threshold: the reference divides
                                0/0 (object_model.cpp:122-125) and never terminates; we stop */

/* result flags */
#define BF_FLAG_ALL_NOISE 1u /* tiny-window guard fired: the reference marks every event of the
                                slice as noise (optimizer_rolling.h:52-53) */
#define BF_FLAG_T_QUANTISED 2u /* slice too long/dense for exact 64-bit packed sums; timestamps were
                                  right-shifted (never at BASELINE sizes) */

/* High bit of bf_event.fr_y marks an event whose `noise` flag is already set (event.h:11):
 * it is warped but not splatted (accel_lib.h:152). */
#define BF_EVENT_NOISE 0x8000u

/* Compact event record, 8 bytes.  Replaces the 152-byte AoS `Event` (event.h:7-34) on the device:
 * fr_x, fr_y as the reference stores them, t_ns = Event::t after set_local_time (event.h:61-63),
 * i.e. timestamp - slice_start as a signed offset in ns. */
typedef struct bf_event {
    uint16_t fr_x;
    uint16_t fr_y;
    int32_t t_ns;
} bf_event;

/* POD image of ObjectModel's state (object_model.h:10-13), same order. */
typedef struct bf_model {
    double cx, cy, dx, dy, rot, div;
    uint32_t cnt;
    uint32_t pad_;
    double total_dx, total_dy, total_rot, total_div;
} bf_model;

/* Everything OptimizerRolling exposes after run() for one slice. */
typedef struct bf_slice_result {
    bf_model model;          /* get_model()                       (optimizer_rolling.h:285-287) */
    int32_t rc;              /* run() return value + BF_RC_* extensions */
    int32_t iters;           /* run()'s local itercount           (optimizer_rolling.h:60) */
    float dividers[4];       /* x, y, rot, div dividers           (optimizer_rolling.h:36) */
    int32_t x_min, x_max, y_min, y_max;   /* bbox                 (optimizer_rolling.h:252-260) */
    int32_t img_rows, img_cols;           /* scale_img_x/y        (optimizer_rolling.h:276-277) */
    double x_shift, y_shift;              /*                      (optimizer_rolling.h:279-282) */
    int32_t n_events;
    uint32_t flags;          /* BF_FLAG_* */
} bf_slice_result;

typedef struct bf_ctx bf_ctx;

/* ---- device / context --------------------------------------------------------------------- */

/* Replaces OpenCLDriver::init() (src/opencl_driver.cpp:14-60): selects the CUDA device.
 * Returns BF_OK or BF_ERR_CUDA. */
int bf_cuda_init(int device);
int bf_device_count(void);
const char *bf_last_error(void);
const char *bf_version(void);

/* Replaces AccelLib::init_gpu (accel_lib.h:71-145), but pooled: the reference re-allocates device
 * buffers for every slice (a fresh OptimizerRolling per slice, dvs_flow.h:210); a context is
 * created once and reused.
 *   sensor_rows/cols   RES_X / RES_Y (common.h:39-40), now run-time
 *   max_scale          largest `scale` that will be used (1, 3 or 5)
 *   max_events         total event capacity of one batch
 *   max_slices         slice capacity of one batch */
bf_ctx *bf_ctx_create(int sensor_rows, int sensor_cols, int max_scale, long long max_events, int max_slices);
void bf_ctx_destroy(bf_ctx *ctx);

/* Tunables.  Keys: "group_size" (CTAs cooperating on one slice; 0 = auto), "iter_cap",
 * "min_events" (the reference's 1000, optimizer_rolling.h:57). */
int bf_ctx_set_option(bf_ctx *ctx, const char *key, long long value);
long long bf_ctx_get_option(bf_ctx *ctx, const char *key);

/* Run everything of this context on a caller-owned stream (a cudaStream_t passed as void*), e.g. the
 * stream a host framework times with its own events or orders NCCL calls on.  NULL restores the
 * context's private stream. */
int bf_ctx_set_stream(bf_ctx *ctx, void *cuda_stream);

/* ---- batched minimisation: N independent slices in one persistent launch -------------------
 * A "slice" is what DVS_flow::recompute hands to OptimizerRolling (dvs_flow.h:196-222):
 * events in iteration order, local times, scale, max_iter and an optional warm-start model. */

int bf_batch_reset(bf_ctx *ctx);

/* set_cloud + set_time + set_maxiter + [set_model] (optimizer_rolling.h:236-299).
 *   noise      nullable per-event flags
 *   init       nullable; NULL == --stm-disable (dvs_flow.h:218-219)
 * Returns the slot index (>= 0) of the slice inside the batch. */
int bf_batch_add(bf_ctx *ctx, const uint16_t *fr_x, const uint16_t *fr_y, const int32_t *t_ns,
                 const uint8_t *noise, int n, int scale, int max_iter, const bf_model *init);
/* Same, events already in the compact 8-byte layout (no host-side packing pass).  Like bf_batch_add, refuses
 * (BF_ERR_ARG) a batch holding a coordinate outside the context's sensor: the device sizes a slice's images
 * from the bounding box of its events. */
int bf_batch_add_packed(bf_ctx *ctx, const bf_event *events, int n, int scale, int max_iter,
                        const bf_model *init);

/* Pinned host staging area the batch is assembled in (so callers can fill it in place):
 * returns the base of the event staging buffer; *capacity = max_events. */
bf_event *bf_batch_staging(bf_ctx *ctx, long long *capacity);
/* Declare a slice that already lives in the staging buffer at [offset, offset+n) (same coordinate check). */
int bf_batch_add_staged(bf_ctx *ctx, long long offset, int n, int scale, int max_iter,
                        const bf_model *init);

/* Asynchronous pieces (all on the context's stream) ... */
int bf_batch_upload(bf_ctx *ctx);                    /* H2D of events + slice table */
int bf_batch_launch(bf_ctx *ctx, int want_events);   /* OptimizerRolling::run() for every slice */
int bf_batch_download(bf_ctx *ctx);                  /* D2H of the result records */
int bf_batch_sync(bf_ctx *ctx);
/* ... and the synchronous composition upload -> launch -> download -> sync. */
int bf_batch_run(bf_ctx *ctx, int want_events);

/* upload -> launch -> download with the event upload streamed in slice-ordered chunks on a second
 * stream and overlapped with the minimisation of the slices that have already landed (one
 * persistent launch).  Requires slices added back to back in order (bf_batch_add / _packed /
 * _staged with increasing offsets).  Asynchronous: follow with bf_batch_sync. */
int bf_batch_run_streamed(bf_ctx *ctx, int want_events);

/* Times `reps` back-to-back launches on the resident batch with CUDA events on the context's
 * stream (inputs already in HBM).  Returns total milliseconds in *ms. */
int bf_batch_time_launches(bf_ctx *ctx, int reps, int want_events, float *ms);
/* Number of kernels launched by this context so far. */
long long bf_ctx_launch_count(bf_ctx *ctx);

/* Device address and byte size of the result records of the current batch (bf_slice_result[n]),
 * valid after bf_batch_launch: lets a multi-GPU caller hand them to its collective (NCCL gather of
 * the per-slice flow) without a host round trip. */
int bf_batch_results_device(bf_ctx *ctx, void **dev_ptr, long long *bytes);

/* Debug aid: with option "profile" = 1 the minimise kernel accumulates clock64 cycles per phase and
 * per CTA; this copies them out as out[ctas][16] (see PF_* in csrc/bf_device.cuh). */
int bf_debug_profile(bf_ctx *ctx, long long *out, int max_ctas);

int bf_batch_size(bf_ctx *ctx);
int bf_batch_result(bf_ctx *ctx, int slot, bf_slice_result *out);
/* Replaces AccelLib::writeout_events (accel_lib.h:310-329): per-event state after run().
 * Needs want_events != 0 at launch.  Any pointer may be NULL. */
int bf_batch_events(bf_ctx *ctx, int slot, double *pr_x, double *pr_y, double *nx, double *ny);

/* One slice, synchronously == one OptimizerRolling::run() (optimizer_rolling.h:48-125).
 * Returns the run() code (BF_RC_*) or < 0. */
int bf_minimize(bf_ctx *ctx, const uint16_t *fr_x, const uint16_t *fr_y, const int32_t *t_ns,
                const uint8_t *noise, int n, int scale, int max_iter, const bf_model *init,
                bf_slice_result *out, double *pr_x, double *pr_y, double *nx, double *ny);

/* ---- OptimizerLocal: the contrast-driven (nx, ny) descent ------------------------------------
 * Replaces OptimizerLocal(LinearEventCloud*, scale)::run() (include/better_flow/optimizer_sampler.h:41-56,
 * src/optimizer_sampler.cpp:4-38): per step every event is projected with the global (nx, ny)
 * (Event::project, event.h:65-70), splatted into a saturating 8-bit count image, the image is
 * Gaussian-blurred (scale x scale) and scored by the mean of its non-zero pixels; nx, ny are moved
 * alternately by +-dn, dn halving and flipping whenever the score does not improve, until
 * hypot(dnx, dny) <= dn_th.  Runs in the same persistent launch as the rolling slices of a batch.
 * t_ns is Event::t as the caller has it (the class never calls set_local_time).  scale: 1, 3 or 5.
 * Result record (bf_slice_result) of such a slice:
 *   model.total_dx = nx, model.total_dy = ny      (get_nx / get_ny)
 *   model.dx = last_score, model.dy = dnx, model.rot = dny, model.div = dn_th
 *   model.cnt = non-zero pixels of the last image, iters = iteration_steps executed
 *   rc = 0 ok / BF_RC_SKIPPED window too small (run() returns 1, optimizer_sampler.cpp:9-13) */
int bf_batch_add_local(bf_ctx *ctx, const uint16_t *fr_x, const uint16_t *fr_y, const int32_t *t_ns, int n, int scale);
/* Switch an already added slice between OptimizerRolling (0) and OptimizerLocal (1) mode. */
int bf_batch_slot_mode(bf_ctx *ctx, int slot, int mode);
/* One cloud, synchronously.  Returns rc or < 0. */
int bf_local_minimize(bf_ctx *ctx, const uint16_t *fr_x, const uint16_t *fr_y, const int32_t *t_ns, int n, int scale,
                      bf_slice_result *out);

/* ---- stage-level entry points (the AccelLib surface; used by the kernel parity tests) ------ */

/* AccelLib::get_time_img (accel_lib.h:211-217 -> 147-178): mean-timestamp image in seconds,
 * (w+scale) x (h+scale) floats, row-major. */
int bf_time_img(bf_ctx *ctx, int n, const double *pr_x, const double *pr_y, const int32_t *t_ns,
                const uint8_t *noise, int w, int h, int scale, int x_sh, int y_sh, float *out);

/* AccelLib::fast_model(get_time_img(...)) (accel_lib.h:337-341 -> object_model.h:31-34):
 * out7 = cx, cy (image units), dx, dy, rot, div, cnt.  gx/gy nullable: the Scharr images of
 * AccelLib::Sobel (accel_lib.h:400-434, 513-543). */
int bf_fast_model(bf_ctx *ctx, int n, const double *pr_x, const double *pr_y, const int32_t *t_ns,
                  const uint8_t *noise, int w, int h, int scale, int x_sh, int y_sh,
                  double *out7, float *gx, float *gy);

/* AccelLib::fast_model(ObjectModel&, cv::Mat&) on a caller-supplied mean-timestamp image
 * (accel_lib.h:337-398 -> object_model.cpp:4-39,103-126) and AccelLib::Sobel (accel_lib.h:400-434):
 * img is rows x cols f32 row-major.  out7 (nullable) = cx, cy, dx, dy, rot, div, cnt; gx/gy nullable. */
int bf_model_from_image(bf_ctx *ctx, int rows, int cols, const float *img, double *out7, float *gx, float *gy);

/* AccelLib::project_4param_reinit (accel_lib.h:263-267 -> event.h:99-110,164-168), in place on
 * pr_x/pr_y; nx/ny nullable outputs. */
int bf_project(bf_ctx *ctx, int n, const uint16_t *fr_x, const uint16_t *fr_y, const int32_t *t_ns,
               double *pr_x, double *pr_y, double *nx, double *ny,
               double dnx, double dny, double cx, double cy, double div, double crl);

/* ---- slice-sharded multi-GPU front ----------------------------------------------
 * The reference has no multi-device path at all (one OpenCL queue on platform[0]/device[0],
 * src/opencl_driver.cpp:25-39).  Slices are independent under --stm-disable semantics
 * (dvs_flow.h:218-219), so a batch is dealt to N devices in blocks of `block` consecutive slices
 * (option "block", default 4), every device minimises its share with its own persistent launch, and
 * the per-slice result records are exchanged with ONE ncclAllGather per batch.  One host process
 * drives all devices (ncclCommInitAll); NCCL is loaded with dlopen when n_devices > 1.
 * Warm-started chains cannot be sharded: init models are not accepted here. */
typedef struct bf_multi bf_multi;

/* devices == NULL means 0..n_devices-1.  Capacities are per device. */
bf_multi *bf_multi_create(int n_devices, const int *devices, int sensor_rows, int sensor_cols, int max_scale,
                          long long max_events_per_device, int max_slices_per_device);
void bf_multi_destroy(bf_multi *m);
int bf_multi_device_count(bf_multi *m);
/* "block", or any bf_ctx_set_option key (applied to every device's context). */
int bf_multi_set_option(bf_multi *m, const char *key, long long value);
/* Owner device index of global slice k: (k / block) % n_devices. */
int bf_multi_owner(int slice, int n_devices, int block);

int bf_multi_reset(bf_multi *m);
/* Appends the next slice (global index = call order); returns that index. */
int bf_multi_add_packed(bf_multi *m, const bf_event *events, int n, int scale, int max_iter);
/* H2D + launch on every device, the all-gather, D2H of the gathered records.  Asynchronous. */
int bf_multi_run(bf_multi *m, int want_events);
int bf_multi_sync(bf_multi *m);
int bf_multi_size(bf_multi *m);
int bf_multi_result(bf_multi *m, int slice, bf_slice_result *out);
/* Context / slot / CUDA device a slice was minimised on (per-event read-back via bf_batch_events). */
int bf_multi_locate(bf_multi *m, int slice, bf_ctx **ctx, int *slot, int *device);
long long bf_multi_launch_count(bf_multi *m);


/* ---- compact upload format (end-to-end path) ----------------------------------------------------------------------
 * The 8-byte bf_event is what the kernel reads; over PCIe a slice travels as 6-byte DELTA records: 12-bit fr_x, 12-bit
 * fr_y, the noise bit and a 23-bit time difference to the previous event of the slice (slices are handed over
 * newest -> oldest, dvs_flow.h:196-198, so local times decrease: dt = t[i-1] - t[i] >= 0), with an absolute local time
 * every 1024 events (one 16-byte block descriptor).  bf_batch_add_delta packs a slice into the context's pinned delta
 * staging; when EVERY slice of a batch was added this way, bf_batch_run_streamed uploads the 6-byte records (25 % fewer
 * bytes: with 8 GPUs on one host the H2D of 8 x 426 MB per step is what bounds the end-to-end rate) and the
 * group of CTAs that takes a slice expands its records into bf_event records in its prologue (the pass that computes
 * the slice's bounding box anyway), so the upload still streams in chunks under the running kernel.
 * Returns the slot, or BF_ERR_ARG when the slice cannot be represented (a coordinate >= 4096, a time step backwards or
 * a gap of 2^23 ns = 8.4 ms or more between consecutive events): add it with bf_batch_add_packed instead (the whole
 * batch then travels as 8-byte records).  The results are bit-identical either way. */
int bf_batch_add_delta(bf_ctx *c, const bf_event *events, int n, int scale, int max_iter, const bf_model *init);
long long bf_batch_upload_bytes(bf_ctx *c);   /* host-to-device bytes of one bf_batch_run_streamed of the current batch */

/* ---- debug images ----------------------------------------------------------------------------------
 * EventFile::projection_img (event_file.h:460-515): the "motion-compensated event image" the reference dumps with
 * --img / --video (dvs_flow.h:256-260) and publishes from its ROS node.  Per non-noise event: x = int(pr_x * scale),
 * y = int(pr_y * scale) (C truncation), rejected unless 0 <= x < scale * (RES_X - 1) and likewise y (:486-487); a
 * saturating 8-bit count is splatted on the scale x scale block centred on (x + scale/2, y + scale/2) (:489-501; with
 * those bounds the block never leaves the image, so the clamps of :494-495 never bite); cv::GaussianBlur(scale x scale)
 * (:504-506); then the image is scaled by 127 / nonzero_average (event_file.cpp:282-294) with cv::convertScaleAbs
 * (:508-509).  out = [RES_X * scale][RES_Y * scale] bytes, row-major; *nz_avg (nullable) receives the nonzero average
 * of the blurred count image.  Pass pr_x / pr_y for the warped image, the events' fr_x / fr_y for show_final = true.
 * An image without any event is returned all-zero (the reference divides by zero there). */
int bf_projection_img(bf_ctx *c, int n, const double *pr_x, const double *pr_y, const uint8_t *noise, int scale,
                      uint8_t *out, double *nz_avg);

/* EventFile::color_time_img (event_file.h:649-747): the second debug image of --img / --video (dvs_flow.h:257-259).  Per
 * non-noise event a phase angle 2 * 3.14 * (t - t_min) / (t_max - t_min) of its local time (t range over ALL events);
 * cos / sin of it are averaged per pixel over the scale x scale blocks the events cover (:696-722; the image spans the
 * full frame, :668-680: (scale * RES_X + scale) x (scale * RES_Y + scale) pixels); the mean direction becomes hue
 * (angle / 2), its length saturation, value 255 (:724-739), converted HSV -> BGR like cv::cvtColor (:742-746).
 * out = [rows][cols][3] bytes, B G R.  The reference accumulates in f32 in event order; here the sums are f64 atomics
 * (order-free), so single pixels may differ by one level from an OpenCV-rendered image (tests allow +-1 on < 1 %). */
int bf_color_time_img(bf_ctx *c, int n, const double *pr_x, const double *pr_y, const int32_t *t_ns, const uint8_t *noise,
                      int scale, uint8_t *out_bgr);

/* ---- device-resident slice ring: DVS_flow's default mode ---------------------------------------
 * Replaces, for the reference's default operating mode (overlapping 50 k-event / 200 ms windows re-minimised every
 * 20 k events / 33 ms, each warm-started from the previous model), the per-slice hand-over of
 * DVS_flow::recompute (dvs_flow.h:184-252): there every slice walks the whole CircularArray (datastructures.h:6-115)
 * to build a LinearEventPtrs, resets / re-times every event, and -- with an accelerator behind AccelLib -- would
 * re-upload the 60-95 % of the window it already sent for the previous slice (accel_lib.h:83-124 uploads per
 * optimiser instance).  Here the ring lives on the device:
 *   bf_ring_push    appends only the NEW events (absolute timestamps, 16-byte records) -- asynchronous H2D;
 *   bf_ring_reserve / bf_ring_commit   the same without the staging copy: `reserve` hands out room for up to n events
 *                   in the ring's pinned staging buffer (waiting for earlier copies that still read that part), the
 *                   caller writes its events there as they arrive, `commit(m <= n)` checks and sends the first m (a
 *                   refused commit leaves the reservation open; a new reservation replaces an open one, whose events
 *                   are then not sent; bf_ring_push is refused while a reservation is open);
 *   bf_ring_slice   enqueues "minimise the newest n events, local time = timestamp - slice_start"
 *                   (OptimizerRolling::set_cloud / set_time / [set_model] / run, optimizer_rolling.h:236-299,48-125):
 *                   a small kernel cuts the packed newest->oldest slice out of the ring (the order of
 *                   dvs_flow.h:196-198), the persistent kernel minimises it.  With chain != 0 the warm start
 *                   (set_model(last_model), dvs_flow.h:218-219) is taken ON THE DEVICE from the previous slice's
 *                   result record, so the call returns without waiting for that slice: a whole warm-start chain is
 *                   stream-ordered device work and the host only enqueues.  The tiny-window guard's noise marking
 *                   (optimizer_rolling.h:49-55) is applied to the ring entries on the device as well.
 *   bf_ring_result  returns the record of one slice (tickets are results of bf_ring_slice, in order; a ticket stays
 *                   readable until `max_pending` later slices have been enqueued).  Records are fetched on demand: the
 *                   first request for a ticket that has not been fetched yet waits for everything enqueued so far.
 *   bf_ring_seed    sets the model the NEXT chained slice starts from (the caller's last_model when a ring is created
 *                   in the middle of a stream -- after a context was re-created, or when the host path handled the
 *                   slices before); without it the first chained slice starts from ObjectModel().  The noise marks of
 *                   the slice before are not carried over.
 * A ring belongs to one context and shares its stream, event buffer and images with the batch entry points (calls
 * are serialised in stream order).  Per-event outputs are not available through the ring (use bf_minimize). */
typedef struct bf_ring bf_ring;
typedef struct bf_ring_event {
    uint16_t fr_x, fr_y;     /* as in bf_event (fr_y may carry BF_EVENT_NOISE) */
    uint32_t reserved;
    uint64_t timestamp;      /* Event::timestamp, ns (event.h:13) */
} bf_ring_event;
bf_ring *bf_ring_create(bf_ctx *c, long long capacity, int max_pending);
void bf_ring_destroy(bf_ring *r);
int bf_ring_push(bf_ring *r, const bf_ring_event *ev, int n);
int bf_ring_reserve(bf_ring *r, int n, bf_ring_event **where);   /* n <= capacity */
int bf_ring_commit(bf_ring *r, int n);
int bf_ring_slice(bf_ring *r, int n, uint64_t slice_start, int scale, int max_iter, int chain);   /* -> ticket */
int bf_ring_result(bf_ring *r, int ticket, bf_slice_result *out);
int bf_ring_seed(bf_ring *r, const bf_model *model);
int bf_ring_sync(bf_ring *r);
long long bf_ring_pushed(bf_ring *r);   /* events appended so far */

#ifdef __cplusplus
}
#endif
#endif /* BF_CUDA_H */
