// bf_device.cuh -- sm_100a device code of the motion-compensation hot path.
//
// Data layout in HBM / L2:
//   events   bf_event[n]        8 B/event, read-only, ld.global.nc
//   state    float2[n]          carried between iterations.  The warped position pr that the next
//            re-projection starts from (event.h:100) is  float(fr) - m / 10000.0  with the f32
//            product m = k * float(t), k = float(float(n) / nz)  (event.h:164-168), so storing the
//            two f32 values (m_x, m_y) reproduces pr exactly at half the bytes of two doubles
//   image    u64[rows_alloc][pitch]  the POINT image: every event adds ONE packed word
//            (count | sum of t) at its centre pixel with one 64-bit integer atomic.  The
//            reference's s x s splat (accel_lib.h:160-165) is recovered exactly in the image pass
//            as an s x s box sum of the point image (integer sums are associative).  A zero border
//            of BF_BORDER pixels surrounds the image so patch loads never need bounds checks.
//   Two such images (and flag arrays) per CTA group: iteration k splats into image k&1; the image
//   pass of iteration k reads image k&1 and, in passing, zeroes the cells of image (k-1)&1 that
//   were live one iteration earlier (coalesced row stores), so every image is all-zero again
//   before it is splatted into the next time.
//   flags    u32[cells]         one generation tag per 8 x (32-2H) pixel cell: an event stamps the
//            current iteration's tag on every cell whose haloed patch contains its pixel; the image
//            pass visits only cells carrying the current tag (the image is 1-5 % occupied).
//
// Reference paths are relative to /root/reference/better_flow_core/.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include "bf_logic.h"

// The passes of the minimise kernel are inlined into every instance of the slice loop (3 scales x own slice / helped
// slice); left to its heuristics the compiler stops inlining them once the kernel grows (measured with the TMA
// variant: event_pass became a real call, its arguments went through a 592-byte stack frame, the event pass took
// +79 %).  BF_PASS_INLINE pins the decision.
#ifndef BF_PASS_INLINE
#define BF_PASS_INLINE __forceinline__
#endif

// ---- compile-time geometry -------------------------------------------------------------------
#ifndef BF_NT
#define BF_NT 512              // threads per CTA (16 warps)
#endif
#define BF_NW (BF_NT / 32)
#ifndef BF_MIN_CTAS
#define BF_MIN_CTAS 1          // resident CTAs per SM the minimise kernel is compiled for (register cap = 65536 / (BF_NT * this))
#endif
#define BF_BORDER 4            // zero border of the stored image (>= scale/2 + 1)
#define BF_CELL_ROWS 8         // output rows per cell
#ifndef BF_LIST_CAP
#define BF_LIST_CAP 4096       // active-cell list entries per scan chunk (a multiple of BF_NT)
#endif

typedef unsigned long long u64;

// The image pass works on CELLS: 8 output rows x (32 - 2H) output columns, H = scale/2 + 1 being the
// halo needed by the box sum (scale/2) plus the 3x3 Scharr (1).  A warp owns a cell: lane l holds
// column l of the (8 + 2H) x 32 point patch in registers, so horizontal neighbours are shuffles.
// OptimizerLocal's image is the box-summed count image followed by a (2 SH + 1)^2 Gaussian blur: its halo is the box
// radius plus the blur radius, 2 SH, which exceeds SH + 1 at scale 5 only -- LOCAL selects that geometry (H = 4, 24
// output columns per cell) for the slices minimised by OptimizerLocal at scale 5; every other case is unchanged.
template <int SH, bool LOCAL = false> struct CellCfg {
    static constexpr int H = (LOCAL && 2 * SH > SH + 1) ? 2 * SH : SH + 1;
    static constexpr int PR = BF_CELL_ROWS + 2 * H;   // patch rows held per lane
    static constexpr int AR = BF_CELL_ROWS + 2;       // mean-time rows (1-row halo for Scharr)
    static constexpr int CW = 32 - 2 * H;             // output columns per cell
};
#define BF_CW_MIN 24   // CellCfg<2, true>::CW (the narrowest cell), for sizing the flag array

struct SliceDesc {
    long long ev_off;   // first event of the slice in the batch arrays
    int n;
    int scale;
    int max_iter;
    int has_init;
    int mode;           // 0: OptimizerRolling::run, 1: OptimizerLocal::run
    int block0;         // compact upload (bf_batch_add_delta): first DeltaBlock of the slice, or -1
    bf_model init;
};

// Compact upload format: up to 1024 consecutive events of one slice as 6-byte delta records (include/bf_cuda.h).
struct DeltaBlock {
    long long first;         // index of the block's first event in the batch arrays
    int count;
    int t0;                  // local time of the first event
};
#define BF_DELTA_BLOCK 1024

// Per-group control block in global memory (one 256-byte record per group).
struct GroupWs {
    unsigned bar;            // monotonically increasing arrival counter
    unsigned pad0[31];
    int cur_slice;
    int victim;              // helper side: group this group is about to help (broadcast inside the helper group)
    unsigned victim_seq;     //              value of the victim's join_seq when it was claimed
    int pad1[1];
    int bbox[2][8];          // [parity]: x_min, x_max, y_min, y_max, t_min, t_max
    // ---- tail helping (see HELPING below) ----
    unsigned help_state;     // HELP_*: written by the group's leader and by the claiming helper (CAS)
    unsigned join_seq;       // bumped (release) by the leader when a JoinRecord has been published
    int grow_iter;           // iteration whose barrier B carried the leader's decision to grow
    int pad2[9];
};

// HELPING.  Iteration counts per slice vary several-fold, so when the slice queue runs dry some
// groups are still in the middle of a long slice while most SMs idle (16-19 % of the SM-time on the
// benchmark batch).  A group that finds the queue empty therefore does not exit: it claims a group
// that is still minimising (CAS on help_state) and JOINS it at an iteration boundary -- the slice is
// then worked on by more CTAs.  All cross-iteration state of a slice lives in global memory (events,
// states, images, flags), so joining only needs the few hundred bytes of per-slice scalars, which
// the victim's leader publishes in a JoinRecord.  Protocol (leader = rank 0, thread 0 of the victim):
//   helper : state 1 -> 2 by CAS, then polls join_seq (joined) / help_state (0: slice ended, give up);
//   victim : leader samples help_state before barrier B of iteration k; if it is 2 it stores
//            grow_iter = k (visible to the whole group after barrier B); if the GD step then says
//            "continue", every member switches to G + G_base CTAs from iteration k+1 on and the leader
//            publishes the record; otherwise the slice ends and help_state = 0 releases the helper.
// Helpers join one group at a time (the leader re-opens the slice once the previous helper is in), up
// to BF_MAX_GROW groups per slice.  Only OptimizerRolling slices are helped.
enum { HELP_CLOSED = 0, HELP_OPEN = 1, HELP_CLAIMED = 2, HELP_JOINED = 3 };
#define BF_MAX_GROW 16         // capacity: a slice is worked on by at most this many groups (run-time limit: KParams::max_grow)

struct JoinRecord {
    SliceDesc sd;
    BfGeom g;
    BfPack pk;
    BfOpt opt;
    BfProj proj;
    int slice;
    int iter_next, buf;
    unsigned tag, bar_target;
    int G_new, rank_base;
};

struct KParams {
    const bf_event *events;
    float2 *state;           // (m_x, m_y) per event, see above
    double2 *pr_out;         // optional output (pr_x, pr_y), may be null
    double2 *nxy;            // optional output (nx, ny), may be null
    const SliceDesc *slices;
    bf_slice_result *results;
    int n_slices;
    int *queue;              // next slice to hand out
    GroupWs *ws;
    double *partials;        // [n_groups][G][BF_NSUMS]
    u64 *images;             // [n_groups][2][img_elems]
    long long img_elems;
    int pitch;               // elements per stored image row
    unsigned *flags;         // [n_groups][2][flag_elems] per-cell generation tags, one array per image
    long long flag_elems;
    unsigned tag_base;       // launch sequence number << 20: tags are never reused, so flags need no clearing
    int G;                   // CTAs per group
    int res_x, res_y;        // sensor rows / cols
    int min_events;          // 1000 (optimizer_rolling.h:57)
    int iter_cap;
    int want_events;
    int tab_rows, tab_cols;  // capacity of the per-slice cell tables in dynamic shared memory (max image rows / cols)
    int bm_words;            // words of the per-CTA stamp bitmap behind the cell tables (0: stamp the global flags directly)
    int allow_help;          // tail helping enabled (needs BF_MAX_GROW x G partial-sum slots per group)
    int part_stride;         // partial-sum records per group (G, or max_grow x G with helping)
    int max_grow;            // a slice is worked on by at most this many groups (<= BF_MAX_GROW)
    JoinRecord *join;        // [n_groups]
    unsigned *groups_done;   // groups that found the slice queue empty
    const unsigned *ready;   // optional: number of slices whose events have landed in HBM (streamed upload)
    const bf_slice_result *chain_src;   // optional: record whose model warm-starts a slice with has_init == 2 (bf_ring_slice)
    bf_event *events_w;      // compact upload: the event buffer, writable (the groups expand their slices into it), else null
    const unsigned short *delta_rec;   //            the 6-byte records (3 x u16 per event), or null
    const DeltaBlock *delta_blocks;
    int tma_off;             // BF_TMA_PATCH: byte offset of the per-warp tile buffers in dynamic shared memory (128-byte aligned)
    int tma_tile_elems;      //               u64 elements per warp buffer
    long long *prof;         // optional [gridDim.x][BF_NPROF] cycle counters per phase (debug), else null
};
#define BF_NPROF 16
enum { PF_EVENT = 0, PF_BAR_A, PF_SCAN, PF_CELLS, PF_REDUCE, PF_BAR_B, PF_SERIAL, PF_PROLOGUE, PF_FINAL, PF_ITERS, PF_SLICES,
       PF_BAR_A_SPIN, PF_BAR_B_SPIN, PF_TOTAL };

// ---- small PTX helpers ------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ void red_release_add_u32(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// The per-event splat: a fire-and-forget 64-bit reduction.  Written as PTX because the compiler only
// turns an atomicAdd whose result is unused into RED when the kernel contains no fences (with the
// helping protocol's fences it emitted the returning ATOMG form: +65 % event-pass time, measured).
__device__ __forceinline__ void red_add_u64(u64 *p, u64 v) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
#ifndef BF_EVICT
#define BF_EVICT 1             // event / state streams are loaded / stored with the streaming (.cs, L2 evict-first) hint: 426 MB of
                               // each pass through L2 per iteration and would otherwise push out the images (+0.8 %, same-box A/B)
#endif
// EVENT LOADS.  The event buffer is read-only for the kernel EXCEPT under the streamed upload
// (bf_batch_run_streamed), where the copy engine is still filling it in slice order while the kernel works on
// the slices that have landed.  A 32-byte sector (4 events) can straddle two slices, so a load that allocates in
// L1 (ld.global.nc, LDG.E.CONSTANT) may cache the not-yet-written first events of the NEXT slice, and a CTA that
// later minimises that slice on the same SM could hit the stale sector.  Events are therefore loaded L2-coherent
// (ld.global.cg = LDG.E.STRONG.GPU: never served from L1); nothing is lost, an event is read once per iteration by
// one thread, so L1 never had a hit to offer.  BF_EV_LD selects the variant for same-box A/B runs:
//   0  ld.global.cs.nc   (round 1: L1-allocating, evict-first; unsafe under the streamed upload)
//   1  ld.global.cg      (L2-coherent; default)
//   2  ld.relaxed.gpu + createpolicy L2::evict_first  (L2-coherent and evict-first in L2)
#ifndef BF_EV_LD
#define BF_EV_LD 1
#endif
__device__ __forceinline__ uint4 ld_nc_u32x4(const void *p) {
    uint4 v;
#if BF_EV_LD == 0
    asm volatile("ld.global.cs.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
#elif BF_EV_LD == 2
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.relaxed.gpu.global.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
#else
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
#endif
    return v;
}
// the per-event state: read and rewritten once per iteration by the same thread
__device__ __forceinline__ float4 ld_state4(const float4 *p) {
#if BF_EVICT
    return __ldcs(p);
#else
    return *p;
#endif
}
__device__ __forceinline__ void st_state4(float4 *p, float4 v) {
#if BF_EVICT
    __stcs(p, v);
#else
    *p = v;
#endif
}

__device__ __forceinline__ uint2 ld_nc_u32x2(const void *p) {   // (one event; same coherence rule as ld_nc_u32x4)
    uint2 v;
#if BF_EV_LD == 0
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
#else
    asm volatile("ld.global.cg.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
#endif
    return v;
}

// INVARIANT of the relaxed protocol below: every load of data that ANOTHER CTA wrote during this launch must bypass
// L1 -- ld.cg / ld.relaxed.gpu (cell flags, image patches, partial sums, bbox, slice ids, join records: __ldcg,
// ld_relaxed_u32, copy_cg) or, for the event buffer under the streamed upload, ld.global.cg (ld_nc_u32x4).  The only
// L1-cached cross-iteration data is the per-event state, which a thread re-reads after writing it itself; when a group
// grows (helping) its members invalidate L1 explicitly (slice_loop).  compute-sanitizer's racecheck does not cover
// global-memory races, so the evidence for this protocol is the BF_STRICT_SYNC validation build (same results).
#ifndef BF_STRICT_SYNC
#define BF_STRICT_SYNC 0
#endif
// Barrier over the G CTAs of one group (all co-resident: cooperative launch).  `target` is the
// CTA-local running arrival total; the counter is zeroed by the host before every launch.
__device__ __forceinline__ long long group_barrier(unsigned *counter, unsigned &target, int G) {
    __syncthreads();
    target += (unsigned)G;
    long long spin = 0;
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        red_release_add_u32(counter, 1u);
        // Poll with a RELAXED (L2) load and NO acquire fence.  An ld.acquire / fence.acq_rel.gpu here
        // compiles to CCTL.IVALL -- a full L1 invalidation that the next load of the SM waits ~20k
        // cycles for (measured) -- and it buys nothing: every cross-CTA read in this kernel (cell flags,
        // image patches, partial sums, bbox, slice id) is an L2-coherent ld.cg / relaxed.gpu load, the
        // producers drained their writes with MEMBAR.GPU before arriving (red.release), and the other
        // threads of this CTA are ordered behind this poll by the bar.sync below.
#if BF_STRICT_SYNC
        // validation build (-DBF_STRICT_SYNC=1, tools/gpu_strict.sh): the formally correct acquire -- the GPU tests
        // must give the same results with it, which is the evidence that the relaxed protocol above loses nothing
        while (ld_acquire_u32(counter) < target) {
        }
        fence_acq_rel_gpu();
#else
        while (ld_relaxed_u32(counter) < target) {
        }
#endif
        spin = clock64() - t0;
    }
    __syncthreads();
    return spin;   // cycles thread 0 spent between its arrival and the release (valid in thread 0)
}

// ---- thread-block-cluster flavour of the group (CLUSTER instances of the kernel) ----------------------------------
// When a group of G <= 16 CTAs is launched as ONE thread-block cluster, its barrier is the hardware cluster barrier
// (barrier.cluster.arrive.release / wait.acquire -> UCGABAR_ARV / UCGABAR_WAIT; every thread arrives, so every thread's
// splats are released: no bar.sync + thread-0 handshake through a global counter) and the per-CTA partial sums are read
// straight out of the peers' shared memory (mapa + ld.shared::cluster: distributed shared memory) instead of going
// through global memory.  Used for the latency-bound case -- ONE slice at a time, the warm-start chain of DVS_flow's
// default mode -- where the ~2 x 3 us of the global-memory barrier and the partial-sum round trip are a third of an
// iteration; for throughput batches the 2-CTA... 4-CTA groups on the global barrier stay (clusters of 8 strand SMs on
// this GPU and larger groups lose).
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double ld_dsmem_f64(const double *own_smem, unsigned peer_rank) {
    unsigned ra;
    double v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"((unsigned)__cvta_generic_to_shared(own_smem)), "r"(peer_rank));
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
    return v;
}
template <bool CLUSTER>
__device__ __forceinline__ long long sync_group(unsigned *counter, unsigned &target, int G) {
    if constexpr (CLUSTER) {
        cluster_barrier();
        return 0;
    } else {
        return group_barrier(counter, target, G);
    }
}
#define BF_MAX_CLUSTER 16

// ---- exact division by the two constants of Event::apply_project (event.h:164-168) ----------
// q = fma(fma(-b, a*r, a), r, a*r) with r = rn(1/b) equals the correctly rounded a / b for EVERY
// finite f32-valued a, for b = 127 and b = 10000 (checked exhaustively over all 2^32 floats,
// tests/test_divconst.py).  3 FP64 instructions instead of a ~30-instruction IEEE divide.
__device__ __forceinline__ double div_const(double a, double b, double r) {
    const double q0 = __dmul_rn(a, r);
    const double rem = __fma_rn(-b, q0, a);
    return __fma_rn(rem, r, q0);
}

// Event::apply_project (event.h:164-168) for one coordinate, in two halves:
//   m  = k * float(t),  k = float(n) / nz   (f64 divide rounded to f32, then an f32 multiply)
//   pr = float(fr) - m / 10000.0            (f64 divide, f64 subtract)
__device__ __forceinline__ float slope_time(double n, float tf) {
    const float k = (float)div_const((double)(float)n, 127.0, 1.0 / 127.0);
    return __fmul_rn(k, tf);
}
__device__ __forceinline__ double warp_from_m(float m, double fr) {
    return __dsub_rn(fr, div_const((double)m, 10000.0, 1.0 / 10000.0));
}
// Exact u32 (< 2^32) -> f64 on the FP64 pipe (2^52 exponent trick) instead of an I2F on the XU pipe.
__device__ __forceinline__ double u32_to_double(unsigned v) {
    return __dsub_rn(__hiloint2double(0x43300000, (int)v), 4503599627370496.0);
}

// Event::project_4param_reinit + apply_project (event.h:99-110,164-168) for one event.
// All FP64 operations individually rounded (no contraction), in the reference's order.
__device__ __forceinline__ void project_event(double &prx, double &pry, double &ex, double &ey, float &mx, float &my,
                                              double frx, double fry, float tf, const BfProj &q) {
    const double rx = __dsub_rn(prx, q.cx), ry = __dsub_rn(pry, q.cy);                           // :100
    const double qx = __dsub_rn(__dmul_rn(q.c, rx), __dmul_rn(q.s, ry));                        // :102
    const double qy = __dadd_rn(__dmul_rn(q.s, rx), __dmul_rn(q.c, ry));                        // :103
    const double dx = __dadd_rn(__dmul_rn(-qx, q.div), __dsub_rn(qx, rx));                      // :105
    const double dy = __dadd_rn(__dmul_rn(-qy, q.div), __dsub_rn(qy, ry));
    ex = __dadd_rn(dx, q.dnx);                                                                  // :107
    ey = __dadd_rn(dy, q.dny);                                                                  // :108
    mx = slope_time(ex, tf);
    my = slope_time(ey, tf);
    prx = warp_from_m(mx, frx);
    pry = warp_from_m(my, fry);
}

// Pixel of an event in the time image, AccelLib::get_time_img_cpu (accel_lib.h:154-158):
//   int x = pr_x * scale + x_sh (f64 arithmetic, C truncation toward zero), rejected unless
//   half <= x < w + half.
// The test is done on the truncated integers, as the reference does it.  Out-of-range values:
// x86's cvttsd2si returns INT_MIN for NaN / overflow (rejected by x < half); cvt.rzi.s32.f64
// saturates to INT_MIN / INT_MAX (both rejected) and maps NaN to 0, which `x < half` rejects too
// except for half == 0 (scale 1), where NaN is tested explicitly.
struct PixelMap {
    double sc, xs, ys;       // scale, x_sh, y_sh as f64
    int half, w, h;
    int pitch;
};
__device__ __forceinline__ void make_pixel_map(PixelMap &m, const BfGeom &g, int pitch) {
    m.sc = (double)g.scale; m.xs = (double)g.x_sh; m.ys = (double)g.y_sh;
    m.half = g.half; m.w = g.w; m.h = g.h;
    m.pitch = pitch;
}
__device__ __forceinline__ bool event_pixel(double prx, double pry, const PixelMap &m, int &x, int &y) {
    const double fx = __dadd_rn(__dmul_rn(prx, m.sc), m.xs);
    const double fy = __dadd_rn(__dmul_rn(pry, m.sc), m.ys);
    x = __double2int_rz(fx);
    y = __double2int_rz(fy);
    bool ok = (unsigned)(x - m.half) < (unsigned)m.w && (unsigned)(y - m.half) < (unsigned)m.h;
    if (m.half == 0) ok = ok && fx == fx && fy == fy;
    return ok;
}
__device__ __forceinline__ long long pixel_offset(int x, int y, int pitch) {
    return (long long)(x + BF_BORDER) * pitch + (y + BF_BORDER);
}

// Stamp `tag` on every cell whose haloed patch contains pixel (x, y): at most 2 x 2 cells.
template <int SH>
__device__ __forceinline__ void mark_cells(unsigned *flags, unsigned tag, int x, int y, int n_ci, int n_cj) {
    typedef CellCfg<SH> C;
    const int ci = x >> 3, cj = y / C::CW;
    const int lx = x & 7, ly = y - cj * C::CW;
    unsigned *f = flags + ci * n_cj + cj;
    *f = tag;
    const int dj = (ly < C::H && cj > 0) ? -1 : ((ly >= C::CW - C::H && cj + 1 < n_cj) ? 1 : 0);
    const int di = (lx < C::H && ci > 0) ? -n_cj : ((lx >= BF_CELL_ROWS - C::H && ci + 1 < n_ci) ? n_cj : 0);
    if (dj != 0) f[dj] = tag;
    if (di != 0) {
        f[di] = tag;
        if (dj != 0) f[di + dj] = tag;
    }
}

// The same stamping with the per-row / per-column parts looked up instead of computed: the cell
// geometry of a slice is fixed for all its iterations, so every CTA tabulates it once per slice in
// shared memory (the computed form is ~45 instructions per event, the table form ~12).
//   row_tab[x] = (first flag index of x's cell row, +-n_cj or 0: offset to the neighbouring cell row to stamp too)
//   col_tab[y] = (cell column of y, +-1 or 0)
// `margin`: how far outside a cell's interior an event still matters to that cell's OUTPUT pixels.  An
// output pixel only contributes when its own mean time is occupied, i.e. when an event lies within the
// box radius SH of it -- so the neighbour cell is stamped for events within SH of the border, not within
// the patch halo H = SH + 1 (the halo is what the cell READS, not what makes it live).  OptimizerLocal's
// blur spreads one more box radius: margin 2 * SH.
template <int SH, bool LOCAL = false>
__device__ __forceinline__ void fill_cell_tables(int2 *row_tab, short2 *col_tab, int rows, int cols, int margin) {
    typedef CellCfg<SH, LOCAL> C;
    const int n_ci = (rows + BF_CELL_ROWS - 1) / BF_CELL_ROWS, n_cj = (cols + C::CW - 1) / C::CW;
    for (int x = threadIdx.x; x < rows; x += blockDim.x) {
        const int ci = x >> 3, lx = x & 7;
        const int di = (lx < margin && ci > 0) ? -n_cj : ((lx >= BF_CELL_ROWS - margin && ci + 1 < n_ci) ? n_cj : 0);
        row_tab[x] = make_int2(ci * n_cj, di);
    }
    for (int y = threadIdx.x; y < cols; y += blockDim.x) {
        const int cj = y / C::CW, ly = y - cj * C::CW;
        const int dj = (ly < margin && cj > 0) ? -1 : ((ly >= C::CW - margin && cj + 1 < n_cj) ? 1 : 0);
        col_tab[y] = make_short2((short)cj, (short)dj);
    }
}
// In the minimise kernel the stamps go to a per-CTA bitmap in shared memory (one bit per cell) instead of the global flag array:
// a slice's events stamp each live cell ~50 times per iteration, and every stamp used to be a 4-byte global
// store on its own L2 sector (more L2 write transactions than the splat itself).  The CTA sets bits while it
// walks its events and writes the tag once per live cell afterwards (flush_stamp_bitmap).
// BF_STAMP_BYTES: the per-CTA stamp map holds one BYTE per cell and a stamp is a plain byte store (no read, no
// atomic: every writer stores the same value) instead of test-then-atomicOr on a bit -- 2 instructions per stamp
// instead of ~6, at 8x the (small) shared-memory footprint.
#ifndef BF_STAMP_BYTES
#define BF_STAMP_BYTES 0
#endif
#if BF_STAMP_BYTES
__device__ __forceinline__ void stamp_bit(unsigned *bm, int f) { reinterpret_cast<unsigned char *>(bm)[f] = 1; }
#else
__device__ __forceinline__ void stamp_bit(unsigned *bm, int f) {
    const unsigned b = 1u << (f & 31);
    if (!(bm[f >> 5] & b)) atomicOr(bm + (f >> 5), b);
}
#endif
__device__ __forceinline__ void mark_cells_bm(unsigned *bm, int x, int y, const int2 *row_tab, const short2 *col_tab) {
    const int2 r = row_tab[x];
    const short2 c = col_tab[y];
    const int f = r.x + (int)c.x;
    const int dj = (int)c.y;
    stamp_bit(bm, f);
    if (dj != 0) stamp_bit(bm, f + dj);
    if (r.y != 0) {
        stamp_bit(bm, f + r.y);
        if (dj != 0) stamp_bit(bm, f + r.y + dj);
    }
}
// CTA-wide: write `tag` to the flag of every cell whose bit is set and leave the bitmap all-zero again.
__device__ __forceinline__ void flush_stamp_bitmap(unsigned *bm, unsigned *flags, unsigned tag, int n_cells) {
    __syncthreads();
#if BF_STAMP_BYTES
    const int words = (n_cells + 3) >> 2;
    for (int w = threadIdx.x; w < words; w += blockDim.x) {
        const unsigned m = bm[w];
        if (m == 0u) continue;
        bm[w] = 0u;
        if (m & 0x000000ffu) flags[w * 4 + 0] = tag;
        if (m & 0x0000ff00u) flags[w * 4 + 1] = tag;
        if (m & 0x00ff0000u) flags[w * 4 + 2] = tag;
        if (m & 0xff000000u) flags[w * 4 + 3] = tag;
    }
#else
    const int words = (n_cells + 31) >> 5;
    for (int w = threadIdx.x; w < words; w += blockDim.x) {
        unsigned m = bm[w];
        if (m == 0u) continue;
        bm[w] = 0u;
        do {
            const int b = __ffs((int)m) - 1;
            m &= m - 1u;
            flags[w * 32 + b] = tag;
        } while (m != 0u);
    }
#endif
}

// ---- compact upload: a group expands its slice's 6-byte delta records into bf_event records -----------------------
// CTA `rank` of the G CTAs of the group takes blocks rank, rank + G, ... of the slice (1024 events each, two rounds of
// BF_NT = 512 events): t[i] = t0 - (dt[1] + ... + dt[i]) by a CTA-wide scan, the 8-byte record is written to the event
// buffer (read back L2-coherently by every later pass) and folded into the thread's bounding box / time range, so this
// pass REPLACES the bounding-box pass over the events.  Records and block table were written by the copy engine during
// this launch (streamed upload): ld.cg only.  Kept out of line: it runs once per slice and must not perturb the register
// allocation of the per-iteration passes.
__device__ __forceinline__ unsigned ld_cg_u16(const unsigned short *p) {
    unsigned short v;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p));
    return (unsigned)v;
}
__device__ __noinline__ void delta_expand_slice(const unsigned short *rec, const DeltaBlock *blocks, bf_event *out, int block0, int n,
                                                int rank, int G, int *scan /* [BF_NW] shared */, int *smm /* 6 ints, shared: min / max pairs */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = (n + BF_DELTA_BLOCK - 1) / BF_DELTA_BLOCK;
    int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN, tmin = INT_MAX, tmax = INT_MIN;
    for (int kb = rank; kb < nb; kb += G) {
        const DeltaBlock *bp = blocks + block0 + kb;
        const long long first = __ldcg(&bp->first);
        const int count = __ldcg(&bp->count), t0 = __ldcg(&bp->t0);
        const unsigned short *r = rec + (size_t)first * 3;
        unsigned carry = 0;
        for (int base = 0; base < BF_DELTA_BLOCK; base += BF_NT) {
            const int i = base + (int)threadIdx.x;
            unsigned w0 = 0, w1 = 0, w2 = 0;
            if (i < count) { w0 = ld_cg_u16(r + 3 * i); w1 = ld_cg_u16(r + 3 * i + 1); w2 = ld_cg_u16(r + 3 * i + 2); }
            const unsigned dt = (w1 >> 9) | (w2 << 7);
            unsigned incl = dt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            __syncthreads();
            if (lane == 31) scan[warp] = (int)incl;
            __syncthreads();
            unsigned before = carry, total = 0;
            for (int w = 0; w < BF_NW; ++w) {
                const unsigned v = (unsigned)scan[w];
                if (w < warp) before += v;
                total += v;
            }
            if (i < count) {
                const int fx = (int)(w0 & 0xfffu), fy = (int)((w0 >> 12) | ((w1 & 0xffu) << 4));
                const int t = t0 - (int)(before + incl);
                bf_event e;
                e.fr_x = (uint16_t)fx; e.fr_y = (uint16_t)(fy | ((w1 & 0x100u) ? BF_EVENT_NOISE : 0u)); e.t_ns = t;
                out[first + i] = e;
                xmin = min(xmin, fx); xmax = max(xmax, fx); ymin = min(ymin, fy); ymax = max(ymax, fy);
                tmin = min(tmin, t); tmax = max(tmax, t);
            }
            carry += total;
        }
    }
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o)); xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o)); ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        tmin = min(tmin, __shfl_xor_sync(0xffffffffu, tmin, o)); tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    }
    if (lane == 0 && rank < nb) {
        atomicMin(&smm[0], xmin); atomicMax(&smm[1], xmax);
        atomicMin(&smm[2], ymin); atomicMax(&smm[3], ymax);
        atomicMin(&smm[4], tmin); atomicMax(&smm[5], tmax);
    }
}

// ---- event pass: clear old pixel, re-project, splat --------------------------------------------
// One thread per event, CTA `rank` of the group owns a contiguous chunk (same chunk every
// iteration, so each thread re-reads the state it wrote itself).
//   first     : pr state does not exist yet (pr = fr, Event::reset, event.h:54-59)
//   project   : apply the warp `q` before splatting
//   img_new   : image receiving this iteration's splats (may be null: final pass)
//   out_nxy   : when non-null, nx/ny are written (final pass for writeout_events)
// Two events per thread per trip with all loads issued up front: the pass is latency-bound
// (one L2 round trip per event otherwise), not bandwidth-bound.
struct EventCtx {
    PixelMap pm;
    BfProj q;
    u64 one;                 // 1 << cnt_shift: the count field of a packed word
    int tq, t_min;
    int n_ci, n_cj;
    bool first, project;
    u64 *img_new;
    unsigned *flags;
    unsigned tag;
    const int2 *row_tab;
    const short2 *col_tab;
    unsigned *bm;            // per-CTA stamp bitmap in shared memory (one bit per cell)
};

// `st` carries the event's state in and, when the event is (re-)projected, its new state out.
template <int SH>
__device__ __forceinline__ void event_one(const EventCtx &c, uint2 e, float2 &st, double2 *pr_slot, double2 *nxy_slot) {
    const unsigned frx_u = e.x & 0xffffu;
    const unsigned fry_raw = e.x >> 16;
    const bool noise = (fry_raw & BF_EVENT_NOISE) != 0;
    const unsigned fry_u = fry_raw & 0x7fffu;
    const int t = (int)e.y;
    // float(fr) of event.h:165-167 is exact for 16-bit coordinates, so its f64 value is fr itself
    const double frx = u32_to_double(frx_u), fry = u32_to_double(fry_u);
    const float tf = (float)t;
    double prx, pry;
    if (c.first) { prx = frx; pry = fry; }                                         // Event::reset (event.h:54-59)
    else { prx = warp_from_m(st.x, frx); pry = warp_from_m(st.y, fry); }           // = the pr computed last time
    int x, y;
    double ex = 0.0, ey = 0.0;
    float mx = 0.0f, my = 0.0f;
    if (c.project) project_event(prx, pry, ex, ey, mx, my, frx, fry, tf, c.q);
    if (c.project || c.first) st = make_float2(mx, my);
    if (pr_slot != nullptr) *pr_slot = make_double2(prx, pry);
    if (nxy_slot != nullptr) *nxy_slot = make_double2(ex, ey);
    if (c.img_new != nullptr && !noise && event_pixel(prx, pry, c.pm, x, y)) {
        const u64 dt = (u64)((long long)t - (long long)c.t_min);
        red_add_u64(c.img_new + pixel_offset(x, y, c.pm.pitch), c.one + (dt >> c.tq));
        mark_cells_bm(c.bm, x, y, c.row_tab, c.col_tab);
    }
}

template <int SH>
__device__ BF_PASS_INLINE void event_pass(const KParams &P, const SliceDesc &sd, const BfGeom &g, const BfPack &pk,
                           const BfProj &q, int rank, int G, bool first, bool project, u64 *img_new,
                           double2 *out_nxy, unsigned *flags, unsigned tag, const int2 *row_tab, const short2 *col_tab,
                           unsigned *bm = nullptr) {
    typedef CellCfg<SH> C;
    const int per = (((sd.n + G - 1) / G) + 31) & ~31;
    const int lo = rank * per;
    const int cnt = min(sd.n, lo + per) - lo;
    if (cnt <= 0) return;
    const bool pro = out_nxy != nullptr;   // per-event outputs (pr, nx/ny) come as a pair
    EventCtx c;
    make_pixel_map(c.pm, g, P.pitch);
    c.q = q;
    c.one = 1ull << pk.cnt_shift; c.tq = pk.q; c.t_min = pk.t_min;
    c.n_ci = (g.rows + BF_CELL_ROWS - 1) / BF_CELL_ROWS;
    c.n_cj = (g.cols + C::CW - 1) / C::CW;
    c.first = first; c.project = project; c.img_new = img_new; c.flags = flags; c.tag = tag;
    c.row_tab = row_tab; c.col_tab = col_tab; c.bm = bm;
    // Events are taken in PAIRS aligned in the batch arrays (one 16-byte load brings two events, one
    // more their two states, one 16-byte store writes the states back); a pair that straddles the
    // chunk boundary has its foreign half predicated off.  Software pipeline: the loads of trip
    // n+1 are issued before trip n is processed.
    const long long gs = sd.ev_off + lo, ge = gs + cnt;          // this CTA's events, batch-global indices
    const long long p0 = gs >> 1, p1 = (ge + 1) >> 1;            // pairs [p0, p1)
    const uint4 *ev4 = reinterpret_cast<const uint4 *>(P.events);
    float4 *st4 = reinterpret_cast<float4 *>(P.state);
    long long p = p0 + (long long)threadIdx.x;
    uint4 e = make_uint4(0u, 0u, 0u, 0u);
    float4 st = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (p < p1) {
        e = ld_nc_u32x4(ev4 + p);
        if (!first) st = ld_state4(st4 + p);   // (when a group grows -- helping -- its members invalidate L1 first, see slice_loop)
    }
    for (; p < p1; p += BF_NT) {
        const long long np = p + BF_NT;
        uint4 ne = make_uint4(0u, 0u, 0u, 0u);
        float4 nst = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (np < p1) {
            ne = ld_nc_u32x4(ev4 + np);
            if (!first) nst = ld_state4(st4 + np);
        }
        const long long i0 = 2 * p, i1 = i0 + 1;
        const bool v0 = i0 >= gs, v1 = i1 < ge;                  // (i0 < ge and i1 >= gs hold by construction)
        float2 w0 = make_float2(st.x, st.y), w1 = make_float2(st.z, st.w);
        if (v0) event_one<SH>(c, make_uint2(e.x, e.y), w0, pro ? P.pr_out + i0 : nullptr, out_nxy ? out_nxy + i0 : nullptr);
        if (v1) event_one<SH>(c, make_uint2(e.z, e.w), w1, pro ? P.pr_out + i1 : nullptr, out_nxy ? out_nxy + i1 : nullptr);
        // the final pass (no splat) only reads the state; every other pass rewrites it
        if (img_new != nullptr) {
            if (v0 && v1) st_state4(st4 + p, make_float4(w0.x, w0.y, w1.x, w1.y));
            else if (v0) P.state[i0] = w0;
            else if (v1) P.state[i1] = w1;
        }
        e = ne; st = nst;
    }
    if (img_new != nullptr) flush_stamp_bitmap(bm, flags, tag, c.n_ci * c.n_cj);   // (cnt is CTA-uniform: every thread gets here)
}

// ---- fast unpack of a packed box sum ---------------------------------------------------------------
// bf_unpack_avg (bf_logic.h) is the specification: s = f32(sum_t / 1e9), mean = s / f32(cnt).  It is
// ~20 % of all instructions of the minimise kernel when written naively (64-bit shifts, an s64 -> f64
// conversion, an IEEE f32 divide with its MUFU.RCP + Newton step).  When BfPack::fast holds
// (q == 0, no offset, sums < 2^52) the same bits are obtained with:
//   * cnt from the high word only (cnt_shift >= 32 always);
//   * sum -> f64 by the 2^52 exponent trick (one LOP3 + one DADD, exact for sums < 2^52);
//   * the divide as Markstein's correction  q0 = s*y, q = fma(fma(-c, q0, s), y, q0)  with
//     y = RN(1/c) read from a per-CTA table: correctly rounded for every f32 s and integer
//     c < BF_RCP_TAB, because  (i) |q0 - s/c| <= 1.5 ulp, so the computed q0 + rem*y differs from
//     the exact quotient by < 2^-23 ulp, and (ii) s/c with c < 2^17 is never closer than 2^-18 ulp
//     to a rounding boundary without lying on it, which it cannot (c*(25-bit odd) has no 24-bit
//     representation).  tests/test_divconst.py checks the sequence against IEEE division.
#ifndef BF_RCP_TAB
#define BF_RCP_TAB 1024
#endif
__device__ __forceinline__ void fill_rcp_table(float2 *tab) {
    for (int c = threadIdx.x; c < BF_RCP_TAB; c += blockDim.x) {
        const float cf = (float)c;
        tab[c] = make_float2(cf, c ? __fdiv_rn(1.0f, cf) : 0.0f);
    }
}
__device__ __forceinline__ float unpack_avg_fast(const BfPack &pk, const float2 *tab, u64 v) {
    const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
    const unsigned cnt = hi >> (pk.cnt_shift - 32);
    const unsigned mhi = (unsigned)(pk.sum_mask >> 32);
    const double a = __dsub_rn(__hiloint2double((int)((hi & mhi) | 0x43300000u), (int)lo), 4503599627370496.0);
    const double q0 = __dmul_rn(a, 1e-9);
    const float s = (float)__fma_rn(__fma_rn(-1000000000.0, q0, a), 1e-9, q0);
    if (cnt < BF_RCP_TAB) {
        const float2 cy = tab[cnt];
        const float f0 = __fmul_rn(s, cy.y);
        return __fmaf_rn(__fmaf_rn(-cy.x, f0, s), cy.y, f0);
    }
    return __fdiv_rn(s, (float)cnt);
}

// ---- TMA tile staging of the cell patches (BF_TMA_PATCH) ----------------------------------------------------------
// The image pass reads, per live cell, a (8 + 2H) x 32 tile of packed words: with BF_TMA_PATCH the tile is fetched by
// the TMA unit (cp.async.bulk.tensor.3d -> UTMALDG) into a per-warp shared-memory buffer and completes on a per-warp
// mbarrier (SYNCS), one cell AHEAD of the arithmetic: while a warp works on cell k out of registers, the tile of its
// cell k + 1 is already in flight, so the L2 / HBM latency of the patch loads (a quarter of the image pass's stall
// samples, profiles/r1j_stall_regions.txt) is off the critical path, and the 12 per-lane LDG.64 with their 64-bit
// address arithmetic become 12 LDS.64 at immediate offsets.  All images of a context form ONE 3-D tensor
// [image][row][column] (one CUtensorMap per scale: the box height is 8 + 2H).
#ifndef BF_TMA_PATCH
#define BF_TMA_PATCH 0
#endif
struct TmaMaps {
    CUtensorMap m[3];          // scale 1, 3, 5
};
struct TmaWarp {               // per-warp staging state (registers)
    const CUtensorMap *map;
    u64 *buf;                  // this warp's tile buffer in shared memory (128-byte aligned)
    unsigned bar;              // shared-memory address of this warp's mbarrier
    unsigned phase;
    int img_index;             // index of the image in the 3-D tensor
};
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned phase) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "BF_MBAR_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra BF_MBAR_WAIT_%=;\n\t}" ::"r"(bar), "r"(phase) : "memory");
}
// lane 0 of a warp: fetch the patch of cell (ci, cj) of image `img_index` into the warp's buffer
template <int SH>
__device__ __forceinline__ void tma_issue_patch(const TmaWarp &t, int ci, int cj) {
    typedef CellCfg<SH> C;
    const int c0 = cj * C::CW - C::H + BF_BORDER, c1 = ci * BF_CELL_ROWS - C::H + BF_BORDER;
    mbar_expect_tx(t.bar, (unsigned)(C::PR * 32 * sizeof(u64)));
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(t.buf)), "l"(t.map), "r"(c0), "r"(c1), "r"(t.img_index), "r"(t.bar) : "memory");
}

// ---- image pass ------------------------------------------------------------------------------------
struct Acc {
    int cnt;
    long long si, sj;
    double sgx, sgy, sigx, sjgx, sigy, sjgy;
};

__device__ __forceinline__ void acc_zero(Acc &a) {
    a.cnt = 0; a.si = 0; a.sj = 0;
    a.sgx = a.sgy = a.sigx = a.sjgx = a.sigy = a.sjgy = 0.0;
}

// `p > 0.000001` with p an f32 promoted to f64 (object_model.cpp:20,114; accel_lib.h:534,599):
// 1e-6f is the largest f32 below the f64 literal 1e-6, so the f32 compare is equivalent.
#define BF_OCC(v) ((v) > 1e-6f)

// One cell, one warp.  Lane l holds column l of the point patch: PR rows of packed words.
//   vertical box sum   (registers)          v[r]  = sum_{|d|<=SH} P[r+d]
//   horizontal box sum (shuffles)           a[r]  = sum_{|d|<=SH} v[r] @ lane l+d
//   unpack -> mean time f32                 A[r], r = 0..AR-1   (row ci*8 - 1 + r of the image)
//   Scharr (accel_lib.h:594-605) + sums     lanes H..31-H, rows 1..8
// Integer packed sums commute, so this equals the reference's per-event s x s splat exactly.
template <int SH, bool MATERIALISE, bool FAST>
__device__ __forceinline__ void cell_compute(Acc &acc, const u64 (&P)[CellCfg<SH>::PR], const BfPack &pk, const float2 *rcp_tab,
                                             int ci, int cj, int rows, int cols, int i0, int j0, float *out_img,
                                             float *out_gx, float *out_gy);

template <int SH, bool MATERIALISE, bool FAST>
__device__ __forceinline__ void cell_process(Acc &acc, const u64 *img, int pitch, const BfPack &pk, const float2 *rcp_tab,
                                             int ci, int cj, int rows, int cols, int i0, int j0, float *out_img,
                                             float *out_gx, float *out_gy) {
    typedef CellCfg<SH> C;
    const int lane = threadIdx.x & 31;
    const u64 *p = img + (long long)(ci * BF_CELL_ROWS - C::H + BF_BORDER) * pitch +
                   (cj * C::CW - C::H + BF_BORDER + lane);
    u64 P[C::PR];
#pragma unroll
    for (int r = 0; r < C::PR; ++r) P[r] = __ldcg(p + (long long)r * pitch);
    cell_compute<SH, MATERIALISE, FAST>(acc, P, pk, rcp_tab, ci, cj, rows, cols, i0, j0, out_img, out_gx, out_gy);
}

// The arithmetic of one cell on the point patch held in registers (lane l = patch column l).
template <int SH, bool MATERIALISE, bool FAST>
__device__ __forceinline__ void cell_compute(Acc &acc, const u64 (&P)[CellCfg<SH>::PR], const BfPack &pk, const float2 *rcp_tab,
                                             int ci, int cj, int rows, int cols, int i0, int j0, float *out_img,
                                             float *out_gx, float *out_gy) {
    typedef CellCfg<SH> C;
    const int lane = threadIdx.x & 31;

    float A[C::AR];
#pragma unroll
    for (int r = 0; r < C::AR; ++r) {
        __syncwarp();   // reconverge after the previous row's data-dependent unpack branch (see group_sums)
        u64 v = 0;
#pragma unroll
        for (int d = 0; d <= 2 * SH; ++d) v += P[r + d];      // patch rows r .. r+2SH  <->  image rows centred on A row r
        u64 a = v;
#pragma unroll
        for (int d = 1; d <= SH; ++d) {
            a += __shfl_up_sync(0xffffffffu, v, d);
            a += __shfl_down_sync(0xffffffffu, v, d);
        }
        A[r] = (a != 0ull) ? (FAST ? unpack_avg_fast(pk, rcp_tab, a) : bf_unpack_avg(pk, a)) : 0.0f;
    }
    // lanes whose +-SH neighbours fall outside the warp hold garbage in A; they are never outputs
    // (outputs are lanes H..31-H) nor neighbours of outputs (lanes H-1..32-H need lanes 0..31 only).
    const bool out_lane = (lane >= C::H) && (lane < 32 - C::H);
    const int j = cj * C::CW + lane - C::H;
    __syncwarp();
    // per-cell partial sums: the column index j is fixed per lane, so its moments are folded in once
    // per cell instead of once per pixel
    int c_cnt = 0, c_si = 0;
    double c_gx = 0.0, c_gy = 0.0;
    float L0 = __shfl_up_sync(0xffffffffu, A[0], 1), R0 = __shfl_down_sync(0xffffffffu, A[0], 1);
    float L1 = __shfl_up_sync(0xffffffffu, A[1], 1), R1 = __shfl_down_sync(0xffffffffu, A[1], 1);
#pragma unroll
    for (int r = 1; r <= BF_CELL_ROWS; ++r) {
        __syncwarp();   // the occupancy / validity branches below diverge per lane
        const float L2 = __shfl_up_sync(0xffffffffu, A[r + 1], 1), R2 = __shfl_down_sync(0xffffffffu, A[r + 1], 1);
        const int i = ci * BF_CELL_ROWS + r - 1;
        const float v = A[r];
        float gx = 0.0f, gy = 0.0f;
        if (out_lane && BF_OCC(v)) {
            c_cnt += 1;
            c_si += i;
            // taps: v(k,l) = T[row-1+l][col-1+k]; column j-1 = (L0,L1,L2), j = (A[r-1],.,A[r+1]), j+1 = (R0,R1,R2)
            if (BF_OCC(L0) && BF_OCC(L1) && BF_OCC(L2) && BF_OCC(A[r - 1]) && BF_OCC(A[r + 1]) && BF_OCC(R0) &&
                BF_OCC(R1) && BF_OCC(R2)) {
                float a = __fmul_rn(L0, 3.0f);
                a = __fadd_rn(a, __fmul_rn(L2, -3.0f));
                a = __fadd_rn(a, __fmul_rn(A[r - 1], 10.0f));
                a = __fadd_rn(a, __fmul_rn(A[r + 1], -10.0f));
                a = __fadd_rn(a, __fmul_rn(R0, 3.0f));
                a = __fadd_rn(a, __fmul_rn(R2, -3.0f));
                float b = __fmul_rn(L0, 3.0f);
                b = __fadd_rn(b, __fmul_rn(L1, 10.0f));
                b = __fadd_rn(b, __fmul_rn(L2, 3.0f));
                b = __fadd_rn(b, __fmul_rn(R0, -3.0f));
                b = __fadd_rn(b, __fmul_rn(R1, -10.0f));
                b = __fadd_rn(b, __fmul_rn(R2, -3.0f));
                gx = a;
                gy = b;
                const double di = (double)(i - i0);
                const double dgx = (double)gx, dgy = (double)gy;
                c_gx += dgx;
                c_gy += dgy;
                acc.sigx += di * dgx;
                acc.sigy += di * dgy;
            }
        }
        if (MATERIALISE) {
            if (out_lane && i < rows && j < cols) {
                const size_t o = (size_t)i * cols + j;
                if (out_img) out_img[o] = v;
                if (out_gx) out_gx[o] = gx;
                if (out_gy) out_gy[o] = gy;
            }
        }
        L0 = L1; L1 = L2; R0 = R1; R1 = R2;
    }
    if (c_cnt != 0) {
        const double dj = (double)(j - j0);
        acc.cnt += c_cnt;
        acc.si += c_si;
        acc.sj += (long long)j * c_cnt;
        acc.sgx += c_gx;
        acc.sgy += c_gy;
        acc.sjgx += dj * c_gx;
        acc.sjgy += dj * c_gy;
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// CTA-wide reduction of the per-thread accumulators; BF_NSUMS doubles are written to `slot`.
// Fixed association order => bit-reproducible.
__device__ void acc_block_reduce(const Acc &a, double *sred /* [BF_NW][BF_NSUMS] */, double *slot) {
    double v[BF_NSUMS] = {(double)a.cnt, (double)a.si, (double)a.sj, a.sgx, a.sgy, a.sigx, a.sjgx, a.sigy, a.sjgy};
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < BF_NSUMS; ++k) {
        v[k] = warp_sum(v[k]);
        if (lane == 0) sred[warp * BF_NSUMS + k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < BF_NSUMS) {
        double s = 0.0;
        for (int w = 0; w < BF_NW; ++w) s += sred[w * BF_NSUMS + threadIdx.x];
        slot[threadIdx.x] = s;
    }
    __syncthreads();
}

// Compact the cells of [base, base + BF_LIST_CAP) whose flag carries `tag` into `list` (chunk-relative
// indices, ascending => deterministic).  Every CTA of the group builds the same list.  Returns the count.
__device__ __forceinline__ int compact_cells(const unsigned *flags, unsigned tag, int base, int n_cells,
                                             unsigned short *list, int *scan) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int PER = BF_LIST_CAP / BF_NT;   // flags per thread per chunk
    unsigned live = 0;
    int mine = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        const int c = base + (int)threadIdx.x * PER + k;
        if (c < n_cells && __ldcg(flags + c) == tag) { live |= 1u << k; ++mine; }
    }
    int incl = mine;
    __syncwarp();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int nb = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += nb;
    }
    __syncthreads();   // the previous list has been fully consumed
    if (lane == 31) scan[warp] = incl;
    __syncthreads();
    int off = incl - mine;
    for (int w = 0; w < warp; ++w) off += scan[w];
    int total = 0;
    for (int w = 0; w < BF_NW; ++w) total += scan[w];
#pragma unroll
    for (int k = 0; k < PER; ++k)
        if (live & (1u << k)) list[off++] = (unsigned short)((int)threadIdx.x * PER + k);
    __syncthreads();
    return total;
}

// Zero the 8 x CW interior of one cell with coalesced row stores (one warp).
template <int SH, bool LOCAL = false>
__device__ __forceinline__ void cell_clear(u64 *img, int pitch, int ci, int cj) {
    typedef CellCfg<SH, LOCAL> C;
    const int lane = threadIdx.x & 31;
    if (lane < C::CW) {
        u64 *p = img + (long long)(ci * BF_CELL_ROWS + BF_BORDER) * pitch + (cj * C::CW + BF_BORDER + lane);
#pragma unroll
        for (int r = 0; r < BF_CELL_ROWS; ++r) p[(long long)r * pitch] = 0ull;
    }
}

// Image pass of one CTA: the warps take live cells  rank*NW + warp, += G*NW  of the compacted list
// (no CTA barrier inside the cell loop).
//
// Clearing: the cells that were live in the OTHER image one iteration ago get their interior zeroed
// here (every splatted pixel lies in the interior of a cell it stamped), so that image is all-zero
// again before it is splatted into.  When the whole cell grid fits one list (n_cells <= BF_LIST_CAP,
// every sensor up to ~VGA) the previous iteration's compacted list is simply kept in shared memory
// (`list_prev`, `n_prev`) and no second scan is needed; larger grids re-scan the other flag array.
// Returns the number of live cells when the grid fits one list (else -1).
template <int SH> __device__ __forceinline__ void local_cell_process(Acc &acc, const u64 *img, int pitch, const BfPack &pk,
                                                                    int ci, int cj, int rows, int cols);

// MODE 0: mean-timestamp image + Scharr + moments (OptimizerRolling); MODE 1: saturated event-count
// image + Gaussian blur + non-zero mean (OptimizerLocal, see local_cell_process).
template <int SH, bool MATERIALISE, int MODE = 0>
__device__ BF_PASS_INLINE int image_pass(Acc &acc, const u64 *img, int pitch, const BfGeom &g, const BfPack &pk,
                          const float2 *rcp_tab, const unsigned *flags, unsigned tag, int rank, int G, unsigned short *list,
                          int *scan, float *out_img, float *out_gx, float *out_gy, u64 *img_clear,
                          const unsigned *flags_clear, unsigned tag_clear, const unsigned short *list_prev,
                          int n_prev, TmaWarp *tma = nullptr) {
    typedef CellCfg<SH, MODE == 1> C;
    const int n_ci = (g.rows + BF_CELL_ROWS - 1) / BF_CELL_ROWS, n_cj = (g.cols + C::CW - 1) / C::CW;
    const int n_cells = n_ci * n_cj;
    const int i0 = g.rows / 2, j0 = g.cols / 2;
    const int warp = threadIdx.x >> 5;
    const bool one_chunk = n_cells <= BF_LIST_CAP;
    int n_live = -1;
    // Live cells are dealt round-robin over the G * BF_NW warps of the group; on grids that need
    // several list chunks the deal continues where the previous chunk stopped (restarting at warp 0
    // for every chunk left the high ranks idle: 26 % of a 1280x720 iteration was barrier wait).
    const int n_warps = G * BF_NW, my_warp = rank * BF_NW + warp;
    int dealt = 0, dealt_clear = 0;
    for (int base = 0; base < n_cells; base += BF_LIST_CAP) {
        // (clearing first: it may borrow `list`, which must hold the CURRENT live cells when this returns)
        if (img_clear != nullptr && !(one_chunk && list_prev != nullptr)) {
            const int total = compact_cells(flags_clear, tag_clear, base, n_cells, list, scan);
            int k0 = (my_warp - dealt_clear) % n_warps;
            if (k0 < 0) k0 += n_warps;
            for (int k = k0; k < total; k += n_warps) {
                const int c = base + (int)list[k];
                cell_clear<SH, MODE == 1>(img_clear, pitch, c / n_cj, c % n_cj);
            }
            dealt_clear = (dealt_clear + total) % n_warps;
        }
        if (img != nullptr) {
            const int total = compact_cells(flags, tag, base, n_cells, list, scan);
            if (one_chunk) n_live = total;
            int k0 = (my_warp - dealt) % n_warps;
            if (k0 < 0) k0 += n_warps;
            dealt = (dealt + total) % n_warps;
            if constexpr (MODE == 1) {
                for (int k = k0; k < total; k += n_warps) {
                    const int c = base + (int)list[k];
                    const int ci = c / n_cj, cj = c - ci * n_cj;
                    local_cell_process<SH>(acc, img, pitch, pk, ci, cj, g.rows, g.cols);
                }
#if BF_TMA_PATCH
            } else if (pk.fast && tma != nullptr) {
                // TMA-staged patches, one cell ahead (see TmaWarp)
                const int lane = threadIdx.x & 31;
                if (k0 < total && lane == 0) {
                    // (the splats this tile is made of were written through the generic proxy by other CTAs and became
                    // visible at barrier A; order this thread's async-proxy reads behind that)
                    asm volatile("fence.proxy.async.global;" ::: "memory");
                    const int c = base + (int)list[k0];
                    tma_issue_patch<SH>(*tma, c / n_cj, c % n_cj);
                }
                for (int k = k0; k < total; k += n_warps) {
                    const int c = base + (int)list[k];
                    const int ci = c / n_cj, cj = c - ci * n_cj;
                    mbar_wait(tma->bar, tma->phase);
                    tma->phase ^= 1u;
                    u64 P[C::PR];
#pragma unroll
                    for (int r = 0; r < C::PR; ++r) P[r] = tma->buf[r * 32 + lane];
                    __syncwarp();   // every lane has its column in registers: the buffer may be refilled
                    if (k + n_warps < total && lane == 0) {
                        const int cn = base + (int)list[k + n_warps];
                        tma_issue_patch<SH>(*tma, cn / n_cj, cn % n_cj);
                    }
                    cell_compute<SH, MATERIALISE, true>(acc, P, pk, rcp_tab, ci, cj, g.rows, g.cols, i0, j0, out_img, out_gx, out_gy);
                }
#endif
            } else if (pk.fast) {
                for (int k = k0; k < total; k += n_warps) {
                    const int c = base + (int)list[k];
                    const int ci = c / n_cj, cj = c - ci * n_cj;
                    cell_process<SH, MATERIALISE, true>(acc, img, pitch, pk, rcp_tab, ci, cj, g.rows, g.cols, i0, j0, out_img, out_gx, out_gy);
                }
            } else {
                for (int k = k0; k < total; k += n_warps) {
                    const int c = base + (int)list[k];
                    const int ci = c / n_cj, cj = c - ci * n_cj;
                    cell_process<SH, MATERIALISE, false>(acc, img, pitch, pk, rcp_tab, ci, cj, g.rows, g.cols, i0, j0, out_img, out_gx, out_gy);
                }
            }
        }
    }
    if (img_clear != nullptr && one_chunk && list_prev != nullptr) {
        for (int k = rank * BF_NW + warp; k < n_prev; k += G * BF_NW) {
            const int c = (int)list_prev[k];
            cell_clear<SH, MODE == 1>(img_clear, pitch, c / n_cj, c % n_cj);
        }
    }
    return n_live;
}

// ====================================================================================================
// OptimizerLocal on the device (reference: src/optimizer_sampler.cpp): the
// contrast-driven coordinate descent over a global (nx, ny).  Same machinery as the rolling path --
// point splat with one 64-bit atomic, generation-tagged live cells, one warp per cell in registers --
// with a different per-cell function and a different (much simpler) per-event warp.
// ====================================================================================================

// OptimizerLocal::iteration_step's image (optimizer_sampler.cpp:124-149) for one cell:
//   count image  C = min(255, s x s box sum of the point counts)   (the reference increments a u8 with
//                    saturation per splatted pixel, :139-145: the result is min(255, #increments));
//   blur         B = cv::GaussianBlur(C, Size(s, s), 0, 0) on CV_8UC1 = binomial [1 2 1]/4 per axis in fixed
//                    point, i.e. (sum + 8) >> 4 for s = 3, BORDER_REFLECT_101 at the image edges; s = 1: B = C;
//   score sums   number and sum of the non-zero B (get_event_score, :192-204) -> acc.cnt, acc.si.
//   scale 5: the same with a 5 x 5 box and the binomial [1 4 6 4 1]/16 per axis, (sum + 128) >> 8, on the LOCAL cell
//   geometry (halo 4 = box radius 2 + blur radius 2: 16 patch rows, 24 output columns).
__device__ __forceinline__ void local_cell_process_s5(Acc &acc, const u64 *img, int pitch, const BfPack &pk, int ci, int cj,
                                                      int rows, int cols) {
    typedef CellCfg<2, true> C;     // H = 4, PR = 16, CW = 24
    constexpr int NC = BF_CELL_ROWS + 4;   // count rows: image rows ci*8 - 2 .. ci*8 + 9
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const u64 *p = img + (long long)(ci * BF_CELL_ROWS - C::H + BF_BORDER) * pitch + (cj * C::CW - C::H + BF_BORDER + lane);
    unsigned pc[C::PR];             // point counts of the patch column (the time sums are not needed here)
#pragma unroll
    for (int r = 0; r < C::PR; ++r) pc[r] = (unsigned)(__ldcg(p + (long long)r * pitch) >> 32) >> (pk.cnt_shift - 32);
    int Cn[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        const unsigned v = pc[k] + pc[k + 1] + pc[k + 2] + pc[k + 3] + pc[k + 4];      // patch rows k .. k+4 <-> image rows centred on count row k
        unsigned a = v;
        a += __shfl_up_sync(FULL, v, 1); a += __shfl_down_sync(FULL, v, 1);
        a += __shfl_up_sync(FULL, v, 2); a += __shfl_down_sync(FULL, v, 2);
        Cn[k] = (int)min(a, 255u);
    }
    // BORDER_REFLECT_101 rows: image row -1 := 1, -2 := 2; row `rows` := rows - 2, rows + 1 := rows - 3
    if (ci == 0) { Cn[1] = Cn[3]; Cn[0] = Cn[4]; }
#pragma unroll
    for (int k = 2; k < NC; ++k) {
        const int ir = ci * BF_CELL_ROWS - 2 + k;
        if (ir == rows) Cn[k] = Cn[k - 2];
        else if (ir == rows + 1 && k >= 4) Cn[k] = Cn[k - 4];
    }
    const int j = cj * C::CW + lane - C::H;
    const bool out_lane = (lane >= C::H) && (lane < 32 - C::H) && j < cols;
    int n_nz = 0, s_nz = 0;
#pragma unroll
    for (int r = 0; r < BF_CELL_ROWS; ++r) {
        const int i = ci * BF_CELL_ROWS + r;
        const int V = Cn[r] + 4 * Cn[r + 1] + 6 * Cn[r + 2] + 4 * Cn[r + 3] + Cn[r + 4];    // count rows i-2 .. i+2
        int L1 = __shfl_up_sync(FULL, V, 1), R1 = __shfl_down_sync(FULL, V, 1);
        int L2 = __shfl_up_sync(FULL, V, 2), R2 = __shfl_down_sync(FULL, V, 2);
        if (j == 0) { L1 = R1; L2 = R2; }              // columns -1, -2 := 1, 2
        else if (j == 1) L2 = V;                        // column -1 := 1
        if (j == cols - 1) { R1 = L1; R2 = L2; }        // columns cols, cols + 1 := cols - 2, cols - 3
        else if (j == cols - 2) R2 = V;                 // column cols := cols - 2
        const int B = (L2 + 4 * L1 + 6 * V + 4 * R1 + R2 + 128) >> 8;
        if (out_lane && i < rows && B > 0) { n_nz += 1; s_nz += B; }
    }
    acc.cnt += n_nz;
    acc.si += s_nz;
}

template <int SH>
__device__ __forceinline__ void local_cell_process(Acc &acc, const u64 *img, int pitch, const BfPack &pk, int ci, int cj,
                                                   int rows, int cols) {
    if constexpr (SH == 2) {
        local_cell_process_s5(acc, img, pitch, pk, ci, cj, rows, cols);
        return;
    }
    typedef CellCfg<SH> C;
    const int lane = threadIdx.x & 31;
    const u64 *p = img + (long long)(ci * BF_CELL_ROWS - C::H + BF_BORDER) * pitch + (cj * C::CW - C::H + BF_BORDER + lane);
    u64 P[C::PR];
#pragma unroll
    for (int r = 0; r < C::PR; ++r) P[r] = __ldcg(p + (long long)r * pitch);
    int Cn[C::AR];   // count image rows ci*8 - 1 .. ci*8 + 8
#pragma unroll
    for (int r = 0; r < C::AR; ++r) {
        u64 v = 0;
#pragma unroll
        for (int d = 0; d <= 2 * SH; ++d) v += P[r + d];
        u64 a = v;
#pragma unroll
        for (int d = 1; d <= SH; ++d) {
            a += __shfl_up_sync(0xffffffffu, v, d);
            a += __shfl_down_sync(0xffffffffu, v, d);
        }
        const unsigned cnt = (unsigned)(a >> 32) >> (pk.cnt_shift - 32);
        Cn[r] = (int)min(cnt, 255u);
    }
    const int j = cj * C::CW + lane - C::H;
    const bool out_lane = (lane >= C::H) && (lane < 32 - C::H) && j < cols;
    int n_nz = 0, s_nz = 0;
    if (SH == 0) {
#pragma unroll
        for (int r = 1; r <= BF_CELL_ROWS; ++r) {
            const int i = ci * BF_CELL_ROWS + r - 1;
            if (out_lane && i < rows && Cn[r] > 0) { n_nz += 1; s_nz += Cn[r]; }
        }
    } else {
        // BORDER_REFLECT_101 rows: image row -1 := row 1, image row `rows` := row rows - 2
        if (ci == 0) Cn[0] = Cn[2];
#pragma unroll
        for (int r = 2; r < C::AR; ++r)
            if (ci * BF_CELL_ROWS - 1 + r == rows) Cn[r] = Cn[r - 2];
#pragma unroll
        for (int r = 1; r <= BF_CELL_ROWS; ++r) {
            const int i = ci * BF_CELL_ROWS + r - 1;
            const int V = Cn[r - 1] + 2 * Cn[r] + Cn[r + 1];
            int L = __shfl_up_sync(0xffffffffu, V, 1), R = __shfl_down_sync(0xffffffffu, V, 1);
            if (j == 0) L = R;               // column -1 := column 1
            if (j == cols - 1) R = L;        // column `cols` := column cols - 2
            const int B = (L + 2 * V + R + 8) >> 4;
            if (out_lane && i < rows && B > 0) { n_nz += 1; s_nz += B; }
        }
    }
    acc.cnt += n_nz;
    acc.si += s_nz;
}

// State of OptimizerLocal::run (optimizer_sampler.cpp:4-38, 90-117).
struct LocalOpt {
    double nx, ny;            // accepted position
    double dnx, dny, dn_th;
    double last_score;
    double cur_nx, cur_ny;    // position of the iteration_step being evaluated
    int phase;                // 0: initial step, 1: compute_new_nx's step, 2: compute_new_ny's step
    int steps, rc;
    unsigned nz_cnt;          // non-zero pixels of the last image
};

__device__ __forceinline__ void local_opt_init(LocalOpt &o, int scale) {
    o.nx = 0; o.ny = 0; o.last_score = 0;                                          // :5-6
    o.dnx = 0.01; o.dny = 0.01;                                                    // :7
    o.dn_th = (127 * 1 * 1000.0) / (double)(10ull * (unsigned long long)scale * 100000000ull);   // NZ*T_DIVIDER*1000 / (10*scale*FROM_MS(MAX_TIME_MS))
    o.cur_nx = 0; o.cur_ny = 0; o.phase = 0; o.steps = 0; o.rc = BF_RC_OK; o.nz_cnt = 0;
}

// Consumes the score of the step at (cur_nx, cur_ny) and sets up the next one; false when run() is over.
__device__ __forceinline__ bool local_opt_advance(LocalOpt &o, double cnt, double sum, int iter_cap) {
    const double score = cnt > 0 ? sum / cnt : 0.0;                                // :192-204
    o.nz_cnt = (unsigned)cnt;
    o.steps += 1;
    if (o.phase == 0) {
        o.last_score = score;                                                      // :16
    } else {
        const double dscore = score - o.last_score;                                // :94-95 / :109-110
        o.last_score = score;
        if (o.phase == 1) { if (dscore <= 0) o.dnx = -o.dnx / 2.0; o.nx = o.cur_nx; }   // :97-101, run() assigns nx
        else { if (dscore <= 0) o.dny = -o.dny / 2.0; o.ny = o.cur_ny; }
    }
    if (o.phase != 1) {
        if (!(hypot(o.dnx, o.dny) > o.dn_th)) return false;                        // while condition, :20
        if (o.steps >= iter_cap) { o.rc = BF_RC_ITER_CAP; return false; }
        o.phase = 1; o.cur_nx = o.nx + o.dnx; o.cur_ny = o.ny;                     // compute_new_nx
    } else {
        o.phase = 2; o.cur_nx = o.nx; o.cur_ny = o.ny + o.dny;                     // compute_new_ny
    }
    return true;
}

// Event pass of one iteration_step (optimizer_sampler.cpp:121, 126-145): Event::project(nx, ny)
// (event.h:65-70 -> 164-168), pixel with the UNtruncated double shifts, window test on the truncated
// ints before the + scale/2 (:133-137), point splat.  No per-event state: the warp is absolute.
template <int SH>
__device__ void local_event_pass(const KParams &P, const SliceDesc &sd, const BfGeom &g, const BfPack &pk, double nx,
                                 double ny, int rank, int G, u64 *img_new, unsigned *flags, unsigned tag, const int2 *row_tab,
                                 const short2 *col_tab, unsigned *bm = nullptr) {
    const int per = (((sd.n + G - 1) / G) + 31) & ~31;
    const int lo = rank * per;
    const int cnt = min(sd.n, lo + per) - lo;
    if (cnt <= 0) return;
    const float kx = (float)((double)(float)nx / 127.0), ky = (float)((double)(float)ny / 127.0);   // event.h:164-165
    // event_c = Event((x_max - x_min) / 2 + x_min, ..., t = 0): project() leaves it at float(fr)  (optimizer_sampler.h:51, .cpp:122)
    const double sc = (double)g.scale;
    const double cx = (double)(float)(unsigned)((g.x_max - g.x_min) / 2 + g.x_min);
    const double cy = (double)(float)(unsigned)((g.y_max - g.y_min) / 2 + g.y_min);
    const double x_shift = __dadd_rn(__dmul_rn(-cx, sc), (double)g.w / 2.0);      // :126-127
    const double y_shift = __dadd_rn(__dmul_rn(-cy, sc), (double)g.h / 2.0);
    const u64 one = 1ull << pk.cnt_shift;
    const long long gs = sd.ev_off + lo, ge = gs + cnt;
    const bf_event *ev = P.events;
    for (long long i = gs + (long long)threadIdx.x; i < ge; i += BF_NT) {
        const uint2 e = ld_nc_u32x2(ev + i);
        const unsigned frx_u = e.x & 0xffffu, fry_u = (e.x >> 16) & 0x7fffu;
        const int t = (int)e.y;
        const float tf = (float)t;
        const double prx = warp_from_m(__fmul_rn(kx, tf), u32_to_double(frx_u));   // event.h:167-168
        const double pry = warp_from_m(__fmul_rn(ky, tf), u32_to_double(fry_u));
        const int x = __double2int_rz(__dadd_rn(__dmul_rn(prx, sc), x_shift));    // :130-131
        const int y = __double2int_rz(__dadd_rn(__dmul_rn(pry, sc), y_shift));
        if ((unsigned)x < (unsigned)g.w && (unsigned)y < (unsigned)g.h) {         // :133-134
            const int xc = x + g.half, yc = y + g.half;                           // :136-137
            const u64 dt = (u64)((long long)t - (long long)pk.t_min);
            red_add_u64(img_new + pixel_offset(xc, yc, P.pitch), one + (dt >> pk.q));
            mark_cells_bm(bm, xc, yc, row_tab, col_tab);
        }
    }
    {
        typedef CellCfg<SH, true> C;
        flush_stamp_bitmap(bm, flags, tag, ((g.rows + BF_CELL_ROWS - 1) / BF_CELL_ROWS) * ((g.cols + C::CW - 1) / C::CW));
    }
}

// Sum the G per-CTA partial records of a group in a fixed order (lane-strided, then butterfly):
// every CTA of the group obtains the bit-identical result.  Call from warp 0; result valid in all lanes.
__device__ __forceinline__ void group_sums(BfSums &s, const double *partials, int G, long long *dbg = nullptr) {
    const int lane = threadIdx.x & 31;
    double v[BF_NSUMS];
#pragma unroll
    for (int k = 0; k < BF_NSUMS; ++k) v[k] = 0.0;
    const long long t0 = dbg ? clock64() : 0;
    for (int r = lane; r < G; r += 32) {
#pragma unroll
        for (int k = 0; k < BF_NSUMS; ++k) v[k] += __ldcg(partials + r * BF_NSUMS + k);
    }
    if (dbg) {
        double z = 0;
#pragma unroll
        for (int k = 0; k < BF_NSUMS; ++k) z += v[k];
        if (z == -1.2345) v[0] = 0;
        if (lane == 0) dbg[0] += clock64() - t0;
    }
    // The lane-strided loop diverges.  Without an explicit reconvergence point the shuffles below run
    // through the compiler's divergent-warp fallback (BRA.DIV handlers): measured 23.8k vs 1.9k cycles.
    __syncwarp();
#pragma unroll
    for (int k = 0; k < BF_NSUMS; ++k) v[k] = warp_sum(v[k]);
    s.cnt = v[0]; s.si = v[1]; s.sj = v[2]; s.sgx = v[3]; s.sgy = v[4];
    s.sigx = v[5]; s.sjgx = v[6]; s.sigy = v[7]; s.sjgy = v[8];
}

// CLUSTER flavour: record r is read from CTA r's shared memory (distributed shared memory).  Warp 0, all lanes.
__device__ __forceinline__ void group_sums_dsmem(BfSums &s, const double *own_record /* shared, BF_NSUMS doubles */, int G) {
    const int lane = threadIdx.x & 31;
    double v[BF_NSUMS];
#pragma unroll
    for (int k = 0; k < BF_NSUMS; ++k) v[k] = lane < G ? ld_dsmem_f64(own_record + k, (unsigned)lane) : 0.0;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < BF_NSUMS; ++k) v[k] = warp_sum(v[k]);
    s.cnt = v[0]; s.si = v[1]; s.sj = v[2]; s.sgx = v[3]; s.sgy = v[4];
    s.sigx = v[5]; s.sjgx = v[6]; s.sigy = v[7]; s.sjgy = v[8];
}

// The same sum for LARGE groups (a single slice on the whole GPU: G = 296), in two steps so that the
// G records are fetched with one L2 round trip instead of G / 32 dependent ones (12 k of the 54 k
// cycles of a DAVIS-240C single-slice iteration): every thread t < G takes record t, warps reduce,
// warp 0 finishes from shared memory.  Fixed order => bit-identical in every CTA.  G <= BF_NT.
__device__ __forceinline__ void group_sums_block_gather(const double *partials, int G, double *sred /* [BF_NW][BF_NSUMS] */) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double v[BF_NSUMS];
#pragma unroll
    for (int k = 0; k < BF_NSUMS; ++k) v[k] = ((int)threadIdx.x < G) ? __ldcg(partials + threadIdx.x * BF_NSUMS + k) : 0.0;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < BF_NSUMS; ++k) {
        v[k] = warp_sum(v[k]);
        if (lane == 0) sred[warp * BF_NSUMS + k] = v[k];
    }
    __syncthreads();
}
__device__ __forceinline__ void group_sums_block_finish(BfSums &s, const double *sred) {   // warp 0, after the gather
    const int lane = threadIdx.x & 31;
    double v[BF_NSUMS];
#pragma unroll
    for (int k = 0; k < BF_NSUMS; ++k) v[k] = warp_sum(lane < BF_NW ? sred[lane * BF_NSUMS + k] : 0.0);
    s.cnt = v[0]; s.si = v[1]; s.sj = v[2]; s.sgx = v[3]; s.sgy = v[4];
    s.sigx = v[5]; s.sjgx = v[6]; s.sigy = v[7]; s.sjgy = v[8];
}
#define BF_BLOCK_GATHER_MIN 48   // groups at least this large gather block-wide

// sin/cos of the accumulated rotation.  |crl| is ~1e-4 rad in practice: a degree-11/10 Taylor
// polynomial is accurate to < 1 ulp for |x| <= 2^-5 and costs ~20 dependent fp64 operations instead
// of the ~200 of the generic sincos; larger angles take the library path.
__device__ __forceinline__ void sincos_small(double x, double &s, double &c) {
    if (fabs(x) <= 0.03125) {
        const double z = x * x;
        double ps = 1.0 / 39916800.0;
        ps = ps * z - 1.0 / 362880.0;   // -fmad=false keeps these as separate mul/add; accuracy is ample
        ps = ps * z + 1.0 / 5040.0;
        ps = ps * z - 1.0 / 120.0;
        ps = ps * z + 1.0 / 6.0;
        s = x - x * z * ps;
        double pc = 1.0 / 3628800.0;
        pc = pc * z - 1.0 / 40320.0;
        pc = pc * z + 1.0 / 720.0;
        pc = pc * z - 1.0 / 24.0;
        pc = pc * z + 0.5;
        c = 1.0 - z * pc;
    } else {
        sincos(x, &s, &c);
    }
}

// Warp-cooperative form of bf_opt_advance (bf_logic.h): identical arithmetic, but the independent
// IEEE fp64 divides run on different lanes (lane & 3 selects dx / dy / rot / div, lane & 1 the
// centre coordinate), which cuts the dependent chain from ~16 divides to 3.  Call from warp 0 with
// `s` identical in all lanes; o and next live in shared memory.  Returns the continue flag (all lanes).
__device__ __forceinline__ bool opt_advance_warp(BfOpt &o, const BfGeom &g, const BfSums &s, int i0, int j0,
                                                 int max_iter, int iter_cap, BfProj &next) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int p = lane & 3;
    const double cnt = s.cnt;
    // centre of mass (object_model.cpp:122-125): even lanes cx, odd lanes cy
    const double c_img = ((lane & 1) ? s.sj : s.si) / cnt;
    __syncwarp();
    const double cx_img = __shfl_sync(FULL, c_img, 0), cy_img = __shfl_sync(FULL, c_img, 1);
    const double ox = cx_img - (double)i0, oy = cy_img - (double)j0;
    // mean gradient / rotation / divergence (object_model.cpp:35-38), see bf_sums_to_model
    double num;
    if (p == 0) num = s.sgx;
    else if (p == 1) num = s.sgy;
    else if (p == 2) { num = s.sigy - s.sjgx; num = num - ox * s.sgy; num = num + oy * s.sgx; }
    else { num = s.sigx + s.sjgy; num = num - ox * s.sgx; num = num - oy * s.sgy; }
    const double val = num / cnt;
    // update_accumulators (object_model.h:48-53)
    float dvd = p == 0 ? o.x_div : p == 1 ? o.y_div : p == 2 ? o.rot_div : o.div_div;
    double tot = p == 0 ? o.m.total_dx : p == 1 ? o.m.total_dy : p == 2 ? o.m.total_rot : o.m.total_div;
    const float old = p == 0 ? o.old_dx : p == 1 ? o.old_dy : p == 2 ? o.old_rot : o.old_div;
    const double qa = val / (double)dvd;
    tot += qa;
    // centre back to sensor units (optimizer_rolling.h:330-331)
    const double cen = (c_img - ((lane & 1) ? g.y_shift : g.x_shift)) / (double)g.scale;
    __syncwarp();   // the per-lane selects above may have been compiled as branches
    const double cx = __shfl_sync(FULL, cen, 0), cy = __shfl_sync(FULL, cen, 1);
    const double tdx = __shfl_sync(FULL, tot, 0), tdy = __shfl_sync(FULL, tot, 1);
    const double trot = __shfl_sync(FULL, tot, 2), tdiv = __shfl_sync(FULL, tot, 3);
    const int iters = o.iters + 1;

    bool stop = false;
    int rc = o.rc;
    bool flip = false;
    if (!(cnt > 0)) { rc = BF_RC_DEGENERATE; stop = true; }
    if (!stop && iters > 1) {
        if (max_iter > 0 && iters > max_iter) stop = true;           // :94-96
        else { flip = (val * (double)old < 0); if (flip) dvd *= 2; } // :98-101
    }
    const float lim = p < 2 ? 320.0f : 32000.0f;                     // :76-79
    const double th = p < 2 ? 1e-5 : (p == 2 ? 1e-4 : 1e-1);         // :81-84
    // val / (2 d) == (val / d) / 2 exactly (power-of-two scaling), so the doubled divider needs no new divide
    const double cq = flip ? qa * 0.5 : qa;
    __syncwarp();
    const bool any_below = (__ballot_sync(FULL, dvd < lim) & 0xfu) != 0u;
    const bool all_conv = (__ballot_sync(FULL, fabs(cq) < th) & 0xfu) == 0xfu;
    if (!stop) {
        if (!any_below) stop = true;
        else if (all_conv) stop = true;
        else if (iters >= iter_cap) { rc = BF_RC_ITER_CAP; stop = true; }
    }
    double sn, cs;
    sincos_small(-trot, sn, cs);
    if (lane < 4) {
        if (p == 0) { o.m.dx = val; o.m.total_dx = tot; o.x_div = dvd; if (!stop) o.old_dx = (float)val; }
        else if (p == 1) { o.m.dy = val; o.m.total_dy = tot; o.y_div = dvd; if (!stop) o.old_dy = (float)val; }
        else if (p == 2) { o.m.rot = val; o.m.total_rot = tot; o.rot_div = dvd; if (!stop) o.old_rot = (float)val; }
        else { o.m.div = val; o.m.total_div = tot; o.div_div = dvd; if (!stop) o.old_div = (float)val; }
    }
    if (lane == 0) {
        o.m.cx = cx; o.m.cy = cy; o.m.cnt = (uint32_t)cnt;
        o.iters = iters; o.rc = rc;
        next.dnx = -tdx; next.dny = -tdy; next.cx = cx; next.cy = cy; next.div = tdiv; next.c = cs; next.s = sn;
    }
    return !stop;
}
