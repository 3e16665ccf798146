// bf_device.cuh -- sm_100a device code of the motion-compensation hot path.
//
// Data layout in HBM / L2:
//   events   bf_event[n]        8 B/event, read-only, ld.global.nc
//   pr       double2[n]         warped position state carried between iterations (event.h:100)
//   image    u64[rows_alloc][pitch]  the POINT image: every event adds ONE packed word
//            (count | sum of t) at its centre pixel with one 64-bit integer atomic.  The
//            reference's s x s splat (accel_lib.h:160-165) is recovered exactly in the image pass
//            as an s x s box sum of the point image (integer sums are associative).  A zero border
//            of BF_BORDER pixels surrounds the image so tile loads never need bounds checks.
//   Two such images per CTA group: iteration k splats into image k&1 while the events clear
//   their iteration k-1 pixels in the other one, so the image pass is read-only.
//
// Reference paths are relative to /root/reference/better_flow_core/.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "bf_logic.h"

// ---- compile-time tiling ------------------------------------------------------------------------
#define BF_NT 512              // threads per CTA (16 warps)
#define BF_TR 30               // output rows per tile
#define BF_TC 126              // output cols per tile
#define BF_AR (BF_TR + 2)      // mean-time tile rows (1-pixel halo for the 3x3 Scharr)
#define BF_AC (BF_TC + 2)      // = 128: one column per thread, 4 row strips of 8
#define BF_STRIP 8
#define BF_BORDER 4            // zero border of the stored image (>= scale/2 + 1 + alignment slack)

typedef unsigned long long u64;

template <int SH> struct TileCfg {
    static constexpr int H = SH + 1;                        // halo of the point tile: box radius + Scharr radius
    static constexpr int OFF = (BF_BORDER - H) & 1;         // extra left column so rows start 16-B aligned
    static constexpr int PR = BF_TR + 2 * H;                // point-tile rows
    static constexpr int PW = BF_TC + 2 * H + 2 * OFF;      // point-tile cols (even)
    static constexpr int CHUNKS = PR * (PW / 2);            // 16-byte chunks per tile
};
#define BF_PTILE_MAX_ELEMS (36 * 134)                       // TileCfg<2>: PR=36, PW=134

struct SliceDesc {
    long long ev_off;   // first event of the slice in the batch arrays
    int n;
    int scale;
    int max_iter;
    int has_init;
    bf_model init;
};

// Per-group control block in global memory (one 256-byte record per group).
struct GroupWs {
    unsigned bar;            // monotonically increasing arrival counter
    unsigned pad0[31];
    int cur_slice;
    int pad1[3];
    int bbox[2][8];          // [parity]: x_min, x_max, y_min, y_max, t_min, t_max
    int pad2[12];
};

struct KParams {
    const bf_event *events;
    double2 *pr;             // state
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
    int G;                   // CTAs per group
    int res_x, res_y;        // sensor rows / cols
    int min_events;          // 1000 (optimizer_rolling.h:57)
    int iter_cap;
    int want_events;
};

// ---- small PTX helpers ------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint2 ld_nc_u32x2(const void *p) {
    uint2 v;
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

// Barrier over the G CTAs of one group (all co-resident: cooperative launch).  `target` is the
// CTA-local running arrival total; the counter is zeroed by the host before every launch.
__device__ __forceinline__ void group_barrier(unsigned *counter, unsigned &target, int G) {
    __syncthreads();
    target += (unsigned)G;
    if (threadIdx.x == 0) {
        red_release_add_u32(counter, 1u);
        while (ld_acquire_u32(counter) < target) {
        }
    }
    __syncthreads();
}

// ---- exact division by the two constants of Event::apply_project (event.h:164-168) ----------
// q = fma(fma(-b, a*r, a), r, a*r) with r = rn(1/b) equals the correctly rounded a / b for EVERY
// finite f32-valued a, for b = 127 and b = 10000 (checked exhaustively over all 2^32 floats,
// tests/test_divconst.py).  3 FP64 instructions instead of a ~30-instruction IEEE divide.
__device__ __forceinline__ double div_const(double a, double b, double r) {
    const double q0 = __dmul_rn(a, r);
    const double rem = __fma_rn(-b, q0, a);
    return __fma_rn(rem, r, q0);
}

// Event::project_4param_reinit + apply_project (event.h:99-110,164-168) for one event.
// All FP64 operations individually rounded (no contraction), in the reference's order.
__device__ __forceinline__ void project_event(double &prx, double &pry, double &ex, double &ey,
                                              float frx, float fry, float tf, const BfProj &q) {
    const double rx = __dsub_rn(prx, q.cx), ry = __dsub_rn(pry, q.cy);                           // :100
    const double qx = __dsub_rn(__dmul_rn(q.c, rx), __dmul_rn(q.s, ry));                        // :102
    const double qy = __dadd_rn(__dmul_rn(q.s, rx), __dmul_rn(q.c, ry));                        // :103
    const double dx = __dadd_rn(__dmul_rn(-qx, q.div), __dsub_rn(qx, rx));                      // :105
    const double dy = __dadd_rn(__dmul_rn(-qy, q.div), __dsub_rn(qy, ry));
    ex = __dadd_rn(dx, q.dnx);                                                                  // :107
    ey = __dadd_rn(dy, q.dny);                                                                  // :108
    const float kx = (float)div_const((double)(float)ex, 127.0, 1.0 / 127.0);                   // :164
    const float ky = (float)div_const((double)(float)ey, 127.0, 1.0 / 127.0);
    prx = __dsub_rn((double)frx, div_const((double)__fmul_rn(kx, tf), 10000.0, 1.0 / 10000.0)); // :167
    pry = __dsub_rn((double)fry, div_const((double)__fmul_rn(ky, tf), 10000.0, 1.0 / 10000.0));
}

// Pixel of an event in the time image, AccelLib::get_time_img_cpu (accel_lib.h:154-158).
// Returns the element offset into the stored (bordered) image, or -1 when the splat is rejected.
__device__ __forceinline__ long long event_pixel(double prx, double pry, const BfGeom &g, int pitch) {
    const double fx = __dadd_rn(__dmul_rn(prx, (double)g.scale), (double)g.x_sh);
    const double fy = __dadd_rn(__dmul_rn(pry, (double)g.scale), (double)g.y_sh);
    if (!(fx == fx) || !(fy == fy)) return -1;   // x86 cvttsd2si(NaN) = INT_MIN -> rejected
    const int x = (int)fx, y = (int)fy;          // truncation toward zero; out-of-range saturates -> rejected
    if ((x >= g.w + g.half) || (x < g.half) || (y >= g.h + g.half) || (y < g.half)) return -1;
    return (long long)(x + BF_BORDER) * pitch + (y + BF_BORDER);
}

// ---- event pass: clear old pixel, re-project, splat --------------------------------------------
// One thread per event, CTA `rank` of the group owns a contiguous chunk (same chunk every
// iteration, so each thread re-reads the state it wrote itself).
//   first     : pr state does not exist yet (pr = fr, Event::reset, event.h:54-59)
//   project   : apply the warp `q` before splatting
//   img_new   : image receiving this iteration's splats (may be null: final pass)
//   img_old   : image holding the previous iteration's splats, cleared here (may be null)
//   out_nxy   : when non-null, nx/ny are written (final pass for writeout_events)
__device__ void event_pass(const KParams &P, const SliceDesc &sd, const BfGeom &g, const BfPack &pk,
                           const BfProj &q, int rank, bool first, bool project, u64 *img_new,
                           u64 *img_old, double2 *out_nxy) {
    const int per = (((sd.n + P.G - 1) / P.G) + 31) & ~31;
    const int lo = rank * per;
    const int hi = min(sd.n, lo + per);
    const bf_event *ev = P.events + sd.ev_off;
    double2 *pr = P.pr + sd.ev_off;
    for (int i = lo + (int)threadIdx.x; i < hi; i += BF_NT) {
        const uint2 e = ld_nc_u32x2(ev + i);
        const unsigned frx_u = e.x & 0xffffu;
        const unsigned fry_raw = e.x >> 16;
        const bool noise = (fry_raw & BF_EVENT_NOISE) != 0;
        const unsigned fry_u = fry_raw & 0x7fffu;
        const int t = (int)e.y;
        double prx, pry;
        if (first) { prx = (double)frx_u; pry = (double)fry_u; }
        else { const double2 s = pr[i]; prx = s.x; pry = s.y; }
        if (img_old != nullptr && !noise) {
            const long long o = event_pixel(prx, pry, g, P.pitch);
            if (o >= 0) img_old[o] = 0ull;
        }
        double ex = 0.0, ey = 0.0;
        if (project) project_event(prx, pry, ex, ey, (float)frx_u, (float)fry_u, (float)t, q);
        if (project || first) pr[i] = make_double2(prx, pry);
        if (out_nxy != nullptr) out_nxy[sd.ev_off + i] = make_double2(ex, ey);
        if (img_new != nullptr && !noise) {
            const long long o = event_pixel(prx, pry, g, P.pitch);
            if (o >= 0) atomicAdd(img_new + o, bf_pack_value(pk, t));
        }
    }
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

// Issue the cp.async loads of one point tile (with halo) into shared memory.
template <int SH>
__device__ __forceinline__ void tile_load(u64 *sP, const u64 *img, int pitch, int tile_r, int tile_c) {
    typedef TileCfg<SH> C;
    const long long base = (long long)(tile_r * BF_TR + BF_BORDER - C::H) * pitch +
                           (tile_c * BF_TC + BF_BORDER - C::H - C::OFF);
    constexpr int CPR = C::PW / 2;   // 16-byte chunks per row
    for (int k = threadIdx.x; k < C::CHUNKS; k += BF_NT) {
        const int r = k / CPR, c = k - r * CPR;
        cp_async16(sP + r * C::PW + 2 * c, img + base + (long long)r * pitch + 2 * c);
    }
}

// Mean-timestamp tile from the point tile: s x s box sum of packed words (separable, rolling over
// rows), then unpack -> (sum_t, count) -> f32 mean.  Thread = (row strip, column).
template <int SH>
__device__ __forceinline__ void tile_mean(float *sA, const u64 *sP, const BfPack &pk) {
    typedef TileCfg<SH> C;
    constexpr int K = 2 * SH + 1;
    const int strip = threadIdx.x >> 7;      // 0..3
    const int ac = threadIdx.x & 127;        // 0..127
    u64 win[K];
#pragma unroll
    for (int k = 0; k < BF_STRIP + 2 * SH; ++k) {
        const int pr = strip * BF_STRIP + k;
        u64 hs = 0;
#pragma unroll
        for (int d = 0; d < K; ++d) hs += sP[pr * C::PW + ac + C::OFF + d];
        win[k % K] = hs;
        if (k >= 2 * SH) {
            u64 v = 0;
#pragma unroll
            for (int d = 0; d < K; ++d) v += win[d];
            sA[(strip * BF_STRIP + k - 2 * SH) * BF_AC + ac] = (v != 0ull) ? bf_unpack_avg(pk, v) : 0.0f;
        }
    }
}

// `p > 0.000001` with p an f32 promoted to f64 (object_model.cpp:20,114; accel_lib.h:534,599):
// 1e-6f is the largest f32 below the f64 literal 1e-6, so the f32 compare is equivalent.
#define BF_OCC(v) ((v) > 1e-6f)

// AccelLib::sobel_point's live part (accel_lib.h:594-605): column-major tap order, f32 multiply
// then f32 add, each rounded.  a(dr, dc) reads the mean-time tile around the centre.
__device__ __forceinline__ bool scharr_at(const float *c, float &gx, float &gy) {
    const float v00 = c[-BF_AC - 1], v01 = c[-1], v02 = c[BF_AC - 1];      // column j-1: rows i-1, i, i+1
    const float v10 = c[-BF_AC], v12 = c[BF_AC];                            // column j
    const float v20 = c[-BF_AC + 1], v21 = c[1], v22 = c[BF_AC + 1];      // column j+1
    if (!(BF_OCC(v00) && BF_OCC(v01) && BF_OCC(v02) && BF_OCC(v10) && BF_OCC(v12) && BF_OCC(v20) &&
          BF_OCC(v21) && BF_OCC(v22)))
        return false;
    float a = __fmul_rn(v00, 3.0f);
    a = __fadd_rn(a, __fmul_rn(v02, -3.0f));
    a = __fadd_rn(a, __fmul_rn(v10, 10.0f));
    a = __fadd_rn(a, __fmul_rn(v12, -10.0f));
    a = __fadd_rn(a, __fmul_rn(v20, 3.0f));
    a = __fadd_rn(a, __fmul_rn(v22, -3.0f));
    float b = __fmul_rn(v00, 3.0f);
    b = __fadd_rn(b, __fmul_rn(v01, 10.0f));
    b = __fadd_rn(b, __fmul_rn(v02, 3.0f));
    b = __fadd_rn(b, __fmul_rn(v20, -3.0f));
    b = __fadd_rn(b, __fmul_rn(v21, -10.0f));
    b = __fadd_rn(b, __fmul_rn(v22, -3.0f));
    gx = a;
    gy = b;
    return true;
}

// Scharr + reduction over the TR x TC output pixels of a tile (ObjectModel::center_of_mass and
// ObjectModel::compute fused, object_model.cpp:103-126,4-39).  When out_* are non-null the mean
// image / gradient images are also materialised (stage-level API and debug images only).
template <bool MATERIALISE>
__device__ __forceinline__ void tile_reduce(Acc &acc, const float *sA, int tile_r, int tile_c, int rows,
                                            int cols, int i0, int j0, float *out_img, float *out_gx,
                                            float *out_gy) {
    const int strip = threadIdx.x >> 7;
    const int oc = threadIdx.x & 127;
    if (oc >= BF_TC) return;
    const int j = tile_c * BF_TC + oc;
#pragma unroll
    for (int k = 0; k < BF_STRIP; ++k) {
        const int orow = strip * BF_STRIP + k;
        if (orow >= BF_TR) break;
        const int i = tile_r * BF_TR + orow;
        const float *c = sA + (orow + 1) * BF_AC + (oc + 1);
        const float v = *c;
        float gx = 0.0f, gy = 0.0f;
        if (BF_OCC(v)) {
            acc.cnt += 1;
            acc.si += i;
            acc.sj += j;
            if (scharr_at(c, gx, gy)) {
                const double di = (double)(i - i0), dj = (double)(j - j0);
                const double dgx = (double)gx, dgy = (double)gy;
                acc.sgx += dgx;
                acc.sgy += dgy;
                acc.sigx += di * dgx;
                acc.sjgx += dj * dgx;
                acc.sigy += di * dgy;
                acc.sjgy += dj * dgy;
            }
        }
        if (MATERIALISE) {
            if (i < rows && j < cols) {
                const size_t o = (size_t)i * cols + j;
                if (out_img) out_img[o] = v;
                if (out_gx) out_gx[o] = gx;
                if (out_gy) out_gy[o] = gy;
            }
        }
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// CTA-wide reduction of the per-thread accumulators; thread 0 writes BF_NSUMS doubles to `slot`.
// Fixed association order => bit-reproducible.
__device__ void acc_block_reduce(const Acc &a, double *sred /* [16][BF_NSUMS] */, double *slot) {
    double v[BF_NSUMS] = {(double)a.cnt, (double)a.si, (double)a.sj, a.sgx, a.sgy, a.sigx, a.sjgx, a.sigy, a.sjgy};
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < BF_NSUMS; ++k) {
        v[k] = warp_sum(v[k]);
        if (lane == 0) sred[warp * BF_NSUMS + k] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < BF_NSUMS) {
        double s = 0.0;
        for (int w = 0; w < BF_NT / 32; ++w) s += sred[w * BF_NSUMS + threadIdx.x];
        slot[threadIdx.x] = s;
    }
    __syncthreads();
}

// Image pass of one CTA: tiles rank, rank+G, ... of the slice's image, double-buffered through
// shared memory with cp.async.
template <int SH, bool MATERIALISE>
__device__ void image_pass(Acc &acc, const u64 *img, int pitch, const BfGeom &g, const BfPack &pk, int rank,
                           int G, u64 *sP0, u64 *sP1, float *sA, float *out_img, float *out_gx, float *out_gy) {
    const int tiles_c = (g.cols + BF_TC - 1) / BF_TC;
    const int tiles_r = (g.rows + BF_TR - 1) / BF_TR;
    const int n_tiles = tiles_r * tiles_c;
    const int i0 = g.rows / 2, j0 = g.cols / 2;
    int t = rank;
    if (t < n_tiles) tile_load<SH>(sP0, img, pitch, t / tiles_c, t % tiles_c);
    cp_async_commit();
    int buf = 0;
    for (; t < n_tiles; t += G) {
        const int tn = t + G;
        if (tn < n_tiles) tile_load<SH>(buf ? sP0 : sP1, img, pitch, tn / tiles_c, tn % tiles_c);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        tile_mean<SH>(sA, buf ? sP1 : sP0, pk);
        __syncthreads();
        tile_reduce<MATERIALISE>(acc, sA, t / tiles_c, t % tiles_c, g.rows, g.cols, i0, j0, out_img, out_gx, out_gy);
        __syncthreads();
        buf ^= 1;
    }
    cp_async_wait<0>();
}

// Sum the G per-CTA partial records of a group in a fixed order (lane-strided, then butterfly):
// every CTA of the group obtains the bit-identical result.  Call from warp 0; result valid in all lanes.
__device__ __forceinline__ void group_sums(BfSums &s, const double *partials, int G) {
    const int lane = threadIdx.x & 31;
    double v[BF_NSUMS];
#pragma unroll
    for (int k = 0; k < BF_NSUMS; ++k) v[k] = 0.0;
    for (int r = lane; r < G; r += 32) {
#pragma unroll
        for (int k = 0; k < BF_NSUMS; ++k) v[k] += __ldcg(partials + r * BF_NSUMS + k);
    }
#pragma unroll
    for (int k = 0; k < BF_NSUMS; ++k) v[k] = warp_sum(v[k]);
    s.cnt = v[0]; s.si = v[1]; s.sj = v[2]; s.sgx = v[3]; s.sgy = v[4];
    s.sigx = v[5]; s.sjgx = v[6]; s.sigy = v[7]; s.sjgy = v[8];
}
