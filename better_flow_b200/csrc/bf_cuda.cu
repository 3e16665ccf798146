// bf_cuda.cu -- kernels + C ABI (include/bf_cuda.h) of the B200-native motion-compensation path.
//
// One persistent cooperative kernel runs OptimizerRolling::run() (optimizer_rolling.h:48-125) for a
// whole batch of independent slices: the grid is cut into groups of G CTAs, each group pulls
// slices from a queue and iterates   event pass -> barrier -> image pass -> barrier -> GD update
// entirely on the device; no host round trip per iteration.  See DESIGN.md.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <chrono>
#include <string>
#include <vector>

#include "bf_device.cuh"

// =================================================================================================
// Kernels
// =================================================================================================

struct __align__(16) Smem {
    double red[BF_NW * BF_NSUMS];
    unsigned short list[2][BF_LIST_CAP];   // live cells of this iteration / of the previous one
    int scan[BF_NW];
    float2 rcp_tab[BF_RCP_TAB];            // (c, RN(1/c)) for the fast unpack
    SliceDesc sd;
    BfGeom g;
    BfPack pk;
    BfProj proj;
    BfOpt opt;
    LocalOpt lopt;
    int cont;
    int guard;
    int minmax[6];
    double cpart[BF_NSUMS];                // CLUSTER instances: this CTA's partial sums, read by the peers through DSMEM
#if BF_TMA_PATCH
    unsigned long long tma_bar[BF_NW];     // one mbarrier per warp (TMA tile staging of the cell patches)
    unsigned tma_phase[BF_NW];
#endif
};

__device__ __forceinline__ void copy_cg(void *dst, const void *src, int bytes);

// What the iteration loop of a rolling slice carries from one iteration to the next, besides the
// per-slice scalars in shared memory (S.sd, S.g, S.pk, S.opt, S.proj).  A helper group enters the
// loop of ANOTHER group's slice with a LoopState rebuilt from that group's JoinRecord.
struct LoopState {
    int vgroup;            // group whose slice / images / flags / partial sums / barrier are used
    int rank, G;           // this CTA's rank among the G CTAs currently working on the slice
    int iter, buf, n_prev;
    unsigned tag, bar_target;
    bool helper;
};

template <int SH, bool CLUSTER = false>
__device__ void slice_loop(const KParams &P, Smem &S, LoopState &L, const TmaMaps *tm) {
    GroupWs *ws = P.ws + L.vgroup;
    u64 *img0 = P.images + (size_t)L.vgroup * 2 * P.img_elems;
    u64 *img1 = img0 + P.img_elems;
    unsigned *flags0 = P.flags + (size_t)L.vgroup * 2 * P.flag_elems;
    unsigned *flags1 = flags0 + P.flag_elems;
    double *partials = P.partials + (size_t)L.vgroup * P.part_stride * BF_NSUMS;
    const int i0 = S.g.rows / 2, j0 = S.g.cols / 2;
    const int rank = L.rank;
    const bool leader = !L.helper && rank == 0 && threadIdx.x == 0;
    const bool may_grow = !CLUSTER && P.allow_help != 0;   // (earlier helpers are full members: they must follow later growth too)

    // per-slice cell tables live behind the fixed part of the shared-memory block
    int2 *row_tab = reinterpret_cast<int2 *>(reinterpret_cast<unsigned char *>(&S) + sizeof(Smem));
    short2 *col_tab = reinterpret_cast<short2 *>(row_tab + P.tab_rows);
    unsigned *bm = reinterpret_cast<unsigned *>(col_tab + P.tab_cols);   // stamp bitmap, all-zero between event passes
    fill_cell_tables<SH>(row_tab, col_tab, S.g.rows, S.g.cols, SH);
    __syncthreads();

    long long *pf = P.prof ? P.prof + (size_t)blockIdx.x * BF_NPROF : nullptr;
    const bool prof = pf != nullptr && threadIdx.x == 0;
    long long tc = prof ? clock64() : 0;
#define PF_MARK(slot)                         \
    if (prof) {                               \
        const long long now_ = clock64();     \
        pf[slot] += now_ - tc;                \
        tc = now_;                            \
    }
    int buf = L.buf;
    int n_prev = L.n_prev;   // live cells of the previous iteration when they are still listed in S.list[buf ^ 1]
    int G = L.G;
    unsigned tag = L.tag, bar_target = L.bar_target;
    for (int iter = L.iter;; ++iter) {
        u64 *img_new = buf ? img1 : img0;
        u64 *img_old = buf ? img0 : img1;
        unsigned *flags_new = buf ? flags1 : flags0;
        unsigned *flags_old = buf ? flags0 : flags1;
        tag += 1;
        event_pass<SH>(P, S.sd, S.g, S.pk, S.proj, rank, G, iter == 0, iter > 0 || S.sd.has_init != 0, img_new, nullptr,
                       flags_new, tag, row_tab, col_tab, bm);
        if (pf) __syncthreads();
        PF_MARK(PF_EVENT);
        {
            const long long sp = sync_group<CLUSTER>(&ws->bar, bar_target, G);   // A: all splats of this iteration are in L2
            if (prof) pf[PF_BAR_A_SPIN] += sp;
        }
        PF_MARK(PF_BAR_A);

        Acc acc;
        acc_zero(acc);
#if BF_TMA_PATCH
        {
            const int warp = threadIdx.x >> 5;
            TmaWarp tw;
            tw.map = &tm->m[SH];
            tw.buf = reinterpret_cast<u64 *>(reinterpret_cast<unsigned char *>(&S) + P.tma_off) + (size_t)warp * P.tma_tile_elems;
            tw.bar = smem_u32(&S.tma_bar[warp]);
            tw.phase = S.tma_phase[warp];
            tw.img_index = L.vgroup * 2 + buf;
            n_prev = image_pass<SH, false>(acc, img_new, P.pitch, S.g, S.pk, S.rcp_tab, flags_new, tag, rank, G, S.list[buf], S.scan,
                                           nullptr, nullptr, nullptr, iter > 0 ? img_old : nullptr, flags_old, tag - 1,
                                           n_prev >= 0 ? S.list[buf ^ 1] : nullptr, n_prev, &tw);
            if ((threadIdx.x & 31) == 0) S.tma_phase[warp] = tw.phase;
        }
#else
        (void)tm;
        n_prev = image_pass<SH, false>(acc, img_new, P.pitch, S.g, S.pk, S.rcp_tab, flags_new, tag, rank, G, S.list[buf], S.scan,
                                       nullptr, nullptr, nullptr, iter > 0 ? img_old : nullptr, flags_old, tag - 1,
                                       n_prev >= 0 ? S.list[buf ^ 1] : nullptr, n_prev);
#endif
        if (pf) __syncthreads();
        PF_MARK(PF_CELLS);
        acc_block_reduce(acc, S.red, CLUSTER ? S.cpart : partials + rank * BF_NSUMS);
        PF_MARK(PF_REDUCE);
        // HELPING, victim side: the leader's decision to let a claimed helper in travels with barrier B
        // (and once the previous helper is in, the slice is opened again for the next one, up to BF_MAX_GROW groups)
        if (leader && may_grow) {
            const unsigned st = ld_relaxed_u32(&ws->help_state);
            if (st == (unsigned)HELP_CLAIMED) ws->grow_iter = iter;
            else if (st == (unsigned)HELP_JOINED && G + P.G <= P.max_grow * P.G) atomicExch(&ws->help_state, (unsigned)HELP_OPEN);
        }
        {
            const long long sp = sync_group<CLUSTER>(&ws->bar, bar_target, G);   // B: all partial sums are visible
            if (prof) pf[PF_BAR_B_SPIN] += sp;
        }
        PF_MARK(PF_BAR_B);
        const bool grow = may_grow && __ldcg(&ws->grow_iter) == iter;

        const bool block_gather = !CLUSTER && G >= BF_BLOCK_GATHER_MIN && G <= BF_NT;
        if (block_gather) group_sums_block_gather(partials, G, S.red);
        if (threadIdx.x < 32) {
            BfSums s;
            if constexpr (CLUSTER) group_sums_dsmem(s, S.cpart, G);
            else if (block_gather) group_sums_block_finish(s, S.red);
            else group_sums(s, partials, G, pf ? pf + 14 : nullptr);
            PF_MARK(PF_SCAN);   // (slot reused: time of the partial-sum gather)
            const bool cont = opt_advance_warp(S.opt, S.g, s, i0, j0, S.sd.max_iter, P.iter_cap, S.proj);
            if (threadIdx.x == 0) S.cont = cont ? 1 : 0;
        }
        __syncthreads();
        PF_MARK(PF_SERIAL);
        if (prof) pf[PF_ITERS] += 1;
        if (!S.cont) break;
        buf ^= 1;
        if (grow) {
            // from the next iteration on the slice is worked on by G + P.G CTAs: ranks G .. G + P.G - 1 are the new helpers
            if (leader) {
                JoinRecord &r = P.join[L.vgroup];
                r.sd = S.sd; r.g = S.g; r.pk = S.pk; r.opt = S.opt; r.proj = S.proj;
                r.slice = S.sd.n;   // (informational)
                r.iter_next = iter + 1; r.buf = buf; r.tag = tag; r.bar_target = bar_target;
                r.G_new = G + P.G; r.rank_base = G;
                __threadfence();
                red_release_add_u32(&ws->join_seq, 1u);
                atomicExch(&ws->help_state, (unsigned)HELP_JOINED);
            }
            G = G + P.G;
            // event chunks are re-dealt among the larger group: drop L1 lines of neighbouring chunks' states this SM may hold
            if (threadIdx.x == 0) fence_acq_rel_gpu();
            __syncthreads();
        }
    }
    // Zero the cells of the image that is still live (its readers all passed barrier B; the next
    // slice's first splat comes two group barriers later).
    {
        Acc none;
        image_pass<SH, false>(none, nullptr, P.pitch, S.g, S.pk, S.rcp_tab, nullptr, 0u, rank, G, S.list[buf ^ 1], S.scan, nullptr,
                              nullptr, nullptr, buf ? img1 : img0, buf ? flags1 : flags0, tag,
                              n_prev >= 0 ? S.list[buf] : nullptr, n_prev);
    }
    // Last re-projection of iteration_step (optimizer_rolling.h:340-344): only needed when the caller
    // wants the per-event state back (writeout_events).
    if (P.want_events)
        event_pass<SH>(P, S.sd, S.g, S.pk, S.proj, rank, G, false, true, nullptr, P.nxy, nullptr, 0u, nullptr, nullptr);
    if (pf) __syncthreads();
    PF_MARK(PF_FINAL);
    if (prof && !L.helper) pf[PF_SLICES] += 1;
#undef PF_MARK
    L.tag = tag; L.bar_target = bar_target; L.G = G;
}

// OptimizerRolling::run for the slice in S.sd, owned by this CTA's group.
template <int SH, bool CLUSTER = false>
__device__ void run_slice(const KParams &P, Smem &S, GroupWs *ws, unsigned &bar_target, unsigned &tag, int group,
                          int rank, const TmaMaps *tm) {
    if (threadIdx.x == 0) {
        bf_opt_init(S.opt, S.sd.has_init ? &S.sd.init : nullptr);
        if (S.sd.has_init) {
            // set_model (optimizer_rolling.h:289-299): note model.cx/cy are used as stored
            const bf_model &m = S.sd.init;
            bf_make_proj(S.proj, -m.total_dx, -m.total_dy, m.cx, m.cy, m.total_div, -m.total_rot);
        }
        if (P.allow_help && rank == 0) {
            ws->grow_iter = -1;
            __threadfence();
            atomicExch(&ws->help_state, (unsigned)HELP_OPEN);
        }
    }
    LoopState L;
    L.vgroup = group; L.rank = rank; L.G = P.G; L.iter = 0; L.buf = 0; L.n_prev = -1;
    L.tag = tag; L.bar_target = bar_target; L.helper = false;
    slice_loop<SH, CLUSTER>(P, S, L, tm);
    tag = L.tag; bar_target = L.bar_target;
    // the slice is over: nobody can join any more (a helper that had claimed but not joined sees 0 and leaves)
    if (P.allow_help && rank == 0 && threadIdx.x == 0) atomicExch(&ws->help_state, (unsigned)HELP_CLOSED);
}

// HELPING, helper side: join the slice group `v` is minimising (its leader has published P.join[v]).
template <int SH>
__device__ void help_slice(const KParams &P, Smem &S, int v, int my_rank, const TmaMaps *tm) {
    const JoinRecord &r = P.join[v];
    if (threadIdx.x == 0) {
        copy_cg(&S.sd, &r.sd, (int)sizeof(SliceDesc));
        copy_cg(&S.g, &r.g, (int)sizeof(BfGeom));
        copy_cg(&S.pk, &r.pk, (int)sizeof(BfPack));
        copy_cg(&S.opt, &r.opt, (int)sizeof(BfOpt));
        copy_cg(&S.proj, &r.proj, (int)sizeof(BfProj));
    }
    __syncthreads();   // slice_loop reads S.g right away
    LoopState L;
    L.vgroup = v; L.rank = __ldcg(&r.rank_base) + my_rank; L.G = __ldcg(&r.G_new);
    L.iter = __ldcg(&r.iter_next); L.buf = __ldcg(&r.buf); L.n_prev = -1;
    L.tag = __ldcg(&r.tag); L.bar_target = __ldcg(&r.bar_target); L.helper = true;
    slice_loop<SH>(P, S, L, tm);   // (its leading __syncthreads publishes the shared-memory copy to the CTA)
}

// OptimizerLocal::run (optimizer_sampler.cpp:4-38) for the slice in S.sd: same double-buffered
// event pass -> barrier -> image pass -> barrier -> control step cycle as run_slice, one cycle per
// iteration_step.
__device__ void help_phase(const KParams &P, Smem &S, GroupWs *ws, unsigned &bar_target, int group, int rank, const TmaMaps *tm);

template <int SH>
__device__ void run_slice_local(const KParams &P, Smem &S, GroupWs *ws, unsigned &bar_target, unsigned &tag, int group,
                                int rank) {
    u64 *img0 = P.images + (size_t)group * 2 * P.img_elems;
    u64 *img1 = img0 + P.img_elems;
    unsigned *flags0 = P.flags + (size_t)group * 2 * P.flag_elems;
    unsigned *flags1 = flags0 + P.flag_elems;
    double *partials = P.partials + (size_t)group * P.part_stride * BF_NSUMS;
    int2 *row_tab = reinterpret_cast<int2 *>(reinterpret_cast<unsigned char *>(&S) + sizeof(Smem));
    short2 *col_tab = reinterpret_cast<short2 *>(row_tab + P.tab_rows);
    unsigned *bm = reinterpret_cast<unsigned *>(col_tab + P.tab_cols);
    fill_cell_tables<SH, true>(row_tab, col_tab, S.g.rows, S.g.cols, 2 * SH);
    if (threadIdx.x == 0) local_opt_init(S.lopt, S.g.scale);
    __syncthreads();
    int buf = 0;
    int n_prev = -1;
    for (int iter = 0;; ++iter) {
        u64 *img_new = buf ? img1 : img0;
        u64 *img_old = buf ? img0 : img1;
        unsigned *flags_new = buf ? flags1 : flags0;
        unsigned *flags_old = buf ? flags0 : flags1;
        tag += 1;
        local_event_pass<SH>(P, S.sd, S.g, S.pk, S.lopt.cur_nx, S.lopt.cur_ny, rank, P.G, img_new, flags_new, tag, row_tab, col_tab, bm);
        group_barrier(&ws->bar, bar_target, P.G);
        Acc acc;
        acc_zero(acc);
        n_prev = image_pass<SH, false, 1>(acc, img_new, P.pitch, S.g, S.pk, S.rcp_tab, flags_new, tag, rank, P.G, S.list[buf],
                                          S.scan, nullptr, nullptr, nullptr, iter > 0 ? img_old : nullptr, flags_old, tag - 1,
                                          n_prev >= 0 ? S.list[buf ^ 1] : nullptr, n_prev);
        acc_block_reduce(acc, S.red, partials + rank * BF_NSUMS);
        group_barrier(&ws->bar, bar_target, P.G);
        const bool block_gather = P.G >= BF_BLOCK_GATHER_MIN && P.G <= BF_NT;
        if (block_gather) group_sums_block_gather(partials, P.G, S.red);
        if (threadIdx.x < 32) {
            BfSums s;
            if (block_gather) group_sums_block_finish(s, S.red);
            else group_sums(s, partials, P.G);
            if (threadIdx.x == 0) S.cont = local_opt_advance(S.lopt, s.cnt, s.si, P.iter_cap) ? 1 : 0;
        }
        __syncthreads();
        if (!S.cont) break;
        buf ^= 1;
    }
    {
        Acc none;
        image_pass<SH, false, 1>(none, nullptr, P.pitch, S.g, S.pk, S.rcp_tab, nullptr, 0u, rank, P.G, S.list[buf ^ 1], S.scan,
                                 nullptr, nullptr, nullptr, buf ? img1 : img0, buf ? flags1 : flags0, tag,
                                 n_prev >= 0 ? S.list[buf] : nullptr, n_prev);
    }
    // results in the slots of bf_slice_result documented in include/bf_cuda.h (bf_batch_add_local)
    if (threadIdx.x == 0) {
        bf_opt_init(S.opt, nullptr);
        S.opt.m.total_dx = S.lopt.nx; S.opt.m.total_dy = S.lopt.ny;
        S.opt.m.dx = S.lopt.last_score; S.opt.m.dy = S.lopt.dnx; S.opt.m.rot = S.lopt.dny; S.opt.m.div = S.lopt.dn_th;
        S.opt.m.cnt = S.lopt.nz_cnt;
        S.opt.iters = S.lopt.steps; S.opt.rc = S.lopt.rc;
        S.opt.x_div = S.opt.y_div = S.opt.rot_div = S.opt.div_div = 0.0f;
    }
    __syncthreads();
}

// Copy `bytes` (a multiple of 8) from global memory through L2 only (the source was written by another SM).
__device__ __forceinline__ void copy_cg(void *dst, const void *src, int bytes) {
    const long long *s8 = reinterpret_cast<const long long *>(src);
    long long *d8 = reinterpret_cast<long long *>(dst);
    for (int k = 0; k < bytes / 8; ++k) d8[k] = __ldcg(s8 + k);
}

// HELPING, helper side (see bf_device.cuh).  Entered by a group that found the slice queue empty.
__device__ void help_phase(const KParams &P, Smem &S, GroupWs *ws, unsigned &bar_target, int group, int rank, const TmaMaps *tm) {
    const int n_groups = (int)gridDim.x / P.G;
    if (rank == 0 && threadIdx.x == 0) atomicAdd(P.groups_done, 1u);
    for (;;) {
        // ---- the group's leader claims a victim (or learns that every group is out of slices)
        if (rank == 0 && threadIdx.x == 0) {
            int v = -1;
            unsigned seq = 0;
            for (;;) {
                for (int k = 1; k < n_groups && v < 0; ++k) {
                    const int g2 = (group + k) % n_groups;
                    GroupWs *w = P.ws + g2;
                    if (ld_relaxed_u32(&w->help_state) != (unsigned)HELP_OPEN) continue;
                    const unsigned sq = ld_relaxed_u32(&w->join_seq);   // sampled BEFORE the claim
                    if (atomicCAS(&w->help_state, (unsigned)HELP_OPEN, (unsigned)HELP_CLAIMED) == (unsigned)HELP_OPEN) { v = g2; seq = sq; }
                }
                if (v >= 0 || ld_relaxed_u32(P.groups_done) >= (unsigned)n_groups) break;
                __nanosleep(2000);
            }
            ws->victim = v;
            ws->victim_seq = seq;
        }
        group_barrier(&ws->bar, bar_target, P.G);
        const int v = __ldcg(&ws->victim);
        if (v < 0) return;
        const unsigned seq0 = __ldcg(&ws->victim_seq);
        // ---- every CTA of the helper group waits for the victim's verdict: joined, or slice over
        if (threadIdx.x == 0) {
            GroupWs *w = P.ws + v;
            int joined = 0;
            for (;;) {
                if (ld_relaxed_u32(&w->join_seq) != seq0) { joined = 1; break; }
                const unsigned st = ld_relaxed_u32(&w->help_state);
                if (st == (unsigned)HELP_CLOSED || st == (unsigned)HELP_OPEN) {
                    // over -- unless the record was published and the slice then finished with us still
                    // polling, which cannot happen (the victim waits for us at its next barrier); re-check anyway
                    fence_acq_rel_gpu();
                    joined = ld_relaxed_u32(&w->join_seq) != seq0 ? 1 : 0;
                    break;
                }
                __nanosleep(200);
            }
            if (joined) fence_acq_rel_gpu();   // acquire the JoinRecord
            S.cont = joined;
        }
        __syncthreads();
        const int joined = S.cont;
        __syncthreads();
        if (joined) {
            const int scale = __ldcg(&P.join[v].sd.scale);
            switch (scale) {
                case 1: help_slice<0>(P, S, v, rank, tm); break;
                case 3: help_slice<1>(P, S, v, rank, tm); break;
                default: help_slice<2>(P, S, v, rank, tm); break;
            }
        }
        __syncthreads();
    }
}

// MINB = resident CTAs per SM the instance is compiled for: 1 -> 128 registers/thread, 2 -> 64
// (twice the warps to hide L2 latency, at the price of a few spills).
// MINB = resident CTAs per SM the instance is compiled for; DELTA = expands the compact upload format in its prologue;
// CLUSTER = every group is one thread-block cluster (hardware barrier, partial sums through distributed shared memory;
// rolling slices only, no tail helping): the single-slice / warm-start-chain instance.
template <int MINB, bool DELTA = false, bool CLUSTER = false>
__global__ void __launch_bounds__(BF_NT, MINB) bf_minimize_kernel(const KParams P, const __grid_constant__ TmaMaps TM) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    const int group = blockIdx.x / P.G;
    const int rank = blockIdx.x - group * P.G;
    GroupWs *ws = P.ws + group;
    unsigned bar_target = 0;
    unsigned tag = P.tag_base;   // advanced once per splatting event pass, identically in every CTA of the group
    int parity = 0;
    long long *pf = P.prof ? P.prof + (size_t)blockIdx.x * BF_NPROF : nullptr;
    const long long t_begin = pf ? clock64() : 0;
    fill_rcp_table(S.rcp_tab);   // made visible by the first group barrier's __syncthreads
#if BF_TMA_PATCH
    if (threadIdx.x < BF_NW) {
        mbar_init(smem_u32(&S.tma_bar[threadIdx.x]), 1u);
        S.tma_phase[threadIdx.x] = 0u;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    if (P.bm_words > 0) {
        unsigned *bm = reinterpret_cast<unsigned *>(smem_raw + sizeof(Smem) + (size_t)P.tab_rows * sizeof(int2) +
                                                    (size_t)P.tab_cols * sizeof(short2));
        for (int w = threadIdx.x; w < P.bm_words; w += BF_NT) bm[w] = 0u;
    }

    for (;;) {
        const long long t_pro = pf ? clock64() : 0;
        // ---- fetch the next slice for this group ------------------------------------------------
        if (rank == 0 && threadIdx.x == 0) {
            const int s = atomicAdd(P.queue, 1);
            // streamed upload: the copy engine is still filling the event buffer in slice order; wait
            // until this slice has landed (the counter is written by a copy that follows the data)
            if (P.ready != nullptr && s < P.n_slices) {
                while (ld_relaxed_u32(P.ready) <= (unsigned)s) __nanosleep(200);
                fence_acq_rel_gpu();   // acquire: the copy engine wrote the slice's events before it bumped the counter
            }
            int *bb = ws->bbox[parity];
            bb[0] = INT_MAX; bb[1] = INT_MIN; bb[2] = INT_MAX; bb[3] = INT_MIN; bb[4] = INT_MAX; bb[5] = INT_MIN;
            ws->cur_slice = s;
        }
        sync_group<CLUSTER>(&ws->bar, bar_target, P.G);
        const int slice = __ldcg(&ws->cur_slice);
        if (slice >= P.n_slices) {
            if constexpr (!CLUSTER) {
                if (!P.allow_help) break;
                help_phase(P, S, ws, bar_target, group, rank, &TM);   // HELPING, helper side: returns when nothing is left to help
            }
            break;
        }
        if (threadIdx.x == 0) {
            S.sd = P.slices[slice];
            if (S.sd.has_init == 2) {
                // warm start from the PREVIOUS slice's result record, which is still on the device (bf_ring_slice:
                // set_model(last_model), dvs_flow.h:218-219, without a host round trip); the launch that wrote it
                // precedes this one in stream order
                copy_cg(&S.sd.init, &P.chain_src->model, (int)sizeof(bf_model));
                S.sd.has_init = 1;
            }
            S.minmax[0] = INT_MAX; S.minmax[1] = INT_MIN; S.minmax[2] = INT_MAX;
            S.minmax[3] = INT_MIN; S.minmax[4] = INT_MAX; S.minmax[5] = INT_MIN;
        }
        __syncthreads();

        // ---- bbox over fr_x / fr_y (optimizer_rolling.h:252-260) and the local-time range ------
        {
            const int n = S.sd.n;
            const int per = (((n + P.G - 1) / P.G) + 31) & ~31;
            const int lo = rank * per, hi = min(n, lo + per);
            const bf_event *ev = P.events + S.sd.ev_off;
            if constexpr (DELTA) {
                // compact upload (bf_batch_add_delta): expand this slice's 6-byte records into the event buffer; the same
                // pass yields the bounding box.  Only the instance launched by a delta-format streamed run carries this
                // code -- the register allocation of the kernel is fragile, and the plain instance must not pay for it.
                delta_expand_slice(P.delta_rec, P.delta_blocks, P.events_w, S.sd.block0, n, rank, P.G, S.scan, S.minmax);
                const bool any = rank < (n + BF_DELTA_BLOCK - 1) / BF_DELTA_BLOCK;
                __syncthreads();
                if (threadIdx.x == 0 && any) {
                    int *bb = ws->bbox[parity];
                    atomicMin(&bb[0], S.minmax[0]); atomicMax(&bb[1], S.minmax[1]);
                    atomicMin(&bb[2], S.minmax[2]); atomicMax(&bb[3], S.minmax[3]);
                    atomicMin(&bb[4], S.minmax[4]); atomicMax(&bb[5], S.minmax[5]);
                }
            } else {
            int xmin = INT_MAX, xmax = INT_MIN, ymin = INT_MAX, ymax = INT_MIN, tmin = INT_MAX, tmax = INT_MIN;
            for (int i = lo + (int)threadIdx.x; i < hi; i += BF_NT) {
                const uint2 e = ld_nc_u32x2(ev + i);
                const int fx = (int)(e.x & 0xffffu), fy = (int)((e.x >> 16) & 0x7fffu), t = (int)e.y;
                xmin = min(xmin, fx); xmax = max(xmax, fx);
                ymin = min(ymin, fy); ymax = max(ymax, fy);
                tmin = min(tmin, t); tmax = max(tmax, t);
            }
            __syncwarp();
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                xmin = min(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
                xmax = max(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
                ymin = min(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
                ymax = max(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
                tmin = min(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
                tmax = max(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
            }
            if ((threadIdx.x & 31) == 0 && hi > lo) {
                atomicMin(&S.minmax[0], xmin); atomicMax(&S.minmax[1], xmax);
                atomicMin(&S.minmax[2], ymin); atomicMax(&S.minmax[3], ymax);
                atomicMin(&S.minmax[4], tmin); atomicMax(&S.minmax[5], tmax);
            }
            __syncthreads();
            if (threadIdx.x == 0 && hi > lo) {
                int *bb = ws->bbox[parity];
                atomicMin(&bb[0], S.minmax[0]); atomicMax(&bb[1], S.minmax[1]);
                atomicMin(&bb[2], S.minmax[2]); atomicMax(&bb[3], S.minmax[3]);
                atomicMin(&bb[4], S.minmax[4]); atomicMax(&bb[5], S.minmax[5]);
            }
            }
        }
        sync_group<CLUSTER>(&ws->bar, bar_target, P.G);

        // ---- geometry + guards (optimizer_rolling.h:248-283, 49-58) ---------------------------
        if (threadIdx.x == 0) {
            const int *bb = ws->bbox[parity];
            // the reference starts the minima at RES_X / RES_Y and the maxima at 0 (:252-253)
            const int x_min = min(P.res_x, __ldcg(&bb[0])), x_max = max(0, __ldcg(&bb[1]));
            const int y_min = min(P.res_y, __ldcg(&bb[2])), y_max = max(0, __ldcg(&bb[3]));
            const int t_min = S.sd.n > 0 ? __ldcg(&bb[4]) : 0, t_max = S.sd.n > 0 ? __ldcg(&bb[5]) : 0;
            bf_make_geom(S.g, x_min, x_max, y_min, y_max, S.sd.scale);
            bf_make_pack(S.pk, S.sd.n, t_min, t_max);
            S.guard = 0;
            if (bf_guard_tiny(S.g, P.res_x, P.res_y)) S.guard = 1;         // :49-55 (and optimizer_sampler.cpp:9-13)
            else if (S.sd.mode == 0 && S.sd.n < P.min_events) S.guard = 2; // :57-58 (OptimizerLocal has no such guard)
            else if (S.sd.n == 0) S.guard = 2;
        }
        __syncthreads();
        const int guard = S.guard;
        if (pf && threadIdx.x == 0) pf[PF_PROLOGUE] += clock64() - t_pro;

        if (!CLUSTER && guard == 0 && S.sd.mode == 1) {
            if constexpr (!CLUSTER) {
                if (S.sd.scale == 1) run_slice_local<0>(P, S, ws, bar_target, tag, group, rank);
                else if (S.sd.scale == 3) run_slice_local<1>(P, S, ws, bar_target, tag, group, rank);
                else run_slice_local<2>(P, S, ws, bar_target, tag, group, rank);
            }
        } else if (guard == 0) {
            switch (S.sd.scale) {
                case 1: run_slice<0, CLUSTER>(P, S, ws, bar_target, tag, group, rank, &TM); break;
                case 3: run_slice<1, CLUSTER>(P, S, ws, bar_target, tag, group, rank, &TM); break;
                default: run_slice<2, CLUSTER>(P, S, ws, bar_target, tag, group, rank, &TM); break;
            }
        } else if (threadIdx.x == 0) {
            bf_opt_init(S.opt, S.sd.has_init ? &S.sd.init : nullptr);
            S.opt.rc = BF_RC_SKIPPED;
        }
        if (S.sd.mode == 1) {
            // OptimizerLocal keeps no per-event state on the device (pr is a closed form of nx, ny)
        } else if (guard != 0 && S.sd.has_init && P.want_events) {
            // set_model already re-projected the events before run() bailed out (dvs_flow.h:218-222)
            if (threadIdx.x == 0) {
                const bf_model &m = S.sd.init;
                bf_make_proj(S.proj, -m.total_dx, -m.total_dy, m.cx, m.cy, m.total_div, -m.total_rot);
            }
            __syncthreads();
            event_pass<0>(P, S.sd, S.g, S.pk, S.proj, rank, P.G, true, true, nullptr, P.nxy, nullptr, 0u, nullptr, nullptr);
        } else if (guard != 0 && P.want_events) {
            BfProj none;
            none.dnx = none.dny = none.cx = none.cy = none.div = none.s = 0; none.c = 1;
            event_pass<0>(P, S.sd, S.g, S.pk, none, rank, P.G, true, false, nullptr, P.nxy, nullptr, 0u, nullptr, nullptr);
        }
        __syncthreads();

        if (rank == 0 && threadIdx.x == 0) {
            bf_slice_result r;
            r.model = S.opt.m;
            r.rc = S.opt.rc;
            r.iters = S.opt.iters;
            r.dividers[0] = S.opt.x_div; r.dividers[1] = S.opt.y_div;
            r.dividers[2] = S.opt.rot_div; r.dividers[3] = S.opt.div_div;
            r.x_min = S.g.x_min; r.x_max = S.g.x_max; r.y_min = S.g.y_min; r.y_max = S.g.y_max;
            r.img_rows = S.g.rows; r.img_cols = S.g.cols;
            r.x_shift = S.g.x_shift; r.y_shift = S.g.y_shift;
            r.n_events = S.sd.n;
            r.flags = ((guard == 1 && S.sd.mode == 0) ? BF_FLAG_ALL_NOISE : 0u) | (S.pk.q > 0 ? BF_FLAG_T_QUANTISED : 0u);
            P.results[slice] = r;
        }
        parity ^= 1;
    }
    if (pf && threadIdx.x == 0) pf[PF_TOTAL] += clock64() - t_begin;
    if constexpr (CLUSTER) cluster_barrier();   // no CTA may exit while a peer can still read its shared memory
}

// ---- stage-level kernels (AccelLib surface; same device functions as the persistent kernel) ----

struct StageParams {
    int n;
    const double *pr_x, *pr_y;
    const int *t;
    const unsigned char *noise;
    BfGeom g;
    BfPack pk;
    u64 *img;
    int pitch;
    unsigned *flags;
    unsigned tag;
    double *partials;
    float *out_img, *out_gx, *out_gy;
    double *out7;
};

// AccelLib::get_time_img_cpu's splat loop (accel_lib.h:151-166) as point splats.
template <int SH>
__global__ void bf_stage_splat_kernel(const StageParams P, int clear) {
    const int n_ci = (P.g.rows + BF_CELL_ROWS - 1) / BF_CELL_ROWS, n_cj = (P.g.cols + CellCfg<SH>::CW - 1) / CellCfg<SH>::CW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += gridDim.x * blockDim.x) {
        if (P.noise && P.noise[i]) continue;
        int x, y;
        PixelMap pm;
        make_pixel_map(pm, P.g, P.pitch);
        if (!event_pixel(P.pr_x[i], P.pr_y[i], pm, x, y)) continue;
        const long long o = pixel_offset(x, y, P.pitch);
        if (clear) P.img[o] = 0ull;
        else {
            atomicAdd(P.img + o, bf_pack_value(P.pk, P.t[i]));
            mark_cells<SH>(P.flags, P.tag, x, y, n_ci, n_cj);
        }
    }
}

template <int SH>
__global__ void __launch_bounds__(BF_NT, 1) bf_stage_image_kernel(const StageParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem &S = *reinterpret_cast<Smem *>(smem_raw);
    Acc acc;
    acc_zero(acc);
    fill_rcp_table(S.rcp_tab);
    __syncthreads();
    image_pass<SH, true>(acc, P.img, P.pitch, P.g, P.pk, S.rcp_tab, P.flags, P.tag, blockIdx.x, gridDim.x, S.list[0], S.scan,
                         P.out_img, P.out_gx, P.out_gy, nullptr, nullptr, 0u, nullptr, -1);
    acc_block_reduce(acc, S.red, P.partials + blockIdx.x * BF_NSUMS);
}

__global__ void bf_stage_finish_kernel(const StageParams P, int G) {
    BfSums s;
    group_sums(s, P.partials, G);
    if (threadIdx.x == 0) {
        bf_model m;
        bf_sums_to_model(m, s, P.g.rows / 2, P.g.cols / 2);
        P.out7[0] = m.cx; P.out7[1] = m.cy; P.out7[2] = m.dx; P.out7[3] = m.dy;
        P.out7[4] = m.rot; P.out7[5] = m.div; P.out7[6] = (double)m.cnt;
    }
}

// ObjectModel::update / AccelLib::Sobel_cpu on a dense f32 image handed in by the caller (debug and
// API-completeness path; the optimiser itself never materialises this image).  One thread per pixel,
// taps straight from global memory (L1/L2 absorb the 9x reuse).
__global__ void __launch_bounds__(256) bf_dense_model_kernel(int rows, int cols, const float *img, float *gx_out,
                                                             float *gy_out, double *partials) {
    __shared__ double sred[8 * BF_NSUMS];
    Acc acc;
    acc_zero(acc);
    const int i0 = rows / 2, j0 = cols / 2;
    const long long P = (long long)rows * cols;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < P; k += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(k / cols), j = (int)(k - (long long)i * cols);
        const float v = img[k];
        float gx = 0.0f, gy = 0.0f;
        if (BF_OCC(v)) {
            acc.cnt += 1; acc.si += i; acc.sj += j;
            if (i >= 1 && i < rows - 1 && j >= 1 && j < cols - 1) {
                const float *c = img + k;
                const float L0 = c[-cols - 1], L1 = c[-1], L2 = c[cols - 1];
                const float U = c[-cols], D = c[cols];
                const float R0 = c[-cols + 1], R1 = c[1], R2 = c[cols + 1];
                if (BF_OCC(L0) && BF_OCC(L1) && BF_OCC(L2) && BF_OCC(U) && BF_OCC(D) && BF_OCC(R0) && BF_OCC(R1) && BF_OCC(R2)) {
                    float a = __fmul_rn(L0, 3.0f);
                    a = __fadd_rn(a, __fmul_rn(L2, -3.0f));
                    a = __fadd_rn(a, __fmul_rn(U, 10.0f));
                    a = __fadd_rn(a, __fmul_rn(D, -10.0f));
                    a = __fadd_rn(a, __fmul_rn(R0, 3.0f));
                    a = __fadd_rn(a, __fmul_rn(R2, -3.0f));
                    float b = __fmul_rn(L0, 3.0f);
                    b = __fadd_rn(b, __fmul_rn(L1, 10.0f));
                    b = __fadd_rn(b, __fmul_rn(L2, 3.0f));
                    b = __fadd_rn(b, __fmul_rn(R0, -3.0f));
                    b = __fadd_rn(b, __fmul_rn(R1, -10.0f));
                    b = __fadd_rn(b, __fmul_rn(R2, -3.0f));
                    gx = a; gy = b;
                    const double di = (double)(i - i0), dj = (double)(j - j0), dgx = (double)gx, dgy = (double)gy;
                    acc.sgx += dgx; acc.sgy += dgy;
                    acc.sigx += di * dgx; acc.sjgx += dj * dgx; acc.sigy += di * dgy; acc.sjgy += dj * dgy;
                }
            }
        }
        if (gx_out) gx_out[k] = gx;
        if (gy_out) gy_out[k] = gy;
    }
    double v[BF_NSUMS] = {(double)acc.cnt, (double)acc.si, (double)acc.sj, acc.sgx, acc.sgy, acc.sigx, acc.sjgx, acc.sigy, acc.sjgy};
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
        for (int w = 0; w < 8; ++w) s += sred[w * BF_NSUMS + threadIdx.x];
        partials[blockIdx.x * BF_NSUMS + threadIdx.x] = s;
    }
}

// AccelLib::project_4param_reinit (accel_lib.h:263-267)
__global__ void bf_stage_project_kernel(int n, const unsigned short *fr_x, const unsigned short *fr_y,
                                        const int *t, double *pr_x, double *pr_y, double *nx, double *ny,
                                        const BfProj q) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double px = pr_x[i], py = pr_y[i], ex, ey;
        float mx, my;
        project_event(px, py, ex, ey, mx, my, (double)fr_x[i], (double)fr_y[i], (float)t[i], q);
        pr_x[i] = px; pr_y[i] = py;
        if (nx) nx[i] = ex;
        if (ny) ny[i] = ey;
    }
}

// ---- EventFile::projection_img (event_file.h:460-515) -------------------------------------------------------------
// Not a performance path (debug images): four plain per-event / per-pixel kernels.
__global__ void bf_proj_splat_kernel(int n, const double *pr_x, const double *pr_y, const unsigned char *noise, int scale,
                                     int res_x, int res_y, unsigned *cnt, int cols) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (noise && noise[i]) continue;
        const double fx = __dmul_rn(pr_x[i], (double)scale), fy = __dmul_rn(pr_y[i], (double)scale);
        if (!(fx == fx) || !(fy == fy)) continue;                      // x86 cvttsd2si gives INT_MIN for NaN: rejected by x < 0
        const int x = __double2int_rz(fx), y = __double2int_rz(fy);    // (saturating: +-huge values are rejected below either way)
        if (x >= scale * (res_x - 1) || x < 0 || y >= scale * (res_y - 1) || y < 0) continue;   // :486-487
        atomicAdd(cnt + (size_t)(x + scale / 2) * cols + (y + scale / 2), 1u);                 // point count; the block splat is the box sum below
    }
}
// saturating scale x scale block splat = min(255, box sum of the point counts)
__global__ void bf_proj_box_kernel(const unsigned *cnt, int rows, int cols, int h, unsigned char *img) {
    const long long P = (long long)rows * cols;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < P; k += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(k / cols), j = (int)(k - (long long)i * cols);
        unsigned s = 0;
        for (int a = max(i - h, 0); a <= min(i + h, rows - 1); ++a)
            for (int b = max(j - h, 0); b <= min(j + h, cols - 1); ++b) s += cnt[(size_t)a * cols + b];
        img[k] = (unsigned char)min(s, 255u);
    }
}
__device__ __forceinline__ int reflect101(int i, int n) {
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}
// cv::GaussianBlur(k x k, sigma 0) on CV_8UC1: OpenCV's fixed table for k <= 7 -- [1 2 1]/4 and [1 4 6 4 1]/16 -- in
// fixed point, rounded once after both passes, BORDER_REFLECT_101 (pinned on the real cv2 in tests/golden); plus the
// nonzero count / sum of the result (EventFile::nonzero_average)
__global__ void bf_proj_blur_kernel(const unsigned char *img, int rows, int cols, int k, unsigned char *out, unsigned long long *nz) {
    const int w3[3] = {1, 2, 1}, w5[5] = {1, 4, 6, 4, 1};
    const int r = k / 2, shift = k == 1 ? 0 : (k == 3 ? 4 : 8);
    unsigned long long my_cnt = 0, my_sum = 0;
    const long long P = (long long)rows * cols;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < P; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q / cols), j = (int)(q - (long long)i * cols);
        int s = 0;
        if (k == 1) s = img[q];
        else
            for (int a = -r; a <= r; ++a) {
                const int wa = k == 3 ? w3[a + r] : w5[a + r];
                const unsigned char *row = img + (size_t)reflect101(i + a, rows) * cols;
                int t = 0;
                for (int b = -r; b <= r; ++b) t += (k == 3 ? w3[b + r] : w5[b + r]) * row[reflect101(j + b, cols)];
                s += wa * t;
            }
        const unsigned char v = (unsigned char)((s + ((1 << shift) >> 1)) >> shift);
        out[q] = v;
        if (v) { my_cnt += 1; my_sum += v; }
    }
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_cnt += __shfl_xor_sync(0xffffffffu, my_cnt, o);
        my_sum += __shfl_xor_sync(0xffffffffu, my_sum, o);
    }
    if ((threadIdx.x & 31) == 0 && my_cnt) { atomicAdd(nz, my_cnt); atomicAdd(nz + 1, my_sum); }
}
// cv::convertScaleAbs(img, img, 127 / avg, 0) for CV_8U: saturate_cast<uchar>(|float(src) * float(alpha)|), round-half-even
__global__ void bf_proj_scale_kernel(unsigned char *img, long long P, const unsigned long long *nz, double *avg_out) {
    const unsigned long long c = nz[0], s = nz[1];
    const double avg = c ? (double)s / (double)c : 0.0;
    if (blockIdx.x == 0 && threadIdx.x == 0 && avg_out) *avg_out = avg;
    if (c == 0) return;
    const float alpha = (float)(127.0 / avg);
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < P; q += (long long)gridDim.x * blockDim.x) {
        const int v = __float2int_rn(fabsf(__fmul_rn((float)img[q], alpha)));
        img[q] = (unsigned char)min(v, 255);
    }
}

// ---- EventFile::color_time_img (event_file.h:649-747) ------------------------------------------------------------
__global__ void bf_color_splat_kernel(int n, const double *pr_x, const double *pr_y, const int *t, const unsigned char *noise,
                                      int scale, int wx, int wy, double x_shift, double y_shift, int t_min, int t_max,
                                      double *sum_cos, double *sum_sin, unsigned *cnt, int cols) {
    const int h = scale / 2;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (noise && noise[i]) continue;
        const double fx = __dadd_rn(__dmul_rn(pr_x[i], (double)scale), x_shift), fy = __dadd_rn(__dmul_rn(pr_y[i], (double)scale), y_shift);
        if (!(fx == fx) || !(fy == fy)) continue;
        int x = __double2int_rz(fx), y = __double2int_rz(fy);
        if (x >= wx || x < 0 || y >= wy || y < 0) continue;                                   // :706-709
        const float angle = (float)(2 * 3.14 * ((double)(t[i] - t_min) / (double)(t_max - t_min)));   // :711
        const double co = cos((double)angle), si = sin((double)angle);
        x += h; y += h;
        for (int jx = x - h; jx <= x + h; ++jx)
            for (int jy = y - h; jy <= y + h; ++jy) {
                const size_t o = (size_t)jx * cols + jy;
                atomicAdd(sum_cos + o, co); atomicAdd(sum_sin + o, si); atomicAdd(cnt + o, 1u);
            }
    }
}
// mean direction -> HSV (:724-739) -> BGR (cv::cvtColor HSV2BGR for 8-bit: the sector table of the float formula, results
// TRUNCATED to 8 bits -- equal to OpenCV 4.13 on 99.65 % of all (H, S) pairs at V = 255, one level off on the rest)
__global__ void bf_color_finish_kernel(const double *sum_cos, const double *sum_sin, const unsigned *cnt, long long P, unsigned char *out) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < P; k += (long long)gridDim.x * blockDim.x) {
        unsigned char b = 0, g = 0, r = 0;
        const unsigned c = cnt[k];
        if (c >= 1u) {
            const float vx = __fdiv_rn((float)sum_cos[k], (float)c), vy = __fdiv_rn((float)sum_sin[k], (float)c);
            const double speed = hypot((double)vx, (double)vy);
            double angle = 0;
            if (speed != 0) angle = (atan2((double)vy, (double)vx) + 3.1416) * 180 / 3.1416;
            const int H = (int)(angle / 2), S = min(255, (int)(speed * 255));
            // HSV -> BGR with V = 255
            const float s = (float)S * (1.0f / 255.0f);
            float hh = (float)H * (6.0f / 180.0f);
            int sector = (int)floorf(hh);
            hh -= (float)sector;
            sector = ((sector % 6) + 6) % 6;
            const float tab[4] = {1.0f, 1.0f - s, 1.0f - s * hh, 1.0f - s * (1.0f - hh)};
            const int sd[6][3] = {{1, 3, 0}, {1, 0, 2}, {3, 0, 1}, {0, 2, 1}, {0, 1, 3}, {2, 1, 0}};
            b = (unsigned char)min(255, max(0, __float2int_rz(__fmul_rn(tab[sd[sector][0]], 255.0f))));
            g = (unsigned char)min(255, max(0, __float2int_rz(__fmul_rn(tab[sd[sector][1]], 255.0f))));
            r = (unsigned char)min(255, max(0, __float2int_rz(__fmul_rn(tab[sd[sector][2]], 255.0f))));
        }
        out[3 * k] = b; out[3 * k + 1] = g; out[3 * k + 2] = r;
    }
}

// ---- device-resident slice ring (include/bf_cuda.h: bf_ring_*) -------------------------------------------------
// Cuts "the newest n events, newest -> oldest, local time = timestamp - start" (what a range-for over the
// reference's CircularArray hands to the optimiser, dvs_flow.h:196-198 + Event::set_local_time, event.h:61-63) out of
// the ring into the packed event buffer, and writes the slice descriptor.  Event g of the stream lives at ring index
// g % cap.  If the previous slice hit the tiny-window guard, its events (stream indices [prev_lo, prev_hi)) are marked
// as noise here -- in the ring too -- exactly as run() marks them in the reference's buffer (optimizer_rolling.h:49-55).
__global__ void bf_ring_build_kernel(bf_ring_event *ring, long long cap, long long head, int n, unsigned long long start,
                                     bf_event *out, SliceDesc *desc, int scale, int max_iter, int chain,
                                     const bf_slice_result *prev, long long prev_lo, long long prev_hi) {
    const bool prev_noise = prev != nullptr && (prev->flags & BF_FLAG_ALL_NOISE) != 0u;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const long long g = head - 1 - i;
        bf_ring_event *slot = ring + (g % cap);
        const bf_ring_event e = *slot;
        unsigned fy = e.fr_y;
        if (prev_noise && g >= prev_lo && g < prev_hi && !(fy & BF_EVENT_NOISE)) {
            fy |= BF_EVENT_NOISE;
            slot->fr_y = (uint16_t)fy;
        }
        long long dt = (long long)(e.timestamp - start);
        dt = dt > (long long)INT_MAX ? (long long)INT_MAX : (dt < (long long)INT_MIN ? (long long)INT_MIN : dt);   // (the host checks the range)
        bf_event o;
        o.fr_x = e.fr_x; o.fr_y = (uint16_t)fy; o.t_ns = (int32_t)dt;
        out[i] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        SliceDesc d;
        d.ev_off = 0; d.n = n; d.scale = scale; d.max_iter = max_iter; d.has_init = chain ? 2 : 0; d.mode = 0; d.block0 = -1;
        d.init.cx = d.init.cy = d.init.dx = d.init.dy = d.init.rot = d.init.div = 0; d.init.cnt = 0; d.init.pad_ = 0;
        d.init.total_dx = d.init.total_dy = d.init.total_rot = d.init.total_div = 0;
        *desc = d;
    }
}

// =================================================================================================
// Host side
// =================================================================================================

static thread_local std::string g_err;
static int g_device = -1;

static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

// used by bf_multi.cpp (same shared object) to report through bf_last_error()
extern "C" void bf_set_error_(const char *msg) { g_err = msg ? msg : ""; }

#define CU(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return fail(BF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct bf_ctx {
    int device = 0;
    int sms = 0;
    int res_x = 0, res_y = 0, max_scale = 3;
    long long max_events = 0;
    int max_slices = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // H2D of the streamed upload
    cudaEvent_t ev_copy = nullptr;
    cudaEvent_t ev_free[2] = {nullptr, nullptr};   // recorded behind the last launch that read event buffer 0 / 1
    unsigned *d_ready = nullptr;          // [2][64]: slices uploaded so far (device, one counter per event buffer), fed from h_ready
    unsigned *h_ready = nullptr;          // [2][64] pinned
    // Two device copies of the batch (events + slice table): bf_batch_run_streamed fills the one the previous launch is
    // NOT reading, so the H2D of batch k+1 runs under the kernel of batch k.  `d_events` / `d_slices` below always point
    // at the current one; the second copy is allocated at the first streamed run.
    bf_event *ev_buf[2] = {nullptr, nullptr};
    struct SliceDesc *sl_buf[2] = {nullptr, nullptr};
    int cur = 0;
    int upload_chunks = 32;
    int upload_delay_us = 0;               // debug: host-side pause on the copy stream before every chunk (emulates a slow host link)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // options
    int opt_group = 0;          // CTAs per slice; 0 = auto (from the batch size)
    int min_group = 4;          // smallest automatic group.  DAVIS-240C, 592 slices, ms per launch (profiles/r2c_ab_group_size.txt):
                                // G = 2: 13.55, 3: 13.26, 4: 13.27, 5: 13.7, 6: 14.1, 8: 14.9, 16: 18.8 -- fewer slices in flight raise the
                                // L2 hit rate (35 -> 39 -> 72 % at G = 2, 4, 16) but every iteration pays its barriers and its serial GD
                                // step over less work per CTA; 4 also halves the image allocation of G = 2
    int ctas_per_sm = 2;        // 1 or 2 resident CTAs per SM
    long long image_budget_mb = 24576;   // cap on the point-image allocation
    int n_groups_alloc = 0;
    int iter_cap = 20000;
    int min_events = 1000;
    int cluster = 0;            // batch launches: 0 = groups on the global-memory barrier; 2..16 = every group is one thread-block cluster of that many CTAs
    int ring_cluster = 0;       // the same for the single-slice launches of bf_ring_slice (the warm-start chain)
    int cluster_max[BF_MAX_CLUSTER + 1] = {0};   // cudaOccupancyMaxActiveClusters per cluster size (0 = not queried yet)
    int smem_pad = 0;           // experiment knob: extra dynamic shared memory per CTA (shrinks the L1 carve-out)
    int tail_help = 1;          // idle groups join slices that are still running when the queue is empty
    int max_grow = 8;           // ... up to this many groups per slice

    // geometry of the stored images
    int pitch = 0, rows_alloc = 0;
    long long img_elems = 0;

    // current launch configuration
    int G = 0, n_groups = 0;

    // host (pinned)
    bf_event *h_events = nullptr;
    SliceDesc *h_slices = nullptr;
    bf_slice_result *h_results = nullptr;
    // device
    bf_event *d_events = nullptr;
    float2 *d_state = nullptr;
    double2 *d_pr_out = nullptr;
    double2 *d_nxy = nullptr;
    SliceDesc *d_slices = nullptr;
    bf_slice_result *d_results = nullptr;
    unsigned char *d_ctrl = nullptr;   // [queue (256 B)][GroupWs x n_groups]
    size_t ctrl_bytes = 0;
    double *d_partials = nullptr;
    JoinRecord *d_join = nullptr;
    u64 *d_images = nullptr;
    size_t images_bytes = 0;
    unsigned *d_flags = nullptr;
    long long flag_elems = 0;
    unsigned launch_seq = 0;     // tag_base = launch_seq << 20; flags are re-zeroed when it wraps
    TmaMaps tmaps;               // BF_TMA_PATCH: tensor maps over d_images, one per scale (re-encoded when the images are re-allocated)
    long long *d_prof = nullptr; // debug phase counters
    int profile = 0;
    // stage scratch
    void *d_stage = nullptr;
    size_t stage_bytes = 0;

    // compact upload format (bf_batch_add_delta): 6-byte records + block descriptors, pinned and on the device
    unsigned short *h_delta = nullptr, *d_delta[2] = {nullptr, nullptr};
    struct DeltaBlock *h_blocks = nullptr, *d_blocks[2] = {nullptr, nullptr};
    long long blocks_cap = 0;
    int n_blocks = 0;
    int delta_slices = 0;                  // slices of the current batch that were added in delta format
    std::vector<int> slice_block0;         // first block descriptor of every slice (+ one past the end)
    std::vector<struct bf_ring *> rings;   // device-resident slice rings of this context (destroyed with it)

    // batch state
    int n_slices = 0;
    long long n_events = 0;
    bool uploaded = false, ran = false, have_events = false;
    long long launches = 0;
};

static size_t smem_bytes() { return sizeof(Smem); }
// per-CTA stamp bitmap: one bit per cell flag (flag_elems is a multiple of 64)
static int bm_words_of(const bf_ctx *c) { return (int)(BF_STAMP_BYTES ? c->flag_elems / 4 : c->flag_elems / 32); }
// minimise kernel: fixed block + the per-slice cell tables (int2 per image row, short2 per image column)
static size_t smem_tables_end(const bf_ctx *c) {
    return sizeof(Smem) + (size_t)c->max_scale * c->res_x * sizeof(int2) + (size_t)c->max_scale * c->res_y * sizeof(short2) +
           (size_t)bm_words_of(c) * sizeof(unsigned);
}
// BF_TMA_PATCH: one (8 + 2H) x 32 tile of packed words per warp behind the tables, 128-byte aligned
static int tma_tile_elems_of(const bf_ctx *c) { return (BF_CELL_ROWS + 2 * (c->max_scale / 2 + 1)) * 32; }
static size_t tma_off_of(const bf_ctx *c) { return (smem_tables_end(c) + 127) & ~(size_t)127; }
static size_t smem_bytes_min(const bf_ctx *c) {
#if BF_TMA_PATCH
    return tma_off_of(c) + (size_t)BF_NW * tma_tile_elems_of(c) * sizeof(u64) + (size_t)c->smem_pad;
#else
    return smem_tables_end(c) + (size_t)c->smem_pad;
#endif
}

static int ensure_device() {
    if (g_device < 0) {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n <= 0)
            return fail(BF_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        g_device = 0;
    }
    CU(cudaSetDevice(g_device));
    return BF_OK;
}

// Launch geometry.  slots = SMs x resident CTAs per SM.  A group of G CTAs works on one slice at a
// time; the best G is the smallest that still keeps every slot busy: barriers and the serial GD
// update cost a fixed ~10 us per iteration per group, so fewer CTAs per slice and more slices in
// flight wins (measured: G=4 > 8 > 16 on DAVIS-240C), while a small batch wants all CTAs on its few
// slices.  Buffers are allocated once for the largest group count (slots / min_group, capped by
// the image-memory budget); G is then chosen per launch from the batch size.
static int max_groups(bf_ctx *c) {
    const int slots = c->sms * c->ctas_per_sm;
    long long by_mem = (long long)((double)c->image_budget_mb * 1048576.0 / (2.0 * (double)c->img_elems * 8.0));
    if (by_mem < 1) by_mem = 1;
    int g = slots / std::max(1, c->min_group);
    if (c->opt_group > 0) g = slots / std::min(slots, c->opt_group);
    g = (int)std::min<long long>(std::max(1, g), by_mem);
    return g;
}

static void pick_launch(bf_ctx *c, int n_slices, long long n_events, int *G, int *n_groups) {
    const int slots = c->sms * c->ctas_per_sm;
    int groups = c->n_groups_alloc;
    if (c->opt_group <= 0) groups = std::min(groups, std::max(1, n_slices));
    int g = std::max(1, slots / groups);
    if (c->opt_group > 0) g = std::min(slots, std::max(c->opt_group, g));
    else {
        // A small batch does not profit from every CTA: each iteration pays two barriers over the G CTAs of
        // a group and one record per CTA in the reduction, so beyond ~3000 events per CTA more CTAs make a
        // slice slower (single DAVIS-240C slice: 0.247 ms at G = 16..32, 0.299 ms at G = 296; measured sweep
        // in tools/sweep_single.py).
        const long long per_slice = n_events / std::max(1, n_slices);
        const int cap = (int)std::min<long long>(slots, std::max<long long>(16, per_slice / 3000));
        g = std::max(std::min(g, cap), std::min(c->min_group, g));
    }
    *G = g;
    *n_groups = std::min(groups, slots / g);
    // With tail helping, groups that get no slice of their own are not wasted: they join the running slices
    // one group per iteration (a single 50 k-event slice: 2.22 ms on a fixed 16 CTAs, 1.74 ms when the idle
    // groups may join).  So a small batch still launches every CTA slot.
    if (c->opt_group <= 0 && c->tail_help) *n_groups = std::max(*n_groups, std::min(c->n_groups_alloc, slots / g));
}

// Every launch gets a fresh range of 2^20 generation tags; when the 12-bit sequence wraps the flag
// arrays are cleared so that an old tag can never alias a new one.
static int next_tag_base(bf_ctx *c, unsigned *tag_base) {
    c->launch_seq += 1;
    if (c->launch_seq >= 4096u) {
        CU(cudaMemsetAsync(c->d_flags, 0, (size_t)c->n_groups_alloc * 2 * (size_t)c->flag_elems * sizeof(unsigned), c->stream));
        c->launch_seq = 1;
    }
    *tag_base = c->launch_seq << 20;
    return BF_OK;
}

static int configure(bf_ctx *c, int n_slices, long long n_events = -1, int fixed_group = 0) {
    if (n_events < 0) n_events = c->n_events;
    const int want = max_groups(c);
    if (want != c->n_groups_alloc || !c->d_images) {
        CU(cudaStreamSynchronize(c->stream));
        if (c->d_images) cudaFree(c->d_images);
        if (c->d_ctrl) cudaFree(c->d_ctrl);
        if (c->d_partials) cudaFree(c->d_partials);
        if (c->d_flags) cudaFree(c->d_flags);
        if (c->d_join) cudaFree(c->d_join);
        c->d_images = nullptr; c->d_ctrl = nullptr; c->d_partials = nullptr; c->d_flags = nullptr; c->d_join = nullptr;
        c->n_groups_alloc = want;
        c->images_bytes = (size_t)want * 2 * (size_t)c->img_elems * sizeof(u64);
        CU(cudaMalloc(&c->d_images, c->images_bytes));
        CU(cudaMemsetAsync(c->d_images, 0, c->images_bytes, c->stream));
        CU(cudaMalloc(&c->d_flags, (size_t)want * 2 * (size_t)c->flag_elems * sizeof(unsigned)));
        CU(cudaMemsetAsync(c->d_flags, 0, (size_t)want * 2 * (size_t)c->flag_elems * sizeof(unsigned), c->stream));
        c->launch_seq = 0;
        c->ctrl_bytes = 256 + (size_t)want * sizeof(GroupWs);
        CU(cudaMalloc(&c->d_ctrl, c->ctrl_bytes));
        // one partial record per CTA slot, whatever the grouping
        // one record per CTA slot, times BF_MAX_GROW: a helped slice is worked on by up to BF_MAX_GROW groups
        CU(cudaMalloc(&c->d_partials, (size_t)c->sms * 4 * BF_MAX_GROW * BF_NSUMS * sizeof(double)));
        CU(cudaMalloc(&c->d_join, (size_t)want * sizeof(JoinRecord)));
#if BF_TMA_PATCH
        {
            // cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
            typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            void *fn = nullptr;
            cudaDriverEntryPointQueryResult qres;
            CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
            if (!fn || qres != cudaDriverEntryPointSuccess) return fail(BF_ERR_CUDA, "cuTensorMapEncodeTiled is not available");
            const cuuint64_t dims[3] = {(cuuint64_t)c->pitch, (cuuint64_t)c->rows_alloc, (cuuint64_t)want * 2};
            const cuuint64_t strides[2] = {(cuuint64_t)c->pitch * sizeof(u64), (cuuint64_t)c->img_elems * sizeof(u64)};
            const cuuint32_t estr[3] = {1, 1, 1};
            for (int k = 0; k < 3; ++k) {
                const cuuint32_t box[3] = {32u, (cuuint32_t)(BF_CELL_ROWS + 2 * (k + 1)), 1u};
                const CUresult r = ((encode_fn)fn)(&c->tmaps.m[k], CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, c->d_images, dims, strides, box, estr,
                                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) return fail(BF_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
            }
        }
#endif
    }
    pick_launch(c, n_slices, n_events, &c->G, &c->n_groups);
    if (fixed_group > 0 && c->opt_group <= 0) {
        // one group of exactly this many CTAs, nobody joins later (LaunchSpec::group)
        c->G = std::min(c->sms * c->ctas_per_sm, fixed_group);
        c->n_groups = 1;
    }
    return BF_OK;
}

extern "C" {

const char *bf_version(void) { return "better_flow_b200 0.1 (sm_100a)"; }
const char *bf_last_error(void) { return g_err.c_str(); }

int bf_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int bf_cuda_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return fail(BF_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (device < 0 || device >= n) return fail(BF_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
    g_device = device;
    CU(cudaSetDevice(device));
    return BF_OK;
}

bf_ctx *bf_ctx_create(int sensor_rows, int sensor_cols, int max_scale, long long max_events, int max_slices) {
    if (sensor_rows <= 0 || sensor_cols <= 0 || sensor_rows > 32767 || sensor_cols > 32767 ||
        (max_scale != 1 && max_scale != 3 && max_scale != 5) || max_events <= 0 || max_slices <= 0) {
        fail(BF_ERR_ARG, "bf_ctx_create: bad arguments (scale must be 1, 3 or 5)");
        return nullptr;
    }
    if (ensure_device() != BF_OK) return nullptr;
    bf_ctx *c = new bf_ctx();
    c->device = g_device;
    c->res_x = sensor_rows; c->res_y = sensor_cols; c->max_scale = max_scale;
    c->max_events = max_events; c->max_slices = max_slices;
    cudaDeviceProp prop;
    auto bail = [&](const char *what, cudaError_t e) -> bf_ctx * {
        fail(BF_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(e));
        bf_ctx_destroy(c);
        return nullptr;
    };
    cudaError_t e;
    if ((e = cudaGetDeviceProperties(&prop, c->device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    c->sms = prop.multiProcessorCount;
    if (!prop.cooperativeLaunch) { fail(BF_ERR_CUDA, "device lacks cooperative launch"); bf_ctx_destroy(c); return nullptr; }
    if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    c->stream = c->own_stream;
    if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaEventCreateWithFlags(&c->ev_copy, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    for (int b = 0; b < 2; ++b)
        if ((e = cudaEventCreateWithFlags(&c->ev_free[b], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMalloc(&c->d_ready, 2 * 64 * sizeof(unsigned))) != cudaSuccess) return bail("cudaMalloc(ready)", e);
    if ((e = cudaMallocHost(&c->h_ready, 2 * 64 * sizeof(unsigned))) != cudaSuccess) return bail("cudaMallocHost(ready)", e);
    if ((e = cudaEventCreate(&c->ev0)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreate(&c->ev1)) != cudaSuccess) return bail("cudaEventCreate", e);

    const int max_rows = max_scale * sensor_rows, max_cols = max_scale * sensor_cols;
    // every cell patch (8 + 2H rows x 32 cols, H <= 3) of every cell that intersects the image must lie
    // inside the allocation: cells start at multiples of 8 rows / CW cols, patches extend H beyond.
    c->rows_alloc = ((max_rows + BF_CELL_ROWS - 1) / BF_CELL_ROWS) * BF_CELL_ROWS + 2 * BF_BORDER;
    c->pitch = max_cols + 32 + 2 * BF_BORDER;
    c->pitch = (c->pitch + 15) & ~15;   // 128-byte rows
    c->img_elems = (long long)c->rows_alloc * c->pitch;
    c->flag_elems = (long long)((max_rows + BF_CELL_ROWS - 1) / BF_CELL_ROWS) * ((max_cols + BF_CW_MIN - 1) / BF_CW_MIN);
    c->flag_elems = (c->flag_elems + 63) & ~63LL;

    if ((e = cudaMallocHost(&c->h_events, (size_t)max_events * sizeof(bf_event))) != cudaSuccess) return bail("cudaMallocHost(events)", e);
    if ((e = cudaMallocHost(&c->h_slices, (size_t)max_slices * sizeof(SliceDesc))) != cudaSuccess) return bail("cudaMallocHost(slices)", e);
    if ((e = cudaMallocHost(&c->h_results, (size_t)max_slices * sizeof(bf_slice_result))) != cudaSuccess) return bail("cudaMallocHost(results)", e);
    // (+2: the event pass reads events / states in aligned pairs, so the pair holding the last event is read whole)
    if ((e = cudaMalloc(&c->d_events, (size_t)(max_events + 2) * sizeof(bf_event))) != cudaSuccess) return bail("cudaMalloc(events)", e);
    // (never hand uninitialised coordinates to the kernel: the pair / sector holding a slice's last event is read whole)
    if ((e = cudaMemset(c->d_events, 0, (size_t)(max_events + 2) * sizeof(bf_event))) != cudaSuccess) return bail("cudaMemset(events)", e);
    if ((e = cudaMalloc(&c->d_state, (size_t)(max_events + 2) * sizeof(float2))) != cudaSuccess) return bail("cudaMalloc(state)", e);
    if ((e = cudaMalloc(&c->d_slices, (size_t)max_slices * sizeof(SliceDesc))) != cudaSuccess) return bail("cudaMalloc(slices)", e);
    c->ev_buf[0] = c->d_events; c->sl_buf[0] = c->d_slices;
    if ((e = cudaMalloc(&c->d_results, (size_t)max_slices * sizeof(bf_slice_result))) != cudaSuccess) return bail("cudaMalloc(results)", e);

    if (smem_bytes_min(c) > 112 * 1024) {
        fail(BF_ERR_ARG, "sensor too large for the per-slice cell tables (%zu bytes of shared memory)", smem_bytes_min(c));
        bf_ctx_destroy(c);
        return nullptr;
    }
    // The attribute is per function AND per device (not per context): only ever raise it, and keep one
    // high-water mark per device -- bf_multi_create makes a context on every device of the process.
    static int smem_attr_of[64] = {0};
    int &smem_attr = smem_attr_of[c->device & 63];
    if ((int)smem_bytes_min(c) > smem_attr) {
        smem_attr = (int)smem_bytes_min(c);
        if ((e = cudaFuncSetAttribute(bf_minimize_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr)) != cudaSuccess) return bail("cudaFuncSetAttribute", e);
        if ((e = cudaFuncSetAttribute(bf_minimize_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr)) != cudaSuccess) return bail("cudaFuncSetAttribute", e);
        if ((e = cudaFuncSetAttribute(bf_minimize_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr)) != cudaSuccess) return bail("cudaFuncSetAttribute", e);
        if ((e = cudaFuncSetAttribute(bf_minimize_kernel<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr)) != cudaSuccess) return bail("cudaFuncSetAttribute", e);
        if ((e = cudaFuncSetAttribute(bf_minimize_kernel<2, false, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1)) != cudaSuccess) return bail("cudaFuncSetAttribute", e);
#if BF_NT <= 256
        if ((e = cudaFuncSetAttribute(bf_minimize_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_attr)) != cudaSuccess) return bail("cudaFuncSetAttribute", e);
#endif
    }
    if ((e = cudaFuncSetAttribute(bf_stage_image_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes())) != cudaSuccess) return bail("cudaFuncSetAttribute", e);
    if ((e = cudaFuncSetAttribute(bf_stage_image_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes())) != cudaSuccess) return bail("cudaFuncSetAttribute", e);
    if ((e = cudaFuncSetAttribute(bf_stage_image_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes())) != cudaSuccess) return bail("cudaFuncSetAttribute", e);
    return c;
}

void bf_ctx_destroy(bf_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    while (!c->rings.empty()) bf_ring_destroy(c->rings.back());   // (each removes itself from the list)
    cudaFreeHost(c->h_events); cudaFreeHost(c->h_slices); cudaFreeHost(c->h_results);
    cudaFree(c->ev_buf[0]); cudaFree(c->ev_buf[1]); cudaFree(c->sl_buf[0]); cudaFree(c->sl_buf[1]);
    cudaFree(c->d_state); cudaFree(c->d_pr_out); cudaFree(c->d_nxy);
    cudaFreeHost(c->h_delta); cudaFree(c->d_delta[0]); cudaFree(c->d_delta[1]);
    cudaFreeHost(c->h_blocks); cudaFree(c->d_blocks[0]); cudaFree(c->d_blocks[1]);
    cudaFree(c->d_results); cudaFree(c->d_ctrl); cudaFree(c->d_partials); cudaFree(c->d_images); cudaFree(c->d_flags);
    cudaFree(c->d_stage); cudaFree(c->d_prof); cudaFree(c->d_join);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->ev_copy) cudaEventDestroy(c->ev_copy);
    for (int b = 0; b < 2; ++b)
        if (c->ev_free[b]) cudaEventDestroy(c->ev_free[b]);
    cudaFree(c->d_ready); cudaFreeHost(c->h_ready);
    delete c;
}

int bf_ctx_set_option(bf_ctx *c, const char *key, long long value) {
    if (!c || !key) return fail(BF_ERR_ARG, "null argument");
    if (!strcmp(key, "group_size")) c->opt_group = (int)value;
    else if (!strcmp(key, "iter_cap")) c->iter_cap = (int)std::max(1LL, value);
    else if (!strcmp(key, "min_events")) c->min_events = (int)value;
    else if (!strcmp(key, "min_group")) c->min_group = (int)std::max(1LL, value);
    else if (!strcmp(key, "image_budget_mb")) c->image_budget_mb = std::max(1LL, value);
    else if (!strcmp(key, "profile")) c->profile = (int)value;
    else if (!strcmp(key, "cluster") || !strcmp(key, "ring_cluster")) {
        if (value != 0 && value != 2 && value != 4 && value != 8 && value != 16) return fail(BF_ERR_ARG, "%s must be 0, 2, 4, 8 or 16", key);
        (key[0] == 'c' ? c->cluster : c->ring_cluster) = (int)value;
    }
    else if (!strcmp(key, "smem_pad")) {
        c->smem_pad = (int)std::max(0LL, std::min(value, 64LL * 1024));
        CU(cudaFuncSetAttribute(bf_minimize_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_min(c)));
        CU(cudaFuncSetAttribute(bf_minimize_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_min(c)));
        CU(cudaFuncSetAttribute(bf_minimize_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_min(c)));
        CU(cudaFuncSetAttribute(bf_minimize_kernel<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_min(c)));
    }
    else if (!strcmp(key, "tail_help")) c->tail_help = value ? 1 : 0;
    else if (!strcmp(key, "max_grow")) c->max_grow = (int)std::min<long long>(BF_MAX_GROW, std::max(1LL, value));
    else if (!strcmp(key, "upload_delay_us")) c->upload_delay_us = (int)std::max(0LL, std::min(value, 1000000LL));
    else if (!strcmp(key, "upload_chunks")) c->upload_chunks = (int)std::min(60LL, std::max(1LL, value));
    else if (!strcmp(key, "ctas_per_sm")) c->ctas_per_sm = (value >= 4 && BF_NT <= 256) ? 4 : (value >= 2 ? 2 : 1);
    else return fail(BF_ERR_ARG, "unknown option '%s'", key);
    return BF_OK;
}

long long bf_ctx_get_option(bf_ctx *c, const char *key) {
    if (!c || !key) return -1;
    if (!strcmp(key, "group_size")) return c->G;          // of the last launch
    if (!strcmp(key, "n_groups")) return c->n_groups;
    if (!strcmp(key, "min_group")) return c->min_group;
    if (!strcmp(key, "image_budget_mb")) return c->image_budget_mb;
    if (!strcmp(key, "iter_cap")) return c->iter_cap;
    if (!strcmp(key, "tail_help")) return c->tail_help;
    if (!strcmp(key, "min_events")) return c->min_events;
    if (!strcmp(key, "sms")) return c->sms;
    if (!strcmp(key, "ctas_per_sm")) return c->ctas_per_sm;
    if (!strcmp(key, "image_bytes")) return c->img_elems * 8;
    if (!strcmp(key, "smem_bytes")) return (long long)smem_bytes_min(c);
    if (!strcmp(key, "cluster")) return c->cluster;
    if (!strcmp(key, "ring_cluster")) return c->ring_cluster;
    return -1;
}

long long bf_ctx_launch_count(bf_ctx *c) { return c ? c->launches : 0; }

int bf_ctx_set_stream(bf_ctx *c, void *cuda_stream) {
    if (!c) return fail(BF_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return BF_OK;
}

// Debug: copies the per-CTA phase cycle counters of the last profiled launch ("profile" option)
// into out[ctas][16]; returns the number of CTAs.
int bf_debug_profile(bf_ctx *c, long long *out, int max_ctas) {
    if (!c || !out || !c->d_prof) return fail(BF_ERR_STATE, "profiling was not enabled");
    const int n = std::min(std::min(max_ctas, 1024), c->n_groups * c->G);
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaMemcpy(out, c->d_prof, (size_t)n * BF_NPROF * sizeof(long long), cudaMemcpyDeviceToHost));
    return n;
}

int bf_batch_results_device(bf_ctx *c, void **dev_ptr, long long *bytes) {
    if (!c || !dev_ptr || !bytes) return fail(BF_ERR_ARG, "bf_batch_results_device: null argument");
    *dev_ptr = c->d_results;
    *bytes = (long long)c->n_slices * (long long)sizeof(bf_slice_result);
    return BF_OK;
}

int bf_batch_reset(bf_ctx *c) {
    if (!c) return fail(BF_ERR_ARG, "null context");
    c->n_slices = 0; c->n_events = 0;
    c->uploaded = c->ran = c->have_events = false;
    c->n_blocks = 0; c->delta_slices = 0; c->slice_block0.clear();
    return BF_OK;
}

static int add_desc(bf_ctx *c, long long off, int n, int scale, int max_iter, const bf_model *init) {
    if (scale != 1 && scale != 3 && scale != 5) return fail(BF_ERR_ARG, "scale %d unsupported (1, 3 or 5)", scale);
    if (scale > c->max_scale) return fail(BF_ERR_ARG, "scale %d exceeds the context's max_scale %d", scale, c->max_scale);
    if (c->n_slices >= c->max_slices) return fail(BF_ERR_ARG, "batch full (%d slices)", c->max_slices);
    SliceDesc &d = c->h_slices[c->n_slices];
    memset(&d, 0, sizeof d);
    d.ev_off = off; d.n = n; d.scale = scale; d.max_iter = max_iter;
    d.block0 = -1;
    d.has_init = init ? 1 : 0;
    if (init) d.init = *init;
    c->uploaded = c->ran = false;
    return c->n_slices++;
}

int bf_batch_add(bf_ctx *c, const uint16_t *fr_x, const uint16_t *fr_y, const int32_t *t_ns,
                 const uint8_t *noise, int n, int scale, int max_iter, const bf_model *init) {
    if (!c || n < 0 || (n > 0 && (!fr_x || !fr_y || !t_ns))) return fail(BF_ERR_ARG, "bf_batch_add: bad arguments");
    if (c->n_events + n > c->max_events) return fail(BF_ERR_ARG, "batch event capacity exceeded (%lld)", c->max_events);
    bf_event *dst = c->h_events + c->n_events;
    unsigned bad = 0;
    for (int i = 0; i < n; ++i) {
        bad |= (unsigned)(fr_x[i] >= c->res_x) | (unsigned)(fr_y[i] >= c->res_y);
        dst[i].fr_x = fr_x[i];
        dst[i].fr_y = (uint16_t)(fr_y[i] | ((noise && noise[i]) ? BF_EVENT_NOISE : 0u));
        dst[i].t_ns = t_ns[i];
    }
    if (bad) return fail(BF_ERR_ARG, "event outside the %dx%d sensor", c->res_x, c->res_y);
    const int slot = add_desc(c, c->n_events, n, scale, max_iter, init);
    if (slot >= 0) c->n_events += n;
    return slot;
}

// OptimizerLocal(LinearEventCloud*, scale) (optimizer_sampler.h:41-56): the slice is minimised by
// OptimizerLocal::run instead of OptimizerRolling::run.
int bf_batch_add_local(bf_ctx *c, const uint16_t *fr_x, const uint16_t *fr_y, const int32_t *t_ns, int n, int scale) {
    if (scale != 1 && scale != 3 && scale != 5) return fail(BF_ERR_ARG, "bf_batch_add_local: scale %d unsupported (1, 3 or 5)", scale);
    const int slot = bf_batch_add(c, fr_x, fr_y, t_ns, nullptr, n, scale, -1, nullptr);
    if (slot >= 0) c->h_slices[slot].mode = 1;
    return slot;
}

int bf_batch_slot_mode(bf_ctx *c, int slot, int mode) {
    if (!c || slot < 0 || slot >= c->n_slices || (mode != 0 && mode != 1)) return fail(BF_ERR_ARG, "bf_batch_slot_mode: bad arguments");
    c->h_slices[slot].mode = mode;
    c->uploaded = c->ran = false;
    return BF_OK;
}

int bf_local_minimize(bf_ctx *c, const uint16_t *fr_x, const uint16_t *fr_y, const int32_t *t_ns, int n, int scale,
                      bf_slice_result *out) {
    int rc;
    if ((rc = bf_batch_reset(c)) != BF_OK) return rc;
    if ((rc = bf_batch_add_local(c, fr_x, fr_y, t_ns, n, scale)) < 0) return rc;
    if ((rc = bf_batch_run(c, 0)) != BF_OK) return rc;
    bf_slice_result r;
    if ((rc = bf_batch_result(c, 0, &r)) != BF_OK) return rc;
    if (out) *out = r;
    return r.rc;
}

int bf_batch_add_packed(bf_ctx *c, const bf_event *events, int n, int scale, int max_iter, const bf_model *init) {
    if (!c || n < 0 || (n > 0 && !events)) return fail(BF_ERR_ARG, "bf_batch_add_packed: bad arguments");
    if (c->n_events + n > c->max_events) return fail(BF_ERR_ARG, "batch event capacity exceeded (%lld)", c->max_events);
    if (!bf_events_in_sensor(events, n, c->res_x, c->res_y)) return fail(BF_ERR_ARG, "event outside the %dx%d sensor", c->res_x, c->res_y);
    memcpy(c->h_events + c->n_events, events, (size_t)n * sizeof(bf_event));
    const int slot = add_desc(c, c->n_events, n, scale, max_iter, init);
    if (slot >= 0) c->n_events += n;
    return slot;
}

int bf_batch_add_delta(bf_ctx *c, const bf_event *events, int n, int scale, int max_iter, const bf_model *init) {
    if (!c || n < 0 || (n > 0 && !events)) return fail(BF_ERR_ARG, "bf_batch_add_delta: bad arguments");
    if (c->n_events + n > c->max_events) return fail(BF_ERR_ARG, "batch event capacity exceeded (%lld)", c->max_events);
    if (c->delta_slices != c->n_slices) return fail(BF_ERR_STATE, "bf_batch_add_delta: the batch already holds slices in 8-byte format");
    if (!bf_events_in_sensor(events, n, c->res_x, c->res_y)) return fail(BF_ERR_ARG, "event outside the %dx%d sensor", c->res_x, c->res_y);
    CU(cudaSetDevice(c->device));
    if (!c->h_delta) {
        c->blocks_cap = c->max_events / BF_DELTA_BLOCK + c->max_slices + 2;
        CU(cudaMallocHost(&c->h_delta, (size_t)(c->max_events + 4) * 6));
        CU(cudaMallocHost(&c->h_blocks, (size_t)c->blocks_cap * sizeof(DeltaBlock)));
        // both device copies (and the second event buffer) now, not inside a later streamed run: cudaMalloc synchronises the device
        for (int b = 0; b < 2; ++b) {
            CU(cudaMalloc(&c->d_delta[b], (size_t)(c->max_events + 4) * 6));
            CU(cudaMalloc(&c->d_blocks[b], (size_t)c->blocks_cap * sizeof(DeltaBlock)));
            if (!c->ev_buf[b]) {
                CU(cudaMalloc(&c->ev_buf[b], (size_t)(c->max_events + 2) * sizeof(bf_event)));
                CU(cudaMemset(c->ev_buf[b], 0, (size_t)(c->max_events + 2) * sizeof(bf_event)));
                CU(cudaMalloc(&c->sl_buf[b], (size_t)c->max_slices * sizeof(SliceDesc)));
            }
        }
    }
    const int nb = (n + BF_DELTA_BLOCK - 1) / BF_DELTA_BLOCK;
    if (c->n_blocks + nb > c->blocks_cap) return fail(BF_ERR_ARG, "bf_batch_add_delta: block table full");
    unsigned short *rec = c->h_delta + (size_t)c->n_events * 3;
    for (int i = 0; i < n; ++i) {
        const bf_event &e = events[i];
        const unsigned fy = e.fr_y & 0x7fffu, nz = (e.fr_y & BF_EVENT_NOISE) ? 1u : 0u;
        long long dt = 0;
        if (i % BF_DELTA_BLOCK != 0) dt = (long long)events[i - 1].t_ns - (long long)e.t_ns;
        if (e.fr_x >= 4096u || fy >= 4096u || dt < 0 || dt >= (1ll << 23))
            return fail(BF_ERR_ARG, "bf_batch_add_delta: slice not representable (coordinate >= 4096, time running backwards or a gap >= 8.4 ms)");
        rec[3 * i] = (unsigned short)(e.fr_x | ((fy & 0xfu) << 12));
        rec[3 * i + 1] = (unsigned short)((fy >> 4) | (nz << 8) | (((unsigned)dt & 0x7fu) << 9));
        rec[3 * i + 2] = (unsigned short)((unsigned)dt >> 7);
    }
    for (int k = 0; k < nb; ++k) {
        DeltaBlock &b = c->h_blocks[c->n_blocks + k];
        b.first = c->n_events + (long long)k * BF_DELTA_BLOCK;
        b.count = std::min(BF_DELTA_BLOCK, n - k * BF_DELTA_BLOCK);
        b.t0 = events[(size_t)k * BF_DELTA_BLOCK].t_ns;
    }
    // (the 8-byte copy is kept too: bf_batch_upload / bf_batch_run use it, and it is the fall-back of the streamed path)
    memcpy(c->h_events + c->n_events, events, (size_t)n * sizeof(bf_event));
    const int slot = add_desc(c, c->n_events, n, scale, max_iter, init);
    if (slot < 0) return slot;
    c->h_slices[slot].block0 = c->n_blocks;
    c->slice_block0.resize((size_t)slot + 2);
    c->slice_block0[(size_t)slot] = c->n_blocks;
    c->n_blocks += nb;
    c->slice_block0[(size_t)slot + 1] = c->n_blocks;
    c->n_events += n;
    c->delta_slices += 1;
    return slot;
}

long long bf_batch_upload_bytes(bf_ctx *c) {
    if (!c) return 0;
    const long long table = (long long)c->n_slices * (long long)sizeof(SliceDesc);
    if (c->n_slices > 0 && c->delta_slices == c->n_slices) return c->n_events * 6 + (long long)c->n_blocks * (long long)sizeof(DeltaBlock) + table;
    return c->n_events * (long long)sizeof(bf_event) + table;
}

bf_event *bf_batch_staging(bf_ctx *c, long long *capacity) {
    if (!c) return nullptr;
    if (capacity) *capacity = c->max_events;
    return c->h_events;
}

int bf_batch_add_staged(bf_ctx *c, long long offset, int n, int scale, int max_iter, const bf_model *init) {
    if (!c || n < 0 || offset < 0 || offset + n > c->max_events) return fail(BF_ERR_ARG, "bf_batch_add_staged: bad range");
    if (!bf_events_in_sensor(c->h_events + offset, n, c->res_x, c->res_y)) return fail(BF_ERR_ARG, "event outside the %dx%d sensor", c->res_x, c->res_y);
    const int slot = add_desc(c, offset, n, scale, max_iter, init);
    if (slot >= 0) c->n_events = std::max(c->n_events, offset + n);
    return slot;
}

int bf_batch_upload(bf_ctx *c) {
    if (!c) return fail(BF_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    if (c->n_events > 0)
        CU(cudaMemcpyAsync(c->d_events, c->h_events, (size_t)c->n_events * sizeof(bf_event), cudaMemcpyHostToDevice, c->stream));
    if (c->n_slices > 0)
        CU(cudaMemcpyAsync(c->d_slices, c->h_slices, (size_t)c->n_slices * sizeof(SliceDesc), cudaMemcpyHostToDevice, c->stream));
    c->uploaded = true;
    return BF_OK;
}

static int launch_impl(bf_ctx *c, int want_events, const unsigned *ready, const unsigned short *delta_rec = nullptr,
                       const DeltaBlock *delta_blocks = nullptr);

int bf_batch_launch(bf_ctx *c, int want_events) {
    if (!c) return fail(BF_ERR_ARG, "null context");
    if (!c->uploaded) return fail(BF_ERR_STATE, "bf_batch_launch before bf_batch_upload");
    return launch_impl(c, want_events, nullptr);
}

// upload -> launch -> download with the event upload STREAMED: the slice table goes first, the kernel
// is launched at once, and the events follow in slice-ordered chunks on a second stream while the
// first slices are already being minimised (each chunk is followed by a 4-byte copy that bumps the
// device-side "slices uploaded" counter the kernel polls).  The upload goes into the device copy of the
// batch that the PREVIOUS launch is not reading (two copies, ping-pong), so back-to-back streamed runs
// overlap the H2D of batch k+1 with the kernel of batch k.  Asynchronous; pair with bf_batch_sync (and
// do not touch the staging buffer before that).
static void CUDART_CB upload_pause(void *us) {
    std::this_thread::sleep_for(std::chrono::microseconds((intptr_t)us));
}

int bf_batch_run_streamed(bf_ctx *c, int want_events) {
    if (!c) return fail(BF_ERR_ARG, "null context");
    CU(cudaSetDevice(c->device));
    if (c->n_slices == 0) { c->uploaded = c->ran = true; return BF_OK; }
    const int nb = c->cur ^ 1;
    if (!c->ev_buf[nb]) {
        CU(cudaMalloc(&c->ev_buf[nb], (size_t)(c->max_events + 2) * sizeof(bf_event)));
        CU(cudaMemset(c->ev_buf[nb], 0, (size_t)(c->max_events + 2) * sizeof(bf_event)));
        CU(cudaMalloc(&c->sl_buf[nb], (size_t)c->max_slices * sizeof(SliceDesc)));
    }
    const bool delta = c->delta_slices == c->n_slices && c->h_delta != nullptr;
    if (delta && !c->d_delta[nb]) {
        CU(cudaMalloc(&c->d_delta[nb], (size_t)(c->max_events + 4) * 6));
        CU(cudaMalloc(&c->d_blocks[nb], (size_t)c->blocks_cap * sizeof(DeltaBlock)));
    }
    unsigned *d_ready = c->d_ready + 64 * nb, *h_ready = c->h_ready + 64 * nb;
    // the copy engine must not write this buffer before the last launch that read it has finished
    CU(cudaStreamWaitEvent(c->copy_stream, c->ev_free[nb], 0));
    // (reset by a COPY of a pinned zero, not by cudaMemsetAsync: a small memset may be executed by a kernel, and a kernel on
    // this stream cannot run while the persistent kernel holds every SM -- the copies queued behind it would then never
    // start and the persistent kernel would wait for them for ever; only DMA operations may go on the copy stream)
    h_ready[63] = 0u;
    CU(cudaMemcpyAsync(d_ready, h_ready + 63, sizeof(unsigned), cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaMemcpyAsync(c->sl_buf[nb], c->h_slices, (size_t)c->n_slices * sizeof(SliceDesc), cudaMemcpyHostToDevice, c->copy_stream));
    CU(cudaEventRecord(c->ev_copy, c->copy_stream));
    CU(cudaStreamWaitEvent(c->stream, c->ev_copy, 0));
    // chunk boundaries on slice boundaries (slices are laid out back to back in add order)
    const int chunks = std::min(c->upload_chunks, c->n_slices);
    int s0 = 0;
    for (int k = 0; k < chunks; ++k) {
        const int s1 = (int)((long long)c->n_slices * (k + 1) / chunks);
        if (s1 <= s0) continue;
        const long long lo = c->h_slices[s0].ev_off;
        const long long hi = c->h_slices[s1 - 1].ev_off + c->h_slices[s1 - 1].n;
        if (c->upload_delay_us > 0) CU(cudaLaunchHostFunc(c->copy_stream, upload_pause, (void *)(intptr_t)c->upload_delay_us));
        if (hi > lo && delta) {
            // compact upload: the chunk's 6-byte records and block descriptors; the group that takes a slice expands it
            // into the event buffer in its prologue (delta_expand_slice).  DMA only on this stream: a second kernel could
            // not run beside the persistent one, which holds every register of every SM.
            const int b0 = c->slice_block0[(size_t)s0], b1 = c->slice_block0[(size_t)s1];
            CU(cudaMemcpyAsync(c->d_delta[nb] + lo * 3, c->h_delta + lo * 3, (size_t)(hi - lo) * 6, cudaMemcpyHostToDevice, c->copy_stream));
            if (b1 > b0)
                CU(cudaMemcpyAsync(c->d_blocks[nb] + b0, c->h_blocks + b0, (size_t)(b1 - b0) * sizeof(DeltaBlock), cudaMemcpyHostToDevice, c->copy_stream));
        } else if (hi > lo) {
            CU(cudaMemcpyAsync(c->ev_buf[nb] + lo, c->h_events + lo, (size_t)(hi - lo) * sizeof(bf_event), cudaMemcpyHostToDevice, c->copy_stream));
        }
        h_ready[k] = (unsigned)s1;
        CU(cudaMemcpyAsync(d_ready, h_ready + k, sizeof(unsigned), cudaMemcpyHostToDevice, c->copy_stream));
        s0 = s1;
    }
    c->cur = nb;
    c->d_events = c->ev_buf[nb];
    c->d_slices = c->sl_buf[nb];
    c->uploaded = true;
    int rc = delta ? launch_impl(c, want_events, d_ready, c->d_delta[nb], c->d_blocks[nb]) : launch_impl(c, want_events, d_ready);
    if (rc != BF_OK) return rc;
    return bf_batch_download(c);
}

// One persistent launch over `n_slices` slice descriptors.
struct LaunchSpec {
    const bf_event *events;
    const SliceDesc *slices;          // device
    bf_slice_result *results;         // device, [n_slices]
    int n_slices;
    long long n_events;               // (launch geometry only)
    const unsigned *ready;
    const bf_slice_result *chain_src; // device record for slices with has_init == 2, or null
    int want_events;
    int group = 0;                    // > 0: ONE group of this many CTAs from the first iteration on (no tail helping)
    int cluster = 0;                  // > 0: launch the CLUSTER instance with groups = thread-block clusters of this many CTAs
    bf_event *events_w = nullptr;     // compact upload: writable event buffer + records + block table (else null)
    const unsigned short *delta_rec = nullptr;
    const DeltaBlock *delta_blocks = nullptr;
};

static int launch_spec(bf_ctx *c, const LaunchSpec &L) {
    CU(cudaSetDevice(c->device));
    int rc = configure(c, L.n_slices, L.n_events, L.group);
    if (rc != BF_OK) return rc;
    if (L.want_events && !c->d_nxy) {
        CU(cudaMalloc(&c->d_nxy, (size_t)c->max_events * sizeof(double2)));
        CU(cudaMalloc(&c->d_pr_out, (size_t)c->max_events * sizeof(double2)));
    }
    CU(cudaMemsetAsync(c->d_ctrl, 0, c->ctrl_bytes, c->stream));
    KParams P;
    P.events = L.events; P.state = c->d_state;
    P.nxy = L.want_events ? c->d_nxy : nullptr; P.pr_out = L.want_events ? c->d_pr_out : nullptr;
    P.slices = L.slices; P.results = L.results; P.n_slices = L.n_slices;
    P.queue = reinterpret_cast<int *>(c->d_ctrl);
    P.ws = reinterpret_cast<GroupWs *>(c->d_ctrl + 256);
    P.partials = c->d_partials; P.images = c->d_images; P.img_elems = c->img_elems; P.pitch = c->pitch;
    P.flags = c->d_flags; P.flag_elems = c->flag_elems;
    if ((rc = next_tag_base(c, &P.tag_base)) != BF_OK) return rc;
    P.G = c->G; P.res_x = c->res_x; P.res_y = c->res_y; P.min_events = c->min_events;
    P.iter_cap = c->iter_cap; P.want_events = L.want_events ? 1 : 0;
    P.ready = L.ready;
    P.chain_src = L.chain_src;
    P.events_w = L.events_w; P.delta_rec = L.delta_rec; P.delta_blocks = L.delta_blocks;
    P.tma_off = (int)tma_off_of(c); P.tma_tile_elems = tma_tile_elems_of(c);
    P.tab_rows = c->max_scale * c->res_x; P.tab_cols = c->max_scale * c->res_y;
    P.bm_words = bm_words_of(c);
    P.allow_help = (c->tail_help && c->n_groups > 1) ? 1 : 0;
    P.max_grow = std::max(1, std::min(c->max_grow, BF_MAX_GROW));
    if (c->opt_group <= 0) {
        // a slice stops profiting from more CTAs at ~800 events per CTA (but take at least 64): measured sweep,
        // profiles/r1f_single_slice_sweep.txt
        const long long per_slice = L.n_events / std::max(1, L.n_slices);
        const long long want_ctas = std::max<long long>(64, per_slice / 800);
        P.max_grow = (int)std::max<long long>(1, std::min<long long>(P.max_grow, (want_ctas + c->G - 1) / c->G));
    }
    P.part_stride = P.allow_help ? P.max_grow * c->G : c->G;
    P.join = c->d_join;
    P.groups_done = reinterpret_cast<unsigned *>(c->d_ctrl + 128);
    P.prof = nullptr;
    if (c->profile) {
        if (!c->d_prof) CU(cudaMalloc(&c->d_prof, (size_t)1024 * BF_NPROF * sizeof(long long)));
        CU(cudaMemsetAsync(c->d_prof, 0, (size_t)1024 * BF_NPROF * sizeof(long long), c->stream));
        P.prof = c->d_prof;
    }
    void *args[] = {&P, &c->tmaps};
    void *kern = c->ctas_per_sm == 2 ? (void *)bf_minimize_kernel<2> : (void *)bf_minimize_kernel<1>;
#if BF_NT <= 256
    if (c->ctas_per_sm == 4) kern = (void *)bf_minimize_kernel<4>;
#endif
    if (L.delta_rec != nullptr) {
        // the instance that expands 6-byte delta records in its prologue (2 CTAs per SM only)
        if (c->ctas_per_sm != 2) return fail(BF_ERR_STATE, "the compact upload format needs ctas_per_sm = 2");
        kern = (void *)bf_minimize_kernel<2, true>;
    }
    if (L.cluster > 0 && L.delta_rec == nullptr && c->ctas_per_sm == 2) {
        // CLUSTER instance: one thread-block cluster per group.  No cooperative launch is needed: clusters never wait for
        // each other (no tail helping), and the CTAs of one cluster are co-scheduled by the hardware.
        const int cs = L.cluster;
        cudaLaunchConfig_t cfg = {};
        cfg.blockDim = dim3(BF_NT); cfg.dynamicSmemBytes = smem_bytes_min(c); cfg.stream = c->stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (c->cluster_max[cs] == 0) {
            cfg.gridDim = dim3((unsigned)cs);
            int n = 0;
            CU(cudaOccupancyMaxActiveClusters(&n, bf_minimize_kernel<2, false, true>, &cfg));
            if (n < 1) return fail(BF_ERR_CUDA, "clusters of %d CTAs cannot be scheduled on this device", cs);
            c->cluster_max[cs] = n;
        }
        c->G = cs;
        c->n_groups = std::max(1, std::min(std::min(L.n_slices, c->n_groups_alloc), c->cluster_max[cs]));
        P.G = cs; P.allow_help = 0; P.max_grow = 1; P.part_stride = cs;
        cfg.gridDim = dim3((unsigned)(c->n_groups * cs));
        CU(cudaLaunchKernelEx(&cfg, bf_minimize_kernel<2, false, true>, P, c->tmaps));
    } else
    CU(cudaLaunchCooperativeKernel(kern, dim3(c->n_groups * c->G), dim3(BF_NT), args, smem_bytes_min(c), c->stream));
    if (c->ev_buf[1]) CU(cudaEventRecord(c->ev_free[c->cur], c->stream));   // (a later streamed upload into this event buffer waits for this; no second buffer = no streamed upload yet)
    c->launches += 1;
    return BF_OK;
}

static int launch_impl(bf_ctx *c, int want_events, const unsigned *ready, const unsigned short *delta_rec, const DeltaBlock *delta_blocks) {
    CU(cudaSetDevice(c->device));
    if (c->n_slices == 0) { c->ran = true; return BF_OK; }
    LaunchSpec L{c->d_events, c->d_slices, c->d_results, c->n_slices, c->n_events, ready, nullptr, want_events};
    L.cluster = c->cluster;
    for (int k = 0; k < c->n_slices && L.cluster > 0; ++k)
        if (c->h_slices[k].mode != 0) L.cluster = 0;          // (the CLUSTER instance runs OptimizerRolling slices only)
    if (delta_rec) { L.events_w = c->d_events; L.delta_rec = delta_rec; L.delta_blocks = delta_blocks; }
    const int rc = launch_spec(c, L);
    if (rc != BF_OK) return rc;
    c->ran = true;
    c->have_events = want_events != 0;
    return BF_OK;
}

int bf_batch_download(bf_ctx *c) {
    if (!c) return fail(BF_ERR_ARG, "null context");
    if (!c->ran) return fail(BF_ERR_STATE, "bf_batch_download before bf_batch_launch");
    if (c->n_slices > 0)
        CU(cudaMemcpyAsync(c->h_results, c->d_results, (size_t)c->n_slices * sizeof(bf_slice_result), cudaMemcpyDeviceToHost, c->stream));
    return BF_OK;
}

int bf_batch_sync(bf_ctx *c) {
    if (!c) return fail(BF_ERR_ARG, "null context");
    CU(cudaStreamSynchronize(c->stream));
    return BF_OK;
}

int bf_batch_run(bf_ctx *c, int want_events) {
    int rc;
    if ((rc = bf_batch_upload(c)) != BF_OK) return rc;
    if ((rc = bf_batch_launch(c, want_events)) != BF_OK) return rc;
    if ((rc = bf_batch_download(c)) != BF_OK) return rc;
    return bf_batch_sync(c);
}

int bf_batch_time_launches(bf_ctx *c, int reps, int want_events, float *ms) {
    if (!c || reps <= 0 || !ms) return fail(BF_ERR_ARG, "bf_batch_time_launches: bad arguments");
    int rc;
    CU(cudaStreamSynchronize(c->stream));
    CU(cudaEventRecord(c->ev0, c->stream));
    for (int r = 0; r < reps; ++r)
        if ((rc = bf_batch_launch(c, want_events)) != BF_OK) return rc;
    CU(cudaEventRecord(c->ev1, c->stream));
    CU(cudaEventSynchronize(c->ev1));
    CU(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return BF_OK;
}

int bf_batch_size(bf_ctx *c) { return c ? c->n_slices : 0; }

int bf_batch_result(bf_ctx *c, int slot, bf_slice_result *out) {
    if (!c || !out || slot < 0 || slot >= c->n_slices) return fail(BF_ERR_ARG, "bf_batch_result: bad slot");
    if (!c->ran) return fail(BF_ERR_STATE, "bf_batch_result before the batch ran");
    *out = c->h_results[slot];
    return BF_OK;
}

int bf_batch_events(bf_ctx *c, int slot, double *pr_x, double *pr_y, double *nx, double *ny) {
    if (!c || slot < 0 || slot >= c->n_slices) return fail(BF_ERR_ARG, "bf_batch_events: bad slot");
    if (!c->ran || !c->have_events) return fail(BF_ERR_STATE, "bf_batch_events needs a launch with want_events");
    const SliceDesc &d = c->h_slices[slot];
    if (d.n == 0) return BF_OK;
    std::vector<double2> tmp((size_t)d.n);
    CU(cudaStreamSynchronize(c->stream));
    if (pr_x || pr_y) {
        CU(cudaMemcpy(tmp.data(), c->d_pr_out + d.ev_off, (size_t)d.n * sizeof(double2), cudaMemcpyDeviceToHost));
        for (int i = 0; i < d.n; ++i) { if (pr_x) pr_x[i] = tmp[i].x; if (pr_y) pr_y[i] = tmp[i].y; }
    }
    if (nx || ny) {
        CU(cudaMemcpy(tmp.data(), c->d_nxy + d.ev_off, (size_t)d.n * sizeof(double2), cudaMemcpyDeviceToHost));
        for (int i = 0; i < d.n; ++i) { if (nx) nx[i] = tmp[i].x; if (ny) ny[i] = tmp[i].y; }
    }
    return BF_OK;
}

int bf_minimize(bf_ctx *c, const uint16_t *fr_x, const uint16_t *fr_y, const int32_t *t_ns, const uint8_t *noise,
                int n, int scale, int max_iter, const bf_model *init, bf_slice_result *out, double *pr_x,
                double *pr_y, double *nx, double *ny) {
    int rc;
    if ((rc = bf_batch_reset(c)) != BF_OK) return rc;
    if ((rc = bf_batch_add(c, fr_x, fr_y, t_ns, noise, n, scale, max_iter, init)) < 0) return rc;
    const int want = (pr_x || pr_y || nx || ny) ? 1 : 0;
    if ((rc = bf_batch_run(c, want)) != BF_OK) return rc;
    bf_slice_result r;
    if ((rc = bf_batch_result(c, 0, &r)) != BF_OK) return rc;
    if (out) *out = r;
    if (want && (rc = bf_batch_events(c, 0, pr_x, pr_y, nx, ny)) != BF_OK) return rc;
    return r.rc;
}

// ---- stage-level entry points -----------------------------------------------------------------

static int stage_alloc(bf_ctx *c, size_t bytes) {
    if (bytes <= c->stage_bytes) return BF_OK;
    if (c->d_stage) cudaFree(c->d_stage);
    c->d_stage = nullptr; c->stage_bytes = 0;
    CU(cudaMalloc(&c->d_stage, bytes));
    c->stage_bytes = bytes;
    return BF_OK;
}

static int stage_image(bf_ctx *c, int n, const double *pr_x, const double *pr_y, const int32_t *t_ns,
                       const uint8_t *noise, int w, int h, int scale, int x_sh, int y_sh, float *out_img,
                       double *out7, float *out_gx, float *out_gy) {
    if (!c || n < 0 || (n > 0 && (!pr_x || !pr_y || !t_ns))) return fail(BF_ERR_ARG, "stage: bad arguments");
    if (scale != 1 && scale != 3 && scale != 5) return fail(BF_ERR_ARG, "scale %d unsupported", scale);
    const int rows = w + scale, cols = h + scale;
    if (w < 0 || h < 0 || rows > c->max_scale * c->res_x || cols > c->max_scale * c->res_y)
        return fail(BF_ERR_ARG, "image %dx%d exceeds the context's capacity", rows, cols);
    CU(cudaSetDevice(c->device));
    int rc = configure(c, 1);
    if (rc != BF_OK) return rc;
    const size_t P = (size_t)rows * cols;
    const int grid = c->sms;
    // scratch layout: pr_x, pr_y (f64) | t (i32) | noise (u8) | partials | out7 | img | gx | gy
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_px = take((size_t)n * 8), o_py = take((size_t)n * 8), o_t = take((size_t)n * 4), o_nz = take((size_t)n);
    const size_t o_part = take((size_t)grid * BF_NSUMS * 8), o_out7 = take(64);
    const size_t o_img = take(P * 4), o_gx = take(P * 4), o_gy = take(P * 4);
    if ((rc = stage_alloc(c, off)) != BF_OK) return rc;
    unsigned char *base = (unsigned char *)c->d_stage;
    if (n > 0) {
        CU(cudaMemcpyAsync(base + o_px, pr_x, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(base + o_py, pr_y, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(base + o_t, t_ns, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        if (noise) CU(cudaMemcpyAsync(base + o_nz, noise, (size_t)n, cudaMemcpyHostToDevice, c->stream));
    }
    StageParams S;
    S.n = n;
    S.pr_x = (const double *)(base + o_px); S.pr_y = (const double *)(base + o_py);
    S.t = (const int *)(base + o_t); S.noise = noise ? (base + o_nz) : nullptr;
    memset(&S.g, 0, sizeof S.g);
    S.g.w = w; S.g.h = h; S.g.rows = rows; S.g.cols = cols; S.g.scale = scale; S.g.half = scale / 2;
    S.g.x_sh = x_sh; S.g.y_sh = y_sh; S.g.x_shift = x_sh; S.g.y_shift = y_sh;
    int32_t t_min = 0, t_max = 0;
    if (n > 0) { t_min = *std::min_element(t_ns, t_ns + n); t_max = *std::max_element(t_ns, t_ns + n); }
    bf_make_pack(S.pk, n, t_min, t_max);
    S.img = c->d_images; S.pitch = c->pitch;
    S.flags = c->d_flags;
    if ((rc = next_tag_base(c, &S.tag)) != BF_OK) return rc;
    S.tag += 1;
    S.partials = (double *)(base + o_part);
    S.out_img = (float *)(base + o_img); S.out_gx = (float *)(base + o_gx); S.out_gy = (float *)(base + o_gy);
    S.out7 = (double *)(base + o_out7);
    const int sb = std::max(1, std::min(4 * c->sms, (n + 255) / 256));
    // the image pass only visits live cells: everything else of the outputs is zero
    CU(cudaMemsetAsync(base + o_img, 0, P * 4, c->stream));
    CU(cudaMemsetAsync(base + o_gx, 0, P * 4, c->stream));
    CU(cudaMemsetAsync(base + o_gy, 0, P * 4, c->stream));
    auto splat = [&](int clear) {
        switch (scale) {
            case 1: bf_stage_splat_kernel<0><<<sb, 256, 0, c->stream>>>(S, clear); break;
            case 3: bf_stage_splat_kernel<1><<<sb, 256, 0, c->stream>>>(S, clear); break;
            default: bf_stage_splat_kernel<2><<<sb, 256, 0, c->stream>>>(S, clear); break;
        }
        c->launches++;
    };
    if (n > 0) splat(0);
    switch (scale) {
        case 1: bf_stage_image_kernel<0><<<grid, BF_NT, smem_bytes(), c->stream>>>(S); break;
        case 3: bf_stage_image_kernel<1><<<grid, BF_NT, smem_bytes(), c->stream>>>(S); break;
        default: bf_stage_image_kernel<2><<<grid, BF_NT, smem_bytes(), c->stream>>>(S); break;
    }
    bf_stage_finish_kernel<<<1, 32, 0, c->stream>>>(S, grid);
    c->launches += 2;
    if (n > 0) splat(1);   // restore the all-zero image
    CU(cudaGetLastError());
    if (out_img) CU(cudaMemcpyAsync(out_img, S.out_img, P * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out_gx) CU(cudaMemcpyAsync(out_gx, S.out_gx, P * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out_gy) CU(cudaMemcpyAsync(out_gy, S.out_gy, P * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out7) CU(cudaMemcpyAsync(out7, S.out7, 7 * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return BF_OK;
}

int bf_time_img(bf_ctx *c, int n, const double *pr_x, const double *pr_y, const int32_t *t_ns, const uint8_t *noise,
                int w, int h, int scale, int x_sh, int y_sh, float *out) {
    if (!out) return fail(BF_ERR_ARG, "bf_time_img: null output");
    return stage_image(c, n, pr_x, pr_y, t_ns, noise, w, h, scale, x_sh, y_sh, out, nullptr, nullptr, nullptr);
}

int bf_fast_model(bf_ctx *c, int n, const double *pr_x, const double *pr_y, const int32_t *t_ns, const uint8_t *noise,
                  int w, int h, int scale, int x_sh, int y_sh, double *out7, float *gx, float *gy) {
    if (!out7) return fail(BF_ERR_ARG, "bf_fast_model: null output");
    return stage_image(c, n, pr_x, pr_y, t_ns, noise, w, h, scale, x_sh, y_sh, nullptr, out7, gx, gy);
}

int bf_model_from_image(bf_ctx *c, int rows, int cols, const float *img, double *out7, float *gx, float *gy) {
    if (!c || rows <= 0 || cols <= 0 || !img) return fail(BF_ERR_ARG, "bf_model_from_image: bad arguments");
    CU(cudaSetDevice(c->device));
    int rc = configure(c, 1);
    if (rc != BF_OK) return rc;
    const size_t P = (size_t)rows * cols;
    const int grid = c->sms;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_img = take(P * 4), o_gx = take(P * 4), o_gy = take(P * 4);
    const size_t o_part = take((size_t)grid * BF_NSUMS * 8), o_out7 = take(64);
    if ((rc = stage_alloc(c, off)) != BF_OK) return rc;
    unsigned char *base = (unsigned char *)c->d_stage;
    CU(cudaMemcpyAsync(base + o_img, img, P * 4, cudaMemcpyHostToDevice, c->stream));
    bf_dense_model_kernel<<<grid, 256, 0, c->stream>>>(rows, cols, (const float *)(base + o_img), gx ? (float *)(base + o_gx) : nullptr,
                                                       gy ? (float *)(base + o_gy) : nullptr, (double *)(base + o_part));
    StageParams S;
    memset(&S, 0, sizeof S);
    S.g.rows = rows; S.g.cols = cols;
    S.partials = (double *)(base + o_part);
    S.out7 = (double *)(base + o_out7);
    bf_stage_finish_kernel<<<1, 32, 0, c->stream>>>(S, grid);
    c->launches += 2;
    CU(cudaGetLastError());
    if (gx) CU(cudaMemcpyAsync(gx, base + o_gx, P * 4, cudaMemcpyDeviceToHost, c->stream));
    if (gy) CU(cudaMemcpyAsync(gy, base + o_gy, P * 4, cudaMemcpyDeviceToHost, c->stream));
    if (out7) CU(cudaMemcpyAsync(out7, S.out7, 7 * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return BF_OK;
}

int bf_project(bf_ctx *c, int n, const uint16_t *fr_x, const uint16_t *fr_y, const int32_t *t_ns, double *pr_x,
               double *pr_y, double *nx, double *ny, double dnx, double dny, double cx, double cy, double div,
               double crl) {
    if (!c || n < 0 || (n > 0 && (!fr_x || !fr_y || !t_ns || !pr_x || !pr_y))) return fail(BF_ERR_ARG, "bf_project: bad arguments");
    if (n == 0) return BF_OK;
    CU(cudaSetDevice(c->device));
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_fx = take((size_t)n * 2), o_fy = take((size_t)n * 2), o_t = take((size_t)n * 4);
    const size_t o_px = take((size_t)n * 8), o_py = take((size_t)n * 8), o_nx = take((size_t)n * 8), o_ny = take((size_t)n * 8);
    int rc;
    if ((rc = stage_alloc(c, off)) != BF_OK) return rc;
    unsigned char *base = (unsigned char *)c->d_stage;
    CU(cudaMemcpyAsync(base + o_fx, fr_x, (size_t)n * 2, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(base + o_fy, fr_y, (size_t)n * 2, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(base + o_t, t_ns, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(base + o_px, pr_x, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CU(cudaMemcpyAsync(base + o_py, pr_y, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    BfProj q;
    bf_make_proj(q, dnx, dny, cx, cy, div, crl);   // cos/sin on the host (glibc), as the reference
    const int sb = std::max(1, std::min(4 * c->sms, (n + 255) / 256));
    bf_stage_project_kernel<<<sb, 256, 0, c->stream>>>(n, (const unsigned short *)(base + o_fx),
                                                        (const unsigned short *)(base + o_fy), (const int *)(base + o_t),
                                                        (double *)(base + o_px), (double *)(base + o_py),
                                                        (double *)(base + o_nx), (double *)(base + o_ny), q);
    c->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(pr_x, base + o_px, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(pr_y, base + o_py, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (nx) CU(cudaMemcpyAsync(nx, base + o_nx, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    if (ny) CU(cudaMemcpyAsync(ny, base + o_ny, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return BF_OK;
}

int bf_projection_img(bf_ctx *c, int n, const double *pr_x, const double *pr_y, const uint8_t *noise, int scale, uint8_t *out,
                      double *nz_avg) {
    if (!c || n < 0 || (n > 0 && (!pr_x || !pr_y)) || !out) return fail(BF_ERR_ARG, "bf_projection_img: bad arguments");
    if (scale != 1 && scale != 3 && scale != 5) return fail(BF_ERR_ARG, "scale %d unsupported (1, 3 or 5)", scale);
    CU(cudaSetDevice(c->device));
    const int rows = c->res_x * scale, cols = c->res_y * scale;
    const size_t P = (size_t)rows * cols;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_px = take((size_t)n * 8), o_py = take((size_t)n * 8), o_nz = take((size_t)n), o_cnt = take(P * 4);
    const size_t o_a = take(P), o_b = take(P), o_acc = take(16), o_avg = take(8);
    int rc;
    if ((rc = stage_alloc(c, off)) != BF_OK) return rc;
    unsigned char *base = (unsigned char *)c->d_stage;
    if (n > 0) {
        CU(cudaMemcpyAsync(base + o_px, pr_x, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(base + o_py, pr_y, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        if (noise) CU(cudaMemcpyAsync(base + o_nz, noise, (size_t)n, cudaMemcpyHostToDevice, c->stream));
    }
    CU(cudaMemsetAsync(base + o_cnt, 0, P * 4, c->stream));
    CU(cudaMemsetAsync(base + o_acc, 0, 16 + 256, c->stream));
    const int gb = std::max(1, std::min(8 * c->sms, (int)((P + 255) / 256)));
    if (n > 0)
        bf_proj_splat_kernel<<<std::max(1, std::min(4 * c->sms, (n + 255) / 256)), 256, 0, c->stream>>>(
            n, (const double *)(base + o_px), (const double *)(base + o_py), noise ? base + o_nz : nullptr, scale, c->res_x, c->res_y,
            (unsigned *)(base + o_cnt), cols);
    bf_proj_box_kernel<<<gb, 256, 0, c->stream>>>((const unsigned *)(base + o_cnt), rows, cols, scale / 2, base + o_a);
    bf_proj_blur_kernel<<<gb, 256, 0, c->stream>>>(base + o_a, rows, cols, scale, base + o_b, (unsigned long long *)(base + o_acc));
    bf_proj_scale_kernel<<<gb, 256, 0, c->stream>>>(base + o_b, (long long)P, (const unsigned long long *)(base + o_acc), (double *)(base + o_avg));
    c->launches += n > 0 ? 4 : 3;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, base + o_b, P, cudaMemcpyDeviceToHost, c->stream));
    if (nz_avg) CU(cudaMemcpyAsync(nz_avg, base + o_avg, 8, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return BF_OK;
}

int bf_color_time_img(bf_ctx *c, int n, const double *pr_x, const double *pr_y, const int32_t *t_ns, const uint8_t *noise, int scale,
                      uint8_t *out_bgr) {
    if (!c || n < 0 || (n > 0 && (!pr_x || !pr_y || !t_ns)) || !out_bgr) return fail(BF_ERR_ARG, "bf_color_time_img: bad arguments");
    if (scale != 1 && scale != 3 && scale != 5) return fail(BF_ERR_ARG, "scale %d unsupported (1, 3 or 5)", scale);
    CU(cudaSetDevice(c->device));
    const int wx = scale * c->res_x, wy = scale * c->res_y, rows = wx + scale, cols = wy + scale;
    const size_t P = (size_t)rows * cols;
    int32_t t_min = INT32_MAX, t_max = INT32_MIN;                                    // event_file.h:663-666
    for (int i = 0; i < n; ++i) { t_min = std::min(t_min, t_ns[i]); t_max = std::max(t_max, t_ns[i]); }
    if (n == 0) { t_min = 0; t_max = 1; }
    const double x_shift = -(double)((c->res_x / 2) * scale) + (double)wx / 2.0;      // :690-691
    const double y_shift = -(double)((c->res_y / 2) * scale) + (double)wy / 2.0;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_px = take((size_t)n * 8), o_py = take((size_t)n * 8), o_t = take((size_t)n * 4), o_nz = take((size_t)n);
    const size_t o_c = take(P * 8), o_s = take(P * 8), o_n = take(P * 4), o_out = take(P * 3);
    int rc;
    if ((rc = stage_alloc(c, off)) != BF_OK) return rc;
    unsigned char *base = (unsigned char *)c->d_stage;
    if (n > 0) {
        CU(cudaMemcpyAsync(base + o_px, pr_x, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(base + o_py, pr_y, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
        CU(cudaMemcpyAsync(base + o_t, t_ns, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
        if (noise) CU(cudaMemcpyAsync(base + o_nz, noise, (size_t)n, cudaMemcpyHostToDevice, c->stream));
    }
    CU(cudaMemsetAsync(base + o_c, 0, o_out - o_c, c->stream));
    if (n > 0)
        bf_color_splat_kernel<<<std::max(1, std::min(4 * c->sms, (n + 255) / 256)), 256, 0, c->stream>>>(
            n, (const double *)(base + o_px), (const double *)(base + o_py), (const int *)(base + o_t), noise ? base + o_nz : nullptr, scale,
            wx, wy, x_shift, y_shift, t_min, t_max, (double *)(base + o_c), (double *)(base + o_s), (unsigned *)(base + o_n), cols);
    bf_color_finish_kernel<<<std::max(1, std::min(8 * c->sms, (int)((P + 255) / 256))), 256, 0, c->stream>>>(
        (const double *)(base + o_c), (const double *)(base + o_s), (const unsigned *)(base + o_n), (long long)P, base + o_out);
    c->launches += n > 0 ? 2 : 1;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out_bgr, base + o_out, P * 3, cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    return BF_OK;
}

// ---- device-resident slice ring ------------------------------------------------------------------------------
struct bf_ring {
    bf_ctx *c = nullptr;
    long long cap = 0;
    int max_pending = 0;
    bf_ring_event *d_ring = nullptr;
    bf_ring_event *h_stage = nullptr;   // pinned staging ring, `stage_cap` entries
    long long stage_cap = 0, stage_pos = 0;
    long long res_off = 0; int res_n = 0;   // the open reservation (bf_ring_reserve), res_n = 0: none
    struct Push { cudaEvent_t ev; long long lo, hi; bool open; };
    std::vector<Push> pushes;           // the last copies issued from the staging ring
    size_t push_next = 0;
    long long pushed = 0;
    SliceDesc *d_desc = nullptr;        // [max_pending]
    bf_slice_result *d_res = nullptr;   // [max_pending + 1]: slot max_pending is the all-zero record (last_model of a fresh DVS_flow)
    bf_slice_result *h_res = nullptr;   // pinned [max_pending]
    int next_ticket = 0;
    int fetched_end = 0;                // tickets below this are in h_res
    int prev_slot = -1;
    long long prev_lo = 0, prev_hi = 0;
};

bf_ring *bf_ring_create(bf_ctx *c, long long capacity, int max_pending) {
    if (!c || capacity <= 0 || capacity + 2 > c->max_events || max_pending < 2) {
        fail(BF_ERR_ARG, "bf_ring_create: capacity must be positive and below the context's max_events, max_pending >= 2");
        return nullptr;
    }
    if (cudaSetDevice(c->device) != cudaSuccess) { fail(BF_ERR_CUDA, "cudaSetDevice failed"); return nullptr; }
    bf_ring *r = new bf_ring();
    r->c = c; r->cap = capacity; r->max_pending = max_pending;
    r->stage_cap = std::max<long long>(2 * capacity, 1 << 18);
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; return x == cudaSuccess; };
    ok(cudaMalloc(&r->d_ring, (size_t)capacity * sizeof(bf_ring_event)));
    ok(cudaMemset(r->d_ring, 0, (size_t)capacity * sizeof(bf_ring_event)));
    ok(cudaMallocHost(&r->h_stage, (size_t)r->stage_cap * sizeof(bf_ring_event)));
    ok(cudaMalloc(&r->d_desc, (size_t)max_pending * sizeof(SliceDesc)));
    ok(cudaMalloc(&r->d_res, (size_t)(max_pending + 1) * sizeof(bf_slice_result)));
    ok(cudaMemset(r->d_res, 0, (size_t)(max_pending + 1) * sizeof(bf_slice_result)));
    ok(cudaMallocHost(&r->h_res, (size_t)max_pending * sizeof(bf_slice_result)));
    { cudaFuncAttributes fa; ok(cudaFuncGetAttributes(&fa, bf_ring_build_kernel)); }   // (loads the kernel now, not inside the first slice)
    r->pushes.resize(16);
    for (auto &p : r->pushes) { p.ev = nullptr; p.lo = p.hi = 0; p.open = false; ok(cudaEventCreateWithFlags(&p.ev, cudaEventDisableTiming)); }
    if (e != cudaSuccess) {
        fail(BF_ERR_CUDA, "bf_ring_create failed: %s", cudaGetErrorString(e));
        bf_ring_destroy(r);
        return nullptr;
    }
    c->rings.push_back(r);
    return r;
}

void bf_ring_destroy(bf_ring *r) {
    if (!r) return;
    cudaSetDevice(r->c->device);
    cudaStreamSynchronize(r->c->stream);
    r->c->rings.erase(std::remove(r->c->rings.begin(), r->c->rings.end(), r), r->c->rings.end());
    cudaFree(r->d_ring); cudaFreeHost(r->h_stage); cudaFree(r->d_desc); cudaFree(r->d_res); cudaFreeHost(r->h_res);
    for (auto &p : r->pushes) if (p.ev) cudaEventDestroy(p.ev);
    delete r;
}

long long bf_ring_pushed(bf_ring *r) { return r ? r->pushed : 0; }

// Waits until no earlier copy still reads staging entries [lo, hi).
static int ring_stage_wait(bf_ring *r, long long lo, long long hi) {
    for (auto &p : r->pushes)
        if (p.open && p.lo < hi && lo < p.hi) { CU(cudaEventSynchronize(p.ev)); p.open = false; }
    return BF_OK;
}

// Checks staging entries [s_off, s_off + n) against the sensor and enqueues their copy to the device ring (the part of
// the device ring they land in may wrap: two copies then); one tracking entry for the whole range.
static int ring_stage_send(bf_ring *r, long long s_off, int n) {
    bf_ctx *c = r->c;
    const bf_ring_event *src = r->h_stage + s_off;
    const unsigned rx = (unsigned)c->res_x, ry = (unsigned)c->res_y;
    unsigned bad = 0;
    for (int i = 0; i < n; ++i) bad |= (unsigned)(src[i].fr_x >= rx) | (unsigned)((src[i].fr_y & 0x7fffu) >= ry);
    if (bad) return fail(BF_ERR_ARG, "event outside the %dx%d sensor", c->res_x, c->res_y);
    int skip = 0;
    if (n > r->cap) { skip = n - (int)r->cap; r->pushed += skip; }   // only the newest `cap` can ever be used
    for (int done_n = skip; done_n < n;) {
        const long long d_off = r->pushed % r->cap;
        const int piece = (int)std::min<long long>(n - done_n, r->cap - d_off);
        CU(cudaMemcpyAsync(r->d_ring + d_off, src + done_n, (size_t)piece * sizeof(bf_ring_event), cudaMemcpyHostToDevice, c->stream));
        r->pushed += piece; done_n += piece;
    }
    bf_ring::Push &p = r->pushes[r->push_next++ % r->pushes.size()];
    if (p.open) CU(cudaEventSynchronize(p.ev));
    p.lo = s_off; p.hi = s_off + n; p.open = true;
    CU(cudaEventRecord(p.ev, c->stream));
    return BF_OK;
}

int bf_ring_reserve(bf_ring *r, int n, bf_ring_event **where) {
    if (!r || !where || n <= 0 || n > r->stage_cap / 2) return fail(BF_ERR_ARG, "bf_ring_reserve: n must be in 1 .. %lld", r ? r->stage_cap / 2 : 0LL);
    CU(cudaSetDevice(r->c->device));
    long long s_off = r->stage_pos % r->stage_cap;
    if (s_off + n > r->stage_cap) { r->stage_pos += r->stage_cap - s_off; s_off = 0; }   // contiguous room: skip the tail
    const int rc = ring_stage_wait(r, s_off, s_off + n);
    if (rc != BF_OK) return rc;
    r->res_off = s_off; r->res_n = n;
    *where = r->h_stage + s_off;
    return BF_OK;
}

int bf_ring_commit(bf_ring *r, int n) {
    if (!r || n < 0 || n > r->res_n) return fail(BF_ERR_ARG, "bf_ring_commit: n exceeds the reservation");
    CU(cudaSetDevice(r->c->device));
    if (n > 0) {
        const int rc = ring_stage_send(r, r->res_off, n);
        if (rc != BF_OK) return rc;
        r->stage_pos += n;
    }
    r->res_n = 0;
    return BF_OK;
}

int bf_ring_push(bf_ring *r, const bf_ring_event *ev, int n) {
    if (!r || n < 0 || (n > 0 && !ev)) return fail(BF_ERR_ARG, "bf_ring_push: bad arguments");
    if (r->res_n > 0) return fail(BF_ERR_STATE, "bf_ring_push: a reservation is open (bf_ring_commit first)");
    const int chunk_max = (int)std::min<long long>(r->stage_cap / 2, 1 << 20);
    if (n > r->cap) { ev += n - r->cap; r->pushed += n - r->cap; n = (int)r->cap; }   // only the newest `cap` can ever be used
    for (int done_n = 0; done_n < n;) {
        const int piece = std::min(n - done_n, chunk_max);
        bf_ring_event *dst = nullptr;
        int rc = bf_ring_reserve(r, piece, &dst);
        if (rc != BF_OK) return rc;
        std::memcpy(dst, ev + done_n, (size_t)piece * sizeof(bf_ring_event));
        if ((rc = bf_ring_commit(r, piece)) != BF_OK) { r->res_n = 0; return rc; }
        done_n += piece;
    }
    return BF_OK;
}

int bf_ring_slice(bf_ring *r, int n, uint64_t slice_start, int scale, int max_iter, int chain) {
    if (!r || n < 0 || n > r->cap || n > r->pushed) return fail(BF_ERR_ARG, "bf_ring_slice: n exceeds the ring's content");
    bf_ctx *c = r->c;
    if (scale != 1 && scale != 3 && scale != 5) return fail(BF_ERR_ARG, "scale %d unsupported (1, 3 or 5)", scale);
    if (scale > c->max_scale) return fail(BF_ERR_ARG, "scale %d exceeds the context's max_scale %d", scale, c->max_scale);
    CU(cudaSetDevice(c->device));
    const int ticket = r->next_ticket++;
    const int slot = ticket % r->max_pending;
    const bf_slice_result *prev = r->prev_slot >= 0 ? r->d_res + r->prev_slot : r->d_res + r->max_pending;
    c->uploaded = c->ran = false;                    // the batch entry points' device copy of the events is overwritten
    const int threads = 256, blocks = std::max(1, std::min(4 * c->sms, (n + threads - 1) / threads));
    bf_ring_build_kernel<<<blocks, threads, 0, c->stream>>>(r->d_ring, r->cap, r->pushed, n, (unsigned long long)slice_start, c->d_events,
                                                            r->d_desc + slot, scale, max_iter, chain ? 1 : 0,
                                                            r->prev_slot >= 0 ? prev : nullptr, r->prev_lo, r->prev_hi);
    CU(cudaGetLastError());
    c->launches += 1;
    LaunchSpec L{c->d_events, r->d_desc + slot, r->d_res + slot, 1, (long long)n, nullptr, prev, 0};
    L.cluster = c->ring_cluster;
    // An independent slice (chain == 0) runs alone on the device: the default launch shape for one slice (16 CTAs, idle
    // groups joining one per iteration up to 64) spends its first iterations growing.  One group of n / 800 CTAs from the
    // start: 0.250 instead of 0.303 ms per 50 k-event slice (profiles/r2z_ring_latency_wide_groups.txt).  Chained slices
    // keep the shape of the host-driven single-slice launch (bf_minimize): the fp64 moments are summed per CTA, another
    // grouping changes their last bits, and a warm-start chain amplifies that at knife-edge exits -- the ring
    // chain stays bit-identical to the chain driven through bf_minimize.
    if (!chain) L.group = (int)std::min<long long>(96, std::max<long long>(16, (n / 800 + 7) / 8 * 8));
    const int rc = launch_spec(c, L);
    if (rc != BF_OK) return rc;
    // (the record stays on the device -- the next slice of a chain reads it there; bf_ring_result fetches on demand)
    r->prev_slot = slot; r->prev_lo = r->pushed - n; r->prev_hi = r->pushed;
    return ticket;
}

int bf_ring_seed(bf_ring *r, const bf_model *model) {
    if (!r || !model) return fail(BF_ERR_ARG, "bf_ring_seed: bad arguments");
    CU(cudaSetDevice(r->c->device));
    // the spare record behind the `max_pending` result slots is what the first slice of a chain reads (all zero after
    // bf_ring_create = set_model(ObjectModel())); the copy is ordered behind the slices already enqueued
    bf_slice_result rec;
    std::memset(&rec, 0, sizeof rec);
    rec.model = *model;
    CU(cudaMemcpyAsync(r->d_res + r->max_pending, &rec, sizeof rec, cudaMemcpyHostToDevice, r->c->stream));   // pageable source: staged before the call returns
    r->prev_slot = -1;
    return BF_OK;
}

int bf_ring_result(bf_ring *r, int ticket, bf_slice_result *out) {
    if (!r || !out || ticket < 0 || ticket >= r->next_ticket || ticket < r->next_ticket - r->max_pending)
        return fail(BF_ERR_ARG, "bf_ring_result: ticket %d is not (or no longer) available", ticket);
    const int slot = ticket % r->max_pending;
    if (ticket >= r->fetched_end) {
        // one copy of the whole (small) record array behind everything enqueued so far, then wait for it
        CU(cudaSetDevice(r->c->device));
        CU(cudaMemcpyAsync(r->h_res, r->d_res, (size_t)r->max_pending * sizeof(bf_slice_result), cudaMemcpyDeviceToHost, r->c->stream));
        CU(cudaStreamSynchronize(r->c->stream));
        r->fetched_end = r->next_ticket;
    }
    *out = r->h_res[slot];
    return BF_OK;
}

int bf_ring_sync(bf_ring *r) {
    if (!r) return fail(BF_ERR_ARG, "null ring");
    CU(cudaStreamSynchronize(r->c->stream));
    return BF_OK;
}

}  // extern "C"

static_assert(sizeof(bf_model) == 88, "bf_model must mirror ObjectModel's 11 scalars");
static_assert(sizeof(bf_slice_result) == 160, "bf_slice_result layout is part of the ABI");
static_assert(sizeof(bf_event) == 8, "bf_event is the 8-byte compact record");
static_assert(sizeof(GroupWs) == 256, "one 256-byte control record per group");
static_assert(sizeof(SliceDesc) % 8 == 0 && sizeof(BfGeom) % 8 == 0 && sizeof(BfPack) % 8 == 0 && sizeof(BfOpt) % 8 == 0 &&
              sizeof(BfProj) % 8 == 0, "JoinRecord members are copied as 8-byte words");
