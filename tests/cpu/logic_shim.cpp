// Host build of better_flow_b200/csrc/bf_logic.h (the scalar logic the GPU runs with one thread per
// CTA) behind a tiny C interface, so the CPU test-suite can exercise it without a GPU.
#include "../../better_flow_b200/csrc/bf_logic.h"

extern "C" {

void lg_geom(int x_min, int x_max, int y_min, int y_max, int scale, int *ints, double *dbls) {
    BfGeom g;
    bf_make_geom(g, x_min, x_max, y_min, y_max, scale);
    ints[0] = g.w; ints[1] = g.h; ints[2] = g.rows; ints[3] = g.cols; ints[4] = g.x_sh; ints[5] = g.y_sh;
    dbls[0] = g.x_shift; dbls[1] = g.y_shift;
}

int lg_guard_tiny(int x_min, int x_max, int y_min, int y_max, int scale, int res_x, int res_y) {
    BfGeom g;
    bf_make_geom(g, x_min, x_max, y_min, y_max, scale);
    return bf_guard_tiny(g, res_x, res_y) ? 1 : 0;
}

void lg_pack(int n, int t_min, int t_max, int *cnt_shift, int *q) {
    BfPack p;
    bf_make_pack(p, n, t_min, t_max);
    *cnt_shift = p.cnt_shift;
    *q = p.q;
}

void lg_pack_full(int n, int t_min, int t_max, int *out4) {
    BfPack p;
    bf_make_pack(p, n, t_min, t_max);
    out4[0] = p.cnt_shift; out4[1] = p.q; out4[2] = p.t_min; out4[3] = p.fast;
}

// Accumulate n packed values (as the device's 64-bit atomic add would) and unpack the mean.
float lg_accumulate(int n_total, int t_min, int t_max, int k, const int *t) {
    BfPack p;
    bf_make_pack(p, n_total, t_min, t_max);
    unsigned long long acc = 0;
    for (int i = 0; i < k; ++i) acc += bf_pack_value(p, t[i]);
    return bf_unpack_avg(p, acc);
}

// Precondition check of the packed / staged batch entry points (8-byte records: u16 fr_x, u16 fr_y | noise bit, i32 t).
int lg_events_in_sensor(const void *events, long long n, int res_x, int res_y) {
    return bf_events_in_sensor(static_cast<const bf_event *>(events), n, res_x, res_y) ? 1 : 0;
}

// Replay of OptimizerRolling::run's control flow on a recorded sequence of per-iteration sums.
// sums: steps x 9 doubles.  Returns the number of steps consumed; out: model(11) + dividers(4) + rc.
int lg_replay(int steps, const double *sums, int x_min, int x_max, int y_min, int y_max, int scale, int i0, int j0,
              int max_iter, int iter_cap, double *out) {
    BfGeom g;
    bf_make_geom(g, x_min, x_max, y_min, y_max, scale);
    BfOpt o;
    bf_opt_init(o, nullptr);
    BfProj next;
    int k = 0;
    for (; k < steps; ++k) {
        BfSums s;
        const double *v = sums + 9 * k;
        s.cnt = v[0]; s.si = v[1]; s.sj = v[2]; s.sgx = v[3]; s.sgy = v[4];
        s.sigx = v[5]; s.sjgx = v[6]; s.sigy = v[7]; s.sjgy = v[8];
        if (!bf_opt_advance(o, g, s, i0, j0, max_iter, iter_cap, next)) { ++k; break; }
    }
    out[0] = o.m.cx; out[1] = o.m.cy; out[2] = o.m.dx; out[3] = o.m.dy; out[4] = o.m.rot; out[5] = o.m.div;
    out[6] = o.m.cnt; out[7] = o.m.total_dx; out[8] = o.m.total_dy; out[9] = o.m.total_rot; out[10] = o.m.total_div;
    out[11] = o.x_div; out[12] = o.y_div; out[13] = o.rot_div; out[14] = o.div_div; out[15] = o.rc; out[16] = o.iters;
    return k;
}

}  // extern "C"
