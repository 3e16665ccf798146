/* TEST INFRASTRUCTURE ONLY -- never linked into, loaded by, or called from the product.
 *
 * CPU restatement ("port") of the reference's motion-compensation hot path, in plain C over
 * SoA arrays.  Each function cites the reference code it follows (paths relative to
 * /root/reference/better_flow_core/).  Parity of this restatement is PINNED: tests/test_oracle.py
 * checks it (a) bit-for-bit against the reference's own sources compiled into oracle/_ref/
 * whenever that library is present, and (b) against golden vectors minted from that library
 * and committed under tests/golden/ (the reference itself ships no tests or fixtures).
 *
 * Build: gcc -O2 -ffp-contract=off, no -march=native, no -ffast-math -- the same x86-64 SSE2
 * rounding (no FMA contraction, FLT_EVAL_METHOD 0) the reference's -O3 build has.
 *
 * accum_mode selects how the time image accumulates (everything else is identical):
 *   0  reference-faithful: f32 += f64 per event, in event order   (accel_lib.h:162-163)
 *   1  exact: per-pixel integer sums of t_ns and counts, converted once -- this is what the
 *      CUDA path computes (order-independent, so it is deterministic under atomics); tests
 *      compare the CUDA path bit-for-bit against mode 1 and within tolerance against mode 0.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double cx, cy, dx, dy, rot, div;
    double cnt; /* uint in the reference (object_model.h:11); a double here so the struct is 11 doubles */
    double total_dx, total_dy, total_rot, total_div;
} bfo_model;

typedef struct {
    int x_min, x_max, y_min, y_max;       /* bbox over fr_x, fr_y */
    int wsize_x, wsize_y;                 /* metric_wsizex/y = scale * extent */
    int img_rows, img_cols;               /* scale_img_x/y = wsize + scale */
    double x_shift, y_shift;
} bfo_setup;

/* OptimizerRolling::set_cloud + set_scale (optimizer_rolling.h:248-283).
 * bbox min starts at RES_X/RES_Y and max at 0 (:252-253); integer /2 on the extent and on
 * scale (:279-282). */
void bfo_setup_slice(int n, const uint16_t *fr_x, const uint16_t *fr_y, int res_x, int res_y,
                     int scale, bfo_setup *s) {
    int x_min = res_x, y_min = res_y, x_max = 0, y_max = 0;
    for (int i = 0; i < n; ++i) {
        if ((int)fr_x[i] > x_max) x_max = fr_x[i];
        if ((int)fr_y[i] > y_max) y_max = fr_y[i];
        if ((int)fr_x[i] < x_min) x_min = fr_x[i];
        if ((int)fr_y[i] < y_min) y_min = fr_y[i];
    }
    s->x_min = x_min; s->x_max = x_max; s->y_min = y_min; s->y_max = y_max;
    s->wsize_x = scale * (x_max - x_min);
    s->wsize_y = scale * (y_max - y_min);
    s->img_rows = s->wsize_x + scale;
    s->img_cols = s->wsize_y + scale;
    s->x_shift = -(double)((x_max - x_min) / 2 + x_min) * (double)scale + (double)s->wsize_x / 2.0 + scale / 2;
    s->y_shift = -(double)((y_max - y_min) / 2 + y_min) * (double)scale + (double)s->wsize_y / 2.0 + scale / 2;
}

/* AccelLib::get_time_img_cpu (accel_lib.h:147-178).  out: (w+scale) rows x (h+scale) cols. */
void bfo_time_img(int n, const double *pr_x, const double *pr_y, const int64_t *t,
                  const uint8_t *noise, int w, int h, int scale, int x_sh, int y_sh,
                  int accum_mode, float *out) {
    const int rows = w + scale, cols = h + scale;
    const size_t P = (size_t)rows * cols;
    float *avg = out;
    memset(avg, 0, P * sizeof(float));
    float *cntf = NULL;
    int64_t *sum_i = NULL;
    int64_t *cnt_i = NULL;
    if (accum_mode == 0) cntf = (float *)calloc(P, sizeof(float));
    else { sum_i = (int64_t *)calloc(P, sizeof(int64_t)); cnt_i = (int64_t *)calloc(P, sizeof(int64_t)); }

    for (int i = 0; i < n; ++i) {
        if (noise && noise[i]) continue;                              /* :152 */
        int x = (int)(pr_x[i] * scale + x_sh);                        /* :154 */
        int y = (int)(pr_y[i] * scale + y_sh);                        /* :155 */
        if ((x >= w + scale / 2) || (x < scale / 2) || (y >= h + scale / 2) || (y < scale / 2))
            continue;                                                 /* :157-158 */
        for (int jx = x - scale / 2; jx <= x + scale / 2; ++jx) {
            for (int jy = y - scale / 2; jy <= y + scale / 2; ++jy) {
                const size_t k = (size_t)jx * cols + jy;
                if (accum_mode == 0) {
                    avg[k] += (double)t[i] / 1000000000.0;            /* :162  f32 <- f32 + f64 */
                    cntf[k] += 1;                                     /* :163 */
                } else {
                    sum_i[k] += t[i];
                    cnt_i[k] += 1;
                }
            }
        }
    }
    if (accum_mode == 0) {
        for (size_t k = 0; k < P; ++k) {                              /* :168-175 */
            if (cntf[k] < 1) continue;
            avg[k] /= cntf[k];
        }
        free(cntf);
    } else {
        /* exact sums -> one rounding to f32 for the sum, then the reference's f32 divide */
        for (size_t k = 0; k < P; ++k) {
            if (cnt_i[k] < 1) continue;
            float s = (float)((double)sum_i[k] / 1000000000.0);
            avg[k] = s / (float)cnt_i[k];
        }
        free(sum_i); free(cnt_i);
    }
}

/* AccelLib::sobel_point (accel_lib.h:545-615), live part only (:594-614): taps visited
 * column-major; returns 0 as soon as a tap is <= 1e-6. (i = column, j = row, as called
 * from Sobel_cpu :536.) */
static int sobel_point(const float *img, int cols, int i, int j, float *dx, float *dy) {
    static const int sharr_x[9] = {3, 0, -3, 10, 0, -10, 3, 0, -3};
    static const int sharr_y[9] = {3, 10, 3, 0, 0, 0, -3, -10, -3};
    int idx = 0;
    float ax = 0, ay = 0;
    *dx = *dy = 0;
    for (int k = 0; k < 3; ++k) {
        for (int l = 0; l < 3; ++l) {
            float val = img[(size_t)(l + j - 1) * cols + (k + i - 1)];
            if (val <= 0.000001) return 0;
            ax += val * sharr_x[idx];
            ay += val * sharr_y[idx];
            idx++;
        }
    }
    *dx = ax; *dy = ay;
    return 1;
}

/* AccelLib::Sobel_cpu (accel_lib.h:513-543): interior pixels only, centre must be > 1e-6. */
void bfo_scharr(int rows, int cols, const float *img, float *gx, float *gy) {
    memset(gx, 0, (size_t)rows * cols * sizeof(float));
    memset(gy, 0, (size_t)rows * cols * sizeof(float));
    for (int i = 1; i < rows - 1; ++i) {
        for (int j = 1; j < cols - 1; ++j) {
            if (img[(size_t)i * cols + j] <= 0.000001) continue;
            float dx = 0, dy = 0;
            if (sobel_point(img, cols, j, i, &dx, &dy)) {
                gx[(size_t)i * cols + j] = dx;
                gy[(size_t)i * cols + j] = dy;
            }
        }
    }
}

/* ObjectModel::center_of_mass (object_model.cpp:103-126) */
static void center_of_mass(bfo_model *m, int rows, int cols, const float *img) {
    double cx = 0, cy = 0;
    unsigned cnt = 0;
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j)
            if (img[(size_t)i * cols + j] > 0.000001) { cx += i; cy += j; cnt++; }
    m->cx = cx / (double)cnt;   /* cnt == 0 -> NaN, as in an NDEBUG reference build (:122) */
    m->cy = cy / (double)cnt;
    m->cnt = cnt;
}

/* ObjectModel::compute(Mat&) (object_model.cpp:4-39) */
static void model_compute(bfo_model *m, int rows, int cols, const float *img, float *gx, float *gy) {
    bfo_scharr(rows, cols, img, gx, gy);
    double dx = 0, dy = 0, rot = 0, div = 0;
    unsigned cnt = 0;
    for (int i = 0; i < rows; ++i) {
        for (int j = 0; j < cols; ++j) {
            const size_t k = (size_t)i * cols + j;
            if (img[k] > 0.000001) {
                double rx = (double)i - m->cx, ry = (double)j - m->cy;
                double g_x = gx[k], g_y = gy[k];
                dx += gx[k];
                dy += gy[k];
                rot += rx * g_y - ry * g_x;      /* Point2d::cross */
                div += rx * g_x + ry * g_y;      /* Point2d::ddot  */
                cnt++;
            }
        }
    }
    m->rot = rot / (double)cnt;
    m->div = div / (double)cnt;
    m->dx = dx / (double)cnt;
    m->dy = dy / (double)cnt;
    m->cnt = cnt;
}

/* ObjectModel::update(Mat) (object_model.h:31-34).  out7: cx,cy,dx,dy,rot,div,cnt.
 * gx/gy nullable (scratch allocated internally when NULL). */
void bfo_model_update(int rows, int cols, const float *img, double *out7, float *gx, float *gy) {
    bfo_model m;
    memset(&m, 0, sizeof m);
    float *tx = gx ? gx : (float *)malloc((size_t)rows * cols * sizeof(float));
    float *ty = gy ? gy : (float *)malloc((size_t)rows * cols * sizeof(float));
    center_of_mass(&m, rows, cols, img);
    model_compute(&m, rows, cols, img, tx, ty);
    out7[0] = m.cx; out7[1] = m.cy; out7[2] = m.dx; out7[3] = m.dy; out7[4] = m.rot; out7[5] = m.div;
    out7[6] = m.cnt;
    if (!gx) free(tx);
    if (!gy) free(ty);
}

/* Event::project_4param_reinit + apply_project (event.h:99-110,164-168), over n events. */
void bfo_project(int n, const uint16_t *fr_x, const uint16_t *fr_y, const int64_t *t,
                 double *pr_x, double *pr_y, double *nx, double *ny,
                 double dnx, double dny, double cx, double cy, double div, double crl) {
    const double c = cos(crl), s = sin(crl);
    const double nz = 127; /* NZ, common.h:60 */
    for (int i = 0; i < n; ++i) {
        double rx = pr_x[i] - cx, ry = pr_y[i] - cy;                 /* :100 */
        double qx = c * rx - s * ry;                                 /* :102 */
        double qy = s * rx + c * ry;                                 /* :103 */
        double dn_x = (-qx) * div + (qx - rx);                       /* :105  -r_*div + (r_-r) */
        double dn_y = (-qy) * div + (qy - ry);
        double ex = dn_x + dnx, ey = dn_y + dny;                     /* :107-108 */
        if (nx) nx[i] = ex;
        if (ny) ny[i] = ey;
        float kx = (float)ex / nz;                                   /* :164  f32 <- f64 / f64 */
        float ky = (float)ey / nz;
        float tf = (float)t[i];
        pr_x[i] = (float)fr_x[i] - kx * tf / 10000.0;                /* :167  (f32*f32) -> f64 / f64 */
        pr_y[i] = (float)fr_y[i] - ky * tf / 10000.0;
    }
}

/* Event::compute_uv (event.h:135-142): 1000000000/(T_DIVIDER*10000) is integer 100000. */
void bfo_compute_uv(int n, const double *nx, const double *ny, double *u, double *v) {
    const double k = 127.0 / (double)(1000000000 / (1 * 10000));
    for (int i = 0; i < n; ++i) {
        double len = hypot(nx[i], ny[i]);
        double speed = len / k;
        u[i] = (len == 0) ? 0 : speed * nx[i] / len;
        v[i] = (len == 0) ? 0 : speed * ny[i] / len;
    }
}

typedef struct {
    int n, scale, accum_mode;
    const uint16_t *fr_x, *fr_y;
    const int64_t *t;
    uint8_t *noise;
    double *pr_x, *pr_y, *nx, *ny;
    bfo_setup su;
    bfo_model model;
    float x_div, y_div, rot_div, div_div;
    float *img, *gx, *gy;
} opt_state;

/* OptimizerRolling::iteration_step (optimizer_rolling.h:305-347) */
static void iteration_step(opt_state *o) {
    /* (i) get_time_img: shifts truncate to int through the callee's signature (accel_lib.h:211) */
    bfo_time_img(o->n, o->pr_x, o->pr_y, o->t, o->noise, o->su.wsize_x, o->su.wsize_y, o->scale,
                 (int)o->su.x_shift, (int)o->su.y_shift, o->accum_mode, o->img);
    /* (ii) fast_model -> ObjectModel::update (accel_lib.h:337-341) */
    center_of_mass(&o->model, o->su.img_rows, o->su.img_cols, o->img);
    model_compute(&o->model, o->su.img_rows, o->su.img_cols, o->img, o->gx, o->gy);
    /* (iii) update_accumulators(rot_divider, div_divider, x_divider, y_divider) (object_model.h:48-53) */
    o->model.total_rot += o->model.rot / o->rot_div;
    o->model.total_div += o->model.div / o->div_div;
    o->model.total_dx += o->model.dx / o->x_div;
    o->model.total_dy += o->model.dy / o->y_div;
    /* (iv) centre back in sensor units with the untruncated shifts (:330-331) */
    double cx = (o->model.cx - o->su.x_shift) / o->scale;
    double cy = (o->model.cy - o->su.y_shift) / o->scale;
    /* (v) re-warp every event from its current position (:340-344) */
    bfo_project(o->n, o->fr_x, o->fr_y, o->t, o->pr_x, o->pr_y, o->nx, o->ny,
                -o->model.total_dx, -o->model.total_dy, cx, cy, o->model.total_div, -o->model.total_rot);
    o->model.cx = cx;                                                 /* (vi) :345-346 */
    o->model.cy = cy;
}

/* OptimizerRolling: set_cloud -> set_time (already applied: t is local) -> set_maxiter ->
 * [set_model] -> run (optimizer_rolling.h:48-125,236-299), as DVS_flow::recompute drives it
 * (dvs_flow.h:210-224).
 *   noise       nullable in/out per-event flags (set to 1 for every event by the tiny-window guard)
 *   init_model  nullable = --stm-disable
 *   pr_out      nullable, 4*n doubles: pr_x, pr_y, nx, ny after run()
 * Returns run()'s value: 0 optimised, 1 skipped. */
int bfo_minimize(int n, const uint16_t *fr_x, const uint16_t *fr_y, const int64_t *t, uint8_t *noise,
                 int res_x, int res_y, int scale, int max_iter, const double *init_model,
                 int accum_mode, double *out_model, int *out_iters, bfo_setup *out_setup,
                 float *out_div, double *pr_out) {
    opt_state o;
    memset(&o, 0, sizeof o);
    o.n = n; o.scale = scale; o.accum_mode = accum_mode;
    o.fr_x = fr_x; o.fr_y = fr_y; o.t = t; o.noise = noise;
    bfo_setup_slice(n, fr_x, fr_y, res_x, res_y, scale, &o.su);
    double *buf = (double *)malloc(sizeof(double) * 4 * (size_t)(n > 0 ? n : 1));
    o.pr_x = buf; o.pr_y = buf + n; o.nx = buf + 2 * (size_t)n; o.ny = buf + 3 * (size_t)n;
    for (int i = 0; i < n; ++i) {                                    /* Event::reset (event.h:54-59) */
        o.pr_x[i] = fr_x[i]; o.pr_y[i] = fr_y[i]; o.nx[i] = 0; o.ny[i] = 0;
    }
    o.x_div = 1; o.y_div = 1; o.rot_div = 10000; o.div_div = 10000;  /* ctor :45 */
    if (init_model) {                                                /* set_model :289-299 */
        memcpy(&o.model, init_model, sizeof(bfo_model));
        bfo_project(n, fr_x, fr_y, t, o.pr_x, o.pr_y, o.nx, o.ny, -o.model.total_dx, -o.model.total_dy,
                    o.model.cx, o.model.cy, o.model.total_div, -o.model.total_rot);
    }

    int rc = 0;
    int itercount = 0;
    if ((o.su.img_rows < scale * res_x / 15) && (o.su.img_cols < scale * res_y / 15)) {   /* :49-55 */
        if (noise) for (int i = 0; i < n; ++i) noise[i] = 1;
        rc = 1;
    } else if (n < 1000) {                                           /* :57-58 */
        rc = 1;
    } else {
        const size_t P = (size_t)o.su.img_rows * o.su.img_cols;
        o.img = (float *)malloc(P * sizeof(float));
        o.gx = (float *)malloc(P * sizeof(float));
        o.gy = (float *)malloc(P * sizeof(float));
        o.x_div = o.y_div = 1.0f;                                    /* :61-63 */
        o.rot_div = 10000;
        o.div_div = 10000;

        iteration_step(&o);                                          /* :73-74 */
        itercount++;
        while (o.x_div < 32 * 10 || o.y_div < 32 * 10 || o.rot_div < 32 * 1000 || o.div_div < 32 * 1000) {
            if (fabs(o.model.dx / o.x_div) < 1e-5 && fabs(o.model.dy / o.y_div) < 1e-5 &&
                fabs(o.model.rot / o.rot_div) < 1e-4 && fabs(o.model.div / o.div_div) < 1e-1)
                break;                                               /* :81-84 */
            float old_dx = o.model.dx, old_dy = o.model.dy;          /* :86-89 (f32!) */
            float old_rot = o.model.rot, old_div = o.model.div;
            iteration_step(&o);                                      /* :91-92 */
            itercount++;
            if (max_iter > 0 && itercount > max_iter) break;         /* :94-96 */
            if (o.model.dx * old_dx < 0) o.x_div *= 2;               /* :98-101 */
            if (o.model.dy * old_dy < 0) o.y_div *= 2;
            if (o.model.rot * old_rot < 0) o.rot_div *= 2;
            if (o.model.div * old_div < 0) o.div_div *= 2;
        }
        free(o.img); free(o.gx); free(o.gy);
    }
    if (out_model) memcpy(out_model, &o.model, sizeof(bfo_model));
    if (out_iters) *out_iters = itercount;
    if (out_setup) *out_setup = o.su;
    if (out_div) { out_div[0] = o.x_div; out_div[1] = o.y_div; out_div[2] = o.rot_div; out_div[3] = o.div_div; }
    if (pr_out) memcpy(pr_out, buf, sizeof(double) * 4 * (size_t)n);
    free(buf);
    return rc;
}

/* =====================================================================================================
 * OptimizerLocal (src/optimizer_sampler.cpp, include/better_flow/optimizer_sampler.h): the contrast-
 * driven coordinate descent over a global (nx, ny).  Never instantiated by DVS_flow / the CLI (only
 * #included, dvs_flow.h:5), but it is the variant BASELINE.json's north-star prose describes
 * ("scores image ... contrast, and gradient-descends over a global (dx,dy) flow"): SURVEY 8a-18 / 8f-3.
 * Pinned against the reference's own class compiled into oracle/_ref (bf_ref_local) and the golden
 * records minted from it; the Gaussian blur against fixtures from the real cv2.GaussianBlur.
 * ===================================================================================================== */

/* cv::GaussianBlur(img, img, Size(k,k), 0, 0) on CV_8UC1 (optimizer_sampler.cpp:147-149): OpenCV's 8-bit
 * path with sigma <= 0 uses the binomial kernels [1 2 1]/4, [1 4 6 4 1]/16 in fixed point = the exact
 * weighted sum rounded half up; BORDER_REFLECT_101. */
static int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}
void bfo_gaussian_blur_u8(int rows, int cols, int k, uint8_t *img) {
    static const int w1[1] = {1}, w3[3] = {1, 2, 1}, w5[5] = {1, 4, 6, 4, 1};
    const int *w = k == 1 ? w1 : k == 3 ? w3 : w5;
    const int r = k / 2, shift = k == 1 ? 0 : k == 3 ? 4 : 8;
    int *tmp = (int *)malloc(sizeof(int) * (size_t)rows * cols);
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) {
            int s = 0;
            for (int d = -r; d <= r; ++d) s += w[d + r] * img[(size_t)i * cols + reflect101(j + d, cols)];
            tmp[(size_t)i * cols + j] = s;
        }
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) {
            int s = 0;
            for (int d = -r; d <= r; ++d) s += w[d + r] * tmp[(size_t)reflect101(i + d, rows) * cols + j];
            img[(size_t)i * cols + j] = (uint8_t)((s + ((1 << shift) >> 1)) >> shift);
        }
    free(tmp);
}

typedef struct {
    int n, scale;
    const uint16_t *fr_x, *fr_y;
    const int64_t *t;
    int wsize_x, wsize_y, img_rows, img_cols;
    double cpr_x, cpr_y;      /* event_c.pr_x / pr_y: event_c has t = 0, so project() leaves it at float(fr) */
    double *pr_x, *pr_y;
    uint8_t *img;
    int steps;
} local_state;

/* OptimizerLocal::iteration_step (optimizer_sampler.cpp:120-153) + get_event_score (:192-204) */
static double local_step(local_state *L, double nx, double ny) {
    const double nz = 127;
    const int s = L->scale;
    for (int i = 0; i < L->n; ++i) {                                    /* :121  Event::project -> apply_project */
        float kx = (float)nx / nz, ky = (float)ny / nz;                 /* event.h:164-165 */
        float tf = (float)L->t[i];
        L->pr_x[i] = (float)L->fr_x[i] - kx * tf / 10000.0;             /* event.h:167-168 */
        L->pr_y[i] = (float)L->fr_y[i] - ky * tf / 10000.0;
    }
    memset(L->img, 0, (size_t)L->img_rows * L->img_cols);               /* :124 */
    const double x_shift = -L->cpr_x * s + (double)L->wsize_x / 2.0;    /* :126-127 */
    const double y_shift = -L->cpr_y * s + (double)L->wsize_y / 2.0;
    for (int i = 0; i < L->n; ++i) {
        int x = L->pr_x[i] * s + x_shift;                               /* :130-131  f64, truncation */
        int y = L->pr_y[i] * s + y_shift;
        if (x >= L->wsize_x || x < 0 || y >= L->wsize_y || y < 0) continue;   /* :133-134 */
        x += s / 2;                                                     /* :136-137 */
        y += s / 2;
        for (int jx = x - s / 2; jx <= x + s / 2; ++jx)                 /* :139-145  saturating u8 count */
            for (int jy = y - s / 2; jy <= y + s / 2; ++jy) {
                uint8_t *p = L->img + (size_t)jx * L->img_cols + jy;
                if (*p < 255) (*p)++;
            }
    }
    if (s > 1) bfo_gaussian_blur_u8(L->img_rows, L->img_cols, s, L->img);   /* :147-149 */
    L->steps++;
    double nz_avg = 0;                                                  /* :192-204  mean of the non-zero pixels */
    long nz_cnt = 0;
    for (size_t k = 0; k < (size_t)L->img_rows * L->img_cols; ++k) {
        if (L->img[k] == 0) continue;
        nz_cnt++;
        nz_avg += L->img[k];
    }
    return nz_cnt == 0 ? 0 : nz_avg / (double)nz_cnt;
}

/* OptimizerLocal(LinearEventCloud*, scale) + run() (optimizer_sampler.h:41-56, optimizer_sampler.cpp:4-38,90-117).
 *   out10    nx, ny, last_score, dnx, dny, dn_th, metric_wsizex, metric_wsizey, scale_img_x, scale_img_y
 *   out_img  nullable: the image of the last iteration_step;  out_pr nullable: 2n doubles pr_x, pr_y
 * Returns run()'s value: 0 ok, 1 window too small. */
int bfo_local_minimize(int n, const uint16_t *fr_x, const uint16_t *fr_y, const int64_t *t, int res_x, int res_y,
                       int scale, double *out10, int *out_steps, uint8_t *out_img, double *out_pr) {
    /* LinearEventCloud::push_back bbox (datastructures.h:141-148): starts at INT_MAX / INT_MIN */
    int x_min = 2147483647, y_min = 2147483647, x_max = -2147483647 - 1, y_max = -2147483647 - 1;
    for (int i = 0; i < n; ++i) {
        if ((int)fr_x[i] > x_max) x_max = fr_x[i];
        if ((int)fr_y[i] > y_max) y_max = fr_y[i];
        if ((int)fr_x[i] < x_min) x_min = fr_x[i];
        if ((int)fr_y[i] < y_min) y_min = fr_y[i];
    }
    local_state L;
    memset(&L, 0, sizeof L);
    L.n = n; L.scale = scale; L.fr_x = fr_x; L.fr_y = fr_y; L.t = t;
    L.wsize_x = scale * (x_max - x_min);                                /* optimizer_sampler.h:48-49 */
    L.wsize_y = scale * (y_max - y_min);
    L.cpr_x = (float)(unsigned)((x_max - x_min) / 2 + x_min);          /* :51  Event(uint, uint, 0); pr = float(fr) */
    L.cpr_y = (float)(unsigned)((y_max - y_min) / 2 + y_min);
    L.img_rows = L.wsize_x + scale;                                     /* update_fields, optimizer_sampler.cpp:207-214 */
    L.img_cols = L.wsize_y + scale;
    double nx = 0, ny = 0, last_score = 0, dscore = 0;                  /* run(), :5-7 */
    double dnx = 0.01, dny = 0.01;
    const double dn_th = (127 * 1 * 1000.0) / (double)(10ULL * (unsigned long long)scale * 100000000ULL);
    int rc = 0;
    if (n == 0 || ((L.img_rows < scale * res_x / 15) && (L.img_cols < scale * res_y / 15))) rc = 1;   /* :9-13 */
    if (rc == 0) {
        L.pr_x = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        L.pr_y = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
        L.img = (uint8_t *)calloc((size_t)L.img_rows * L.img_cols, 1);
        last_score = local_step(&L, nx, ny);                            /* :16 */
        while (hypot(dnx, dny) > dn_th) {                               /* :20 */
            {                                                           /* compute_new_nx, :90-102 */
                double nx_new = nx + dnx;
                double new_score = local_step(&L, nx_new, ny);
                dscore = new_score - last_score;
                last_score = new_score;
                if (dscore <= 0) dnx = -dnx / 2.0;
                nx = nx_new;
            }
            {                                                           /* compute_new_ny, :105-117 */
                double ny_new = ny + dny;
                double new_score = local_step(&L, nx, ny_new);
                dscore = new_score - last_score;
                last_score = new_score;
                if (dscore <= 0) dny = -dny / 2.0;
                ny = ny_new;
            }
        }
        if (out_img) memcpy(out_img, L.img, (size_t)L.img_rows * L.img_cols);
        if (out_pr) { memcpy(out_pr, L.pr_x, sizeof(double) * n); memcpy(out_pr + n, L.pr_y, sizeof(double) * n); }
        free(L.pr_x); free(L.pr_y); free(L.img);
    }
    if (out10) {
        out10[0] = nx; out10[1] = ny; out10[2] = last_score; out10[3] = dnx; out10[4] = dny; out10[5] = dn_th;
        out10[6] = L.wsize_x; out10[7] = L.wsize_y; out10[8] = L.img_rows; out10[9] = L.img_cols;
    }
    if (out_steps) *out_steps = L.steps;
    return rc;
}
