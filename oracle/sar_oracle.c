/* sar_oracle.c — CPU restatement of the reference render path.  TEST INFRASTRUCTURE ONLY
 * (see sar_oracle.h).  Build: gcc -std=c11 -O2 -ffp-contract=off -fno-fast-math -pthread.
 * Every function cites the reference lines (src/lib.rs @ e571d19) it follows. */
#define _GNU_SOURCE
#include "sar_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- Rust `as` cast semantics: saturating, NaN -> 0 ----------------------- */
static inline uint32_t f64_as_u32(double v)   /* lib.rs:800-802 */
{
    if (!(v > 0.0)) return 0u;                 /* NaN, -x, ±0 */
    if (v >= 4294967295.0) return 4294967295u;
    return (uint32_t)v;                        /* truncates toward zero */
}
static inline uint16_t f64_as_u16(double v)   /* lib.rs:862-866 */
{
    if (!(v > 0.0)) return 0u;
    if (v >= 65535.0) return 65535u;
    return (uint16_t)v;
}
static inline uint16_t f32_as_u16(float v)    /* lib.rs:895 */
{
    if (!(v > 0.0f)) return 0u;
    if (v >= 65535.0f) return 65535u;
    return (uint16_t)v;
}

/* ---- attractors::PolynomialSprott2Degree::next_point, lib.rs:585-620 ------ */
static inline double sum_coefficients(const double m[10], const double c[10])
{
    double sum = 0.;                           /* lib.rs:589 */
    for (int i = 0; i < 10; ++i)
        sum += m[i] * c[i];                    /* lib.rs:596, left to right */
    return sum;
}
void orc_next_point(const double coef[3][10], double p[3])
{
    const double x = p[0], y = p[1], z = p[2];
    const double m[10] = {1., x, x * x, x * y, x * z, y, y * y, y * z, z, z * z}; /* lib.rs:602-613 */
    p[0] = sum_coefficients(m, coef[0]);
    p[1] = sum_coefficients(m, coef[1]);
    p[2] = sum_coefficients(m, coef[2]);
}

/* Attractor kinds of include/sar.h.  Kind 0 is the reference's only Attractor impl (above).  Kind 1,
 * PolynomialSprott3Degree, is this repository's extension "behind the same trait" (lib.rs:71-77,
 * README.md:8): the same serial sum continued with ten cubic terms — defined here and in include/sar.h,
 * there is no reference code to follow for it. */
void orc_next_point_cfg(const sar_config *cfg, double p[3])
{
    if (cfg->attractor_kind != SAR_ATTRACTOR_SPROTT3) { orc_next_point(cfg->coef, p); return; }
    const double x = p[0], y = p[1], z = p[2];
    const double xx = x * x, xy = x * y, xz = x * z, yy = y * y, yz = y * z, zz = z * z;
    const double m[20] = {1., x, xx, xy, xz, y, yy, yz, z, zz,
                          xx * x, xx * y, xx * z, xy * y, xy * z, xz * z, yy * y, yy * z, yz * z, zz * z};
    for (int k = 0; k < 3; ++k) {
        double sum = 0.;
        for (int i = 0; i < 10; ++i) sum += m[i] * cfg->coef[k][i];
        for (int i = 0; i < 10; ++i) sum += m[10 + i] * cfg->coef3[k][i];
        p[k] = sum;
    }
}

/* ---- EulerAxisRotation::to_rotation_matrix, lib.rs:179-195 (release) ------ */
void orc_rotation_matrix(const double axis[3], double rotation, double m[3][3])
{
    const double x = axis[0], y = axis[1], z = axis[2];   /* no normalize(): lib.rs:182-183 is debug-only */
    const double c = cos(rotation);
    const double c1 = 1. - c;
    const double s = sin(rotation);
    m[0][0] = c + x * x * c1; m[0][1] = x * y * c1 - z * s; m[0][2] = x * z * c1 + y * s;
    m[1][0] = y * x * c1 + z * s; m[1][1] = c + y * y * c1; m[1][2] = y * z * c1 - x * s;
    m[2][0] = z * x * c1 - y * s; m[2][1] = z * y * c1 + x * s; m[2][2] = c + z * z * c1;
}

/* ---- Matrix3x3::mul_right, lib.rs:208-215 --------------------------------- */
void orc_mul_right(const double m[3][3], const double v[3], double out[3])
{
    out[0] = m[0][0] * v[0] + m[0][1] * v[1] + m[0][2] * v[2];
    out[1] = m[1][0] * v[0] + m[1][1] * v[1] + m[1][2] * v[2];
    out[2] = m[2][0] * v[0] + m[2][1] * v[1] + m[2][2] * v[2];
}

/* Vec3::magnitude, lib.rs:129-131 */
static inline double magnitude(const double v[3])
{
    return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
}

/* ---- color transforms, lib.rs:511-516 and 520-558 ------------------------- */
double orc_color_transform(const sar_config *cfg, const double delta[3], const double p[3])
{
    if (cfg->ct_kind == SAR_CT_ADJUSTED_VELOCITY)
        return (magnitude(delta) + cfg->ct_offset) * cfg->ct_factor;      /* lib.rs:514 */
    if (cfg->ct_kind == SAR_CT_SCREEN_BLEND) {   /* include/sar.h: a closure-style ColorTransform (lib.rs:245), left to right */
        double t = p[0] * cfg->ct_weights[0];
        t = t + p[1] * cfg->ct_weights[1];
        t = t + p[2] * cfg->ct_weights[2];
        t = t + magnitude(delta) * cfg->ct_weights[3];
        return (t + cfg->ct_offset) * cfg->ct_factor;
    }

    /* poisson_saturne: cos/sin(91π/360) as literals, lib.rs:529-536 */
    static const double COS = 0.7009092642998508981833083453238941729068756103515625;
    static const double SIN = 0.7132504491541815649924274111981503665447235107421875;
    const double x2 = (p[0] + cfg->center_camera[0]) * COS + (p[2] + cfg->center_camera[1]) * SIN; /* lib.rs:538-539 */
    double part;
    if (x2 < -0.0839 || 10.55 * x2 + p[1] < 0.46 - 1.0941 || 1.0426 * x2 + p[1] < 0.179 - 0.1576 ||
        0.5139 * x2 - p[1] > -0.04 - 0.04092)                              /* lib.rs:542-545 */
        part = 0.;
    else
        part = 1.;
    const double color = (part + magnitude(delta)) / 2.;                   /* lib.rs:556 */
    return (color - 0.1) / 0.9;                                            /* lib.rs:557 */
}

/* ---- Palette::interpolate, lib.rs:442-472 --------------------------------- */
void orc_palette_interpolate(const sar_config *cfg, double value, double rgb[3])
{
    if (value < 0.) value = 0.;                /* lib.rs:443-449; NaN falls through unchanged */
    else if (value >= 1.) value = 0.999999;
    const uint32_t len = cfg->palette_len;
    value = value * (double)len;               /* count_f64 = list.len()-1 after the push, lib.rs:421,451 */
    /* `value.floor() as usize`: saturating, NaN -> 0 (lib.rs:453) */
    const double fl = floor(value);
    size_t n = (!(fl > 0.)) ? 0 : (fl >= (double)(len - 1) ? (size_t)(len - 1) : (size_t)fl);
    const double t = fmod(value, 1.);          /* lib.rs:454 */
    const double t1 = 1.0 - t;                 /* lib.rs:455 */
    const size_t n2 = (n + 1 < len) ? n + 1 : len - 1;   /* list[len] duplicates list[len-1], lib.rs:418 */
    for (int c = 0; c < 3; ++c)
        rgb[c] = sqrt(cfg->palette_rgb[n2][c] * t + cfg->palette_rgb[n][c] * t1); /* lib.rs:468-470 */
}

/* ---- Runtime, lib.rs:649-738 ---------------------------------------------- */
int orc_runtime_new(uint32_t w, uint32_t h, orc_runtime **out)
{
    orc_runtime *rt = (orc_runtime *)calloc(1, sizeof *rt);
    if (!rt) return -1;
    const size_t n = (size_t)w * h;
    rt->w = w; rt->h = h;
    rt->count = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    rt->steps = (double *)malloc((n ? n : 1) * sizeof(double));
    rt->zbuf = (float *)malloc((n ? n : 1) * sizeof(float));
    if (!rt->count || !rt->steps || !rt->zbuf) { orc_runtime_free(rt); return -1; }
    orc_runtime_reset(rt);
    *out = rt;
    return 0;
}
void orc_runtime_free(orc_runtime *rt)
{
    if (!rt) return;
    free(rt->count); free(rt->steps); free(rt->zbuf); free(rt);
}
void orc_runtime_reset(orc_runtime *rt)        /* lib.rs:682-699 */
{
    const size_t n = (size_t)rt->w * rt->h;
    memset(rt->count, 0, n * sizeof(uint32_t));
    for (size_t i = 0; i < n; ++i) rt->steps[i] = 0.;
    for (size_t i = 0; i < n; ++i) rt->zbuf[i] = -1.f;
    rt->max = 0;
}
int orc_runtime_merge(orc_runtime *a, const orc_runtime *b)   /* lib.rs:708-738 */
{
    if (a->w != b->w || a->h != b->h) return -1;              /* assert_eq!, lib.rs:709-710 */
    const size_t n = (size_t)a->w * a->h;
    for (size_t i = 0; i < n; ++i) {           /* per-pixel independent; traversal order is irrelevant */
        a->count[i] += b->count[i];            /* lib.rs:719 (wrapping in release) */
        if (a->count[i] > a->max) a->max = a->count[i];       /* lib.rs:721-723 */
        if (b->zbuf[i] > a->zbuf[i]) {         /* lib.rs:728: strict, ties keep self */
            a->steps[i] = b->steps[i];
            a->zbuf[i] = b->zbuf[i];
        }
    }
    return 0;
}

/* ---- render(), lib.rs:747-838 --------------------------------------------- */
void orc_render(const sar_config *cfg, orc_runtime *rt, const double init[3], orc_stats *st)
{
    double cur[3] = {init[0], init[1], init[2]};               /* lib.rs:748 (value injected) */
    for (int i = 0; i < 1000; ++i) orc_next_point_cfg(cfg, cur);   /* lib.rs:750-752 */

    double R[3][3];
    orc_rotation_matrix(cfg->axis, cfg->rotation, R);          /* lib.rs:755 */
    const double sin_v = sin(cfg->angle);                      /* lib.rs:756 */
    const double cos_v = cos(cfg->angle);                      /* lib.rs:757 */
    const double ccx = cfg->center_camera[0], ccy = cfg->center_camera[1], ccz = cfg->center_camera[2];
    const double width = (double)cfg->width;                   /* lib.rs:760 */
    const double height = (double)cfg->height;                 /* lib.rs:762 */
    const double width_scaled = width * cfg->scale;            /* lib.rs:763 */
    const double scale_adjusted_mid = 0.5 / cfg->scale;        /* lib.rs:764 */
    const uint32_t W = rt->w;

    double prev[3] = {cur[0], cur[1], cur[2]};                 /* lib.rs:766-767 */
    uint64_t recorded = 0, wins = 0, nans = 0, ties = 0;

    for (uint64_t it = 0; it < cfg->iterations; ++it) {        /* lib.rs:769 */
        orc_next_point_cfg(cfg, cur);                        /* lib.rs:770 */
        double s[3];
        orc_mul_right(R, cur, s);                              /* lib.rs:773 */
        const double x2 = (s[0] + ccx) * cos_v + (s[2] + ccy) * sin_v;   /* lib.rs:776-777 */
        const double z2 = (s[0] + ccx) * sin_v - (s[2] + ccy) * cos_v;   /* lib.rs:778-779 */
        const double fi = (scale_adjusted_mid - x2) * width_scaled;       /* lib.rs:783 */
        const double fj = height / 2. - (s[1] + ccz) * width_scaled;      /* lib.rs:786 */
        if (cur[0] != cur[0]) ++nans;
        if (fi >= width || fj >= height || fi < 0. || fj < 0.) {          /* lib.rs:789: NaN passes */
            prev[0] = cur[0]; prev[1] = cur[1]; prev[2] = cur[2];         /* lib.rs:793 */
            continue;
        }
        const uint32_t i = f64_as_u32(fi), j = f64_as_u32(fj);            /* lib.rs:800-802 */
        const size_t idx = (size_t)j * W + i;
        ++recorded;
        const uint32_t c = ++rt->count[idx];                   /* lib.rs:811, wrapping */
        if (c > rt->max) rt->max = c;                          /* lib.rs:813-815 */
        const float zf = (float)z2;                            /* `z2 as f32`, round-to-nearest-even */
        if (zf > rt->zbuf[idx]) {                              /* lib.rs:821, strict */
            const double delta[3] = {cur[0] - prev[0], cur[1] - prev[1], cur[2] - prev[2]}; /* lib.rs:822 */
            rt->steps[idx] = orc_color_transform(cfg, delta, s);           /* lib.rs:826-830 */
            rt->zbuf[idx] = zf;                                /* lib.rs:832 */
            ++wins;
        } else if (zf == rt->zbuf[idx]) {
            ++ties;
        }
        prev[0] = cur[0]; prev[1] = cur[1]; prev[2] = cur[2];  /* lib.rs:836 */
    }
    if (st) { st->recorded += recorded; st->z_wins += wins; st->nan_iters += nans; st->z_ties += ties; }
}

void orc_render_jobs(const sar_config *cfg, orc_runtime *rt, const double *init_xyz,
                     uint64_t n_jobs, orc_stats *st)
{
    for (uint64_t k = 0; k < n_jobs; ++k) orc_render(cfg, rt, init_xyz + 3 * k, st);
}

/* ---- the same n_jobs render() calls on n_threads OS threads, result identical to the serial loop ----
 * Thread t renders the CONTIGUOUS job slice [t*n/T, (t+1)*n/T) in list order — thread 0 straight into
 * `rt` (which may already hold earlier renders), the others into private reset Runtimes (as the
 * reference's workers do, lib.rs:938-951) — and the caller merges them IN THREAD ORDER with
 * Runtime::merge (lib.rs:708-738).  Counts add (commutative).  For z: each private Runtime holds, per
 * pixel, the greatest z of its slice and, among equal z, the earliest iteration of the earliest job
 * (strict `>`, lib.rs:821); merge's strict `>` (lib.rs:728) keeps the earlier slice on ties.  So the
 * merged (zbuf, steps) are those of the serial run, bit for bit, whatever T is.  max: counts only
 * grow, merge re-derives the running max from the summed counts (lib.rs:721-723).
 * Stats: recorded / nan_iters are order-independent; z_wins / z_ties count the private passes. */
typedef struct mt_worker {
    const sar_config *cfg;
    orc_runtime *rt;
    const double *init_xyz;
    uint64_t first, n;
    orc_stats st;
    pthread_t tid;
} mt_worker;
static void *mt_worker_main(void *arg)
{
    mt_worker *w = (mt_worker *)arg;
    for (uint64_t k = 0; k < w->n; ++k) orc_render(w->cfg, w->rt, w->init_xyz + 3 * (w->first + k), &w->st);
    return NULL;
}
typedef struct mt_merge {
    orc_runtime *dst;
    orc_runtime *const *src;      /* src[1..n_src) merged into dst, in order, over pixels [p0,p1) */
    uint32_t n_src;
    size_t p0, p1;
    uint32_t max;
    pthread_t tid;
} mt_merge;
static void *mt_merge_main(void *arg)  /* lib.rs:708-738 restricted to a pixel range; per-pixel independent */
{
    mt_merge *m = (mt_merge *)arg;
    orc_runtime *a = m->dst;
    uint32_t mx = 0;
    for (uint32_t s = 1; s < m->n_src; ++s) {
        const orc_runtime *b = m->src[s];
        for (size_t i = m->p0; i < m->p1; ++i) {
            a->count[i] += b->count[i];                        /* lib.rs:719 */
            if (b->zbuf[i] > a->zbuf[i]) { a->steps[i] = b->steps[i]; a->zbuf[i] = b->zbuf[i]; }  /* lib.rs:728-735 */
        }
    }
    for (size_t i = m->p0; i < m->p1; ++i) if (a->count[i] > mx) mx = a->count[i];   /* lib.rs:721-723 */
    m->max = mx;
    return NULL;
}
int orc_render_jobs_mt(const sar_config *cfg, orc_runtime *rt, const double *init_xyz, uint64_t n_jobs,
                       uint32_t n_threads, orc_stats *st)
{
    if (n_threads == 0) return -1;
    if ((uint64_t)n_threads > n_jobs) n_threads = (uint32_t)(n_jobs ? n_jobs : 1);
    if (n_threads == 1) { orc_render_jobs(cfg, rt, init_xyz, n_jobs, st); return 0; }
    mt_worker *ws = (mt_worker *)calloc(n_threads, sizeof *ws);
    orc_runtime **rts = (orc_runtime **)calloc(n_threads, sizeof *rts);
    mt_merge *ms = (mt_merge *)calloc(n_threads, sizeof *ms);
    if (!ws || !rts || !ms) { free(ws); free(rts); free(ms); return -1; }
    int rc = 0;
    rts[0] = rt;
    for (uint32_t t = 1; t < n_threads && rc == 0; ++t)
        if (orc_runtime_new(rt->w, rt->h, &rts[t]) != 0) rc = -1;
    uint32_t started = 0;
    for (uint32_t t = 0; t < n_threads && rc == 0; ++t) {
        ws[t].cfg = cfg; ws[t].rt = rts[t]; ws[t].init_xyz = init_xyz;
        ws[t].first = n_jobs * t / n_threads;
        ws[t].n = n_jobs * (t + 1) / n_threads - ws[t].first;
        if (pthread_create(&ws[t].tid, NULL, mt_worker_main, &ws[t]) != 0) { rc = -1; break; }
        ++started;
    }
    for (uint32_t t = 0; t < started; ++t) pthread_join(ws[t].tid, NULL);
    if (rc == 0) {
        /* merge in thread order; pixels are independent, so the pixel range is split over the threads */
        const size_t npix = (size_t)rt->w * rt->h;
        uint32_t m_started = 0;
        for (uint32_t t = 0; t < n_threads; ++t) {
            ms[t].dst = rt; ms[t].src = rts; ms[t].n_src = n_threads;
            ms[t].p0 = npix * t / n_threads; ms[t].p1 = npix * (t + 1) / n_threads;
            if (pthread_create(&ms[t].tid, NULL, mt_merge_main, &ms[t]) != 0) { rc = -1; break; }
            ++m_started;
        }
        for (uint32_t t = 0; t < m_started; ++t) pthread_join(ms[t].tid, NULL);
        if (rc == 0) {
            uint32_t mx = rt->max;
            for (uint32_t t = 0; t < n_threads; ++t) if (ms[t].max > mx) mx = ms[t].max;
            rt->max = mx;
            if (st) for (uint32_t t = 0; t < n_threads; ++t) {
                st->recorded += ws[t].st.recorded; st->z_wins += ws[t].st.z_wins;
                st->nan_iters += ws[t].st.nan_iters; st->z_ties += ws[t].st.z_ties;
            }
        }
    }
    for (uint32_t t = 1; t < n_threads; ++t) orc_runtime_free(rts[t]);
    free(ws); free(rts); free(ms);
    return rc;
}

/* ---- colorize(), lib.rs:841-904 ------------------------------------------- */
void orc_colorize(const sar_config *cfg, const orc_runtime *rt, uint16_t *out, double *out_f64)
{
    const size_t n = (size_t)rt->w * rt->h;
    const double u16_max = 65535.;                             /* lib.rs:848 */
    const double bo = cfg->bright_offset, bf = cfg->bright_factor;
    if (cfg->render_kind == SAR_RENDER_GAS) {                  /* lib.rs:853-874 */
        for (size_t p = 0; p < n; ++p) {
            double rgb[3];
            orc_palette_interpolate(cfg, rt->steps[p], rgb);   /* lib.rs:857 */
            /* lib.rs:860: f64::from(count+1).log(f64::from(max+1)) == ln(a)/ln(b); u32 adds wrap */
            const double factor = log((double)(uint32_t)(rt->count[p] + 1u)) /
                                  log((double)(uint32_t)(rt->max + 1u));
            double v[4];
            v[0] = (rgb[0] * factor + bo) * bf;                /* lib.rs:862-864 (before * u16_max) */
            v[1] = (rgb[1] * factor + bo) * bf;
            v[2] = (rgb[2] * factor + bo) * bf;
            v[3] = cfg->transparent ? factor : 1.0;            /* lib.rs:865-869 */
            out[4 * p + 0] = f64_as_u16(v[0] * u16_max);
            out[4 * p + 1] = f64_as_u16(v[1] * u16_max);
            out[4 * p + 2] = f64_as_u16(v[2] * u16_max);
            out[4 * p + 3] = cfg->transparent ? f64_as_u16(factor * u16_max) : 65535u;
            if (out_f64) { out_f64[4 * p] = v[0]; out_f64[4 * p + 1] = v[1]; out_f64[4 * p + 2] = v[2]; out_f64[4 * p + 3] = v[3]; }
        }
    } else {                                                   /* lib.rs:875-900 */
        float mx = 0.0f, mn = 3.40282347e+38f;                 /* fold seed (0.0, f32::MAX), lib.rs:882 */
        for (size_t p = 0; p < n; ++p) {
            const float z = rt->zbuf[p];
            if (z != -1.0f) { mx = fmaxf(mx, z); mn = fminf(mn, z); }
        }
        const float diff = mx - mn;                            /* lib.rs:883 */
        for (size_t p = 0; p < n; ++p) {
            float z = rt->zbuf[p];
            z = (z == -1.0f) ? 0.0f : (z - mn) / diff;         /* lib.rs:889-894 */
            const uint16_t g = f32_as_u16(z * 65535.0f);       /* lib.rs:895 */
            out[4 * p + 0] = g; out[4 * p + 1] = g; out[4 * p + 2] = g; out[4 * p + 3] = 65535u;
            if (out_f64) { out_f64[4 * p] = z; out_f64[4 * p + 1] = z; out_f64[4 * p + 2] = z; out_f64[4 * p + 3] = 1.0; }
        }
    }
}

/* ---- render_parallel(), lib.rs:1051-1082 on OS threads --------------------- */
typedef struct par_shared {
    const sar_config *cfg;           /* iterations already divided (lib.rs:1058) */
    const double *init_xyz;
    uint64_t total_jobs;
    atomic_ullong counter;           /* job_counter, lib.rs:1062 */
} par_shared;
typedef struct par_worker {
    par_shared *sh;
    orc_runtime *rt;                 /* private Runtime, lib.rs:938 */
    pthread_t tid;
} par_worker;

static void *par_worker_main(void *arg)
{
    par_worker *w = (par_worker *)arg;
    par_shared *sh = w->sh;
    for (;;) {
        /* fetch_update(v>0 -> v-1), lib.rs:962-982 */
        unsigned long long v = atomic_load(&sh->counter);
        int got = 0;
        while (v > 0) {
            if (atomic_compare_exchange_weak(&sh->counter, &v, v - 1)) { got = 1; break; }
        }
        if (!got) break;                                       /* lib.rs:984-986 */
        const uint64_t job = sh->total_jobs - v;               /* deterministic job -> start point */
        orc_render(sh->cfg, w->rt, sh->init_xyz + 3 * job, NULL);   /* lib.rs:987 */
    }
    return NULL;
}

int orc_render_parallel(const sar_config *cfg_in, uint32_t n_threads, uint64_t jobs_per_thread,
                        const double *init_xyz, uint16_t *rgba, orc_runtime **merged)
{
    if (n_threads == 0 || jobs_per_thread == 0) return -1;
    sar_config cfg = *cfg_in;
    cfg.iterations = cfg_in->iterations / n_threads / jobs_per_thread;    /* lib.rs:1058 */
    par_shared sh;
    sh.cfg = &cfg; sh.init_xyz = init_xyz;
    sh.total_jobs = (uint64_t)n_threads * jobs_per_thread;                /* lib.rs:1062 */
    atomic_init(&sh.counter, sh.total_jobs);

    par_worker *ws = (par_worker *)calloc(n_threads, sizeof *ws);
    if (!ws) return -1;
    int rc = 0;
    for (uint32_t t = 0; t < n_threads; ++t) {
        ws[t].sh = &sh;
        if (orc_runtime_new(cfg.width, cfg.height, &ws[t].rt) != 0) { rc = -1; n_threads = t; break; }  /* lib.rs:950-951 */
    }
    uint32_t started = 0;
    if (rc == 0)
        for (uint32_t t = 0; t < n_threads; ++t) {
            if (pthread_create(&ws[t].tid, NULL, par_worker_main, &ws[t]) != 0) { rc = -1; break; }
            ++started;
        }
    for (uint32_t t = 0; t < started; ++t) pthread_join(ws[t].tid, NULL);
    if (rc == 0) {
        for (uint32_t t = 1; t < n_threads; ++t) orc_runtime_merge(ws[0].rt, ws[t].rt);  /* lib.rs:1072-1076 */
        if (rgba) orc_colorize(&cfg, ws[0].rt, rgba, NULL);                              /* lib.rs:1080 */
    }
    for (uint32_t t = 0; t < n_threads; ++t) {
        if (t == 0 && rc == 0 && merged) { *merged = ws[0].rt; continue; }
        orc_runtime_free(ws[t].rt);
    }
    free(ws);
    return rc;
}

/* ---- known-answer helper: screen-space bounding box, lib.rs:329-333 -------- */
void orc_screen_bbox(const sar_config *cfg, const double init[3], uint64_t n, double box[6])
{
    double cur[3] = {init[0], init[1], init[2]};
    for (int i = 0; i < 1000; ++i) orc_next_point_cfg(cfg, cur);
    double R[3][3];
    orc_rotation_matrix(cfg->axis, cfg->rotation, R);
    box[0] = box[2] = box[4] = INFINITY;
    box[1] = box[3] = box[5] = -INFINITY;
    for (uint64_t it = 0; it < n; ++it) {
        orc_next_point_cfg(cfg, cur);
        double s[3];
        orc_mul_right(R, cur, s);
        for (int c = 0; c < 3; ++c) {
            if (s[c] < box[2 * c]) box[2 * c] = s[c];
            if (s[c] > box[2 * c + 1]) box[2 * c + 1] = s[c];
        }
    }
}

/* The first pass the reference's author sketches at lib.rs:326-334, over a list of start points: the
 * union of the per-trajectory boxes of every trajectory that stays finite; the others are counted. */
void orc_screen_bbox_jobs(const sar_config *cfg, const double *init_xyz, uint64_t n_jobs, uint64_t n,
                          double box[6], uint64_t *diverged)
{
    box[0] = box[2] = box[4] = INFINITY;
    box[1] = box[3] = box[5] = -INFINITY;
    uint64_t bad = 0;
    double R[3][3];
    orc_rotation_matrix(cfg->axis, cfg->rotation, R);
    for (uint64_t k = 0; k < n_jobs; ++k) {
        double cur[3] = {init_xyz[3 * k], init_xyz[3 * k + 1], init_xyz[3 * k + 2]};
        for (int i = 0; i < 1000; ++i) orc_next_point_cfg(cfg, cur);
        double b[6] = {INFINITY, -INFINITY, INFINITY, -INFINITY, INFINITY, -INFINITY};
        for (uint64_t it = 0; it < n; ++it) {
            orc_next_point_cfg(cfg, cur);
            double s[3];
            orc_mul_right(R, cur, s);
            for (int c = 0; c < 3; ++c) {
                if (s[c] < b[2 * c]) b[2 * c] = s[c];
                if (s[c] > b[2 * c + 1]) b[2 * c + 1] = s[c];
            }
        }
        int ok = isfinite(cur[0]) && isfinite(cur[1]) && isfinite(cur[2]);
        for (int c = 0; c < 6; ++c) ok = ok && isfinite(b[c]);
        if (!ok) { ++bad; continue; }
        for (int c = 0; c < 3; ++c) {
            if (b[2 * c] < box[2 * c]) box[2 * c] = b[2 * c];
            if (b[2 * c + 1] > box[2 * c + 1]) box[2 * c + 1] = b[2 * c + 1];
        }
    }
    if (diverged) *diverged = bad;
}

/* ---- start-point generator (same definition as include/sar.h) -------------- */
static inline uint64_t splitmix64_at(uint64_t seed, uint64_t n)   /* n-th output of the stream */
{
    uint64_t z = seed + (n + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
void orc_seed_points(uint64_t seed, uint64_t first, uint64_t n, double *out)
{
    for (uint64_t k = 0; k < n; ++k)
        for (uint64_t c = 0; c < 3; ++c) {
            const uint64_t u = splitmix64_at(seed, 3 * (first + k) + c);
            /* [0,1) with 53 bits, then `* 0.1` as lib.rs:748 */
            out[3 * k + c] = ((double)(u >> 11) * 0x1.0p-53) * 0.1;
        }
}

/* ---- presets --------------------------------------------------------------- */
static void config_defaults(sar_config *c)     /* Config::new lib.rs:289-307, Colors::default lib.rs:480-491 */
{
    c->iterations = 10000000ull;
    c->width = 1920; c->height = 1080;
    c->render_kind = SAR_RENDER_GAS;
    c->transparent = 1;
    c->angle = 0.0;
    c->silent = 1;
    c->palette_len = 6;
    static const double r[6] = {1., 0.5, 1., 0.5, 0.5, 1.};
    static const double g[6] = {1., 1., 0.5, 1., 0.5, 0.5};
    static const double b[6] = {0.5, 0.5, 0.5, 1., 1., 1.};
    for (int i = 0; i < 6; ++i) { c->palette_rgb[i][0] = r[i]; c->palette_rgb[i][1] = g[i]; c->palette_rgb[i][2] = b[i]; }
    c->bright_offset = -0.15;                  /* lib.rs:400 */
    c->bright_factor = 5. / 3.;                /* lib.rs:401 */
}
void orc_config_poisson_saturne(sar_config *c) /* lib.rs:310-352 */
{
    memset(c, 0, sizeof *c);
    static const double x[10] = {0.021, 1.182, -1.183, 0.128, -1.12, -0.641, -1.152, -0.834, -0.97, 0.722};
    static const double y[10] = {0.243038, -0.825, -1.2, -0.835443, -0.835443, -0.364557, 0.458, 0.622785, -0.394937, -1.032911};
    static const double z[10] = {-0.455696, 0.673, 0.915, -0.258228, -0.495, -0.264, -0.432, -0.416, -0.877, -0.3};
    memcpy(c->coef[0], x, sizeof x); memcpy(c->coef[1], y, sizeof y); memcpy(c->coef[2], z, sizeof z);
    c->center_camera[0] = -0.005; c->center_camera[1] = 0.262; c->center_camera[2] = -0.366 + 0.12;
    c->axis[0] = 0.304289493528802; c->axis[1] = 0.760492682863655; c->axis[2] = 0.573636455813981;
    c->rotation = 1.78268191887446;
    c->scale = 1.;
    c->ct_kind = SAR_CT_POISSON_SATURNE;
    config_defaults(c);
}
void orc_config_solar_sail(sar_config *c)      /* lib.rs:355-386 */
{
    memset(c, 0, sizeof *c);
    static const double x[10] = {0.744304, -0.546835, 0.121519, -0.653165, 0.399, 0.379, 0.44, 1.014, -0.805063, 0.377};
    static const double y[10] = {-0.683, 0.531646, -0.04557, -1.2, -0.546835, 0.091139, 0.744304, -0.273418, -0.349367, -0.531646};
    static const double z[10] = {0.712, 0.744304, -0.577215, 0.966, 0.04557, 1.063291, 0.01519, -0.425316, 0.212658, -0.01519};
    memcpy(c->coef[0], x, sizeof x); memcpy(c->coef[1], y, sizeof y); memcpy(c->coef[2], z, sizeof z);
    c->center_camera[0] = 0.28; c->center_camera[1] = -0.12; c->center_camera[2] = 0.22;
    c->axis[0] = 0.02466; c->axis[1] = 0.4618; c->axis[2] = -0.54789;
    c->rotation = 2.2195;
    c->scale = 1.7;
    c->ct_kind = SAR_CT_ADJUSTED_VELOCITY;
    c->ct_factor = -0.2; c->ct_offset = 0.8;   /* lib.rs:381-384 */
    config_defaults(c);
}

/* ---- output conversion + raw containers, src/bin/main.rs:40-100 ------------
 * main.rs:52-57 picks DynamicImage::{as is, to_rgb16, to_rgba8, to_rgb8}; the PAM / BMP encoders of the
 * `image` crate (0.25, Cargo.toml:15 — NOT vendored under /root/reference: third-party, restated from its
 * published source, "parity unpinned (third-party)") then write image.as_bytes():
 *   u16 -> u8 sample: `(c16 as u32 + 128) / 257` (image/src/color.rs, FromPrimitive<u16> for u8)
 *   PAM: header "P7\nWIDTH..\nHEIGHT..\nDEPTH..\nMAXVAL..\nTUPLTYPE ..\nENDHDR\n", 16-bit samples big-endian
 *   BMP: 8-bit only; Rgb8 -> BITMAPINFOHEADER (40), Rgba8 -> BITMAPV4HEADER (108) BI_BITFIELDS; BGR(A), bottom-up, rows padded to 4. */
static size_t put_u16le(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); return 2; }
static size_t put_u32le(uint8_t *p, uint32_t v) { put_u16le(p, v & 0xFFFFu); put_u16le(p + 2, v >> 16); return 4; }
/* PNG (main.rs:78-89) with the compressor left out: signature, IHDR, one IDAT = zlib stream of "stored" deflate
 * blocks (RFC 1950 / 1951) around the filter-type-0 scanlines (16-bit samples big-endian, PNG spec), IEND.  CRC-32
 * and Adler-32 by their bitwise / bytewise definitions. */
static uint32_t png_crc(const uint8_t *p, size_t n)
{
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) {
        c ^= p[i];
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
    }
    return c ^ 0xFFFFFFFFu;
}
static void be32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
static size_t png_encode(const uint16_t *rgba, uint32_t w, uint32_t h, int wide, int alpha, uint8_t *out)
{
    const size_t nch = alpha ? 4 : 3, bpp = nch * (wide ? 2 : 1), raw_row = 1 + (size_t)w * bpp, raw_len = raw_row * h;
    const size_t n_blocks = (raw_len + 65534) / 65535, zlen = 2 + raw_len + 5 * n_blocks + 4;
    const size_t total = 8 + 25 + 12 + zlen + 12;
    if (!out) return total;
    uint8_t *raw = (uint8_t *)malloc(raw_len ? raw_len : 1);
    for (uint32_t y = 0; y < h; ++y) {
        uint8_t *r = raw + (size_t)y * raw_row;
        *r++ = 0;                                              /* filter type 0 */
        for (uint32_t x = 0; x < w; ++x)
            for (size_t c = 0; c < nch; ++c) {
                const uint16_t v = rgba[4 * ((size_t)y * w + x) + c];
                if (wide) { *r++ = (uint8_t)(v >> 8); *r++ = (uint8_t)v; }
                else *r++ = (uint8_t)(((uint32_t)v + 128u) / 257u);
            }
    }
    uint8_t *p = out;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    memcpy(p, sig, 8); p += 8;
    be32(p, 13); memcpy(p + 4, "IHDR", 4); be32(p + 8, w); be32(p + 12, h);
    p[16] = wide ? 16 : 8; p[17] = alpha ? 6 : 2; p[18] = 0; p[19] = 0; p[20] = 0;
    be32(p + 21, png_crc(p + 4, 17)); p += 25;
    be32(p, (uint32_t)zlen); memcpy(p + 4, "IDAT", 4);
    uint8_t *z = p + 8;
    z[0] = 0x78; z[1] = 0x01;
    uint8_t *q = z + 2;
    uint32_t a = 1, b = 0;
    for (size_t blk = 0; blk < n_blocks; ++blk) {
        const size_t first = blk * 65535, len = raw_len - first < 65535 ? raw_len - first : 65535;
        q[0] = blk + 1 == n_blocks; q[1] = (uint8_t)len; q[2] = (uint8_t)(len >> 8); q[3] = (uint8_t)~len; q[4] = (uint8_t)(~len >> 8);
        memcpy(q + 5, raw + first, len);
        q += 5 + len;
    }
    for (size_t i = 0; i < raw_len; ++i) { a = (a + raw[i]) % 65521u; b = (b + a) % 65521u; }
    be32(q, (b << 16) | a); q += 4;
    be32(q, png_crc(p + 4, 4 + zlen)); q += 4;
    be32(q, 0); memcpy(q + 4, "IEND", 4); be32(q + 8, png_crc(q + 4, 4));
    free(raw);
    return total;
}

size_t orc_encode(const uint16_t *rgba, uint32_t w, uint32_t h, uint32_t fmt, uint32_t container, uint8_t *out)
{
    if (container > SAR_FILE_PNG) return 0;                     /* unknown / variable-size containers: unsupported here */
    if (container == SAR_FILE_PNG)
        return png_encode(rgba, w, h, fmt == SAR_PIX_RGBA16 || fmt == SAR_PIX_RGB16, fmt == SAR_PIX_RGBA16 || fmt == SAR_PIX_RGBA8, out);
    const int wide = fmt == SAR_PIX_RGBA16 || fmt == SAR_PIX_RGB16, alpha = fmt == SAR_PIX_RGBA16 || fmt == SAR_PIX_RGBA8;
    const size_t nch = alpha ? 4 : 3, bpp = nch * (wide ? 2 : 1);
    size_t stride = (size_t)w * bpp, n = 0;
    if (container == SAR_FILE_BMP) {
        if (wide) return 0;                                    /* BmpEncoder: unsupported color type -> the reference panics */
        stride = (stride + 3) / 4 * 4;
        const uint32_t dib = alpha ? 108u : 40u, off = 14u + dib, img = (uint32_t)(stride * h);
        if (out) {
            uint8_t *p = out;
            *p++ = 'B'; *p++ = 'M'; p += put_u32le(p, img + off); p += put_u16le(p, 0); p += put_u16le(p, 0); p += put_u32le(p, off);
            p += put_u32le(p, dib); p += put_u32le(p, w); p += put_u32le(p, h); p += put_u16le(p, 1); p += put_u16le(p, alpha ? 32 : 24);
            p += put_u32le(p, alpha ? 3u : 0u); p += put_u32le(p, img);
            for (int k = 0; k < 4; ++k) p += put_u32le(p, 0);
            if (alpha) {
                p += put_u32le(p, 0xFFu << 16); p += put_u32le(p, 0xFFu << 8); p += put_u32le(p, 0xFFu); p += put_u32le(p, 0xFFu << 24);
                p += put_u32le(p, 0x73524742u);
                for (int k = 0; k < 12; ++k) p += put_u32le(p, 0);
            }
        }
        n = off;
    } else if (container == SAR_FILE_PAM) {
        char head[160];
        n = (size_t)snprintf(head, sizeof head, "P7\nWIDTH %u\nHEIGHT %u\nDEPTH %u\nMAXVAL %u\nTUPLTYPE %s\nENDHDR\n",
                             w, h, (unsigned)nch, wide ? 65535u : 255u, alpha ? "RGB_ALPHA" : "RGB");
        if (out) memcpy(out, head, n);
    }
    if (out) {
        for (uint32_t y = 0; y < h; ++y) {
            uint8_t *row = out + n + (size_t)(container == SAR_FILE_BMP ? h - 1 - y : y) * stride;
            memset(row, 0, stride);
            for (uint32_t x = 0; x < w; ++x) {
                const uint16_t *px = rgba + 4 * ((size_t)y * w + x);
                uint8_t *o = row + (size_t)x * bpp;
                for (size_t c = 0; c < nch; ++c) {
                    if (wide) {
                        if (container == SAR_FILE_PAM) { o[2 * c] = (uint8_t)(px[c] >> 8); o[2 * c + 1] = (uint8_t)px[c]; }
                        else { o[2 * c] = (uint8_t)px[c]; o[2 * c + 1] = (uint8_t)(px[c] >> 8); }
                    } else {
                        const uint8_t v = (uint8_t)(((uint32_t)px[c] + 128u) / 257u);
                        const size_t k = container == SAR_FILE_BMP && c < 3 ? 2 - c : c;     /* B,G,R[,A] */
                        o[k] = v;
                    }
                }
            }
        }
    }
    return n + stride * h;
}
