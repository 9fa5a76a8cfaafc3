/* sar_oracle.h — CPU restatement of the reference's render path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the oracle the CUDA path is checked against.  It is NOT part of the
 * product: only tests/, __graft_entry__.smoke() and bench.py — its cpu_baseline /
 * `--impl reference` legs and the N-rank parity frames it checks outside every
 * timed region — may build, load or call it.  libsar_b200.so never links or
 * calls anything in this directory.
 *
 * It restates, function by function, `src/lib.rs` of Icelk/strange-attractor-
 * renderer @ e571d19 (citations below are into that file), with Rust's
 * semantics spelled out in C: release build (no axis normalisation,
 * lib.rs:181-183; wrapping u32 `+= 1`, lib.rs:811), saturating float→int `as`
 * casts with NaN→0 (lib.rs:800-802, 862-866, 895), `x.log(b)` = ln x / ln b
 * (lib.rs:860), `%` = fmod (lib.rs:454), and no floating-point contraction
 * (built with -ffp-contract=off; rustc never fuses a*b+c).
 *
 * The one thing it cannot restate is the start point: the reference draws it
 * from an OS-seeded SmallRng (lib.rs:656, 748; crate `rand` 0.9, not vendored,
 * no lockfile).  Every entry point therefore takes the start point(s)
 * explicitly.
 *
 * Pinning (see oracle/README.md, tests/test_reference_images.py, tests/test_oracle_golden.py): the
 * reference's own tests hold no numeric vector for this path (one doctest that only builds a
 * Config, lib.rs:9-15), the Rust toolchain is absent here (no `oracle/_ref`), and the reference
 * seeds itself from the OS, so it cannot be run on given inputs at all.  What it does publish is
 * three output images (the PNGs under media, README.md:72-77), and the oracle is pinned on those pixel by
 * pixel: colorize is inverted on every fully informative pixel (2 133 250), recovering integer
 * count, palette position and Runtime.max such that lib.rs:853-868 returns the PNG's 16-bit channels
 * EXACTLY (orc_colorize gives the sampled pixels back bit for bit); the recovered count field equals
 * the oracle's own 1e9-iteration render at Poisson noise per pixel (chi-square 0.99-1.01; one pixel
 * of shift: > 60), the recovered palette positions equal the oracle's `steps` (median 1e-5), and
 * solar-sail's max is k x (1e9/12/12), the render_parallel decomposition feeding the NaN sink.
 * In the strict sense of "reference run here on seeded inputs" the oracle remains unpinned — the
 * reference has no seeded mode; a `>` vs `>=` on a z tie is below the noise floor of its outputs.
 * Also: the bounding-box known answer (lib.rs:329-333), and a second independent restatement
 * (tests/pyref.py) that agrees bit for bit.
 */
#ifndef SAR_ORACLE_H
#define SAR_ORACLE_H

#include <stdint.h>
#include "../include/sar.h"   /* sar_config only: the POD mirror of Config */

#ifdef __cplusplus
extern "C" {
#endif

/* Runtime, lib.rs:631-646 (without the rng). Row-major, idx = y*w + x. */
typedef struct orc_runtime {
    uint32_t  w, h;
    uint32_t *count;   /* lib.rs:633 */
    double   *steps;   /* lib.rs:635 */
    float    *zbuf;    /* lib.rs:639 */
    uint32_t  max;     /* lib.rs:643 */
} orc_runtime;

typedef struct orc_stats {
    uint64_t recorded;     /* iterations that passed the bounds test (lib.rs:789) */
    uint64_t z_wins;       /* iterations that took the branch at lib.rs:821 */
    uint64_t nan_iters;    /* iterations executed with a NaN state */
    uint64_t z_ties;       /* candidate z == stored z (lost by the strict >) */
} orc_stats;

/* PolynomialSprott2Degree::next_point, lib.rs:585-620 (in place). */
void orc_next_point(const double coef[3][10], double p[3]);
/* next_point of cfg->attractor_kind (0: the above; 1: the cubic extension defined in include/sar.h). */
void orc_next_point_cfg(const sar_config *cfg, double p[3]);
/* EulerAxisRotation::to_rotation_matrix, lib.rs:179-195, release semantics. */
void orc_rotation_matrix(const double axis[3], double rotation, double m[3][3]);
/* Matrix3x3::mul_right, lib.rs:208-215. */
void orc_mul_right(const double m[3][3], const double v[3], double out[3]);
/* ColorTransform::transform for the two shipped kinds, lib.rs:511-516, 520-558. */
double orc_color_transform(const sar_config *cfg, const double delta[3], const double screen[3]);
/* Palette::interpolate, lib.rs:442-472. */
void orc_palette_interpolate(const sar_config *cfg, double value, double rgb[3]);

/* Runtime::new / reset / merge, lib.rs:660, 682, 708. */
int  orc_runtime_new(uint32_t w, uint32_t h, orc_runtime **out);
void orc_runtime_free(orc_runtime *rt);
void orc_runtime_reset(orc_runtime *rt);
int  orc_runtime_merge(orc_runtime *dst, const orc_runtime *src);

/* render(), lib.rs:747-838, with the start point (value of
 * `rng.random::<Vec3>() * 0.1`, lib.rs:748) passed in.  stats may be NULL. */
void orc_render(const sar_config *cfg, orc_runtime *rt, const double init[3], orc_stats *stats);
/* n_jobs render() calls in list order on one Runtime ("continues the building
 * of the image", lib.rs:742-743) — the sequential semantics the GPU matches. */
void orc_render_jobs(const sar_config *cfg, orc_runtime *rt, const double *init_xyz,
                     uint64_t n_jobs, orc_stats *stats);
/* The same n_jobs render() calls, computed on n_threads OS threads (contiguous job slices into
 * private Runtimes, merged in thread order with Runtime::merge): bit-identical to orc_render_jobs
 * for count, zbuf, steps and max, whatever n_threads is — see the proof sketch in sar_oracle.c.
 * It exists so that the BASELINE configurations (1e9 iterations) can be checked at full size.
 * Returns 0, or -1 on allocation / thread failure. */
int orc_render_jobs_mt(const sar_config *cfg, orc_runtime *rt, const double *init_xyz,
                       uint64_t n_jobs, uint32_t n_threads, orc_stats *stats);
/* colorize(), lib.rs:841-904.  rgba_f64 (optional) gets the pre-cast values. */
void orc_colorize(const sar_config *cfg, const orc_runtime *rt, uint16_t *rgba_u16,
                  double *rgba_f64);

/* render_parallel(), lib.rs:1051-1082, on n_threads OS threads: per-thread
 * private Runtime, dynamic job counter (lib.rs:962-982), job k starts at
 * init_xyz[3k], merge in thread order on the caller, colorize.  merged
 * (optional) receives the merged Runtime (caller frees).  Returns 0 or -1. */
int orc_render_parallel(const sar_config *cfg, uint32_t n_threads, uint64_t jobs_per_thread,
                        const double *init_xyz, uint16_t *rgba_u16, orc_runtime **merged);

/* Trajectory bounding box in screen space (R·p) over n iterations after the
 * warm-up: the known answer of lib.rs:329-333.  box = {xmin,xmax,ymin,ymax,zmin,zmax}. */
void orc_screen_bbox(const sar_config *cfg, const double init[3], uint64_t n, double box[6]);

/* The same over a list of start points (the auto-framing first pass of the TODO at lib.rs:326-334):
 * union of the boxes of the trajectories that stay finite; *diverged counts the rest. */
void orc_screen_bbox_jobs(const sar_config *cfg, const double *init_xyz, uint64_t n_jobs, uint64_t n,
                          double box[6], uint64_t *diverged);

/* Output conversion + raw containers (src/bin/main.rs:40-100 + the `image` 0.25 crate it calls, which is
 * third-party and not vendored: restated from its published source, parity unpinned).  fmt / container
 * are the SAR_PIX_* / SAR_FILE_* values of include/sar.h.  Returns the byte count (0: unsupported, as the
 * reference's BmpEncoder on 16-bit images); out may be NULL to query it. */
size_t orc_encode(const uint16_t *rgba_u16, uint32_t w, uint32_t h, uint32_t fmt, uint32_t container, uint8_t *out);

/* Same generator as sar_seed_points (include/sar.h), restated independently. */
void orc_seed_points(uint64_t seed, uint64_t first, uint64_t n, double *out_xyz);

/* Config presets restated from lib.rs:289-307, 310-352, 355-386, 397-404, 480-491. */
void orc_config_poisson_saturne(sar_config *cfg);
void orc_config_solar_sail(sar_config *cfg);

#ifdef __cplusplus
}
#endif
#endif
