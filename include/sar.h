/* sar.h — C-ABI boundary of the B200-native strange-attractor render path.
 *
 * One shared library, `libsar_b200.so`, replaces exactly one path of
 * Icelk/strange-attractor-renderer: iterate the polynomial Sprott map, project,
 * scatter into the count / z / Δp buffers, log tone-map and palette-colourise.
 * Every entry point below names the reference interface (file:line into the
 * reference repository, `src/lib.rs` unless stated) it stands in for.  The
 * reference has no FFI of its own; the seam is its public Rust API
 *   Runtime::{new,reset,merge}   lib.rs:660,682,708
 *   render()                     lib.rs:747
 *   colorize()                   lib.rs:841
 *   ParallelRenderer::{new,shutdown}, render_parallel()   lib.rs:919,1020,1051
 * and INTEGRATION.md shows the Rust `extern "C"` block a maintainer adds.
 *
 * Conventions
 *   - plain pointers and sizes only; all structs are POD, little-endian, natural
 *     alignment (`#[repr(C)]` on the Rust side);
 *   - the caller owns every host buffer passed in or out; opaque handles own
 *     device memory; nothing allocated on one side is freed on the other
 *     (except via the matching sar_*_free / sar_host_free);
 *   - every function returning `int` returns SAR_OK (0) or a negative sar_status;
 *     sar_last_error() gives a thread-local message.  Nothing unwinds across
 *     the boundary.  (The reference panics instead — lib.rs:678,709-710,1024 —
 *     the Rust shim turns non-zero statuses back into those panics.)
 *   - calls on one handle must be serialised by the caller (the reference takes
 *     `&mut Runtime` / `&mut ParallelRenderer`, lib.rs:747,1052); different
 *     handles may be used from different threads.  Every call sets the CUDA
 *     device it needs itself; no thread-local CUDA state is assumed.
 *   - there is NO CPU fallback: without a usable CUDA device every compute
 *     entry point fails with SAR_ERR_CUDA.
 *
 * Determinism contract (what "parity" means)
 *   The reference seeds every Runtime from the OS (lib.rs:656) and draws one
 *   start point per render() (lib.rs:748), so it is not reproducible.  Here the
 *   start points are an explicit input: either a caller-supplied list
 *   (`init_xyz`, n_jobs×3 f64, the value of `rng.random::<Vec3>() * 0.1`
 *   BEFORE the 1000 warm-up steps of lib.rs:750-752) or a documented
 *   counter-based generator (sar_seed_points).  Given that list, the result of
 *   sar_render(cfg, rt, init_xyz, n_jobs) is bit-identical — count (u32),
 *   zbuf (f32) and steps (f64) — to calling the reference's render() n_jobs
 *   times in list order on one non-reset Runtime.
 */
#ifndef SAR_B200_H
#define SAR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAR_ABI_VERSION 2u
#define SAR_MAX_PALETTE 16u        /* palette entries carried across the boundary */
#define SAR_WARMUP_ITERATIONS 1000u /* lib.rs:750 */

typedef enum sar_status {
    SAR_OK = 0,
    SAR_ERR_INVALID = -1,   /* NULL pointer, zero size, bad enum, palette_len out of range */
    SAR_ERR_DIMS = -2,      /* dimension mismatch (reference: assert_eq! in merge, lib.rs:709-710) */
    SAR_ERR_CUDA = -3,      /* CUDA runtime error or no usable device */
    SAR_ERR_NOMEM = -4,     /* host or device allocation failed */
    SAR_ERR_UNSUPPORTED = -5 /* e.g. an Attractor/ColorTransform with no device form */
} sar_status;

/* config::RenderKind, lib.rs:234-239 */
enum { SAR_RENDER_GAS = 0, SAR_RENDER_DEPTH = 1 };
/* ColorTransform kinds (lib.rs:241-249).  0 and 1 are the two instantiations the
 * reference ships (lib.rs:503-559).  2 = ScreenBlend, a device form of the
 * closures the trait also accepts (lib.rs:245), built from exact operations only:
 *   value = ((((s.x*w0) + s.y*w1) + s.z*w2) + |delta|*w3 + ct_offset) * ct_factor
 * with s = screen_space, w = ct_weights, evaluated left to right, no FMA. */
enum { SAR_CT_POISSON_SATURNE = 0, SAR_CT_ADJUSTED_VELOCITY = 1, SAR_CT_SCREEN_BLEND = 2 };
/* Attractor kinds (lib.rs:71-77; README.md:8 "Adding more should be relatively
 * easy").  0 = PolynomialSprott2Degree, the one the reference ships
 * (lib.rs:575-620).  1 = PolynomialSprott3Degree, the cubic member of the same
 * family: each coordinate continues the serial sum of lib.rs:588-600 with ten
 * more terms, coef3[k][0..9] times [x³, x²y, x²z, xy², xyz, xz², y³, y²z, yz², z³],
 * every cubic monomial formed as (quadratic monomial)*(variable): x³ = (x*x)*x,
 * x²y = (x*x)*y, x²z = (x*x)*z, xy² = (x*y)*y, xyz = (x*y)*z, xz² = (x*z)*z,
 * y³ = (y*y)*y, y²z = (y*y)*z, yz² = (y*z)*z, z³ = (z*z)*z. */
enum { SAR_ATTRACTOR_SPROTT2 = 0, SAR_ATTRACTOR_SPROTT3 = 1 };

/* Config<PolynomialSprott2Degree, {Function|AdjustedVelocity}>, lib.rs:265-287,
 * flattened with Colors (lib.rs:475-479), Palette (lib.rs:408-411),
 * BrighnessConstants (lib.rs:390-396), View (lib.rs:253-261) and
 * EulerAxisRotation (lib.rs:170-175). */
typedef struct sar_config {
    uint64_t iterations;                 /* lib.rs:267 (usize) */
    uint32_t width;                      /* lib.rs:269 */
    uint32_t height;                     /* lib.rs:271 */
    uint32_t render_kind;                /* lib.rs:273, SAR_RENDER_* */
    uint32_t transparent;                /* lib.rs:275, 0/1 */
    uint32_t silent;                     /* lib.rs:280, 0/1 */
    uint32_t ct_kind;                    /* lib.rs:286, SAR_CT_* */
    double   angle;                      /* lib.rs:277, RADIANS (lib.rs:745) */
    double   coef[3][10];                /* attractor.x / .y / .z, lib.rs:577-579 */
    double   center_camera[3];           /* view.center_camera, lib.rs:257 */
    double   axis[3];                    /* view.rotation.axis, lib.rs:172 — NOT normalised (release build, lib.rs:181-183) */
    double   rotation;                   /* view.rotation.rotation, lib.rs:174 */
    double   scale;                      /* view.scale, lib.rs:260 */
    double   ct_offset;                  /* AdjustedVelocity.offset, lib.rs:508 */
    double   ct_factor;                  /* AdjustedVelocity.factor, lib.rs:509 */
    uint32_t palette_len;                /* Palette::count(), lib.rs:435; 1..SAR_MAX_PALETTE */
    uint32_t attractor_kind;             /* SAR_ATTRACTOR_* (0 = the reference's PolynomialSprott2Degree) */
    double   palette_rgb[SAR_MAX_PALETTE][3]; /* Palette list WITHOUT the duplicated sentinel of lib.rs:418 */
    double   bright_offset;              /* colors.brighness.offset, lib.rs:394 */
    double   bright_factor;              /* colors.brighness.factor, lib.rs:395 */
    double   coef3[3][10];               /* cubic coefficients, SAR_ATTRACTOR_SPROTT3 only */
    double   ct_weights[4];              /* SAR_CT_SCREEN_BLEND only */
} sar_config;

typedef struct sar_runtime  sar_runtime;   /* Runtime, lib.rs:631-646 (device resident) */
typedef struct sar_renderer sar_renderer;  /* ParallelRenderer, lib.rs:908-915 */

/* ---- library ---------------------------------------------------------- */
uint32_t    sar_abi_version(void);
const char *sar_last_error(void);            /* thread-local, never NULL */
int         sar_device_count(int *count);    /* SAR_ERR_CUDA when no driver/device */
/* Default number of concurrent trajectory lanes on `device` (SM count × 896):
 * the GPU's answer to available_parallelism(), lib.rs:920-922. */
int         sar_default_threads(int device, uint32_t *threads);
/* Options.  Tuning knobs of the iterate kernel that never change results
 * (DESIGN.md §5): "traj_per_thread" (1, 2 or 4) — how many trajectories one GPU
 * thread carries side by side; "pipeline" (0 / 1) — make the depth test of an
 * iteration after the arithmetic of the next one; "tile_scatter" (0 / 1, default
 * 1) — images of at most 25 600 pixels accumulate in per-block shared-memory
 * tiles that are added into the global buffers at the end of the launch;
 * "sync_timeout_ms" — see the frame protocol below.  "diagnostic_mode": the product library accepts only 0;
 * the roofline-experiment variants of the iterate kernel (incomplete results by
 * design) exist only in the separately built libsar_b200_diag.so
 * (-DSAR_DIAGNOSTICS, tools/sweep_iterate.py) — SAR_ERR_UNSUPPORTED here. */
int         sar_set_option(const char *name, int64_t value);

/* ---- Config presets --------------------------------------------------- */
/* Config::new defaults, lib.rs:289-307 + Colors::default lib.rs:480-491, around
 * the caller's attractor/view/transform fields (which are left untouched). */
int sar_config_defaults(sar_config *cfg);
/* Config::poisson_saturne(), lib.rs:310-352 */
int sar_config_poisson_saturne(sar_config *cfg);
/* Config::solar_sail(), lib.rs:355-386 */
int sar_config_solar_sail(sar_config *cfg);

/* ---- start points ----------------------------------------------------- */
/* Stand-in for `runtime.rng.random::<Vec3>() * 0.1` (lib.rs:748, 161-166):
 * point k, coordinate c = ((u >> 11) * 2^-53) * 0.1 with u the (3k+c)-th
 * output of a SplitMix64 stream seeded with `seed`; written to
 * out_xyz[3*(k-first) + c] for k in [first, first+n). */
int sar_seed_points(uint64_t seed, uint64_t first, uint64_t n, double *out_xyz);

/* ---- auto-framing first pass ------------------------------------------ */
/* The reference author's TODO (lib.rs:326-334): "Add option to make first-pass
 * to get these values, to then compute center_camera".  n_jobs trajectories
 * (start points init_xyz, or sar_seed_points(seed) when NULL), each 1000
 * warm-up steps then `iterations` steps; box = {xmin,xmax,ymin,ymax,zmin,zmax}
 * of screen_space = R·p (lib.rs:773) — the quantity of the comment's table —
 * over every trajectory that stays bounded; `diverged` counts the others (for
 * solar_sail ~38 % of the start points; they still render, into the NaN sink).
 * center_camera = minus the box mid-points in the pairing the projection uses
 * (x, screen z, screen y; lib.rs:776-786); scale = 0.95 x the largest scale
 * that keeps the box in view from every view angle at cfg's aspect ratio
 * (cfg->width/height; 1:1 when 0).  The box equals the oracle's bit for bit. */
typedef struct sar_autoframe_result {
    double   box[6];
    double   center_camera[3];
    double   scale;
    uint64_t diverged;
    uint64_t n_jobs;
} sar_autoframe_result;
int sar_autoframe(const sar_config *cfg, int device, uint64_t seed, const double *init_xyz,
                  uint64_t n_jobs, uint64_t iterations, sar_autoframe_result *out);

/* ---- Runtime ---------------------------------------------------------- */
/* Runtime::new, lib.rs:660 (allocate + reset) on CUDA device `device`. */
int  sar_runtime_new(uint32_t width, uint32_t height, int device, sar_runtime **out);
void sar_runtime_free(sar_runtime *rt);
/* Runtime::reset, lib.rs:682: count=0, steps=0.0, zbuf=-1.0, max=0. */
int  sar_runtime_reset(sar_runtime *rt);
/* Runtime::merge, lib.rs:708: count+=, max, `other` wins a pixel iff its z is
 * strictly greater (ties keep dst).  SAR_ERR_DIMS on mismatch.  src may live
 * on another device. */
int  sar_runtime_merge(sar_runtime *dst, const sar_runtime *src);
int  sar_runtime_dims(const sar_runtime *rt, uint32_t *width, uint32_t *height, int *device);
/* The three textures + max in the reference's layout (row-major, idx=y*w+x,
 * as image::ImageBuffer; lib.rs:633-643).  Any pointer may be NULL. */
int  sar_runtime_download(const sar_runtime *rt, uint32_t *count, double *steps,
                          float *zbuf, uint32_t *max);
/* Inverse of download (all three arrays required): the checkpoint/resume of
 * the reference's progressive accumulation (lib.rs:742-743). */
int  sar_runtime_upload(sar_runtime *rt, const uint32_t *count, const double *steps,
                        const float *zbuf);

/* ---- render() --------------------------------------------------------- */
/* n_jobs reference render() calls (lib.rs:747) accumulated into `rt`: job k
 * starts at init_xyz[3k..3k+3], runs SAR_WARMUP_ITERATIONS unrecorded steps and
 * then cfg->iterations recorded steps.  Does not reset `rt`.  Blocking. */
int sar_render(const sar_config *cfg, sar_runtime *rt,
               const double *init_xyz, uint64_t n_jobs);
/* Same, start points = sar_seed_points(seed, first_job, n_jobs) generated on
 * the device (no host→device traffic). */
int sar_render_seeded(const sar_config *cfg, sar_runtime *rt,
                      uint64_t seed, uint64_t first_job, uint64_t n_jobs);

/* ---- colorize() ------------------------------------------------------- */
/* colorize, lib.rs:841: interleaved RGBA u16 (FinalImage::into_raw, lib.rs:625),
 * width*height*4 values into caller memory.  rgba_f32 (optional, may be NULL)
 * receives the pre-quantisation channel values `(c*factor+offset)*bf` and
 * alpha as f32 (Gas) or z/1/1/1-normalised grey (Depth). */
int sar_colorize(const sar_config *cfg, const sar_runtime *rt,
                 uint16_t *rgba_u16, float *rgba_f32);

/* ---- ParallelRenderer / render_parallel ------------------------------- */
/* ParallelRenderer::new, lib.rs:919.  `devices`/`n_devices`: CUDA ordinals of
 * this process's GPUs (NULL/0 = device 0).  `threads_per_device` plays the
 * role of available_parallelism() (lib.rs:920-922): the number of concurrent
 * trajectory lanes; 0 = default (sar_default_threads). */
int  sar_renderer_new(const int *devices, int n_devices, uint32_t threads_per_device,
                      sar_renderer **out);
/* ParallelRenderer::shutdown, lib.rs:1020. */
void sar_renderer_shutdown(sar_renderer *r);
/* num_threads(), lib.rs:1015 — total over the renderer's devices.  With an
 * explicit threads_per_device it is fixed.  In auto mode (0) it depends on
 * jobs_per_thread: every job already gets its own lane and all jobs have the
 * same length, so the renderer keeps the number of JOBS at the device's lane
 * count: num_threads = lanes / jobs_per_thread (multiple of 32, >= 32), and
 * num_threads × jobs_per_thread jobs of iterations/num_threads/jobs_per_thread
 * steps run, exactly as lib.rs:1058-1062 prescribes for that num_threads. */
int  sar_renderer_num_threads(const sar_renderer *r, uint64_t *num_threads);   /* jobs_per_thread = 1 */
int  sar_renderer_num_threads_for(const sar_renderer *r, uint64_t jobs_per_thread, uint64_t *num_threads);
/* The decomposition sar_render_parallel will use for a frame of `iterations`:
 * num_threads and iterations/num_threads/jobs_per_thread (lib.rs:1058).  In
 * auto mode num_threads is additionally capped so that every job gets at least
 * 64 recorded steps (multiple of 32, >= 32): a GPU has ~10^5 lanes where the
 * reference has ~10 threads, and a small render — the reference's default is
 * 1e7 iterations — must not round to 0 steps per job.  Either output may be NULL. */
int  sar_renderer_plan(const sar_renderer *r, uint64_t iterations, uint64_t jobs_per_thread,
                       uint64_t *num_threads, uint64_t *iterations_per_job);
/* render_parallel, lib.rs:1051: iterations/num_threads/jobs_per_thread per job
 * (integer division, lib.rs:1058), num_threads*jobs_per_thread jobs
 * (lib.rs:1062), merge (lib.rs:1072-1076), colorize (lib.rs:1080).
 * Start points: init_xyz (n_jobs×3) if non-NULL, else sar_seed_points(seed).
 * rgba_u16: width*height*4, caller-owned host memory.  Blocking. */
int  sar_render_parallel(sar_renderer *r, const sar_config *cfg, uint64_t jobs_per_thread,
                         uint64_t seed, const double *init_xyz, uint16_t *rgba_u16);
/* A sequence of frames: the per-frame loop of the reference's binary around
 * render_parallel (src/bin/main.rs:496-512; angles as AngleIter yields them,
 * main.rs:107-176, in RADIANS).  Frame f is a one-device render_parallel of
 * `cfg` with angle = angles_rad[f]; frames round-robin over the renderer's
 * devices (independent replicas) and each frame's device→host copy overlaps
 * the next frame's render.  Start points: by default frame f takes the next
 * num_threads*jobs_per_thread points of the seed stream (fresh points every
 * frame, like the reference).  With SAR_SEQ_SHARED_POINTS every frame uses
 * points [0, jobs): the trajectories are then identical across frames and the
 * 1000-step warm-up (lib.rs:750-752) runs once for the whole sequence.
 * Output: rgba_frames (n_frames × width*height*4 u16, caller-owned, may be
 * NULL) and/or cb(user, frame, pixels), called in frame order from the calling
 * thread; without rgba_frames the pixels pointer is only valid inside cb. */
#define SAR_SEQ_SHARED_POINTS 1u
typedef void (*sar_frame_callback)(void *user, uint32_t frame, const uint16_t *rgba_u16);
int  sar_render_sequence(sar_renderer *r, const sar_config *cfg, const double *angles_rad,
                         uint32_t n_frames, uint64_t jobs_per_thread, uint64_t seed, uint32_t flags,
                         uint16_t *rgba_frames, sar_frame_callback cb, void *user);
/* ---- output conversion + raw encoders (src/bin/main.rs:40-100) ----------
 * write_image_matches converts the FinalImage by (transparent, 8bit)
 * (main.rs:52-57): RGBA16 as is, to_rgb16(), to_rgba8(), to_rgb8(), and hands
 * image.as_bytes() to an `image`-crate encoder (PAM main.rs:62-68, BMP
 * main.rs:70-76, PNG main.rs:78-89).  Here the conversion runs on the device
 * and the containers are written around it; the PNG branch exists both with
 * the compressor (sar_runtime_encode_png, below) and without (SAR_FILE_PNG).
 * Third-party arithmetic: the u16 -> u8 narrowing of to_rgba8/to_rgb8 is
 * image 0.25's `FromPrimitive<u16> for u8`, (c + 128) / 257 =
 * round(c*255/65535); that crate is not vendored in the reference, so this is
 * a restatement of its published source — parity unpinned (third-party).
 * Bytes produced:
 *   SAR_FILE_RAW  samples in native little-endian order, rows top-down
 *   SAR_FILE_PAM  "P7\nWIDTH w\nHEIGHT h\nDEPTH d\nMAXVAL m\nTUPLTYPE RGB|RGB_ALPHA\nENDHDR\n"
 *                 + samples, 16-bit ones most significant byte first
 *   SAR_FILE_BMP  8-bit formats only (the reference's BmpEncoder panics on a
 *                 16-bit image): BITMAPINFOHEADER 24 bpp / BITMAPV4HEADER 32 bpp
 *                 BI_BITFIELDS, B,G,R[,A], rows bottom-up padded to 4 bytes
 *   SAR_FILE_PNG  a complete PNG (main.rs:78-89) WITHOUT the compressor: IHDR, one
 *                 IDAT whose zlib stream holds the filter-type-0 scanlines in
 *                 "stored" deflate blocks, IEND.  Decodes to exactly the pixels
 *                 the reference's PNG decodes to; the file is as large as the raw
 *                 image (the reference deflates with CompressionType::Default —
 *                 feed SAR_FILE_RAW bytes to a PNG encoder for that).  Scanline
 *                 packing and per-chunk CRC-32 / Adler-32 partial sums run on the
 *                 device; the host folds the sums into the 20 trailing bytes. */
enum { SAR_PIX_RGBA16 = 0, SAR_PIX_RGB16 = 1, SAR_PIX_RGBA8 = 2, SAR_PIX_RGB8 = 3 };
/* SAR_FILE_PNG_DEFLATE: the compressed PNG of sar_runtime_encode_png as a frame format of sar_render_sequence_encoded
 * (frames differ in size: callback only, frames_out must be NULL); the fixed-size entry points reject it. */
enum { SAR_FILE_RAW = 0, SAR_FILE_PAM = 1, SAR_FILE_BMP = 2, SAR_FILE_PNG = 3, SAR_FILE_PNG_DEFLATE = 4 };
/* bytes of one encoded image (header + pixels); 0 for an unsupported combination */
size_t sar_encoded_size(uint32_t width, uint32_t height, uint32_t pixel_format, uint32_t container);
/* the container header alone (host only, no GPU); out may be NULL to query header_bytes */
int  sar_encode_header(uint32_t width, uint32_t height, uint32_t pixel_format, uint32_t container,
                       uint8_t *out, size_t out_bytes, size_t *header_bytes);
/* Convert the Runtime's device-resident image (the result of the last colorize
 * on it) and copy header + pixels to `out` (sar_encoded_size bytes).  Blocking. */
int  sar_runtime_encode(sar_runtime *rt, uint32_t pixel_format, uint32_t container,
                        uint8_t *out, size_t out_bytes, void *stream);
/* A complete, COMPRESSED PNG of the Runtime's device-resident image — what the reference's PngEncoder
 * (main.rs:78-89, CompressionType::Default, adaptive filter) is for: a file of a few MB instead of the raw image.
 * The compressor runs on the device (csrc/sar_deflate.cu): Sub-filtered scanlines (so the untouched part of a frame is
 * zeros), one deflate block per 16 KB with run-length matches and its own dynamic Huffman code, blocks made
 * independently by one warp each and joined on byte boundaries; CRC-32 / Adler-32 partial sums on the device, folded
 * by the host.  A deflate stream is specified by what it decodes to (RFC 1950/1951), not by its bytes: the file
 * decodes to exactly the pixels of the reference's PNG; its size is within a few per cent of the reference's
 * (reference poisson-saturne.png 3.63 MB, this 3.72 MB).  Pixel formats as above (16-bit samples big-endian).
 * sar_png_bound: worst-case bytes for out_capacity (0 for an unsupported size); *out_bytes: bytes written.  Blocking. */
size_t sar_png_bound(uint32_t width, uint32_t height, uint32_t pixel_format);
int  sar_runtime_encode_png(sar_runtime *rt, uint32_t pixel_format, uint8_t *out, size_t out_capacity,
                            size_t *out_bytes, void *stream);
/* File::create(path) + write_all, main.rs:102-104 */
int  sar_write_file(const char *path, const uint8_t *bytes, size_t n_bytes);
/* sar_render_sequence with every frame converted on the device and handed over
 * encoded: frames_out (n_frames x sar_encoded_size bytes, may be NULL) and/or
 * cb(user, frame, bytes, n_bytes) in frame order — what the reference's encoder
 * side threads (main.rs:508-511) would write to disk. */
typedef void (*sar_frame_bytes_callback)(void *user, uint32_t frame, const uint8_t *bytes, size_t n_bytes);
int  sar_render_sequence_encoded(sar_renderer *r, const sar_config *cfg, const double *angles_rad,
                                 uint32_t n_frames, uint64_t jobs_per_thread, uint64_t seed, uint32_t flags,
                                 uint32_t pixel_format, uint32_t container, uint8_t *frames_out,
                                 sar_frame_bytes_callback cb, void *user);

/* The merged Runtime of the last sar_render_parallel on the renderer's first
 * device (valid until the next call / shutdown); for inspection and tests. */
int  sar_renderer_runtime(sar_renderer *r, sar_runtime **rt);

/* ---- device-resident / multi-process plumbing -------------------------
 * Used by bench.py (kernel-only timing) and by the one-process-per-GPU
 * driver.  `stream` is a cudaStream_t passed as void*
 * (NULL = the runtime's own stream); *_async calls do not synchronise.
 * `threads` = concurrent trajectory lanes (0 = default, sar_default_threads);
 * lane L runs jobs L, L+threads, ... one after the other.
 * `first_job` positions the call in the seed stream only.  Order key of job k
 * of a call = job_base + k, and the call advances job_base by n_jobs
 * (sar_runtime_set_job_base; reset() zeroes it); earlier keys keep exact z
 * ties.  Keys are 32-bit: a call that would pass 2^32 jobs since the last reset
 * fails with SAR_ERR_INVALID instead of letting late jobs share a key. */
int sar_render_seeded_async(const sar_config *cfg, sar_runtime *rt, uint64_t seed,
                            uint64_t first_job, uint64_t n_jobs, uint32_t threads,
                            void *stream);
/* d_init_xyz: DEVICE pointer to this call's n_jobs×3 start points. */
int sar_render_device_async(const sar_config *cfg, sar_runtime *rt, const double *d_init_xyz,
                            uint64_t first_job, uint64_t n_jobs, uint32_t threads, void *stream);
int sar_runtime_reset_async(sar_runtime *rt, void *stream);
/* Bring Runtime.max (lib.rs:643) up to date in the runtime's device scalars, over
 * rows [row0,row0+rows) (rows=0: whole image).  The render kernel keeps a running
 * max like lib.rs:813-815 does, so for a whole image whose counts all came from
 * render calls since the last reset this is a one-thread kernel (it folds the NaN
 * debt of pixel (0,0) in); otherwise a reduction over the accumulators.  The Depth
 * min/max (lib.rs:877-882) is reduced by the colourise calls when a Depth image is
 * asked for. */
int sar_runtime_max_async(sar_runtime *rt, uint32_t row0, uint32_t rows, void *stream);
/* Read back / force Runtime.max (synchronises `stream`). */
int sar_runtime_get_max(sar_runtime *rt, uint32_t *max_out, void *stream);
int sar_runtime_set_max(sar_runtime *rt, uint32_t max, void *stream);
/* colorize rows [row0,row0+rows) (rows=0: all) with the CURRENT device max into
 * the device-resident image of `dst` (NULL = rt's own; a peer = remote store
 * over NVLink).  sar_runtime_image_download copies image rows to the host. */
typedef struct sar_peer sar_peer;
int sar_colorize_rows_async(const sar_config *cfg, sar_runtime *rt, uint32_t row0, uint32_t rows,
                            sar_peer *dst, void *stream);
int sar_runtime_image_download(sar_runtime *rt, uint32_t row0, uint32_t rows, uint16_t *rgba_u16,
                               void *stream);
int sar_stream_synchronize(sar_runtime *rt, void *stream);
/* job order counter of the runtime (reset() zeroes it).  A multi-rank frame sets it to the
 * rank's first global job index before rendering, so that keys are global over the ranks. */
int sar_runtime_get_job_base(const sar_runtime *rt, uint64_t *job_base);
int sar_runtime_set_job_base(sar_runtime *rt, uint64_t job_base);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t sar_launch_count(void);

/* Pinned host memory for callers that want full-rate PCIe copies. */
int  sar_host_alloc(size_t bytes, void **out);
void sar_host_free(void *p);

/* Cross-process peer access (one process per GPU, NVLink P2P via CUDA IPC).
 * A Runtime's accumulators and image live in ONE device allocation; export
 * gives its 64-byte cudaIpcMemHandle and each rank opens every peer's handle
 * (the frame protocol below reads and writes peer memory through them).
 * sar_runtime_merge_peers_async() is the plain building block: it reduces rows
 * [row0,row0+rows) of `rt` with the given Runtimes' accumulators by direct
 * loads — counts add, the Δp record with the greatest (z, earlier job) wins,
 * the deterministic form of Runtime::merge (lib.rs:708-738) — with no
 * synchronisation of its own (the in-process multi-device renderer uses it). */
#define SAR_IPC_HANDLE_BYTES 64u
int  sar_runtime_ipc_export(const sar_runtime *rt, uint8_t out[SAR_IPC_HANDLE_BYTES]);
int  sar_peer_open(const uint8_t handle[SAR_IPC_HANDLE_BYTES], uint32_t width, uint32_t height,
                   int local_device, sar_peer **out);
void sar_peer_close(sar_peer *p);
int  sar_runtime_merge_peers_async(sar_runtime *rt, sar_peer *const *peers, int n_peers,
                                   uint32_t row0, uint32_t rows, void *stream);
/* One frame of render_parallel (lib.rs:1051-1082) over N ranks (one process per
 * GPU), with NO host round trip: rank r renders its slice of the jobs into its
 * own Runtime, then reduces and colourises ROW STRIPE r of the frame.  Every
 * Runtime holds, inside its exported allocation, one flag per (event kind,
 * source rank); a rank announces an event by storing the frame epoch into that
 * flag on its targets (remote stores over NVLink) and waits by polling its own
 * memory.  Waits and signals are the prologues / epilogues of the kernels below,
 * not launches of their own.  `peers_by_rank[r]` = sar_peer of rank r (entry
 * my_rank is ignored, may be NULL); `epoch` = frame number, starting at 1 and
 * growing by 1 per frame on every rank.  Per frame, on one stream, each rank
 * enqueues:
 *   sar_frame_reset_async      wait until every peer has finished reading this
 *                              rank's accumulators of frame epoch-1, then
 *                              Runtime::reset (lib.rs:682; job_base = 0)
 *   sar_runtime_set_job_base(first global job of the rank) + a render call
 *   sar_frame_export_async     counts -> pixel-order u32 array peers can read
 *                              with coalesced loads; RENDER_DONE to every rank
 *   sar_frame_merge_async      wait RENDER_DONE from all; rows [row0,row0+rows):
 *                              counts add, the record with the greatest
 *                              (z, earlier job) wins — Runtime::merge
 *                              (lib.rs:708-738) made order-independent; the
 *                              stripe's share of Runtime.max and of the Depth
 *                              min/max (lib.rs:877-882) goes to every rank;
 *                              MAX_READY + MERGE_DONE
 *   sar_frame_colorize_async   wait MAX_READY from all (the log base of
 *                              lib.rs:860 and the Depth range are global) and
 *                              IMAGE_FREE(epoch-1) from the owner; colorize
 *                              (lib.rs:841) of the stripe straight into rank
 *                              `owner_rank`'s image; IMAGE_DONE to the owner
 * and the owner additionally
 *   sar_frame_image_wait_async wait IMAGE_DONE from all: the image is complete
 *   (sar_runtime_image_download, or any use of the image on `stream`)
 *   sar_frame_image_release_async  IMAGE_FREE to every rank.
 * The N-rank frame is bit-identical to the same job list rendered on one GPU.
 * A wait that sees no progress for "sync_timeout_ms" (sar_set_option, default
 * 10 000) records an error instead of hanging the GPU; every later kernel of the
 * protocol on that Runtime then does nothing.  sar_runtime_sync_error reads the
 * error (0 = none, else 1 + the event kind that timed out) and optionally clears
 * it: check it whenever a frame's result is consumed. */
int  sar_frame_reset_async(sar_runtime *rt, int n_ranks, uint32_t epoch, void *stream);
int  sar_frame_export_async(sar_runtime *rt, sar_peer *const *peers_by_rank, int n_ranks, int my_rank,
                            uint32_t epoch, void *stream);
int  sar_frame_merge_async(sar_runtime *rt, sar_peer *const *peers_by_rank, int n_ranks, int my_rank,
                           uint32_t row0, uint32_t rows, uint32_t epoch, void *stream);
int  sar_frame_colorize_async(const sar_config *cfg, sar_runtime *rt, sar_peer *const *peers_by_rank,
                              int n_ranks, int my_rank, int owner_rank, uint32_t row0, uint32_t rows,
                              uint32_t epoch, void *stream);
int  sar_frame_image_wait_async(sar_runtime *rt, int n_ranks, uint32_t epoch, void *stream);
int  sar_frame_image_release_async(sar_runtime *rt, sar_peer *const *peers_by_rank, int n_ranks, int my_rank,
                                   uint32_t epoch, void *stream);
int  sar_runtime_sync_error(sar_runtime *rt, uint32_t *error, int clear);

#ifdef __cplusplus
}
#endif
#endif /* SAR_B200_H */
