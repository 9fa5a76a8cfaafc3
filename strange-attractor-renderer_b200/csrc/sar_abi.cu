// sar_abi.cu — the extern "C" boundary declared in include/sar.h, and the host-side logic
// behind it: Config → kernel constants (the reference does this at lib.rs:754-764 at the
// top of render()), Runtime ownership (lib.rs:631-699), render()/colorize()/
// render_parallel() orchestration (lib.rs:747, 841, 1051).  No CPU compute fallback:
// every entry point that needs a GPU fails with SAR_ERR_CUDA when there is none.
#include "../../include/sar.h"
#include "sar_device.cuh"
#include "sar_deflate.cuh"

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace sar;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err = "";

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define SAR_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) {                                                                    \
            (void)cudaGetLastError();                                                               \
            return fail(e_ == cudaErrorMemoryAllocation ? SAR_ERR_NOMEM : SAR_ERR_CUDA, "%s: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                \
        }                                                                                           \
    } while (0)

// ---------------------------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------------------------
struct sar_runtime {
    int device = 0;
    uint32_t w = 0, h = 0;
    size_t npix = 0;
    size_t nslots = 0;               // power of two >= npix: size of the scrambled `fast` array
    SlotMap slots = {1u, 0u};
    // one device allocation: rec | fast | image | cnt | scal   (so one IPC handle exports it all)
    void *block = nullptr;
    size_t block_bytes = 0;
    ulonglong2 *rec = nullptr;
    unsigned long long *fast = nullptr;
    uint16_t *image = nullptr;       // RGBA u16, FinalImage (lib.rs:625)
    uint32_t *cnt = nullptr;         // counts in pixel order: what peers read in the multi-GPU exchange
    Scalars *scal = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;            // render_parallel: stripe-wise device→host copy behind the colourise
    cudaEvent_t stripe_done[4] = {nullptr, nullptr, nullptr, nullptr};
    uint64_t job_base = 0;
    uint32_t host_max = 0;           // Runtime.max as last read back / forced by the host ...
    bool host_max_valid = false;     // ... valid until the accumulators change
    bool max_tracked = false;        // every count change since the last reset came from the iterate kernel, which keeps
                                     // scal->max current: the max "reduction" is then one thread (launch_fold_max)
    bool depth_valid = false;        // scal->zmax_key / zmin_key (Depth fold, lib.rs:877-882) match the accumulators
    int sm_count = 148;
    // lazily allocated scratch
    double *d_init = nullptr; size_t d_init_cap = 0;
    void *d_scratch = nullptr; size_t d_scratch_cap = 0;
};

struct sar_peer {
    int local_device = 0;
    uint32_t w = 0, h = 0;
    SlotMap slots = {1u, 0u};
    void *block = nullptr;           // cudaIpcOpenMemHandle mapping (or a borrowed in-process pointer)
    bool ipc = false;
    ulonglong2 *rec = nullptr;
    unsigned long long *fast = nullptr;
    uint16_t *image = nullptr;
    uint32_t *cnt = nullptr;
    Scalars *scal = nullptr;
};

struct seq_device {                  // per-device pipeline state of sar_render_sequence
    sar_runtime *rt[2] = {nullptr, nullptr};        // two Runtimes, frames alternate: frame f renders while frame f-1 is
                                                    // colourised with ITS max read back by the host and copied out
    uint8_t *stage[2] = {nullptr, nullptr};         // pinned host staging (when the caller gives no frame array)
    size_t stage_bytes = 0;
    uint8_t *enc[2] = {nullptr, nullptr};           // device: converted pixels (formats other than RGBA16 native) [+ PNG partial sums]
    size_t enc_bytes = 0;
    uint8_t *h_sums[2] = {nullptr, nullptr};        // pinned: PNG partial checksums of the frame in each slot
    size_t sums_cap = 0;
    uint32_t *h_max = nullptr;                      // pinned: Runtime.max of the frame in each slot
    cudaEvent_t max_ready[2] = {nullptr, nullptr}, rendered[2] = {nullptr, nullptr}, copied[2] = {nullptr, nullptr};
    cudaStream_t copy_stream = nullptr;
    cudaStream_t pay_stream = nullptr;              // compressed PNG: the exact-size copy of a finished frame's stream
    double *warm = nullptr; size_t warm_cap = 0;    // warmed states shared by all frames (SAR_SEQ_SHARED_POINTS)
    size_t img_bytes = 0;
};

struct sar_renderer {
    std::vector<int> devices;
    std::vector<sar_runtime *> rts;
    std::vector<seq_device> seq;
    uint32_t threads_per_device = 0;
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t slots_for(size_t npix)
{
    size_t p = 1;
    while (p < npix) p <<= 1;
    return p;
}

// scramble the fast array only while the scrambled working set still fits the L2 comfortably
static SlotMap slotmap_for(size_t npix)
{
    SlotMap m;
    m.mask = (uint32_t)(slots_for(npix) - 1);
    m.mult = npix <= ((size_t)1 << 23) ? SLOT_SCRAMBLE : 1u;
    return m;
}

struct Layout { size_t fast, image, cnt, scal, total; };   // rec sits at offset 0
static Layout layout(size_t npix)
{
    Layout l;
    l.fast = align_up(npix * sizeof(ulonglong2), 256);
    l.image = l.fast + align_up(slots_for(npix) * sizeof(unsigned long long), 256);
    l.cnt = l.image + align_up(npix * 4 * sizeof(uint16_t), 256);
    l.scal = l.cnt + align_up(npix * sizeof(uint32_t) + 16, 256);      // pixel-order counts of the multi-GPU exchange (+ one group of slack)
    l.total = l.scal + 1024;
    return l;
}

static cudaStream_t pick(const sar_runtime *rt, void *stream) { return stream ? (cudaStream_t)stream : rt->stream; }

static int check_dims(uint32_t w, uint32_t h)
{
    if (w == 0 || h == 0) return fail(SAR_ERR_INVALID, "width and height must be non-zero (got %ux%u)", w, h);
    if ((unsigned long long)w * h > (1ull << 31)) return fail(SAR_ERR_INVALID, "width*height must be <= 2^31 (got %ux%u)", w, h);
    return SAR_OK;
}

static int check_config(const sar_config *cfg, const sar_runtime *rt)
{
    if (!cfg) return fail(SAR_ERR_INVALID, "cfg is NULL");
    if (cfg->palette_len == 0 || cfg->palette_len > SAR_MAX_PALETTE)   // Palette::new panics on an empty list, lib.rs:415-418
        return fail(SAR_ERR_INVALID, "palette_len must be in 1..%u (got %u)", SAR_MAX_PALETTE, cfg->palette_len);
    if (cfg->ct_kind > SAR_CT_SCREEN_BLEND)
        return fail(SAR_ERR_UNSUPPORTED, "ct_kind %u has no device implementation", cfg->ct_kind);
    if (cfg->attractor_kind > SAR_ATTRACTOR_SPROTT3)
        return fail(SAR_ERR_UNSUPPORTED, "attractor_kind %u has no device implementation", cfg->attractor_kind);
    if (cfg->render_kind > SAR_RENDER_DEPTH) return fail(SAR_ERR_INVALID, "render_kind %u", cfg->render_kind);
    if (rt && (cfg->width != rt->w || cfg->height != rt->h))
        return fail(SAR_ERR_DIMS, "config is %ux%u but runtime is %ux%u", cfg->width, cfg->height, rt->w, rt->h);
    return SAR_OK;
}

// ln(n), n < LNLUT_LEN, from the host libm (see ColorParams); one table per device, built once.
static const unsigned int LNLUT_LEN = 1u << 20;
static double *g_lnlut[64] = {nullptr};
static std::mutex g_lnlut_mutex;
static int ensure_lnlut(int device)
{
    if (device < 0 || device >= 64) return fail(SAR_ERR_INVALID, "device ordinal %d out of range", device);
    std::lock_guard<std::mutex> lock(g_lnlut_mutex);
    if (g_lnlut[device]) return SAR_OK;
    std::vector<double> h(LNLUT_LEN);
    for (unsigned int n = 0; n < LNLUT_LEN; ++n) h[n] = std::log((double)n);     // f64::ln, lib.rs:860
    SAR_CUDA(cudaSetDevice(device));
    double *d = nullptr;
    SAR_CUDA(cudaMalloc((void **)&d, LNLUT_LEN * sizeof(double)));
    SAR_CUDA(cudaMemcpy(d, h.data(), LNLUT_LEN * sizeof(double), cudaMemcpyHostToDevice));
    g_lnlut[device] = d;
    return SAR_OK;
}


// ---------------------------------------------------------------------------------------------
// Config → kernel constants.  Host libm sin/cos so that the oracle and the device consume the
// same f64 values (the reference calls the same glibc functions through Rust's f64::sin/cos).
// ---------------------------------------------------------------------------------------------
static void rotation_matrix(const double ax[3], double rot, double m[3][3])   // lib.rs:179-195, release: no normalize
{
    const double x = ax[0], y = ax[1], z = ax[2];
    const double c = std::cos(rot);
    const double c1 = 1. - c;
    const double s = std::sin(rot);
    m[0][0] = c + x * x * c1;     m[0][1] = x * y * c1 - z * s; m[0][2] = x * z * c1 + y * s;
    m[1][0] = y * x * c1 + z * s; m[1][1] = c + y * y * c1;     m[1][2] = y * z * c1 - x * s;
    m[2][0] = z * x * c1 - y * s; m[2][1] = z * y * c1 + x * s; m[2][2] = c + z * z * c1;
}

static void make_iter_params(const sar_config *cfg, sar_runtime *rt, IterParams &p)
{
    memset(&p, 0, sizeof p);
    for (int k = 0; k < 3; ++k) {
        for (int i = 0; i < 10; ++i) p.c[k][i] = cfg->coef[k][i];
        volatile double zero = 0.0, one = 1.0;                  // `sum = 0.; sum += 1. * c0`, lib.rs:589-603
        p.c[k][0] = zero + one * cfg->coef[k][0];
    }
    rotation_matrix(cfg->axis, cfg->rotation, p.m);                              // lib.rs:755
    p.sv = std::sin(cfg->angle);                                                 // lib.rs:756
    p.cv = std::cos(cfg->angle);                                                 // lib.rs:757
    p.ccx = cfg->center_camera[0]; p.ccy = cfg->center_camera[1]; p.ccz = cfg->center_camera[2];
    const double width = (double)cfg->width, height = (double)cfg->height;       // lib.rs:760-762
    p.ws = width * cfg->scale;                                                   // lib.rs:763
    p.sam = 0.5 / cfg->scale;                                                    // lib.rs:764
    p.half_h = height / 2.;                                                      // lib.rs:786
    p.ct_offset = cfg->ct_offset; p.ct_factor = cfg->ct_factor;
    for (int i = 0; i < 4; ++i) p.ct_w[i] = cfg->ct_weights[i];
    for (int k = 0; k < 3; ++k) for (int i = 0; i < 10; ++i) p.c3[k][i] = cfg->coef3[k][i];
    p.attractor_kind = cfg->attractor_kind;
    p.fast = rt->fast; p.rec = rt->rec; p.scal = rt->scal;
    p.W = cfg->width; p.H = cfg->height; p.ct_kind = cfg->ct_kind;
    p.slots = rt->slots;
    p.iterations = cfg->iterations;
    p.warmup = SAR_WARMUP_ITERATIONS;                                            // lib.rs:750
#ifdef SAR_DIAGNOSTICS
    p.diag_hot = (unsigned int)get_diag_hot(); p.diag_tab_entries = 2048;
#endif
}

static void make_color_params(const sar_config *cfg, const sar_runtime *rt, ColorParams &c, uint32_t row0, uint32_t rows,
                              const uint32_t *host_max = nullptr)
{
    memset(&c, 0, sizeof c);
    c.slots = rt->slots;
    c.lnlut = g_lnlut[rt->device];
    c.lnlut_len = c.lnlut ? LNLUT_LEN : 0u;
    if (host_max) {                                              // blocking callers read max back: exact ln(max+1) from the host libm
        c.host_lnmax_valid = 1u;
        c.ln_max1_host = std::log((double)(uint32_t)(*host_max + 1u));
    }
    for (uint32_t i = 0; i < cfg->palette_len; ++i)
        for (int k = 0; k < 3; ++k) c.pal[i][k] = cfg->palette_rgb[i][k];
    for (int k = 0; k < 3; ++k) c.pal[cfg->palette_len][k] = cfg->palette_rgb[cfg->palette_len - 1][k];  // lib.rs:418
    c.pal_len = (double)cfg->palette_len;                                        // lib.rs:421
    c.bright_offset = cfg->bright_offset; c.bright_factor = cfg->bright_factor;
    c.palette_len = cfg->palette_len; c.transparent = cfg->transparent; c.render_kind = cfg->render_kind;
    c.W = cfg->width; c.H = cfg->height; c.row0 = row0; c.rows = rows;
}

// Concurrent trajectory lanes per SM: 7 warps per scheduler is where the L2 atomic-with-return
// rate saturates (profiles/r1_sweep.md); more lanes only add 1000-step warm-ups (lib.rs:750).
static const uint32_t LANES_PER_SM = 896u;
static uint32_t default_lanes(const sar_runtime *rt) { return (uint32_t)rt->sm_count * LANES_PER_SM; }

// ---------------------------------------------------------------------------------------------
extern "C" {

uint32_t sar_abi_version(void) { return SAR_ABI_VERSION; }
const char *sar_last_error(void) { return g_err.c_str(); }
uint64_t sar_launch_count(void) { return launch_count(); }

int sar_set_option(const char *name, int64_t value)
{
    if (!name) return fail(SAR_ERR_INVALID, "name is NULL");
    if (strcmp(name, "traj_per_thread") == 0) {
        if (!set_traj_per_thread((int)value)) return fail(SAR_ERR_INVALID, "traj_per_thread must be 1, 2 or 4");
        return SAR_OK;
    }
    if (strcmp(name, "sync_timeout_ms") == 0) {
        if (value < 1 || value > 3600000) return fail(SAR_ERR_INVALID, "sync_timeout_ms must be in 1..3600000");
        set_sync_timeout_ms(value);
        return SAR_OK;
    }
    if (strcmp(name, "tile_scatter") == 0) {
        if (!set_tile_scatter((int)value)) return fail(SAR_ERR_INVALID, "tile_scatter must be 0 or 1");
        return SAR_OK;
    }
    if (strcmp(name, "pipeline") == 0) {
        if (!set_pipeline((int)value)) return fail(SAR_ERR_INVALID, "pipeline must be 0 or 1");
        return SAR_OK;
    }
#ifdef SAR_DIAGNOSTICS
    if (strcmp(name, "diag_hot") == 0) { set_diag_hot((int)value); return SAR_OK; }
#endif
    if (strcmp(name, "diagnostic_mode") == 0) {
#ifdef SAR_DIAGNOSTICS
        if (!set_mode((int)value)) return fail(SAR_ERR_INVALID, "diagnostic_mode must be 0, 1, 2, 4 or 5");
        return SAR_OK;
#else
        if (value == 0) return SAR_OK;
        return fail(SAR_ERR_UNSUPPORTED, "this build has no diagnostic kernels (build with -DSAR_DIAGNOSTICS: libsar_b200_diag.so)");
#endif
    }
    return fail(SAR_ERR_INVALID, "unknown option '%s'", name);
}

int sar_default_threads(int device, uint32_t *threads)
{
    if (!threads) return fail(SAR_ERR_INVALID, "threads is NULL");
    int sm = 0;
    SAR_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device));
    *threads = (uint32_t)sm * LANES_PER_SM;
    return SAR_OK;
}

int sar_device_count(int *count)
{
    if (!count) return fail(SAR_ERR_INVALID, "count is NULL");
    *count = 0;
    SAR_CUDA(cudaGetDeviceCount(count));
    return SAR_OK;
}

// ---- presets (restated from lib.rs:289-307, 310-352, 355-386, 397-404, 480-491) -------------
int sar_config_defaults(sar_config *c)
{
    if (!c) return fail(SAR_ERR_INVALID, "cfg is NULL");
    c->iterations = 10000000ull;          // lib.rs:291
    c->width = 1920; c->height = 1080;    // lib.rs:292-293
    c->render_kind = SAR_RENDER_GAS;      // lib.rs:295
    c->transparent = 1;                   // lib.rs:296
    c->angle = 0.0;                       // lib.rs:297
    c->silent = 1;                        // lib.rs:299
    c->palette_len = 6;
    const double r[6] = {1., 0.5, 1., 0.5, 0.5, 1.}, g[6] = {1., 1., 0.5, 1., 0.5, 0.5}, b[6] = {0.5, 0.5, 0.5, 1., 1., 1.};  // lib.rs:483-487
    memset(c->palette_rgb, 0, sizeof c->palette_rgb);
    for (int i = 0; i < 6; ++i) { c->palette_rgb[i][0] = r[i]; c->palette_rgb[i][1] = g[i]; c->palette_rgb[i][2] = b[i]; }
    c->bright_offset = -0.15;             // lib.rs:400
    c->bright_factor = 5. / 3.;           // lib.rs:401
    return SAR_OK;
}

int sar_config_poisson_saturne(sar_config *c)
{
    if (!c) return fail(SAR_ERR_INVALID, "cfg is NULL");
    memset(c, 0, sizeof *c);
    const double x[10] = {0.021, 1.182, -1.183, 0.128, -1.12, -0.641, -1.152, -0.834, -0.97, 0.722};
    const double y[10] = {0.243038, -0.825, -1.2, -0.835443, -0.835443, -0.364557, 0.458, 0.622785, -0.394937, -1.032911};
    const double z[10] = {-0.455696, 0.673, 0.915, -0.258228, -0.495, -0.264, -0.432, -0.416, -0.877, -0.3};
    memcpy(c->coef[0], x, sizeof x); memcpy(c->coef[1], y, sizeof y); memcpy(c->coef[2], z, sizeof z);
    c->center_camera[0] = -0.005; c->center_camera[1] = 0.262; c->center_camera[2] = -0.366 + 0.12;   // lib.rs:335-340
    c->axis[0] = 0.304289493528802; c->axis[1] = 0.760492682863655; c->axis[2] = 0.573636455813981;
    c->rotation = 1.78268191887446;
    c->scale = 1.;
    c->ct_kind = SAR_CT_POISSON_SATURNE;
    return sar_config_defaults(c);
}

int sar_config_solar_sail(sar_config *c)
{
    if (!c) return fail(SAR_ERR_INVALID, "cfg is NULL");
    memset(c, 0, sizeof *c);
    const double x[10] = {0.744304, -0.546835, 0.121519, -0.653165, 0.399, 0.379, 0.44, 1.014, -0.805063, 0.377};
    const double y[10] = {-0.683, 0.531646, -0.04557, -1.2, -0.546835, 0.091139, 0.744304, -0.273418, -0.349367, -0.531646};
    const double z[10] = {0.712, 0.744304, -0.577215, 0.966, 0.04557, 1.063291, 0.01519, -0.425316, 0.212658, -0.01519};
    memcpy(c->coef[0], x, sizeof x); memcpy(c->coef[1], y, sizeof y); memcpy(c->coef[2], z, sizeof z);
    c->center_camera[0] = 0.28; c->center_camera[1] = -0.12; c->center_camera[2] = 0.22;
    c->axis[0] = 0.02466; c->axis[1] = 0.4618; c->axis[2] = -0.54789;
    c->rotation = 2.2195;
    c->scale = 1.7;
    c->ct_kind = SAR_CT_ADJUSTED_VELOCITY;
    c->ct_factor = -0.2; c->ct_offset = 0.8;                   // lib.rs:381-384
    return sar_config_defaults(c);
}

// ---- start points ----------------------------------------------------------------------------
int sar_seed_points(uint64_t seed, uint64_t first, uint64_t n, double *out)
{
    if (!out && n) return fail(SAR_ERR_INVALID, "out_xyz is NULL");
    for (uint64_t k = 0; k < n; ++k)
        for (uint64_t c = 0; c < 3; ++c) {
            uint64_t z = seed + (3 * (first + k) + c + 1) * 0x9E3779B97F4A7C15ull;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z ^= z >> 31;
            out[3 * k + c] = ((double)(z >> 11) * 0x1.0p-53) * 0.1;
        }
    return SAR_OK;
}

// ---- Runtime -----------------------------------------------------------------------------------
int sar_runtime_new(uint32_t width, uint32_t height, int device, sar_runtime **out)
{
    if (!out) return fail(SAR_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (int rc = check_dims(width, height)) return rc;
    int ndev = 0;
    SAR_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(SAR_ERR_CUDA, "CUDA device %d not available (%d visible)", device, ndev);
    if (int rc = ensure_lnlut(device)) return rc;
    SAR_CUDA(cudaSetDevice(device));
    sar_runtime *rt = new (std::nothrow) sar_runtime();
    if (!rt) return fail(SAR_ERR_NOMEM, "host allocation failed");
    rt->device = device; rt->w = width; rt->h = height; rt->npix = (size_t)width * height;
    rt->nslots = slots_for(rt->npix); rt->slots = slotmap_for(rt->npix);
    const Layout lay = layout(rt->npix);
    cudaError_t e = cudaMalloc(&rt->block, lay.total);
    if (e != cudaSuccess) { (void)cudaGetLastError(); delete rt; return fail(SAR_ERR_NOMEM, "cudaMalloc(%zu bytes): %s", lay.total, cudaGetErrorString(e)); }
    rt->block_bytes = lay.total;
    rt->rec = (ulonglong2 *)rt->block;
    rt->fast = (unsigned long long *)((char *)rt->block + lay.fast);
    rt->image = (uint16_t *)((char *)rt->block + lay.image);
    rt->cnt = (uint32_t *)((char *)rt->block + lay.cnt);
    rt->scal = (Scalars *)((char *)rt->block + lay.scal);
    e = cudaStreamCreateWithFlags(&rt->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { cudaFree(rt->block); delete rt; return fail(SAR_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    cudaMemsetAsync(rt->scal, 0, 1024, rt->stream);         // flags/epochs start at 0; reset never touches them
    cudaDeviceGetAttribute(&rt->sm_count, cudaDevAttrMultiProcessorCount, device);
    *out = rt;
    int rc = sar_runtime_reset(rt);
    if (rc) { sar_runtime_free(rt); *out = nullptr; }
    return rc;
}

void sar_runtime_free(sar_runtime *rt)
{
    if (!rt) return;
    cudaSetDevice(rt->device);
    if (rt->stream) { cudaStreamSynchronize(rt->stream); cudaStreamDestroy(rt->stream); }
    if (rt->copy_stream) { cudaStreamSynchronize(rt->copy_stream); cudaStreamDestroy(rt->copy_stream); }
    for (auto &e : rt->stripe_done) if (e) cudaEventDestroy(e);
    cudaFree(rt->block); cudaFree(rt->d_init); cudaFree(rt->d_scratch);
    delete rt;
}

int sar_runtime_reset_async(sar_runtime *rt, void *stream)
{
    if (!rt) return fail(SAR_ERR_INVALID, "runtime is NULL");
    SAR_CUDA(cudaSetDevice(rt->device));
    launch_reset(rt->fast, rt->rec, rt->scal, rt->npix, rt->nslots, pick(rt, stream));
    SAR_CUDA(cudaGetLastError());
    rt->job_base = 0;
    rt->host_max_valid = false;
    rt->max_tracked = true;
    rt->depth_valid = false;
    return SAR_OK;
}
int sar_runtime_reset(sar_runtime *rt)
{
    if (int rc = sar_runtime_reset_async(rt, nullptr)) return rc;
    SAR_CUDA(cudaStreamSynchronize(rt->stream));
    return SAR_OK;
}

int sar_runtime_dims(const sar_runtime *rt, uint32_t *w, uint32_t *h, int *device)
{
    if (!rt) return fail(SAR_ERR_INVALID, "runtime is NULL");
    if (w) *w = rt->w;
    if (h) *h = rt->h;
    if (device) *device = rt->device;
    return SAR_OK;
}
int sar_runtime_get_job_base(const sar_runtime *rt, uint64_t *jb)
{
    if (!rt || !jb) return fail(SAR_ERR_INVALID, "NULL argument");
    *jb = rt->job_base;
    return SAR_OK;
}
int sar_runtime_set_job_base(sar_runtime *rt, uint64_t jb)
{
    if (!rt) return fail(SAR_ERR_INVALID, "runtime is NULL");
    rt->job_base = jb;
    return SAR_OK;
}

static int ensure_scratch(sar_runtime *rt, size_t bytes)
{
    if (rt->d_scratch_cap >= bytes) return SAR_OK;
    cudaFree(rt->d_scratch); rt->d_scratch = nullptr; rt->d_scratch_cap = 0;
    SAR_CUDA(cudaMalloc(&rt->d_scratch, bytes));
    rt->d_scratch_cap = bytes;
    return SAR_OK;
}

int sar_runtime_download(const sar_runtime *crt, uint32_t *count, double *steps, float *zbuf, uint32_t *max)
{
    sar_runtime *rt = const_cast<sar_runtime *>(crt);
    if (!rt) return fail(SAR_ERR_INVALID, "runtime is NULL");
    SAR_CUDA(cudaSetDevice(rt->device));
    const size_t n = rt->npix;
    const size_t oc = 0, os = align_up(n * 4, 256), oz = os + align_up(n * 8, 256), total = oz + align_up(n * 4, 256);
    if (int rc = ensure_scratch(rt, total)) return rc;
    char *base = (char *)rt->d_scratch;
    launch_unpack(rt->fast, rt->rec, rt->scal, n, rt->slots, (uint32_t *)(base + oc), (double *)(base + os), (float *)(base + oz), rt->stream);
    SAR_CUDA(cudaGetLastError());
    if (count) SAR_CUDA(cudaMemcpyAsync(count, base + oc, n * 4, cudaMemcpyDeviceToHost, rt->stream));
    if (steps) SAR_CUDA(cudaMemcpyAsync(steps, base + os, n * 8, cudaMemcpyDeviceToHost, rt->stream));
    if (zbuf) SAR_CUDA(cudaMemcpyAsync(zbuf, base + oz, n * 4, cudaMemcpyDeviceToHost, rt->stream));
    if (max) {
        if (int rc = sar_runtime_max_async(rt, 0, 0, nullptr)) return rc;
        if (int rc = sar_runtime_get_max(rt, max, nullptr)) return rc;
    }
    SAR_CUDA(cudaStreamSynchronize(rt->stream));
    return SAR_OK;
}

int sar_runtime_upload(sar_runtime *rt, const uint32_t *count, const double *steps, const float *zbuf)
{
    if (!rt || !count || !steps || !zbuf) return fail(SAR_ERR_INVALID, "NULL argument");
    SAR_CUDA(cudaSetDevice(rt->device));
    const size_t n = rt->npix;
    const size_t oc = 0, os = align_up(n * 4, 256), oz = os + align_up(n * 8, 256), total = oz + align_up(n * 4, 256);
    if (int rc = ensure_scratch(rt, total)) return rc;
    char *base = (char *)rt->d_scratch;
    SAR_CUDA(cudaMemcpyAsync(base + oc, count, n * 4, cudaMemcpyHostToDevice, rt->stream));
    SAR_CUDA(cudaMemcpyAsync(base + os, steps, n * 8, cudaMemcpyHostToDevice, rt->stream));
    SAR_CUDA(cudaMemcpyAsync(base + oz, zbuf, n * 4, cudaMemcpyHostToDevice, rt->stream));
    launch_pack(rt->fast, rt->rec, rt->scal, n, rt->slots, (const uint32_t *)(base + oc), (const double *)(base + os), (const float *)(base + oz), rt->stream);
    SAR_CUDA(cudaGetLastError());
    SAR_CUDA(cudaStreamSynchronize(rt->stream));
    rt->host_max_valid = false;
    rt->max_tracked = false;
    rt->depth_valid = false;
    if (rt->job_base == 0) rt->job_base = 1;   // uploaded records carry job key 0: they keep every future tie
    return SAR_OK;
}

int sar_runtime_merge(sar_runtime *dst, const sar_runtime *src)
{
    if (!dst || !src) return fail(SAR_ERR_INVALID, "NULL argument");
    if (dst->w != src->w || dst->h != src->h)                   // assert_eq!, lib.rs:709-710
        return fail(SAR_ERR_DIMS, "merge: %ux%u vs %ux%u", dst->w, dst->h, src->w, src->h);
    SAR_CUDA(cudaSetDevice(src->device));
    SAR_CUDA(cudaStreamSynchronize(src->stream));
    SAR_CUDA(cudaSetDevice(dst->device));
    const unsigned long long *sfast = src->fast;
    const ulonglong2 *srec = src->rec;
    const Scalars *sscal = src->scal;
    if (src->device != dst->device) {   // stage the source accumulators on dst's device
        if (int rc = ensure_scratch(dst, src->block_bytes)) return rc;
        SAR_CUDA(cudaMemcpyPeerAsync(dst->d_scratch, dst->device, src->block, src->device, src->block_bytes, dst->stream));
        const Layout lay = layout(src->npix);
        srec = (const ulonglong2 *)dst->d_scratch;
        sfast = (const unsigned long long *)((char *)dst->d_scratch + lay.fast);
        sscal = (const Scalars *)((char *)dst->d_scratch + lay.scal);
    }
    launch_merge(dst->fast, dst->rec, dst->scal, sfast, srec, sscal, dst->npix, dst->slots, dst->stream);
    SAR_CUDA(cudaGetLastError());
    dst->host_max_valid = false;
    dst->max_tracked = false;
    dst->depth_valid = false;
    SAR_CUDA(cudaStreamSynchronize(dst->stream));
    if (src->job_base > dst->job_base) dst->job_base = src->job_base;
    return SAR_OK;
}

// ---- render ------------------------------------------------------------------------------------
static int render_launch(const sar_config *cfg, sar_runtime *rt, const double *d_init, uint64_t seed,
                         uint64_t first_job, uint64_t n_jobs, uint32_t threads, cudaStream_t s,
                         bool init_is_warm = false)
{
    if (int rc = check_config(cfg, rt)) return rc;
    SAR_CUDA(cudaSetDevice(rt->device));
    IterParams p;
    make_iter_params(cfg, rt, p);
    p.init = d_init; p.seed = seed; p.first_job = first_job; p.n_jobs = n_jobs;
    if (init_is_warm) p.warmup = 0;
    // Order keys (who keeps an exact z tie): this call's job k gets key job_base + k — independent of
    // first_job, which only positions the call in the seed stream.  32-bit keys: refuse rather than
    // let late jobs share a key (ties between them would then resolve by scheduling order).
    if (rt->job_base + n_jobs > (1ull << 32))
        return fail(SAR_ERR_INVALID, "job order keys exhausted (%llu jobs since the last reset, 2^32 at most): reset or "
                    "download/upload the Runtime", (unsigned long long)rt->job_base);
    p.job_key0 = (unsigned int)rt->job_base;
    if (launch_iterate(p, threads ? threads : default_lanes(rt), s)) rt->max_tracked = false;   // tile path: max by reduction
    SAR_CUDA(cudaGetLastError());
    rt->host_max_valid = false;
    rt->depth_valid = false;
    rt->job_base += n_jobs;
    return SAR_OK;
}

int sar_render_seeded_async(const sar_config *cfg, sar_runtime *rt, uint64_t seed, uint64_t first_job,
                            uint64_t n_jobs, uint32_t threads, void *stream)
{
    if (!rt) return fail(SAR_ERR_INVALID, "runtime is NULL");
    return render_launch(cfg, rt, nullptr, seed, first_job, n_jobs, threads, pick(rt, stream));
}
int sar_render_device_async(const sar_config *cfg, sar_runtime *rt, const double *d_init_xyz, uint64_t first_job,
                            uint64_t n_jobs, uint32_t threads, void *stream)
{
    if (!rt) return fail(SAR_ERR_INVALID, "runtime is NULL");
    if (!d_init_xyz && n_jobs) return fail(SAR_ERR_INVALID, "d_init_xyz is NULL");
    return render_launch(cfg, rt, d_init_xyz, 0, first_job, n_jobs, threads, pick(rt, stream));
}

static int upload_init(sar_runtime *rt, const double *init_xyz, uint64_t n_jobs, cudaStream_t s)
{
    const size_t bytes = (size_t)n_jobs * 3 * sizeof(double);
    if (rt->d_init_cap < bytes) {
        cudaFree(rt->d_init); rt->d_init = nullptr; rt->d_init_cap = 0;
        SAR_CUDA(cudaMalloc((void **)&rt->d_init, bytes));
        rt->d_init_cap = bytes;
    }
    SAR_CUDA(cudaMemcpyAsync(rt->d_init, init_xyz, bytes, cudaMemcpyHostToDevice, s));
    return SAR_OK;
}

int sar_render(const sar_config *cfg, sar_runtime *rt, const double *init_xyz, uint64_t n_jobs)
{
    if (!rt) return fail(SAR_ERR_INVALID, "runtime is NULL");
    if (!init_xyz && n_jobs) return fail(SAR_ERR_INVALID, "init_xyz is NULL");
    if (int rc = check_config(cfg, rt)) return rc;
    if (n_jobs == 0) return SAR_OK;
    SAR_CUDA(cudaSetDevice(rt->device));
    if (int rc = upload_init(rt, init_xyz, n_jobs, rt->stream)) return rc;
    if (int rc = render_launch(cfg, rt, rt->d_init, 0, 0, n_jobs, 0, rt->stream)) return rc;
    SAR_CUDA(cudaStreamSynchronize(rt->stream));
    return SAR_OK;
}

int sar_render_seeded(const sar_config *cfg, sar_runtime *rt, uint64_t seed, uint64_t first_job, uint64_t n_jobs)
{
    if (int rc = sar_render_seeded_async(cfg, rt, seed, first_job, n_jobs, 0, nullptr)) return rc;
    SAR_CUDA(cudaStreamSynchronize(rt->stream));
    return SAR_OK;
}

// ---- auto-framing first pass (lib.rs:326-334) ----------------------------------------------------
static double dkey_to_double(unsigned long long k)
{
    const unsigned long long b = (k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k;
    double v;
    memcpy(&v, &b, sizeof v);
    return v;
}

int sar_autoframe(const sar_config *cfg, int device, uint64_t seed, const double *init_xyz, uint64_t n_jobs, uint64_t iterations,
                  sar_autoframe_result *out)
{
    if (!cfg || !out) return fail(SAR_ERR_INVALID, "NULL argument");
    if (n_jobs == 0 || iterations == 0) return fail(SAR_ERR_INVALID, "n_jobs and iterations must be non-zero");
    if (int rc = check_config(cfg, nullptr)) return rc;
    int ndev = 0;
    SAR_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(SAR_ERR_CUDA, "CUDA device %d not available (%d visible)", device, ndev);
    SAR_CUDA(cudaSetDevice(device));
    sar_runtime dummy;                                   // make_iter_params only reads the accumulator pointers (unused here)
    sar_config c = *cfg;
    c.iterations = iterations;
    if (c.width == 0) c.width = 1;
    if (c.height == 0) c.height = 1;
    IterParams p;
    make_iter_params(&c, &dummy, p);
    BBoxAccum h;
    for (int k = 0; k < 3; ++k) { h.lo[k] = ~0ull; h.hi[k] = 0ull; }
    h.diverged = 0;
    BBoxAccum *d_acc = nullptr;
    double *d_init = nullptr;
    cudaStream_t s = nullptr;
    int rc = SAR_OK;
    do {
        if (cudaMalloc((void **)&d_acc, sizeof h) != cudaSuccess) { rc = fail(SAR_ERR_NOMEM, "cudaMalloc"); break; }
        if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { rc = fail(SAR_ERR_CUDA, "cudaStreamCreate"); break; }
        if (cudaMemcpyAsync(d_acc, &h, sizeof h, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = fail(SAR_ERR_CUDA, "cudaMemcpy"); break; }
        if (init_xyz) {
            const size_t bytes = (size_t)n_jobs * 3 * sizeof(double);
            if (cudaMalloc((void **)&d_init, bytes) != cudaSuccess) { rc = fail(SAR_ERR_NOMEM, "cudaMalloc(%zu)", bytes); break; }
            if (cudaMemcpyAsync(d_init, init_xyz, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) { rc = fail(SAR_ERR_CUDA, "cudaMemcpy"); break; }
        }
        p.init = d_init; p.seed = seed; p.first_job = 0; p.n_jobs = n_jobs;
        launch_bbox(p, d_acc, s);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(&h, d_acc, sizeof h, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) { rc = fail(SAR_ERR_CUDA, "autoframe: %s", cudaGetErrorString(e)); break; }
    } while (0);
    cudaFree(d_acc); cudaFree(d_init);
    if (s) cudaStreamDestroy(s);
    (void)cudaGetLastError();
    if (rc) return rc;
    memset(out, 0, sizeof *out);
    out->n_jobs = n_jobs; out->diverged = h.diverged;
    if (h.diverged == n_jobs) return SAR_OK;             // nothing bounded: box stays 0, caller sees diverged == n_jobs
    for (int k = 0; k < 3; ++k) { out->box[2 * k] = dkey_to_double(h.lo[k]); out->box[2 * k + 1] = dkey_to_double(h.hi[k]); }
    // center_camera is ADDED to screen_space before projecting, x with x, y with screen z, z with screen y
    // (lib.rs:776-786): centring means minus the mid-points, paired that way (cf. the values at lib.rs:335-340)
    out->center_camera[0] = -(out->box[0] + out->box[1]) / 2.;
    out->center_camera[1] = -(out->box[4] + out->box[5]) / 2.;
    out->center_camera[2] = -(out->box[2] + out->box[3]) / 2.;
    // largest scale that keeps the box in view from EVERY view angle (the view rotates about the vertical axis through
    // the centre, lib.rs:776-779): horizontal |x2| < 0.5/scale (lib.rs:783), vertical |y| < height/(2 width scale) (lib.rs:786)
    const double rx = std::hypot((out->box[1] - out->box[0]) / 2., (out->box[5] - out->box[4]) / 2.);
    const double ry = (out->box[3] - out->box[2]) / 2.;
    const double aspect = (double)c.height / (double)c.width;
    double sc = rx > 0. ? 0.5 / rx : INFINITY;
    if (ry > 0. && 0.5 * aspect / ry < sc) sc = 0.5 * aspect / ry;
    out->scale = std::isfinite(sc) ? 0.95 * sc : 1.0;
    return SAR_OK;
}

// ---- max / colorize ----------------------------------------------------------------------------
// Full reduction: Runtime.max and the Depth fold over a row range, from the accumulators.
static int runtime_max_full(sar_runtime *rt, uint32_t row0, uint32_t rows, cudaStream_t s)
{
    const unsigned int init[3] = {0u, ZKEY_ZERO, ZKEY_FLT_MAX};   // max, zmax_key, zmin_key (fold seed lib.rs:882)
    SAR_CUDA(cudaMemcpyAsync(&rt->scal->max, init, sizeof init, cudaMemcpyHostToDevice, s));
    launch_max(rt->fast, rt->rec, rt->scal, (size_t)row0 * rt->w, (size_t)rows * rt->w, rt->slots, s);
    SAR_CUDA(cudaGetLastError());
    const bool whole = row0 == 0 && rows == rt->h;
    rt->max_tracked = whole;          // a whole-image reduction is a valid starting point for the render's own tracking
    rt->depth_valid = whole;
    return SAR_OK;
}

int sar_runtime_max_async(sar_runtime *rt, uint32_t row0, uint32_t rows, void *stream)
{
    if (!rt) return fail(SAR_ERR_INVALID, "runtime is NULL");
    if (rows == 0) { row0 = 0; rows = rt->h; }
    if ((uint64_t)row0 + rows > rt->h) return fail(SAR_ERR_INVALID, "rows [%u,%u) outside image height %u", row0, row0 + rows, rt->h);
    SAR_CUDA(cudaSetDevice(rt->device));
    cudaStream_t s = pick(rt, stream);
    rt->host_max_valid = false;
    if (rt->max_tracked && row0 == 0 && rows == rt->h) {
        // the iterate kernel kept scal->max current (lib.rs:813-815 is a running max): only the NaN debt is left to fold.
        // The Depth fold is not touched here; the colourise calls compute it when a Depth image is asked for.
        launch_fold_max(rt->fast, rt->scal, rt->slots, s);
        SAR_CUDA(cudaGetLastError());
        return SAR_OK;
    }
    return runtime_max_full(rt, row0, rows, s);
}
// Depth images need the min/max of the touched z values (lib.rs:877-882): reduce them if they are stale.
static int ensure_depth_fold(const sar_config *cfg, sar_runtime *rt, cudaStream_t s)
{
    if (cfg->render_kind != SAR_RENDER_DEPTH || rt->depth_valid) return SAR_OK;
    return runtime_max_full(rt, 0, rt->h, s);
}
int sar_runtime_get_max(sar_runtime *rt, uint32_t *max_out, void *stream)
{
    if (!rt || !max_out) return fail(SAR_ERR_INVALID, "NULL argument");
    SAR_CUDA(cudaSetDevice(rt->device));
    cudaStream_t s = pick(rt, stream);
    SAR_CUDA(cudaMemcpyAsync(max_out, &rt->scal->max, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    SAR_CUDA(cudaStreamSynchronize(s));
    rt->host_max = *max_out; rt->host_max_valid = true;
    return SAR_OK;
}
int sar_runtime_set_max(sar_runtime *rt, uint32_t max, void *stream)
{
    if (!rt) return fail(SAR_ERR_INVALID, "runtime is NULL");
    SAR_CUDA(cudaSetDevice(rt->device));
    cudaStream_t s = pick(rt, stream);
    SAR_CUDA(cudaMemcpyAsync(&rt->scal->max, &max, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    SAR_CUDA(cudaStreamSynchronize(s));
    rt->host_max = max; rt->host_max_valid = true;
    return SAR_OK;
}

int sar_colorize_rows_async(const sar_config *cfg, sar_runtime *rt, uint32_t row0, uint32_t rows, sar_peer *dst, void *stream)
{
    if (!rt) return fail(SAR_ERR_INVALID, "runtime is NULL");
    if (int rc = check_config(cfg, rt)) return rc;
    if (rows == 0) { row0 = 0; rows = rt->h; }
    if ((uint64_t)row0 + rows > rt->h) return fail(SAR_ERR_INVALID, "rows [%u,%u) outside image height %u", row0, row0 + rows, rt->h);
    if (dst && (dst->w != rt->w || dst->h != rt->h)) return fail(SAR_ERR_DIMS, "peer image is %ux%u, runtime %ux%u", dst->w, dst->h, rt->w, rt->h);
    SAR_CUDA(cudaSetDevice(rt->device));
    if (int rc = ensure_depth_fold(cfg, rt, pick(rt, stream))) return rc;
    ColorParams cp;
    make_color_params(cfg, rt, cp, row0, rows, rt->host_max_valid ? &rt->host_max : nullptr);
    launch_colorize(cp, rt->fast, rt->rec, rt->scal, dst ? dst->image : rt->image, nullptr, pick(rt, stream));
    SAR_CUDA(cudaGetLastError());
    return SAR_OK;
}

int sar_runtime_image_download(sar_runtime *rt, uint32_t row0, uint32_t rows, uint16_t *rgba, void *stream)
{
    if (!rt || !rgba) return fail(SAR_ERR_INVALID, "NULL argument");
    if (rows == 0) { row0 = 0; rows = rt->h; }
    if ((uint64_t)row0 + rows > rt->h) return fail(SAR_ERR_INVALID, "rows outside image");
    SAR_CUDA(cudaSetDevice(rt->device));
    cudaStream_t s = pick(rt, stream);
    const size_t off = (size_t)row0 * rt->w * 4, cnt = (size_t)rows * rt->w * 4;
    SAR_CUDA(cudaMemcpyAsync(rgba + off, rt->image + off, cnt * sizeof(uint16_t), cudaMemcpyDeviceToHost, s));
    SAR_CUDA(cudaStreamSynchronize(s));
    return SAR_OK;
}

int sar_colorize(const sar_config *cfg, const sar_runtime *crt, uint16_t *rgba_u16, float *rgba_f32)
{
    sar_runtime *rt = const_cast<sar_runtime *>(crt);
    if (!rt) return fail(SAR_ERR_INVALID, "runtime is NULL");
    if (!rgba_u16 && !rgba_f32) return fail(SAR_ERR_INVALID, "no output buffer");
    if (int rc = check_config(cfg, rt)) return rc;
    SAR_CUDA(cudaSetDevice(rt->device));
    if (int rc = sar_runtime_max_async(rt, 0, 0, nullptr)) return rc;
    if (int rc = ensure_depth_fold(cfg, rt, rt->stream)) return rc;
    float *d_f32 = nullptr;
    if (rgba_f32) {
        if (int rc = ensure_scratch(rt, rt->npix * 4 * sizeof(float))) return rc;
        d_f32 = (float *)rt->d_scratch;
    }
    uint32_t host_max = 0;
    if (int rc = sar_runtime_get_max(rt, &host_max, nullptr)) return rc;
    ColorParams cp;
    make_color_params(cfg, rt, cp, 0, rt->h, &host_max);
    launch_colorize(cp, rt->fast, rt->rec, rt->scal, rt->image, d_f32, rt->stream);
    SAR_CUDA(cudaGetLastError());
    if (rgba_u16) SAR_CUDA(cudaMemcpyAsync(rgba_u16, rt->image, rt->npix * 4 * sizeof(uint16_t), cudaMemcpyDeviceToHost, rt->stream));
    if (rgba_f32) SAR_CUDA(cudaMemcpyAsync(rgba_f32, d_f32, rt->npix * 4 * sizeof(float), cudaMemcpyDeviceToHost, rt->stream));
    SAR_CUDA(cudaStreamSynchronize(rt->stream));
    return SAR_OK;
}

int sar_stream_synchronize(sar_runtime *rt, void *stream)
{
    if (!rt) return fail(SAR_ERR_INVALID, "runtime is NULL");
    SAR_CUDA(cudaSetDevice(rt->device));
    SAR_CUDA(cudaStreamSynchronize(pick(rt, stream)));
    return SAR_OK;
}

// ---- pinned host memory ------------------------------------------------------------------------
int sar_host_alloc(size_t bytes, void **out)
{
    if (!out) return fail(SAR_ERR_INVALID, "out is NULL");
    *out = nullptr;
    SAR_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return SAR_OK;
}
void sar_host_free(void *p) { if (p) cudaFreeHost(p); }

// ---- cross-process peers ------------------------------------------------------------------------
int sar_runtime_ipc_export(const sar_runtime *rt, uint8_t out[SAR_IPC_HANDLE_BYTES])
{
    if (!rt || !out) return fail(SAR_ERR_INVALID, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == SAR_IPC_HANDLE_BYTES, "IPC handle size");
    SAR_CUDA(cudaSetDevice(rt->device));
    cudaIpcMemHandle_t h;
    SAR_CUDA(cudaIpcGetMemHandle(&h, rt->block));
    memcpy(out, &h, sizeof h);
    return SAR_OK;
}

static void peer_view(sar_peer *p, void *block, uint32_t w, uint32_t h)
{
    const Layout lay = layout((size_t)w * h);
    p->w = w; p->h = h; p->block = block;
    p->slots = slotmap_for((size_t)w * h);
    p->rec = (ulonglong2 *)block;
    p->fast = (unsigned long long *)((char *)block + lay.fast);
    p->image = (uint16_t *)((char *)block + lay.image);
    p->cnt = (uint32_t *)((char *)block + lay.cnt);
    p->scal = (Scalars *)((char *)block + lay.scal);
}

int sar_peer_open(const uint8_t handle[SAR_IPC_HANDLE_BYTES], uint32_t width, uint32_t height, int local_device, sar_peer **out)
{
    if (!handle || !out) return fail(SAR_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (int rc = check_dims(width, height)) return rc;
    SAR_CUDA(cudaSetDevice(local_device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    void *block = nullptr;
    SAR_CUDA(cudaIpcOpenMemHandle(&block, h, cudaIpcMemLazyEnablePeerAccess));
    sar_peer *p = new (std::nothrow) sar_peer();
    if (!p) { cudaIpcCloseMemHandle(block); return fail(SAR_ERR_NOMEM, "host allocation failed"); }
    p->local_device = local_device; p->ipc = true;
    peer_view(p, block, width, height);
    *out = p;
    return SAR_OK;
}

void sar_peer_close(sar_peer *p)
{
    if (!p) return;
    if (p->ipc && p->block) { cudaSetDevice(p->local_device); cudaIpcCloseMemHandle(p->block); }
    delete p;
}

int sar_runtime_merge_peers_async(sar_runtime *rt, sar_peer *const *peers, int n_peers, uint32_t row0, uint32_t rows, void *stream)
{
    if (!rt || (n_peers > 0 && !peers)) return fail(SAR_ERR_INVALID, "NULL argument");
    if (n_peers < 0 || n_peers > 16) return fail(SAR_ERR_INVALID, "n_peers must be in 0..16 (got %d)", n_peers);
    if (rows == 0) { row0 = 0; rows = rt->h; }
    if ((uint64_t)row0 + rows > rt->h) return fail(SAR_ERR_INVALID, "rows outside image");
    PeerList pl;
    memset(&pl, 0, sizeof pl);
    pl.n = n_peers;
    for (int i = 0; i < n_peers; ++i) {
        if (!peers[i]) return fail(SAR_ERR_INVALID, "peer %d is NULL", i);
        if (peers[i]->w != rt->w || peers[i]->h != rt->h)
            return fail(SAR_ERR_DIMS, "peer %d is %ux%u, runtime %ux%u", i, peers[i]->w, peers[i]->h, rt->w, rt->h);
        pl.fast[i] = peers[i]->fast; pl.rec[i] = peers[i]->rec; pl.scal[i] = peers[i]->scal;
    }
    SAR_CUDA(cudaSetDevice(rt->device));
    launch_merge_peers(rt->fast, rt->rec, rt->scal, pl, (size_t)row0 * rt->w, (size_t)rows * rt->w, rt->slots, pick(rt, stream));
    SAR_CUDA(cudaGetLastError());
    rt->host_max_valid = false;
    rt->max_tracked = false;
    rt->depth_valid = false;
    return SAR_OK;
}

// ---- cross-GPU frame protocol: one process per GPU, device-side synchronisation -------
static int make_frame_sync(sar_runtime *rt, sar_peer *const *peers_by_rank, int n_ranks, int my_rank, uint32_t epoch,
                           FrameSync &S, PeerList *pl)
{
    if (!rt) return fail(SAR_ERR_INVALID, "runtime is NULL");
    if (n_ranks < 1 || n_ranks > SYNC_MAX_RANKS || my_rank < 0 || my_rank >= n_ranks)
        return fail(SAR_ERR_INVALID, "bad rank %d of %d (at most %d ranks)", my_rank, n_ranks, SYNC_MAX_RANKS);
    if (n_ranks > 1 && !peers_by_rank) return fail(SAR_ERR_INVALID, "peers_by_rank is NULL");
    memset(&S, 0, sizeof S);
    if (pl) memset(pl, 0, sizeof *pl);
    S.n_ranks = n_ranks; S.my_rank = my_rank; S.epoch = epoch;
    for (int r = 0; r < n_ranks; ++r) {
        if (r == my_rank) { S.scal[r] = rt->scal; continue; }
        const sar_peer *p = peers_by_rank[r];
        if (!p) return fail(SAR_ERR_INVALID, "peer %d is NULL", r);
        if (p->w != rt->w || p->h != rt->h) return fail(SAR_ERR_DIMS, "peer %d is %ux%u, runtime %ux%u", r, p->w, p->h, rt->w, rt->h);
        S.scal[r] = p->scal;
        if (pl) { pl->fast[pl->n] = p->fast; pl->rec[pl->n] = p->rec; pl->scal[pl->n] = p->scal; pl->cnt[pl->n] = p->cnt; ++pl->n; }
    }
    return SAR_OK;
}
static int check_rows(const sar_runtime *rt, uint32_t &row0, uint32_t &rows)
{
    if (rows == 0) { row0 = 0; rows = rt->h; }
    if ((uint64_t)row0 + rows > rt->h) return fail(SAR_ERR_INVALID, "rows [%u,%u) outside image height %u", row0, row0 + rows, rt->h);
    return SAR_OK;
}

int sar_frame_reset_async(sar_runtime *rt, int n_ranks, uint32_t epoch, void *stream)
{
    if (!rt) return fail(SAR_ERR_INVALID, "runtime is NULL");
    if (n_ranks < 1 || n_ranks > SYNC_MAX_RANKS) return fail(SAR_ERR_INVALID, "bad n_ranks %d", n_ranks);
    SAR_CUDA(cudaSetDevice(rt->device));
    launch_frame_reset(rt->fast, rt->rec, rt->scal, rt->npix, rt->nslots, n_ranks, epoch, pick(rt, stream));
    SAR_CUDA(cudaGetLastError());
    rt->job_base = 0;
    rt->host_max_valid = false;
    rt->max_tracked = false;          // the frame protocol reduces its stripe maxima itself
    rt->depth_valid = false;
    return SAR_OK;
}

int sar_frame_export_async(sar_runtime *rt, sar_peer *const *peers_by_rank, int n_ranks, int my_rank, uint32_t epoch, void *stream)
{
    FrameSync S;
    if (int rc = make_frame_sync(rt, peers_by_rank, n_ranks, my_rank, epoch, S, nullptr)) return rc;
    SAR_CUDA(cudaSetDevice(rt->device));
    launch_frame_export(rt->fast, rt->scal, rt->cnt, rt->npix, rt->slots, S, pick(rt, stream));
    SAR_CUDA(cudaGetLastError());
    return SAR_OK;
}

int sar_frame_merge_async(sar_runtime *rt, sar_peer *const *peers_by_rank, int n_ranks, int my_rank, uint32_t row0, uint32_t rows,
                          uint32_t epoch, void *stream)
{
    FrameSync S;
    PeerList pl;
    if (int rc = make_frame_sync(rt, peers_by_rank, n_ranks, my_rank, epoch, S, &pl)) return rc;
    if (int rc = check_rows(rt, row0, rows)) return rc;
    SAR_CUDA(cudaSetDevice(rt->device));
    launch_frame_merge(rt->fast, rt->rec, rt->cnt, rt->scal, pl, (size_t)row0 * rt->w, (size_t)rows * rt->w, rt->slots, S, pick(rt, stream));
    SAR_CUDA(cudaGetLastError());
    rt->host_max_valid = false;
    rt->max_tracked = false;
    rt->depth_valid = false;
    return SAR_OK;
}

int sar_frame_colorize_async(const sar_config *cfg, sar_runtime *rt, sar_peer *const *peers_by_rank, int n_ranks, int my_rank,
                             int owner_rank, uint32_t row0, uint32_t rows, uint32_t epoch, void *stream)
{
    FrameSync S;
    if (int rc = make_frame_sync(rt, peers_by_rank, n_ranks, my_rank, epoch, S, nullptr)) return rc;
    if (int rc = check_config(cfg, rt)) return rc;
    if (int rc = check_rows(rt, row0, rows)) return rc;
    if (owner_rank < 0 || owner_rank >= n_ranks) return fail(SAR_ERR_INVALID, "bad owner rank %d", owner_rank);
    SAR_CUDA(cudaSetDevice(rt->device));
    ColorParams cp;
    make_color_params(cfg, rt, cp, row0, rows);
    uint16_t *image = owner_rank == my_rank ? rt->image : peers_by_rank[owner_rank]->image;
    launch_frame_colorize(cp, rt->cnt, rt->rec, rt->scal, image, S, owner_rank, pick(rt, stream));
    SAR_CUDA(cudaGetLastError());
    rt->host_max_valid = false;
    return SAR_OK;
}

int sar_frame_image_wait_async(sar_runtime *rt, int n_ranks, uint32_t epoch, void *stream)
{
    if (!rt) return fail(SAR_ERR_INVALID, "runtime is NULL");
    if (n_ranks < 1 || n_ranks > SYNC_MAX_RANKS) return fail(SAR_ERR_INVALID, "bad n_ranks %d", n_ranks);
    SAR_CUDA(cudaSetDevice(rt->device));
    launch_wait(rt->scal, SYNC_IMAGE_DONE, n_ranks, epoch, pick(rt, stream));
    SAR_CUDA(cudaGetLastError());
    return SAR_OK;
}

int sar_frame_image_release_async(sar_runtime *rt, sar_peer *const *peers_by_rank, int n_ranks, int my_rank, uint32_t epoch, void *stream)
{
    FrameSync S;
    if (int rc = make_frame_sync(rt, peers_by_rank, n_ranks, my_rank, epoch, S, nullptr)) return rc;
    SAR_CUDA(cudaSetDevice(rt->device));
    launch_signal(S, SYNC_IMAGE_FREE, pick(rt, stream));
    SAR_CUDA(cudaGetLastError());
    return SAR_OK;
}

int sar_runtime_sync_error(sar_runtime *rt, uint32_t *error, int clear)
{
    if (!rt || !error) return fail(SAR_ERR_INVALID, "NULL argument");
    SAR_CUDA(cudaSetDevice(rt->device));
    SAR_CUDA(cudaMemcpy(error, &rt->scal->sync_error, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (clear && *error) SAR_CUDA(cudaMemset(&rt->scal->sync_error, 0, sizeof(uint32_t)));
    return SAR_OK;
}

// ---- ParallelRenderer / render_parallel ---------------------------------------------------------
int sar_renderer_new(const int *devices, int n_devices, uint32_t threads_per_device, sar_renderer **out)
{
    if (!out) return fail(SAR_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int ndev = 0;
    SAR_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return fail(SAR_ERR_CUDA, "no CUDA device");
    sar_renderer *r = new (std::nothrow) sar_renderer();
    if (!r) return fail(SAR_ERR_NOMEM, "host allocation failed");
    if (!devices || n_devices <= 0) r->devices.push_back(0);
    else
        for (int i = 0; i < n_devices; ++i) {
            if (devices[i] < 0 || devices[i] >= ndev) { delete r; return fail(SAR_ERR_CUDA, "CUDA device %d not available (%d visible)", devices[i], ndev); }
            r->devices.push_back(devices[i]);
        }
    if (r->devices.size() > 16) { delete r; return fail(SAR_ERR_INVALID, "at most 16 devices"); }
    r->threads_per_device = threads_per_device;
    r->rts.assign(r->devices.size(), nullptr);     // Runtimes are created on first use (Runtime::empty, lib.rs:938)
    r->seq.resize(r->devices.size());
    *out = r;
    return SAR_OK;
}

static void seq_release(sar_renderer *r, size_t d)
{
    seq_device &q = r->seq[d];
    cudaSetDevice(r->devices[d]);
    if (q.copy_stream) { cudaStreamSynchronize(q.copy_stream); cudaStreamDestroy(q.copy_stream); }
    if (q.pay_stream) { cudaStreamSynchronize(q.pay_stream); cudaStreamDestroy(q.pay_stream); }
    for (int k = 0; k < 2; ++k) {
        sar_runtime_free(q.rt[k]);
        if (q.stage[k]) cudaFreeHost(q.stage[k]);
        if (q.h_sums[k]) cudaFreeHost(q.h_sums[k]);
        cudaFree(q.enc[k]);
        if (q.max_ready[k]) cudaEventDestroy(q.max_ready[k]);
        if (q.rendered[k]) cudaEventDestroy(q.rendered[k]);
        if (q.copied[k]) cudaEventDestroy(q.copied[k]);
    }
    if (q.h_max) cudaFreeHost(q.h_max);
    cudaFree(q.warm);
    q = seq_device();
}

void sar_renderer_shutdown(sar_renderer *r)
{
    if (!r) return;
    for (size_t d = 0; d < r->seq.size(); ++d) seq_release(r, d);
    for (sar_runtime *rt : r->rts) sar_runtime_free(rt);
    delete r;
}

static uint32_t renderer_lanes(const sar_renderer *r, size_t d)
{
    if (r->threads_per_device) return r->threads_per_device;
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, r->devices[d]);
    return (uint32_t)sm * LANES_PER_SM;
}

// "Threads" of device d for a given jobs_per_thread.  With an explicit threads_per_device this
// is that number (each lane then runs jobs_per_thread jobs one after the other, like a reference
// worker, lib.rs:956-988).  In auto mode the GPU needs no over-decomposition for load balance —
// every job has the same length and gets its own lane — so the job count is kept at the
// device's lane count: threads = lanes / jobs_per_thread (multiple of 32, at least 32), which
// keeps the 1000-step warm-up per job (lib.rs:750) from swamping short frames.  A GPU has ~10^5
// lanes where the reference has ~10 threads, so auto mode also never uses more threads than give
// every job at least MIN_JOB_ITERATIONS recorded steps: a small render (the reference's default is
// 1e7 iterations, and `iterations / threads / jobs` rounds to 0 below threads*jobs) must not come
// out empty.  `iterations` = the whole frame's (UINT64_MAX: no cap).
static const uint64_t MIN_JOB_ITERATIONS = 64;
static uint32_t renderer_threads_for(const sar_renderer *r, size_t d, uint64_t jobs_per_thread, uint64_t iterations = ~0ull)
{
    const uint32_t lanes = renderer_lanes(r, d);
    if (r->threads_per_device) return lanes;
    uint64_t t = lanes / jobs_per_thread;
    const uint64_t share = iterations / r->devices.size() / jobs_per_thread / MIN_JOB_ITERATIONS;   // per device
    if (share < t) t = share;
    t &= ~31ull;
    return (uint32_t)(t < 32 ? 32 : t);
}

int sar_renderer_plan(const sar_renderer *r, uint64_t iterations, uint64_t jobs_per_thread, uint64_t *num_threads,
                      uint64_t *iterations_per_job)
{
    if (!r) return fail(SAR_ERR_INVALID, "renderer is NULL");
    if (jobs_per_thread == 0) return fail(SAR_ERR_INVALID, "jobs_per_thread must be non-zero");
    uint64_t n = 0;
    for (size_t d = 0; d < r->devices.size(); ++d) n += renderer_threads_for(r, d, jobs_per_thread, iterations);
    if (num_threads) *num_threads = n;
    if (iterations_per_job) *iterations_per_job = iterations / n / jobs_per_thread;     // lib.rs:1058
    return SAR_OK;
}

int sar_renderer_num_threads_for(const sar_renderer *r, uint64_t jobs_per_thread, uint64_t *num_threads)
{
    if (!num_threads) return fail(SAR_ERR_INVALID, "NULL argument");
    return sar_renderer_plan(r, ~0ull, jobs_per_thread, num_threads, nullptr);
}

int sar_renderer_num_threads(const sar_renderer *r, uint64_t *num_threads)
{
    return sar_renderer_num_threads_for(r, 1, num_threads);
}

int sar_renderer_runtime(sar_renderer *r, sar_runtime **rt)
{
    if (!r || !rt) return fail(SAR_ERR_INVALID, "NULL argument");
    if (!r->rts[0]) return fail(SAR_ERR_INVALID, "no render_parallel call has been made yet");
    *rt = r->rts[0];
    return SAR_OK;
}

int sar_render_parallel(sar_renderer *r, const sar_config *cfg_in, uint64_t jobs_per_thread, uint64_t seed,
                        const double *init_xyz, uint16_t *rgba_u16)
{
    if (!r || !cfg_in || !rgba_u16) return fail(SAR_ERR_INVALID, "NULL argument");
    if (jobs_per_thread == 0) return fail(SAR_ERR_INVALID, "jobs_per_thread must be non-zero");   // NonZeroUsize in the CLI, main.rs:503
    if (int rc = check_config(cfg_in, nullptr)) return rc;
    if (int rc = check_dims(cfg_in->width, cfg_in->height)) return rc;
    const size_t nd = r->devices.size();
    uint64_t num_threads = 0;
    sar_renderer_plan(r, cfg_in->iterations, jobs_per_thread, &num_threads, nullptr);
    sar_config cfg = *cfg_in;
    cfg.iterations = cfg_in->iterations / num_threads / jobs_per_thread;        // lib.rs:1058
    const uint64_t total_jobs = jobs_per_thread * num_threads;                  // lib.rs:1062

    // set_width_height + reset on every worker (lib.rs:950-951)
    for (size_t d = 0; d < nd; ++d) {
        if (r->rts[d] && (r->rts[d]->w != cfg.width || r->rts[d]->h != cfg.height)) { sar_runtime_free(r->rts[d]); r->rts[d] = nullptr; }
        if (!r->rts[d]) { if (int rc = sar_runtime_new(cfg.width, cfg.height, r->devices[d], &r->rts[d])) return rc; }
        else if (int rc = sar_runtime_reset_async(r->rts[d], nullptr)) return rc;
    }
    // device d renders the contiguous job slice [first, first+n): lanes_d * jobs_per_thread jobs
    uint64_t first = 0;
    for (size_t d = 0; d < nd; ++d) {
        sar_runtime *rt = r->rts[d];
        const uint32_t threads = renderer_threads_for(r, d, jobs_per_thread, cfg_in->iterations);
        const uint64_t n = (uint64_t)threads * jobs_per_thread;
        // explicit thread count: that many lanes, jobs_per_thread jobs each; auto: one lane per job
        const uint32_t lanes = r->threads_per_device ? threads : (uint32_t)(n < renderer_lanes(r, d) ? n : renderer_lanes(r, d));
        rt->job_base = first;                      // global order keys: device d's jobs follow device d-1's
        if (init_xyz) {
            SAR_CUDA(cudaSetDevice(rt->device));
            if (int rc = upload_init(rt, init_xyz + 3 * first, n, rt->stream)) return rc;
            if (int rc = render_launch(&cfg, rt, rt->d_init, 0, first, n, lanes, rt->stream)) return rc;
        } else if (int rc = render_launch(&cfg, rt, nullptr, seed, first, n, lanes, rt->stream)) return rc;
        first += n;
    }
    // merge every other device into the first (lib.rs:1070-1076), deterministically: device 0 reads the other
    // devices' accumulators in place over NVLink when peer access can be enabled (only the touched records move),
    // else from a staged copy of the whole block
    sar_runtime *rt0 = r->rts[0];
    for (size_t d = 1; d < nd; ++d) {
        sar_runtime *src = r->rts[d];
        SAR_CUDA(cudaSetDevice(src->device));
        SAR_CUDA(cudaStreamSynchronize(src->stream));
        SAR_CUDA(cudaSetDevice(rt0->device));
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, rt0->device, src->device) != cudaSuccess) can = 0;
        if (can) {
            const cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
            (void)cudaGetLastError();
        }
        sar_peer view;
        if (can) {
            peer_view(&view, src->block, cfg.width, cfg.height);
        } else {
            if (int rc = ensure_scratch(rt0, src->block_bytes)) return rc;
            SAR_CUDA(cudaMemcpyPeerAsync(rt0->d_scratch, rt0->device, src->block, src->device, src->block_bytes, rt0->stream));
            peer_view(&view, rt0->d_scratch, cfg.width, cfg.height);
        }
        sar_peer *pv = &view;
        if (int rc = sar_runtime_merge_peers_async(rt0, &pv, 1, 0, 0, nullptr)) return rc;
    }
    rt0->job_base = total_jobs;
    if (int rc = sar_runtime_max_async(rt0, 0, 0, nullptr)) return rc;
    if (int rc = ensure_depth_fold(&cfg, rt0, rt0->stream)) return rc;
    uint32_t host_max = 0;
    if (int rc = sar_runtime_get_max(rt0, &host_max, nullptr)) return rc;
    // colorize (lib.rs:1080) stripe by stripe, each stripe's copy to the host starting as soon as it is coloured
    if (!rt0->copy_stream) {
        SAR_CUDA(cudaStreamCreateWithFlags(&rt0->copy_stream, cudaStreamNonBlocking));
        for (auto &e : rt0->stripe_done) SAR_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    const uint32_t n_stripes = rt0->h >= 64 ? (uint32_t)(sizeof rt0->stripe_done / sizeof rt0->stripe_done[0]) : 1u;
    for (uint32_t k = 0; k < n_stripes; ++k) {
        const uint32_t row0 = (uint32_t)((uint64_t)rt0->h * k / n_stripes), row1 = (uint32_t)((uint64_t)rt0->h * (k + 1) / n_stripes);
        ColorParams cp;
        make_color_params(&cfg, rt0, cp, row0, row1 - row0, &host_max);
        launch_colorize(cp, rt0->fast, rt0->rec, rt0->scal, rt0->image, nullptr, rt0->stream);
        SAR_CUDA(cudaGetLastError());
        SAR_CUDA(cudaEventRecord(rt0->stripe_done[k], rt0->stream));
        SAR_CUDA(cudaStreamWaitEvent(rt0->copy_stream, rt0->stripe_done[k], 0));
        const size_t off = (size_t)row0 * rt0->w * 4, cnt = (size_t)(row1 - row0) * rt0->w * 4;
        SAR_CUDA(cudaMemcpyAsync(rgba_u16 + off, rt0->image + off, cnt * sizeof(uint16_t), cudaMemcpyDeviceToHost, rt0->copy_stream));
    }
    SAR_CUDA(cudaStreamSynchronize(rt0->copy_stream));
    return SAR_OK;
}

// ---- output conversion + raw encoders (src/bin/main.rs:40-100) -----------------------------------
// `write_image_matches` converts the FinalImage (main.rs:52-57: RGBA16 as is, to_rgb16, to_rgba8, to_rgb8)
// and hands `image.as_bytes()` to an encoder of the `image` crate (PAM main.rs:62-68, BMP :70-76, PNG :78-89).
// Here the conversion runs on the device (convert_kernel) and the two RAW containers are written around it:
// the host only formats the header.  The PNG branch: stored blocks (SAR_FILE_PNG) or compressed on the device (sar_runtime_encode_png).
struct OutSpec {
    uint32_t fmt = SAR_PIX_RGBA16, container = SAR_FILE_RAW;
    uint32_t order = ORDER_NATIVE;
    size_t row_stride = 0, payload = 0, header = 0, trailer = 0, total = 0;
    uint8_t head[160] = {0};
    // PNG only: the raw (filtered scanline) stream, its stored blocks and the partial-checksum chunks
    size_t raw_row = 0, raw_len = 0, n_blocks = 0, n_crc = 0, n_adler = 0, sums_bytes = 0;
    // compressed PNG (SAR_FILE_PNG_DEFLATE / sar_runtime_encode_png): device scratch layout and bounds
    bool deflate = false;
    size_t n_chunks = 0, pay_bound = 0, n_crc_max = 0, off_chunks = 0, off_sizes = 0, off_offsets = 0, off_pay = 0, off_crc = 0,
           off_adler = 0, scratch = 0;
};

// ---- CRC-32 (PNG / zlib polynomial, reflected) and Adler-32 folding of the device's partial sums -----------
static uint32_t g_crc_table[256];
static uint32_t g_crc_shift_chunk[32];          // operator: CRC register after PNG_CHUNK zero bytes, per input bit
static std::once_flag g_crc_once;
static uint32_t crc_zero_byte(uint32_t c) { return g_crc_table[c & 0xFFu] ^ (c >> 8); }
static uint32_t gf2_apply(const uint32_t *mat, uint32_t v)
{
    uint32_t r = 0;
    for (int i = 0; v; v >>= 1, ++i) if (v & 1u) r ^= mat[i];
    return r;
}
static void crc_init()
{
    for (uint32_t n = 0; n < 256; ++n) {
        uint32_t c = n;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        g_crc_table[n] = c;
    }
    uint32_t m[32], sq[32];
    for (int i = 0; i < 32; ++i) m[i] = crc_zero_byte(1u << i);          // one zero byte
    for (size_t len = 1; len < PNG_CHUNK; len <<= 1) {                     // square up to PNG_CHUNK (a power of two) zero bytes
        for (int i = 0; i < 32; ++i) sq[i] = gf2_apply(m, m[i]);
        memcpy(m, sq, sizeof m);
    }
    memcpy(g_crc_shift_chunk, m, sizeof m);
}
static uint32_t crc_bytes(uint32_t c, const uint8_t *p, size_t n)          // register update, no pre/post conditioning
{
    for (size_t i = 0; i < n; ++i) c = g_crc_table[(c ^ p[i]) & 0xFFu] ^ (c >> 8);
    return c;
}
static void put_be32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; }
// the 20 bytes after the IDAT payload: Adler-32 of the raw stream, CRC-32 of the IDAT chunk, the IEND chunk
static void png_fold(const uint32_t *crc_part, size_t n_crc, size_t payload_len, const unsigned long long *ad, size_t n_adler,
                     size_t raw_len, uint8_t *out)
{
    std::call_once(g_crc_once, crc_init);
    // Adler-32 (zlib): a = 1 + sum d, b = sum of the running a, both mod 65521
    unsigned long long a = 1, b = 0;
    for (size_t i = 0; i < n_adler; ++i) {
        const size_t len = (i + 1) * PNG_CHUNK <= raw_len ? PNG_CHUNK : raw_len - i * PNG_CHUNK;
        b = (b + (unsigned long long)len * a + ad[2 * i + 1]) % 65521ull;
        a = (a + ad[2 * i]) % 65521ull;
    }
    const uint32_t adler = (uint32_t)((b << 16) | a);
    // CRC-32 over "IDAT" + zlib header + payload + Adler-32: linear in (register, data), so the register is advanced over
    // each chunk with the precomputed zero-bytes operator and the chunk's own contribution is XORed in
    const uint8_t pre[6] = {'I', 'D', 'A', 'T', 0x78, 0x01};
    uint32_t c = crc_bytes(0xFFFFFFFFu, pre, 6);
    for (size_t i = 0; i < n_crc; ++i) {
        const size_t len = (i + 1) * PNG_CHUNK <= payload_len ? PNG_CHUNK : payload_len - i * PNG_CHUNK;
        if (len == PNG_CHUNK) c = gf2_apply(g_crc_shift_chunk, c);
        else for (size_t k = 0; k < len; ++k) c = crc_zero_byte(c);
        c ^= crc_part[i];
    }
    uint8_t ab[4];
    put_be32(ab, adler);
    c = crc_bytes(c, ab, 4) ^ 0xFFFFFFFFu;
    memcpy(out, ab, 4);
    put_be32(out + 4, c);
    const uint8_t iend[12] = {0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xAE, 0x42, 0x60, 0x82};
    memcpy(out + 8, iend, 12);
}
static void png_trailer(const OutSpec &o, const uint8_t *sums, uint8_t *out)
{
    png_fold(reinterpret_cast<const uint32_t *>(sums), o.n_crc, o.payload,
             reinterpret_cast<const unsigned long long *>(sums + align_up(o.n_crc * 4, 8)), o.n_adler, o.raw_len, out);
}
static int make_outspec(uint32_t w, uint32_t h, uint32_t fmt, uint32_t container, OutSpec &o)
{
    if (fmt > SAR_PIX_RGB8) return fail(SAR_ERR_INVALID, "pixel_format %u", fmt);
    if (container > SAR_FILE_PNG) return fail(SAR_ERR_INVALID, "container %u", container);
    const bool wide = fmt == SAR_PIX_RGBA16 || fmt == SAR_PIX_RGB16, alpha = fmt == SAR_PIX_RGBA16 || fmt == SAR_PIX_RGBA8;
    const size_t bpp = (alpha ? 4 : 3) * (wide ? 2 : 1);
    o.fmt = fmt; o.container = container;
    o.row_stride = (size_t)w * bpp;
    o.order = ORDER_NATIVE;
    o.header = 0;
    if (container == SAR_FILE_PAM) {
        // image 0.25 pnm encoder, ArbitraryMap subtype: "P7\nWIDTH w\nHEIGHT h\nDEPTH d\nMAXVAL m\nTUPLTYPE t\nENDHDR\n",
        // 16-bit samples most significant byte first
        o.order = wide ? ORDER_BIG_ENDIAN : ORDER_NATIVE;
        o.header = (size_t)snprintf((char *)o.head, sizeof o.head, "P7\nWIDTH %u\nHEIGHT %u\nDEPTH %u\nMAXVAL %u\nTUPLTYPE %s\nENDHDR\n",
                                    w, h, alpha ? 4u : 3u, wide ? 65535u : 255u, alpha ? "RGB_ALPHA" : "RGB");
    } else if (container == SAR_FILE_BMP) {
        // image 0.25 bmp encoder: 8-bit only (a 16-bit image makes the reference panic in write_image's unwrap, main.rs:36);
        // Rgb8 -> BITMAPINFOHEADER, 24 bpp; Rgba8 -> BITMAPV4HEADER, 32 bpp BI_BITFIELDS, BGRA; rows bottom-up, padded to 4 bytes
        if (wide) return fail(SAR_ERR_UNSUPPORTED, "BMP holds 8-bit samples only (the reference's BmpEncoder rejects 16-bit images)");
        o.order = ORDER_BMP;
        o.row_stride = align_up((size_t)w * bpp, 4);
        const uint32_t dib = alpha ? 108u : 40u, off = 14u + dib;
        const uint64_t img = (uint64_t)o.row_stride * h;
        if (img + off > 0xFFFFFFFFull) return fail(SAR_ERR_INVALID, "image too large for a BMP file");
        uint8_t *p = o.head;
        auto u16le = [&](uint32_t v) { *p++ = (uint8_t)v; *p++ = (uint8_t)(v >> 8); };
        auto u32le = [&](uint32_t v) { u16le(v & 0xFFFFu); u16le(v >> 16); };
        *p++ = 'B'; *p++ = 'M'; u32le((uint32_t)(img + off)); u16le(0); u16le(0); u32le(off);
        u32le(dib); u32le(w); u32le(h); u16le(1); u16le(alpha ? 32 : 24); u32le(alpha ? 3u : 0u); u32le((uint32_t)img);
        u32le(0); u32le(0); u32le(0); u32le(0);
        if (alpha) {
            u32le(0xFFu << 16); u32le(0xFFu << 8); u32le(0xFFu); u32le(0xFFu << 24);   // R, G, B, A masks
            u32le(0x73524742u);                                                          // "sRGB"
            for (int k = 0; k < 12; ++k) u32le(0);                                       // endpoints + gamma
        }
        o.header = (size_t)(p - o.head);
    }
    o.payload = o.row_stride * h;
    o.trailer = 0;
    if (container == SAR_FILE_PNG) {
        // signature, IHDR, then ONE IDAT chunk holding a zlib stream of stored blocks (no compression)
        std::call_once(g_crc_once, crc_init);
        o.order = wide ? ORDER_BIG_ENDIAN : ORDER_NATIVE;
        o.raw_row = 1 + (size_t)w * bpp;
        o.raw_len = o.raw_row * h;
        o.n_blocks = (o.raw_len + 65534) / 65535;
        o.payload = o.raw_len + 5 * o.n_blocks;
        if (2 + o.payload + 4 > 0x7FFFFFFFull) return fail(SAR_ERR_INVALID, "image too large for one IDAT chunk");
        o.n_crc = (o.payload + PNG_CHUNK - 1) / PNG_CHUNK;
        o.n_adler = (o.raw_len + PNG_CHUNK - 1) / PNG_CHUNK;
        o.sums_bytes = align_up(o.n_crc * 4, 8) + o.n_adler * 16;
        uint8_t *p = o.head;
        const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
        memcpy(p, sig, 8); p += 8;
        put_be32(p, 13); p += 4;
        uint8_t *ihdr = p;
        memcpy(p, "IHDR", 4); p += 4;
        put_be32(p, w); p += 4; put_be32(p, h); p += 4;
        *p++ = wide ? 16 : 8; *p++ = alpha ? 6 : 2; *p++ = 0; *p++ = 0; *p++ = 0;      // bit depth, colour type, deflate, adaptive, no interlace
        put_be32(p, crc_bytes(0xFFFFFFFFu, ihdr, 17) ^ 0xFFFFFFFFu); p += 4;
        put_be32(p, (uint32_t)(2 + o.payload + 4)); p += 4;
        memcpy(p, "IDAT", 4); p += 4;
        *p++ = 0x78; *p++ = 0x01;                                                       // zlib header: deflate, 32 K window, no dictionary
        o.header = (size_t)(p - o.head);
        o.trailer = 20;
    }
    o.total = o.header + o.payload + o.trailer;
    return SAR_OK;
}

size_t sar_encoded_size(uint32_t width, uint32_t height, uint32_t pixel_format, uint32_t container)
{
    OutSpec o;
    if (check_dims(width, height) || make_outspec(width, height, pixel_format, container, o)) return 0;
    return o.total;
}

int sar_encode_header(uint32_t width, uint32_t height, uint32_t pixel_format, uint32_t container, uint8_t *out, size_t out_bytes,
                      size_t *header_bytes)
{
    OutSpec o;
    if (int rc = check_dims(width, height)) return rc;
    if (int rc = make_outspec(width, height, pixel_format, container, o)) return rc;
    if (header_bytes) *header_bytes = o.header;
    if (out) {
        if (out_bytes < o.header) return fail(SAR_ERR_INVALID, "header needs %zu bytes", o.header);
        memcpy(out, o.head, o.header);
    }
    return SAR_OK;
}

int sar_runtime_encode(sar_runtime *rt, uint32_t pixel_format, uint32_t container, uint8_t *out, size_t out_bytes, void *stream)
{
    if (!rt || !out) return fail(SAR_ERR_INVALID, "NULL argument");
    OutSpec o;
    if (int rc = make_outspec(rt->w, rt->h, pixel_format, container, o)) return rc;
    if (out_bytes < o.total) return fail(SAR_ERR_INVALID, "output needs %zu bytes (got %zu)", o.total, out_bytes);
    SAR_CUDA(cudaSetDevice(rt->device));
    cudaStream_t s = pick(rt, stream);
    const size_t sums_off = align_up(o.payload, 256);
    if (int rc = ensure_scratch(rt, sums_off + o.sums_bytes)) return rc;
    uint8_t *d_pay = (uint8_t *)rt->d_scratch;
    std::vector<uint8_t> sums(o.sums_bytes);
    if (o.container == SAR_FILE_PNG) {
        launch_png_pack(rt->image, d_pay, rt->w, rt->h, o.fmt, o.raw_row, o.raw_len, o.n_blocks, s);
        launch_png_sums(d_pay, o.payload, o.raw_len, (uint32_t *)(d_pay + sums_off),
                        (unsigned long long *)(d_pay + sums_off + align_up(o.n_crc * 4, 8)), o.n_crc, o.n_adler, s);
        SAR_CUDA(cudaGetLastError());
        SAR_CUDA(cudaMemcpyAsync(sums.data(), d_pay + sums_off, o.sums_bytes, cudaMemcpyDeviceToHost, s));
    } else {
        launch_convert(rt->image, d_pay, rt->w, rt->h, o.fmt, o.order, o.row_stride, s);
        SAR_CUDA(cudaGetLastError());
    }
    memcpy(out, o.head, o.header);
    SAR_CUDA(cudaMemcpyAsync(out + o.header, d_pay, o.payload, cudaMemcpyDeviceToHost, s));
    SAR_CUDA(cudaStreamSynchronize(s));
    if (o.container == SAR_FILE_PNG) png_trailer(o, sums.data(), out + o.header + o.payload);
    return SAR_OK;
}

// ---- PNG with the compressor (main.rs:78-89): Sub-filtered scanlines, one run-length + dynamic-Huffman deflate block per
// 16 KB, all on the device (sar_deflate.cu); the host patches the IDAT length and folds the checksums.
static int png_plan(uint32_t w, uint32_t h, uint32_t fmt, OutSpec &o)
{
    if (int rc = check_dims(w, h)) return rc;
    if (int rc = make_outspec(w, h, fmt, SAR_FILE_PNG, o)) return rc;
    o.deflate = true;
    o.n_chunks = (o.raw_len + dfl::CHUNK - 1) / dfl::CHUNK;
    o.pay_bound = o.raw_len + 5 * o.n_chunks;                         // every block at worst stored
    if (2 + o.pay_bound + 4 > 0x7FFFFFFFull) return fail(SAR_ERR_INVALID, "image too large for one IDAT chunk");
    o.n_crc_max = (o.pay_bound + PNG_CHUNK - 1) / PNG_CHUNK;
    o.n_adler = (o.raw_len + PNG_CHUNK - 1) / PNG_CHUNK;
    o.off_chunks = align_up(o.raw_len, 256);
    o.off_sizes = o.off_chunks + align_up(o.n_chunks * dfl::CHUNK_CAP, 256);
    o.off_offsets = o.off_sizes + align_up(o.n_chunks * sizeof(uint32_t), 256);
    o.off_pay = o.off_offsets + align_up((o.n_chunks + 1) * sizeof(unsigned long long), 256);
    o.off_crc = o.off_pay + align_up(o.pay_bound, 256);
    o.off_adler = o.off_crc + align_up(o.n_crc_max * sizeof(uint32_t), 256);
    o.scratch = o.off_adler + o.n_adler * 2 * sizeof(unsigned long long);
    // sequence driver: what one frame needs on the host (header + worst-case stream + trailer) and the pinned sums block
    // [total u64][crc partials][adler partials]
    o.payload = o.pay_bound;
    o.total = o.header + o.pay_bound + 20;
    o.sums_bytes = 8 + align_up(o.n_crc_max * sizeof(uint32_t), 8) + o.n_adler * 2 * sizeof(unsigned long long);
    return SAR_OK;
}
static void png_deflate_launch(const sar_runtime *rt, const OutSpec &o, uint8_t *base, cudaStream_t s)
{
    launch_png_deflate(rt->image, rt->w, rt->h, o.fmt, o.raw_row, o.raw_len, base, base + o.off_chunks,
                       (uint32_t *)(base + o.off_sizes), (unsigned long long *)(base + o.off_offsets), base + o.off_pay,
                       (uint32_t *)(base + o.off_crc), (unsigned long long *)(base + o.off_adler), o.n_crc_max, o.n_adler, s);
}

size_t sar_png_bound(uint32_t width, uint32_t height, uint32_t pixel_format)
{
    OutSpec o;
    if (png_plan(width, height, pixel_format, o)) return 0;
    return o.total;
}

int sar_runtime_encode_png(sar_runtime *rt, uint32_t pixel_format, uint8_t *out, size_t out_capacity, size_t *out_bytes, void *stream)
{
    if (!rt || !out || !out_bytes) return fail(SAR_ERR_INVALID, "NULL argument");
    *out_bytes = 0;
    OutSpec o;
    if (int rc = png_plan(rt->w, rt->h, pixel_format, o)) return rc;
    if (out_capacity < o.header + 20) return fail(SAR_ERR_INVALID, "output capacity %zu is below the fixed parts of a PNG", out_capacity);
    SAR_CUDA(cudaSetDevice(rt->device));
    cudaStream_t s = pick(rt, stream);
    if (int rc = ensure_scratch(rt, o.scratch)) return rc;
    uint8_t *base = (uint8_t *)rt->d_scratch;
    png_deflate_launch(rt, o, base, s);
    SAR_CUDA(cudaGetLastError());
    unsigned long long total = 0;
    SAR_CUDA(cudaMemcpyAsync(&total, base + o.off_offsets + o.n_chunks * sizeof(unsigned long long), sizeof total, cudaMemcpyDeviceToHost, s));
    SAR_CUDA(cudaStreamSynchronize(s));
    if (total == 0 || total > o.pay_bound) return fail(SAR_ERR_CUDA, "deflate produced %llu bytes (bound %zu)", total, o.pay_bound);
    const size_t need = o.header + (size_t)total + 20;
    if (out_capacity < need) return fail(SAR_ERR_INVALID, "output needs %zu bytes (got %zu; sar_png_bound gives the worst case)", need, out_capacity);
    const size_t n_crc = ((size_t)total + PNG_CHUNK - 1) / PNG_CHUNK;
    std::vector<uint32_t> crc(n_crc);
    std::vector<unsigned long long> adler(2 * o.n_adler);
    memcpy(out, o.head, o.header);
    put_be32(out + o.header - 10, (uint32_t)(2 + total + 4));        // IDAT length: zlib header + deflate stream + Adler-32
    SAR_CUDA(cudaMemcpyAsync(out + o.header, base + o.off_pay, (size_t)total, cudaMemcpyDeviceToHost, s));
    SAR_CUDA(cudaMemcpyAsync(crc.data(), base + o.off_crc, n_crc * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    SAR_CUDA(cudaMemcpyAsync(adler.data(), base + o.off_adler, adler.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    SAR_CUDA(cudaStreamSynchronize(s));
    png_fold(crc.data(), n_crc, (size_t)total, adler.data(), o.n_adler, o.raw_len, out + o.header + (size_t)total);
    *out_bytes = need;
    return SAR_OK;
}

int sar_write_file(const char *path, const uint8_t *bytes, size_t n_bytes)
{
    if (!path || (!bytes && n_bytes)) return fail(SAR_ERR_INVALID, "NULL argument");
    FILE *f = fopen(path, "wb");                                 // File::create(path).unwrap(), main.rs:102-104
    if (!f) return fail(SAR_ERR_INVALID, "cannot create '%s'", path);
    const size_t n = fwrite(bytes, 1, n_bytes, f);
    const int rc = fclose(f);
    if (n != n_bytes || rc != 0) return fail(SAR_ERR_INVALID, "short write to '%s'", path);
    return SAR_OK;
}

// ---- frame sequences -------------------------------------------------------------------------
// The per-frame loop of the reference's binary (src/bin/main.rs:496-512): for each angle,
// config.angle = angle; image = render_parallel(...); hand the image to an encoder thread.
// Frames are independent, so they round-robin over the renderer's devices (replicas, no
// collective); on each device the frame's device→host copy overlaps the next frame's render.
static int seq_prepare(sar_renderer *r, size_t d, const sar_config &cfg, bool need_stage, const OutSpec &o)
{
    seq_device &q = r->seq[d];
    const size_t bytes = (size_t)cfg.width * cfg.height * 4 * sizeof(uint16_t);
    SAR_CUDA(cudaSetDevice(r->devices[d]));
    if (q.img_bytes != bytes || (q.rt[0] && (q.rt[0]->w != cfg.width || q.rt[0]->h != cfg.height))) {
        seq_release(r, d);
        q.img_bytes = bytes;
    }
    if (!q.copy_stream) SAR_CUDA(cudaStreamCreateWithFlags(&q.copy_stream, cudaStreamNonBlocking));
    if (!q.pay_stream) SAR_CUDA(cudaStreamCreateWithFlags(&q.pay_stream, cudaStreamNonBlocking));
    if (!q.h_max) SAR_CUDA(cudaHostAlloc((void **)&q.h_max, 2 * sizeof(uint32_t), cudaHostAllocPortable));
    for (int k = 0; k < 2; ++k) {
        if (!q.rt[k]) if (int rc = sar_runtime_new(cfg.width, cfg.height, r->devices[d], &q.rt[k])) return rc;
        if (need_stage && (!q.stage[k] || q.stage_bytes < o.total)) {
            if (q.stage[k]) cudaFreeHost(q.stage[k]);
            q.stage[k] = nullptr;
            SAR_CUDA(cudaHostAlloc((void **)&q.stage[k], o.total, cudaHostAllocPortable));
        }
        const bool convert = !(o.fmt == SAR_PIX_RGBA16 && o.order == ORDER_NATIVE) || o.container == SAR_FILE_PNG;
        const size_t enc_need = o.deflate ? o.scratch : align_up(o.payload, 256) + o.sums_bytes;
        if (convert && (!q.enc[k] || q.enc_bytes < enc_need)) {
            cudaFree(q.enc[k]);
            q.enc[k] = nullptr;
            SAR_CUDA(cudaMalloc((void **)&q.enc[k], enc_need));
        }
        if (o.sums_bytes && (!q.h_sums[k] || q.sums_cap < o.sums_bytes)) {
            if (q.h_sums[k]) cudaFreeHost(q.h_sums[k]);
            q.h_sums[k] = nullptr;
            SAR_CUDA(cudaHostAlloc((void **)&q.h_sums[k], o.sums_bytes, cudaHostAllocPortable));
        }
        if (!q.max_ready[k]) SAR_CUDA(cudaEventCreateWithFlags(&q.max_ready[k], cudaEventDisableTiming));
        if (!q.rendered[k]) SAR_CUDA(cudaEventCreateWithFlags(&q.rendered[k], cudaEventDisableTiming));
        if (!q.copied[k]) SAR_CUDA(cudaEventCreateWithFlags(&q.copied[k], cudaEventDisableTiming));
    }
    // capacities of what was (re)allocated above
    if (need_stage && q.stage_bytes < o.total) q.stage_bytes = o.total;
    const size_t enc_cap = o.deflate ? o.scratch : align_up(o.payload, 256) + o.sums_bytes;
    if ((!(o.fmt == SAR_PIX_RGBA16 && o.order == ORDER_NATIVE) || o.container == SAR_FILE_PNG) && q.enc_bytes < enc_cap) q.enc_bytes = enc_cap;
    if (o.sums_bytes && q.sums_cap < o.sums_bytes) q.sums_cap = o.sums_bytes;
    return SAR_OK;
}

// frames_out: n_frames x o.total bytes (may be NULL); cb16 / cb8: at most one of them
static int sequence_core(sar_renderer *r, const sar_config *cfg_in, const double *angles_rad, uint32_t n_frames,
                         uint64_t jobs_per_thread, uint64_t seed, uint32_t flags, const OutSpec &o, uint8_t *frames_out,
                         sar_frame_callback cb16, sar_frame_bytes_callback cb8, void *user)
{
    if (!r || !cfg_in || (!angles_rad && n_frames)) return fail(SAR_ERR_INVALID, "NULL argument");
    if (!frames_out && !cb16 && !cb8) return fail(SAR_ERR_INVALID, "need a frame array or a callback");
    uint8_t *const rgba_frames = frames_out;
    const bool pngz = o.deflate;                                   // compressed PNG: frames of different sizes, callback only
    const bool png = o.container == SAR_FILE_PNG && !pngz;
    const bool convert = !(o.fmt == SAR_PIX_RGBA16 && o.order == ORDER_NATIVE) || png || pngz;
    if (pngz && (frames_out || !cb8)) return fail(SAR_ERR_INVALID, "compressed PNG frames differ in size: pass a callback and no frame array");
    if (jobs_per_thread == 0) return fail(SAR_ERR_INVALID, "jobs_per_thread must be non-zero");
    if (flags & ~SAR_SEQ_SHARED_POINTS) return fail(SAR_ERR_INVALID, "unknown flags 0x%x", flags);
    if (int rc = check_config(cfg_in, nullptr)) return rc;
    if (int rc = check_dims(cfg_in->width, cfg_in->height)) return rc;
    if (n_frames == 0) return SAR_OK;
    const size_t nd = r->devices.size();
    const bool shared = (flags & SAR_SEQ_SHARED_POINTS) != 0;
    const size_t frame_u16 = o.total;          // bytes of one delivered frame

    // every device renders whole frames with its own lanes: the decomposition of a frame is that
    // of a one-device render_parallel (lib.rs:1058-1062)
    std::vector<uint32_t> threads(nd), lanes(nd);
    std::vector<uint64_t> jobs(nd);
    std::vector<sar_config> cfgs(nd, *cfg_in);
    for (size_t d = 0; d < nd; ++d) {
        threads[d] = renderer_threads_for(r, d, jobs_per_thread, cfg_in->iterations * nd);   // a whole frame per device
        jobs[d] = (uint64_t)threads[d] * jobs_per_thread;
        lanes[d] = r->threads_per_device ? threads[d] : (uint32_t)(jobs[d] < renderer_lanes(r, d) ? jobs[d] : renderer_lanes(r, d));
        cfgs[d].iterations = cfg_in->iterations / threads[d] / jobs_per_thread;
        if (int rc = seq_prepare(r, d, *cfg_in, rgba_frames == nullptr, o)) return rc;
        if (shared) {   // warm the one shared list of start points once (lib.rs:748-752)
            seq_device &q = r->seq[d];
            const size_t need = (size_t)jobs[d] * 3 * sizeof(double);
            if (q.warm_cap < need) { cudaFree(q.warm); q.warm = nullptr; q.warm_cap = 0; SAR_CUDA(cudaMalloc((void **)&q.warm, need)); q.warm_cap = need; }
            IterParams p;
            make_iter_params(&cfgs[d], q.rt[0], p);
            p.init = nullptr; p.seed = seed; p.first_job = 0; p.n_jobs = jobs[d];
            launch_warm(p, q.warm, q.rt[0]->stream);
            SAR_CUDA(cudaGetLastError());
            SAR_CUDA(cudaStreamSynchronize(q.rt[0]->stream));
        }
    }

    auto finalize = [&](uint32_t g) -> int {     // wait for frame g's host copy, hand it to the caller
        const size_t d = g % nd;
        const int slot = (int)((g / nd) % 2);
        SAR_CUDA(cudaSetDevice(r->devices[d]));
        SAR_CUDA(cudaEventSynchronize(r->seq[d].copied[slot]));
        uint8_t *bytes = rgba_frames ? rgba_frames + (size_t)g * frame_u16 : r->seq[d].stage[slot];
        if (pngz) {
            // the frame's stream length, CRC and Adler partial sums have arrived; fetch exactly that many bytes
            seq_device &q = r->seq[d];
            const uint8_t *hs = q.h_sums[slot];
            const unsigned long long total = *reinterpret_cast<const unsigned long long *>(hs);
            if (total == 0 || total > o.pay_bound) return fail(SAR_ERR_CUDA, "deflate produced %llu bytes (bound %zu)", total, o.pay_bound);
            memcpy(bytes, o.head, o.header);
            put_be32(bytes + o.header - 10, (uint32_t)(2 + total + 4));
            SAR_CUDA(cudaMemcpyAsync(bytes + o.header, q.enc[slot] + o.off_pay, (size_t)total, cudaMemcpyDeviceToHost, q.pay_stream));
            SAR_CUDA(cudaStreamSynchronize(q.pay_stream));
            png_fold(reinterpret_cast<const uint32_t *>(hs + 8), ((size_t)total + PNG_CHUNK - 1) / PNG_CHUNK, (size_t)total,
                     reinterpret_cast<const unsigned long long *>(hs + 8 + align_up(o.n_crc_max * sizeof(uint32_t), 8)), o.n_adler, o.raw_len,
                     bytes + o.header + (size_t)total);
            cb8(user, g, bytes, o.header + (size_t)total + 20);
            return SAR_OK;
        }
        if (png) png_trailer(o, r->seq[d].h_sums[slot], bytes + o.header + o.payload);   // Adler-32, IDAT CRC, IEND
        if (cb16) cb16(user, g, reinterpret_cast<const uint16_t *>(bytes));
        if (cb8) cb8(user, g, bytes, o.total);
        return SAR_OK;
    };

    // Frame g's second half: its Runtime.max has been copied to the host — ln(max + 1), the log base of
    // lib.rs:860, is computed by the host libm like every blocking entry point does, so the frame is
    // bit-exact whatever max is (a solar-sail frame's NaN sink is far beyond the ln table) —, then
    // colourise and start the copy out.  Both Runtimes of a device share ONE compute stream: the next frame's
    // render is already queued behind this frame's max, so the device never waits for the host, and no two of
    // the big kernels run side by side (a reset or colourise streaming 100 MB through the L2 while the other
    // Runtime's iterate kernel works out of it cost 24 %; measured, tools/seq_bench.py).
    auto colourise = [&](uint32_t g) -> int {
        const size_t d = g % nd;
        const int slot = (int)((g / nd) % 2);
        seq_device &q = r->seq[d];
        sar_runtime *rt = q.rt[slot];
        SAR_CUDA(cudaSetDevice(rt->device));
        // compressed PNG: the slot's previous frame still sits in the slot's device buffers until finalize() has fetched it
        if (pngz && g >= 2 * nd) if (int rc = finalize(g - 2 * (uint32_t)nd)) return rc;
        SAR_CUDA(cudaSetDevice(rt->device));
        SAR_CUDA(cudaEventSynchronize(q.max_ready[slot]));
        sar_config cfg = cfgs[d];
        cfg.angle = angles_rad[g];
        ColorParams cp;
        make_color_params(&cfg, rt, cp, 0, rt->h, &q.h_max[slot]);
        cudaStream_t cs = q.rt[0]->stream;                                      // the device's compute stream
        SAR_CUDA(cudaStreamWaitEvent(cs, q.copied[slot], 0));                   // the slot's previous image has left the device
        launch_colorize(cp, rt->fast, rt->rec, rt->scal, rt->image, nullptr, cs);           // colorize, lib.rs:1080
        SAR_CUDA(cudaGetLastError());
        const size_t sums_off = align_up(o.payload, 256);
        if (pngz) {                                                             // main.rs:78-89, compressor included
            png_deflate_launch(rt, o, q.enc[slot], cs);
            SAR_CUDA(cudaGetLastError());
        } else if (png) {                                                       // main.rs:78-89 minus the compressor
            launch_png_pack(rt->image, q.enc[slot], rt->w, rt->h, o.fmt, o.raw_row, o.raw_len, o.n_blocks, cs);
            launch_png_sums(q.enc[slot], o.payload, o.raw_len, (uint32_t *)(q.enc[slot] + sums_off),
                            (unsigned long long *)(q.enc[slot] + sums_off + align_up(o.n_crc * 4, 8)), o.n_crc, o.n_adler, cs);
            SAR_CUDA(cudaGetLastError());
        } else if (convert) {                                                   // main.rs:52-57 on the device
            launch_convert(rt->image, q.enc[slot], rt->w, rt->h, o.fmt, o.order, o.row_stride, cs);
            SAR_CUDA(cudaGetLastError());
        }
        SAR_CUDA(cudaEventRecord(q.rendered[slot], cs));
        SAR_CUDA(cudaStreamWaitEvent(q.copy_stream, q.rendered[slot], 0));
        // the slot's previous frame (g - 2 nd) is handed to the caller here — its copy finished a frame ago, so the
        // host does not stall, and the slot's staging buffer is free again before this frame's copy is queued
        if (pngz) {
            // only the stream's length and the partial checksums now; the stream itself is fetched by finalize(), exact size
            uint8_t *hs = q.h_sums[slot];
            SAR_CUDA(cudaMemcpyAsync(hs, q.enc[slot] + o.off_offsets + o.n_chunks * sizeof(unsigned long long), 8, cudaMemcpyDeviceToHost, q.copy_stream));
            SAR_CUDA(cudaMemcpyAsync(hs + 8, q.enc[slot] + o.off_crc, o.n_crc_max * sizeof(uint32_t), cudaMemcpyDeviceToHost, q.copy_stream));
            SAR_CUDA(cudaMemcpyAsync(hs + 8 + align_up(o.n_crc_max * sizeof(uint32_t), 8), q.enc[slot] + o.off_adler,
                                     o.n_adler * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, q.copy_stream));
            SAR_CUDA(cudaEventRecord(q.copied[slot], q.copy_stream));
            return SAR_OK;
        }
        if (g >= 2 * nd) if (int rc = finalize(g - 2 * (uint32_t)nd)) return rc;
        uint8_t *dst = rgba_frames ? rgba_frames + (size_t)g * frame_u16 : q.stage[slot];
        memcpy(dst, o.head, o.header);
        SAR_CUDA(cudaMemcpyAsync(dst + o.header, convert ? (const void *)q.enc[slot] : (const void *)rt->image, o.payload,
                                 cudaMemcpyDeviceToHost, q.copy_stream));
        if (png) SAR_CUDA(cudaMemcpyAsync(q.h_sums[slot], q.enc[slot] + sums_off, o.sums_bytes, cudaMemcpyDeviceToHost, q.copy_stream));
        SAR_CUDA(cudaEventRecord(q.copied[slot], q.copy_stream));
        return SAR_OK;
    };
    for (uint32_t f = 0; f < n_frames; ++f) {
        const size_t d = f % nd;
        const int slot = (int)((f / nd) % 2);
        seq_device &q = r->seq[d];
        sar_runtime *rt = q.rt[slot];
        SAR_CUDA(cudaSetDevice(rt->device));
        sar_config cfg = cfgs[d];
        cfg.angle = angles_rad[f];                                             // main.rs:497
        cudaStream_t cs = q.rt[0]->stream;                                      // the device's compute stream (both Runtimes)
        if (int rc = sar_runtime_reset_async(rt, cs)) return rc;                // Runtime::reset per frame, lib.rs:951
        if (shared) {
            if (int rc = render_launch(&cfg, rt, q.warm, 0, 0, jobs[d], lanes[d], cs, true)) return rc;
        } else {
            // fresh start points per frame, like the reference: frame f takes the next jobs[d] points of the stream
            if (int rc = render_launch(&cfg, rt, nullptr, seed, (uint64_t)f * jobs[d], jobs[d], lanes[d], cs)) return rc;
        }
        if (int rc = sar_runtime_max_async(rt, 0, 0, cs)) return rc;
        if (int rc = ensure_depth_fold(&cfg, rt, cs)) return rc;
        SAR_CUDA(cudaMemcpyAsync(&q.h_max[slot], &rt->scal->max, sizeof(uint32_t), cudaMemcpyDeviceToHost, cs));
        SAR_CUDA(cudaEventRecord(q.max_ready[slot], cs));
        if (f >= nd) if (int rc = colourise(f - (uint32_t)nd)) return rc;       // the device's previous frame, one behind
    }
    for (uint32_t g = n_frames > nd ? n_frames - (uint32_t)nd : 0; g < n_frames; ++g)
        if (int rc = colourise(g)) return rc;
    for (uint32_t g = n_frames > 2 * nd ? n_frames - 2 * (uint32_t)nd : 0; g < n_frames; ++g)
        if (int rc = finalize(g)) return rc;
    for (size_t d = 0; d < nd; ++d) {
        SAR_CUDA(cudaSetDevice(r->devices[d]));
        for (int k = 0; k < 2; ++k) SAR_CUDA(cudaStreamSynchronize(r->seq[d].rt[k]->stream));
    }
    return SAR_OK;
}

int sar_render_sequence(sar_renderer *r, const sar_config *cfg_in, const double *angles_rad, uint32_t n_frames,
                        uint64_t jobs_per_thread, uint64_t seed, uint32_t flags, uint16_t *rgba_frames,
                        sar_frame_callback cb, void *user)
{
    if (!cfg_in) return fail(SAR_ERR_INVALID, "NULL argument");
    OutSpec o;
    if (int rc = check_dims(cfg_in->width, cfg_in->height)) return rc;
    if (int rc = make_outspec(cfg_in->width, cfg_in->height, SAR_PIX_RGBA16, SAR_FILE_RAW, o)) return rc;
    return sequence_core(r, cfg_in, angles_rad, n_frames, jobs_per_thread, seed, flags, o, reinterpret_cast<uint8_t *>(rgba_frames),
                         cb, nullptr, user);
}

int sar_render_sequence_encoded(sar_renderer *r, const sar_config *cfg_in, const double *angles_rad, uint32_t n_frames,
                                uint64_t jobs_per_thread, uint64_t seed, uint32_t flags, uint32_t pixel_format,
                                uint32_t container, uint8_t *frames_out, sar_frame_bytes_callback cb, void *user)
{
    if (!cfg_in) return fail(SAR_ERR_INVALID, "NULL argument");
    OutSpec o;
    if (int rc = check_dims(cfg_in->width, cfg_in->height)) return rc;
    if (container == SAR_FILE_PNG_DEFLATE) {
        if (int rc = png_plan(cfg_in->width, cfg_in->height, pixel_format, o)) return rc;
    } else if (int rc = make_outspec(cfg_in->width, cfg_in->height, pixel_format, container, o)) return rc;
    return sequence_core(r, cfg_in, angles_rad, n_frames, jobs_per_thread, seed, flags, o, frames_out, nullptr, cb, user);
}

}  // extern "C"
