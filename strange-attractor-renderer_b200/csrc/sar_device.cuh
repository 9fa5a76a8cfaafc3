// sar_device.cuh — device-side data layout and kernel launch interface (internal).
//
// HBM layout of one Runtime (reference: Runtime{count,steps,zbuf,max}, lib.rs:631-646),
// W*H pixels, row-major idx = y*W + x exactly like image::ImageBuffer:
//
//   fast[slot(idx)]  u64   (slot = (idx * M) & (P-1), P = next power of two >= W*H.  For images up
//                    to 2^23 pixels M = 0x9E3779B1: a bijective scramble, so that the pixels
//                    sharing a 32-byte sector are far apart in the image and a dense region does
//                    not serialise its atomics on a few sectors (+15 % on the atomic ceiling at
//                    2048^2).  Larger images keep M = 1 (natural order): scrambling would spread
//                    the ~20 % live pixels over 57 % of the sectors and push the working set out
//                    of the 126 MB L2 (-20 % at 4096^2).  profiles/r1_sweep.md)
//                    bits 31..0  = count (u32, lib.rs:633)
//                    bits 63..32 = zhint: an order-preserving key of a z value that is
//                                  <= the pixel's recorded z ("a candidate below this loses")
//   rec[idx]   16 B  .x (low 8)  = steps as f64 bits (lib.rs:635)
//                    .y (high 8) = zkey(zbuf) << 32 | ~job   (zbuf f32, lib.rs:639)
//
// One 64-bit ATOMG.ADD per recorded iteration both increments the count (lib.rs:811)
// and returns the hint for the depth test (lib.rs:821); the 16-byte record is only
// touched on the rare winning path, with a 128-bit compare-and-swap ordered on .y.
// ~job (0xFFFFFFFF - job index) makes ties in z resolve to the earlier render() call,
// and within a call program order keeps the earlier iteration: exactly the outcome of
// running the reference's render() job after job on one Runtime (lib.rs:742-743).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sar {

// order-preserving map f32 -> u32 (a > b  <=>  zkey(a) > zkey(b) for non-NaN a,b; -0 is canonicalised first)
__host__ __device__ inline uint32_t zkey_from_bits(uint32_t b) { return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__host__ __device__ inline uint32_t zbits_from_key(uint32_t k) { return (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k; }

// f32 compare of two keys: -0.0 == +0.0.  canon_key folds zkey(-0.0) onto zkey(+0.0); rec_order does
// the same on a record's high word (zkey << 32 | ~job), so records compare as (z, earlier job) with
// the reference's notion of "z strictly greater" (lib.rs:728, 821).  Hints always hold canonical keys.
__host__ __device__ inline uint32_t canon_key(uint32_t k) { return k == 0x7FFFFFFFu ? 0x80000000u : k; }
__host__ __device__ inline unsigned long long rec_order(unsigned long long y)
{
    return (uint32_t)(y >> 32) == 0x7FFFFFFFu ? y + (1ull << 32) : y;
}

// pixel index -> slot of the `fast` array (see the layout comment above)
struct SlotMap { uint32_t mult, mask; };
constexpr uint32_t SLOT_SCRAMBLE = 0x9E3779B1u;
__host__ __device__ inline uint32_t slot_of(uint32_t idx, SlotMap m) { return (idx * m.mult) & m.mask; }

constexpr uint32_t ZKEY_SENTINEL = 0x407FFFFFu;   // zkey(-1.0f): Runtime::reset fills zbuf with -1.0 (lib.rs:693)
constexpr uint32_t ZKEY_POS_INF  = 0xFF800000u;   // zkey(+inf); +NaN keys are larger, -NaN keys are < zkey(-inf)
constexpr uint32_t ZKEY_ZERO     = 0x80000000u;   // zkey(+0.0f)
constexpr uint32_t ZKEY_NEG_ZERO = 0x7FFFFFFFu;   // zkey(-0.0f): directly below ZKEY_ZERO, but f32 `>` sees the two zeros equal
constexpr uint32_t ZKEY_FLT_MAX  = 0xFF7FFFFFu;   // zkey(f32::MAX)
constexpr unsigned long long FAST_RESET = (unsigned long long)(ZKEY_SENTINEL + 1u) << 32;  // count 0, hint just above the sentinel
constexpr unsigned long long REC_HI_RESET = ((unsigned long long)ZKEY_SENTINEL << 32) | 0xFFFFFFFFull;

constexpr unsigned int TILE_BLOCK = 896;                 // lanes of a tile-path block (one block per SM)
constexpr size_t TILE_SMEM_MAX = 200 * 1024;             // shared memory a tile may take: W*H*8 bytes -> up to 25 600 pixels

constexpr int SYNC_MAX_RANKS = 16;
enum SyncKind {                       // what a flag announces (see sar_runtime_signal_async)
    SYNC_RENDER_DONE = 0,             // rank r's trajectories of this frame are all in its accumulators
    SYNC_MERGE_DONE = 1,              // rank r has finished reading every peer's accumulators
    SYNC_MAX_READY = 2,               // stripe_max[r] holds rank r's stripe maximum
    SYNC_IMAGE_DONE = 3,              // rank r has stored its colourised stripe into this rank's image
    SYNC_IMAGE_FREE = 4,              // rank r (the image owner) is done with the image of the frame
    SYNC_KINDS = 5
};

struct Scalars {                      // device-resident scalar state of a Runtime
    unsigned long long nan_sink;      // iterations of NaN trajectories, owed to count[(0,0)]
    unsigned int max;                 // Runtime.max (lib.rs:643): kept current by the iterate kernel (without the NaN debt);
                                      // final after launch_fold_max() / launch_max()
    unsigned int zmax_key, zmin_key;  // Depth colourise fold (lib.rs:877-882)
    unsigned int pad;
    // Cross-GPU synchronisation over peer memory (one process per GPU).  flag[k][r] is
    // written ONLY by rank r (remotely, over NVLink) and polled locally; values are frame epochs
    // and only grow.  Never touched by reset.
    unsigned int sync_error;          // a wait timed out
    unsigned int pad2[3];
    unsigned int flag[SYNC_KINDS][SYNC_MAX_RANKS];
    unsigned int stripe_max[SYNC_MAX_RANKS];   // rank r's share of Runtime.max for the current frame (lib.rs:721-723)
    unsigned int stripe_zmax[SYNC_MAX_RANKS];  // ... and of the Depth fold (lib.rs:877-882), as canonical z keys
    unsigned int stripe_zmin[SYNC_MAX_RANKS];
    unsigned int done_counter[4];              // blocks that have finished the export / merge / colourise kernel of the frame
};
static_assert(sizeof(Scalars) <= 1024, "Scalars must fit its slot of the runtime allocation");

struct IterParams {                   // everything the iterate kernel reads; lives in the constant bank
    double c[3][10];                  // attractor coefficients; c[k][0] pre-reduced to 0.0 + 1.0*c0 (lib.rs:589-596)
    double m[3][3];                   // rotation matrix (lib.rs:755), host computed
    double ccx, ccy, ccz;             // center_camera (lib.rs:758)
    double cv, sv;                    // cos/sin(angle) (lib.rs:756-757), host computed
    double sam, ws, half_h;           // scale_adjusted_mid, width_scaled, height/2 (lib.rs:763-764, 786)
    double ct_offset, ct_factor;      // AdjustedVelocity (lib.rs:507-510); also ScreenBlend
    double ct_w[4];                   // ScreenBlend weights (include/sar.h)
    double c3[3][10];                 // cubic coefficients of attractor kind 1 (include/sar.h)
    unsigned long long *fast;
    ulonglong2 *rec;
    Scalars *scal;
    const double *init;               // n_jobs x 3 start points, or nullptr -> generated from seed
    unsigned long long seed;
    unsigned long long first_job;     // index of this launch's job 0 in the seed stream
    unsigned long long n_jobs;
    unsigned long long iterations;    // recorded iterations per job
    unsigned int job_key0;            // order key of job 0 (Runtime job counter)
    unsigned int W, H;
    unsigned int ct_kind;
    unsigned int attractor_kind;
    SlotMap slots;                    // pixel -> slot map of the fast array
#ifdef SAR_DIAGNOSTICS
    unsigned int diag_hot, diag_tab_entries;   // MODE 5 cost model: hot fraction (per 1024) and table entries per block
#endif
    unsigned int warmup;              // unrecorded steps before the recorded ones: 1000 (lib.rs:750), or 0 when `init` holds warmed states
};

struct ColorParams {
    double pal[17][3];                // palette + duplicated last entry (lib.rs:416-424)
    double pal_len;                   // count_f64 (lib.rs:421)
    double bright_offset, bright_factor;
    unsigned int palette_len, transparent, render_kind;
    unsigned int W, H, row0, rows;
    SlotMap slots;
    // ln(n) for n < lnlut_len, computed on the HOST with the platform libm — the function the
    // reference's f64::ln resolves to — so that `ln(count+1)/ln(max+1)` (lib.rs:860) carries the
    // same bits as a CPU run; larger arguments fall back to the device log (<= 1 ulp).
    const double *lnlut;
    unsigned int lnlut_len;
    unsigned int host_lnmax_valid;    // ln_max1_host was computed on the host from the current max
    double ln_max1_host;
};

// output conversion (src/bin/main.rs:52-57): pixel formats and sample orders of launch_convert
enum { PIX_RGBA16 = 0, PIX_RGB16 = 1, PIX_RGBA8 = 2, PIX_RGB8 = 3 };
enum { ORDER_NATIVE = 0, ORDER_BIG_ENDIAN = 1, ORDER_BMP = 2 };
void launch_convert(const uint16_t *rgba, uint8_t *out, unsigned int W, unsigned int H, unsigned int fmt, unsigned int order,
                    size_t row_stride, cudaStream_t s);

// auto-framing first pass (lib.rs:326-334): screen-space bounding box as order-preserving u64 keys
// (key(v) = bits ^ sign-fill, so unsigned min/max = f64 min/max); lo starts at ~0, hi at 0
struct BBoxAccum { unsigned long long lo[3], hi[3], diverged; };
void launch_bbox(const IterParams &p, BBoxAccum *acc, cudaStream_t s);

// PNG with stored deflate blocks: payload writer and per-chunk partial checksums (sar_kernels.cu)
constexpr size_t PNG_CHUNK = 4096;
void launch_png_pack(const uint16_t *rgba, uint8_t *out, unsigned int W, unsigned int H, unsigned int fmt, size_t raw_row,
                     size_t raw_len, size_t n_blocks, cudaStream_t s);
void launch_png_sums(const uint8_t *payload, size_t payload_len, size_t raw_len, uint32_t *crc, unsigned long long *adler,
                     size_t n_crc, size_t n_adler, cudaStream_t s);

// PNG with the compressor (sar_deflate.cu): filter, one deflate block per 16 KB, compaction, partial checksums
void launch_png_deflate(const uint16_t *rgba, unsigned int W, unsigned int H, unsigned int fmt, size_t raw_row, size_t raw_len,
                        uint8_t *raw, uint8_t *chunks, uint32_t *sizes, unsigned long long *offsets, uint8_t *pay,
                        uint32_t *crc, unsigned long long *adler, size_t n_crc_max, size_t n_adler, cudaStream_t s);
void bump_launches(unsigned int n);

// launchers (sar_kernels.cu); every one bumps the launch counter
void launch_reset(unsigned long long *fast, ulonglong2 *rec, Scalars *scal, size_t npix, size_t nslots, cudaStream_t s);
// returns true when the shared-memory tile path ran (small images; Runtime.max then needs the full reduction)
bool launch_iterate(const IterParams &p, unsigned int lanes, cudaStream_t s);
// the warm-up alone (lib.rs:748-752): start points -> states after p.warmup steps, out[3*job..]
void launch_warm(const IterParams &p, double *out, cudaStream_t s);
void launch_fold_max(const unsigned long long *fast, Scalars *scal, SlotMap slots, cudaStream_t s);
void launch_max(const unsigned long long *fast, const ulonglong2 *rec, Scalars *scal, size_t pix0, size_t npix, SlotMap slots, cudaStream_t s);
void launch_colorize(const ColorParams &cp, const unsigned long long *fast, const ulonglong2 *rec, const Scalars *scal,
                     uint16_t *rgba_u16, float *rgba_f32, cudaStream_t s);
void launch_unpack(const unsigned long long *fast, const ulonglong2 *rec, const Scalars *scal, size_t npix, SlotMap slots,
                   uint32_t *count, double *steps, float *zbuf, cudaStream_t s);
void launch_pack(unsigned long long *fast, ulonglong2 *rec, Scalars *scal, size_t npix, SlotMap slots,
                 const uint32_t *count, const double *steps, const float *zbuf, cudaStream_t s);
// Runtime::merge (lib.rs:708-738): z-only compare, ties keep dst
void launch_merge(unsigned long long *dfast, ulonglong2 *drec, Scalars *dscal,
                  const unsigned long long *sfast, const ulonglong2 *srec, const Scalars *sscal,
                  size_t npix, SlotMap slots, cudaStream_t s);
// deterministic merge with other Runtimes by direct loads: counts add, the record with the greatest (z, ~job) wins
struct PeerList { const unsigned long long *fast[16]; const ulonglong2 *rec[16]; const Scalars *scal[16]; const uint32_t *cnt[16]; int n; };
void launch_merge_peers(unsigned long long *dfast, ulonglong2 *drec, Scalars *dscal, const PeerList &peers,
                        size_t pix0, size_t npix, SlotMap slots, cudaStream_t s);
void launch_seed_points(unsigned long long seed, unsigned long long first, unsigned long long n, double *out, cudaStream_t s);
// cross-GPU frame protocol: waits and signals are prologues / epilogues of these kernels
struct FrameSync { Scalars *scal[SYNC_MAX_RANKS]; int n_ranks, my_rank; unsigned int epoch; };   // every rank's scalars, by rank
void launch_frame_reset(unsigned long long *fast, ulonglong2 *rec, Scalars *scal, size_t npix, size_t nslots,
                        int n_ranks, unsigned int epoch, cudaStream_t s);
void launch_frame_export(const unsigned long long *fast, Scalars *scal, uint32_t *cnt, size_t npix, SlotMap slots,
                         const FrameSync &S, cudaStream_t s);
void launch_frame_merge(unsigned long long *dfast, ulonglong2 *drec, uint32_t *dcnt, Scalars *dscal, const PeerList &peers,
                        size_t pix0, size_t npix, SlotMap slots, const FrameSync &S, cudaStream_t s);
void launch_frame_colorize(const ColorParams &cp, const uint32_t *cnt, const ulonglong2 *rec, Scalars *scal, uint16_t *rgba_u16,
                           const FrameSync &S, int owner, cudaStream_t s);
void launch_signal(const FrameSync &S, int kind, cudaStream_t s);
void launch_wait(Scalars *mine, int kind, int n_ranks, unsigned int epoch, cudaStream_t s);
void set_sync_timeout_ms(long long ms);
unsigned long long launch_count();
bool set_mode(int mode);     // SAR_DIAGNOSTICS builds only: 0 = product path; 1, 2, 4 = roofline experiments (incomplete results)
#ifdef SAR_DIAGNOSTICS
void set_diag_hot(int per_1024);
int get_diag_hot();
#endif
bool set_tile_scatter(int on);      // 1 (default): images that fit a shared-memory tile use per-block private histograms
bool set_pipeline(int on);          // tuning: depth test one iteration behind its atomic (0/1); never changes results
bool set_traj_per_thread(int nt);   // tuning: trajectories carried per thread (1, 2 or 4); never changes results

}  // namespace sar
