// sar_kernels.cu — hand-written sm_100a kernels for the iterate → project → scatter →
// tone-map/colourise path of Icelk/strange-attractor-renderer (src/lib.rs:747-904).
//
// Arithmetic contract: every f64 operation on the trajectory path is an explicit
// round-to-nearest __dmul_rn/__dadd_rn/__dsub_rn (never contracted to FMA, whatever the
// compiler flags), in the reference's exact association order, so that the chaotic map
// reproduces the CPU trajectory bit for bit.
#include "sar_device.cuh"

#include <atomic>
#include <math_constants.h>
#include <type_traits>

#ifndef SAR_DEFAULT_NT
#define SAR_DEFAULT_NT 1
#endif
#ifndef SAR_DEFAULT_PIPE
#define SAR_DEFAULT_PIPE 0
#endif

namespace sar {

static std::atomic<unsigned long long> g_launches{0};
unsigned long long launch_count() { return g_launches.load(); }
void bump_launches(unsigned int n) { g_launches += n; }

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t zkey_of(float z) { return zkey_from_bits(__float_as_uint(z)); }

// 128-bit compare-and-swap (ATOMG.E.CAS.128): returns the previous value.
__device__ __forceinline__ ulonglong2 cas128(ulonglong2 *addr, ulonglong2 expect, ulonglong2 desired)
{
    ulonglong2 old;
    asm volatile(
        "{\n\t"
        ".reg .b128 c, n, o;\n\t"
        "mov.b128 c, {%2, %3};\n\t"
        "mov.b128 n, {%4, %5};\n\t"
        "atom.global.relaxed.gpu.cas.b128 o, [%6], c, n;\n\t"
        "mov.b128 {%0, %1}, o;\n\t"
        "}"
        : "=l"(old.x), "=l"(old.y)
        : "l"(expect.x), "l"(expect.y), "l"(desired.x), "l"(desired.y), "l"(addr)
        : "memory");
    return old;
}

// count += 1 and fetch of the packed (zhint, count) word, predicated; the result is UNDEFINED when
// !p (callers zero the candidate key of such lanes instead, so it is never looked at).  Written as
// one predicated instruction with a single definition of the result because
// `old = ~0; if (p) old = atomicAdd(..)` compiles to the atomic into a temporary plus a MOV that
// waits for the L2 round trip right behind it (ncu: 52 % of all stall samples sat on that MOV),
// which serialises the lanes a thread carries and defeats the deferred depth test.
__device__ __forceinline__ unsigned long long atom_inc_if(bool p, unsigned long long *addr)
{
    unsigned long long old;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.u32 p, %2, 0;\n\t"
        "@p atom.global.add.u64 %0, [%1], 1;\n\t"
        "}"
        : "=l"(old)
        : "l"(addr), "r"((unsigned int)p)
        : "memory");
    return old;
}

__device__ __forceinline__ unsigned long long splitmix64_at(unsigned long long seed, unsigned long long n)
{
    unsigned long long z = seed + (n + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// stand-in for `rng.random::<f64>() * 0.1` (lib.rs:748): 53-bit uniform in [0,1) times 0.1
__device__ __forceinline__ double seed_coord(unsigned long long seed, unsigned long long n)
{
    const unsigned long long u = splitmix64_at(seed, n);
    return __dmul_rn(__dmul_rn((double)(u >> 11), 0x1.0p-53), 0.1);
}

// PolynomialSprott2Degree::next_point (lib.rs:585-620).  One coordinate: the serial
// left-to-right sum of lib.rs:588-600; c[0] already holds 0.0 + 1.0*c0.
__device__ __forceinline__ double sprott_sum(const double (&c)[10], double x, double y, double z,
                                             double xx, double xy, double xz, double yy, double yz, double zz)
{
    double s = c[0];
    s = __dadd_rn(s, __dmul_rn(x, c[1]));
    s = __dadd_rn(s, __dmul_rn(xx, c[2]));
    s = __dadd_rn(s, __dmul_rn(xy, c[3]));
    s = __dadd_rn(s, __dmul_rn(xz, c[4]));
    s = __dadd_rn(s, __dmul_rn(y, c[5]));
    s = __dadd_rn(s, __dmul_rn(yy, c[6]));
    s = __dadd_rn(s, __dmul_rn(yz, c[7]));
    s = __dadd_rn(s, __dmul_rn(z, c[8]));
    s = __dadd_rn(s, __dmul_rn(zz, c[9]));
    return s;
}
// The ten cubic monomials of AK 1 continue the same serial sum (see next_point<AK>).
__device__ __forceinline__ double cubic_tail(double s, const double (&c)[10], double xxx, double xxy, double xxz, double xyy,
                                             double xyz, double xzz, double yyy, double yyz, double yzz, double zzz)
{
    s = __dadd_rn(s, __dmul_rn(xxx, c[0]));
    s = __dadd_rn(s, __dmul_rn(xxy, c[1]));
    s = __dadd_rn(s, __dmul_rn(xxz, c[2]));
    s = __dadd_rn(s, __dmul_rn(xyy, c[3]));
    s = __dadd_rn(s, __dmul_rn(xyz, c[4]));
    s = __dadd_rn(s, __dmul_rn(xzz, c[5]));
    s = __dadd_rn(s, __dmul_rn(yyy, c[6]));
    s = __dadd_rn(s, __dmul_rn(yyz, c[7]));
    s = __dadd_rn(s, __dmul_rn(yzz, c[8]));
    s = __dadd_rn(s, __dmul_rn(zzz, c[9]));
    return s;
}
// Attractor::next_point (lib.rs:71-77) for the attractor kinds behind sar_config.attractor_kind:
//   AK 0  PolynomialSprott2Degree (lib.rs:575-620): 3 x 10 coefficients over [1,x,x²,xy,xz,y,y²,yz,z,z²]
//   AK 1  PolynomialSprott3Degree (the cubic member of the same family, README.md:8 "Adding more should
//         be relatively easy"): the ten cubic monomials [x³,x²y,x²z,xy²,xyz,xz²,y³,y²z,yz²,z³] — each
//         formed as (quadratic monomial)·(variable) — appended to the same left-to-right sum.
template <int AK>
__device__ __forceinline__ void next_point(const IterParams &P, double x, double y, double z, double &nx, double &ny, double &nz)
{
    const double xx = __dmul_rn(x, x), xy = __dmul_rn(x, y), xz = __dmul_rn(x, z);
    const double yy = __dmul_rn(y, y), yz = __dmul_rn(y, z), zz = __dmul_rn(z, z);
    nx = sprott_sum(P.c[0], x, y, z, xx, xy, xz, yy, yz, zz);
    ny = sprott_sum(P.c[1], x, y, z, xx, xy, xz, yy, yz, zz);
    nz = sprott_sum(P.c[2], x, y, z, xx, xy, xz, yy, yz, zz);
    if (AK == 1) {
        const double xxx = __dmul_rn(xx, x), xxy = __dmul_rn(xx, y), xxz = __dmul_rn(xx, z), xyy = __dmul_rn(xy, y), xyz = __dmul_rn(xy, z);
        const double xzz = __dmul_rn(xz, z), yyy = __dmul_rn(yy, y), yyz = __dmul_rn(yy, z), yzz = __dmul_rn(yz, z), zzz = __dmul_rn(zz, z);
        nx = cubic_tail(nx, P.c3[0], xxx, xxy, xxz, xyy, xyz, xzz, yyy, yyz, yzz, zzz);
        ny = cubic_tail(ny, P.c3[1], xxx, xxy, xxz, xyy, xyz, xzz, yyy, yyz, yzz, zzz);
        nz = cubic_tail(nz, P.c3[2], xxx, xxy, xxz, xyy, xyz, xzz, yyy, yyz, yzz, zzz);
    }
}
// for the kernels off the hot path (warm-up, auto-framing): one uniform branch per step
__device__ __forceinline__ void next_point_any(const IterParams &P, double x, double y, double z, double &nx, double &ny, double &nz)
{
    if (P.attractor_kind == 1u) next_point<1>(P, x, y, z, nx, ny, nz);
    else next_point<0>(P, x, y, z, nx, ny, nz);
}

// Vec3::magnitude (lib.rs:129-131)
__device__ __forceinline__ double magnitude(double x, double y, double z)
{
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}

// ColorTransform::transform(delta, screen_space, view) for the two shipped kinds (lib.rs:511-516,
// 520-558).  `x / 2.` is written `x * 0.5` (identical in IEEE arithmetic, one instruction).
__device__ __forceinline__ double transform_ds(const IterParams &P, double dx, double dy, double dz,
                                               double sx, double sy, double sz)
{
    const double mag = magnitude(dx, dy, dz);
    if (P.ct_kind == 1u) return __dmul_rn(__dadd_rn(mag, P.ct_offset), P.ct_factor);   // AdjustedVelocity, lib.rs:514
    if (P.ct_kind == 2u) {                                                               // ScreenBlend (include/sar.h): a ColorTransform
        double t = __dmul_rn(sx, P.ct_w[0]);                                             // closure (lib.rs:245) made of exact operations
        t = __dadd_rn(t, __dmul_rn(sy, P.ct_w[1]));
        t = __dadd_rn(t, __dmul_rn(sz, P.ct_w[2]));
        t = __dadd_rn(t, __dmul_rn(mag, P.ct_w[3]));
        return __dmul_rn(__dadd_rn(t, P.ct_offset), P.ct_factor);
    }
    // color_transforms::poisson_saturne, lib.rs:520-558 (COS/SIN literals lib.rs:529-536)
    const double COS = 0.7009092642998508981833083453238941729068756103515625;
    const double SIN = 0.7132504491541815649924274111981503665447235107421875;
    const double x2 = __dadd_rn(__dmul_rn(__dadd_rn(sx, P.ccx), COS), __dmul_rn(__dadd_rn(sz, P.ccy), SIN));
    const bool out = (x2 < -0.0839) ||
                     (__dadd_rn(__dmul_rn(10.55, x2), sy) < (0.46 - 1.0941)) ||
                     (__dadd_rn(__dmul_rn(1.0426, x2), sy) < (0.179 - 0.1576)) ||
                     (__dsub_rn(__dmul_rn(0.5139, x2), sy) > (-0.04 - 0.04092));
    const double part = out ? 0. : 1.;
    const double color = __dmul_rn(__dadd_rn(part, mag), 0.5);               // `/ 2.`, lib.rs:556
    return __ddiv_rn(__dsub_rn(color, 0.1), 0.9);                            // lib.rs:557
}
// The winning branch of the depth test (lib.rs:821-833), off the hot loop.
//   value = color_transform.transform(delta, screen_space, view)   lib.rs:826-828
//   steps[idx] = value; zbuf[idx] = z2 as f32                       lib.rs:830-832
// made atomic and order-independent: the record is replaced iff (zkey, ~job) is strictly
// greater than the stored one.  The candidate comes straight from the hot loop's registers
// (delta = current - previous point, lib.rs:822; screen_space, lib.rs:773).  `key` is the
// canonical key (-0.0 already folded onto +0.0, which is how f32 `>` sees them); the record itself
// keeps zkey(-0.0) so that zbuf reads back -0.0 exactly where the reference stores it, and records
// are ordered with rec_order(), which folds the two zeros again.
__device__ __forceinline__ void record_win(const IterParams &P, unsigned int idx, uint32_t key, uint32_t job_inv,
                                           unsigned long long old, double dx, double dy, double dz,
                                           double sx, double sy, double sz, bool raise_hint = true)
{
    unsigned long long hi = ((unsigned long long)key << 32) | job_inv;
    if (key == ZKEY_ZERO) {
        // z2 as f32 is +0.0 or -0.0; which one is not kept by the hot loop: redo lib.rs:778-779 from screen_space
        const double z2 = __dsub_rn(__dmul_rn(__dadd_rn(sx, P.ccx), P.sv), __dmul_rn(__dadd_rn(sz, P.ccy), P.cv));
        if (__float_as_uint(__double2float_rn(z2)) == 0x80000000u) hi = ((unsigned long long)ZKEY_NEG_ZERO << 32) | job_inv;
    }
    ulonglong2 *r = P.rec + idx;
    // Current record: known without a load if the pixel was untouched when our atomic hit it,
    // otherwise loaded now (may be stale or torn — the CAS validates it) so that the round trip
    // overlaps the arithmetic below.
    ulonglong2 cur = make_ulonglong2(0ull, REC_HI_RESET);
    if ((uint32_t)(old >> 32) != ZKEY_SENTINEL + 1u) {
        cur.y = __ldcg(&r->y);
        cur.x = __ldcg(&r->x);
    }
    // Best-effort raise of the hint so later candidates below this z skip this path: one
    // compare-and-swap against the value our own atomic produced, result ignored.  If other hits
    // landed in between it fails and the hint stays low, which only costs a later visit here.
    if (raise_hint) {
        const unsigned long long expect = old + 1ull;
        const unsigned long long want = ((unsigned long long)key << 32) | (expect & 0xFFFFFFFFull);
        if ((uint32_t)(expect >> 32) < key) (void)atomicCAS(P.fast + slot_of(idx, P.slots), expect, want);
    }
    const double value = transform_ds(P, dx, dy, dz, sx, sy, sz);
    const ulonglong2 want = make_ulonglong2((unsigned long long)__double_as_longlong(value), hi);
    while (rec_order(want.y) > rec_order(cur.y)) {
        const ulonglong2 prev = cas128(r, cur, want);
        if (prev.x == cur.x && prev.y == cur.y) break;
        cur = prev;
    }
}

__device__ __noinline__ void record_win_call(const IterParams *Pp, unsigned int idx, uint32_t key, uint32_t job_inv,
                                             unsigned long long old, double dx, double dy, double dz,
                                             double sx, double sy, double sz, bool raise_hint)
{
    record_win(*Pp, idx, key, job_inv, old, dx, dy, dz, sx, sy, sz, raise_hint);
}

// Everything that is not a plain in-view hit: out of view, non-finite coordinates, and the
// corner pixel.  Evaluates lib.rs:789-802 literally.  Returns action << 32 | idx with action
// 0 = not recorded (continue), 1 = record at idx, 2 = the state is NaN: this and every later iteration of the job hits
// count[(0,0)] and can never win the depth test (NaN is absorbing).
__device__ __noinline__ unsigned long long classify_rare(const IterParams *Pp, double sx, double sy, double sz,
                                                         double nx, double ny, double nz)
{
    const IterParams &P = *Pp;
    // the pixel coordinates again (lib.rs:776-786), same instructions on the same inputs as the hot loop
    const double a = __dadd_rn(sx, P.ccx);
    const double b = __dadd_rn(sz, P.ccy);
    const double x2 = __dadd_rn(__dmul_rn(a, P.cv), __dmul_rn(b, P.sv));
    const double fi = __dmul_rn(__dsub_rn(P.sam, x2), P.ws);
    const double fj = __dsub_rn(P.half_h, __dmul_rn(__dadd_rn(sy, P.ccz), P.ws));
    const double w = (double)P.W, h = (double)P.H;
    if (fi >= w || fj >= h || fi < 0. || fj < 0.) return 0ull;  // lib.rs:789; NaN passes every test
    if (nx != nx || ny != ny || nz != nz) return 2ull << 32;    // all of screen_space is NaN -> i = j = 0
    const unsigned int i = (fi != fi) ? 0u : __double2uint_rz(fi);   // `as u32`: truncates, NaN -> 0 (lib.rs:800-802)
    const unsigned int j = (fj != fj) ? 0u : __double2uint_rz(fj);
    return (1ull << 32) | (unsigned long long)(j * P.W + i);
}

// ---------------------------------------------------------------------------------------------
// iterate → project → scatter: render() (lib.rs:747-838).
// One LANE = one trajectory = one reference render() call (a job): start point, 1000 warm-up steps
// (lib.rs:750-752), `iterations` recorded steps.  A THREAD carries NT lanes at once (jobs tid,
// tid + nthreads, ...): their arithmetic is written branch-free and side by side so that the
// compiler interleaves the NT dependency chains and every constant fetched from the parameter
// bank (47 f64: coefficients, rotation matrix, projection scalars — they do not fit the uniform
// register file, so each iteration re-fetches them) feeds NT multiplications instead of one.
//
// MODE (only in SAR_DIAGNOSTICS builds; the product library holds MODE 0 alone): 1 = arithmetic
// only (no memory traffic); 2 = count with a fire-and-forget reduction, no depth test; 4 = the
// product's atomic with the win path removed; 5 = cost model of a per-SM shared-memory table for
// hot pixels (a pseudo-random fraction of the hits pays a tag load + two shared-memory atomics
// instead of the L2 atomic).  Modes != 0 leave the Runtime in a state that is only good for timing
// (tools/sweep_iterate.py, profiles/r2_iterate_variants.md).
// ---------------------------------------------------------------------------------------------
constexpr unsigned int IDX_RARE = 0xFFFFFFFFu;   // candidate that needs classify_rare (W*H <= 2^31, so never a pixel)

// A recorded hit between its count atomic and its depth test: what the winning branch needs.
struct Cand {
    double dx, dy, dz;            // delta = current - previous point (lib.rs:822)
    double sx, sy, sz;            // screen_space (lib.rs:773)
    unsigned long long old;       // what the atomic returned: zhint << 32 | count; ~0 = no hit
    unsigned int idx;
    uint32_t key;                 // canonical z key; 0 = cannot win
};

// PIPE = 1: the depth test of iteration i is made after the arithmetic of iteration i+1, so the L2
// round trip of the atomic is covered by ~130 FP64 instructions of the same warp instead of by
// other warps alone — with NT lanes per thread there are only 3-4 warps per scheduler.  Within a
// lane tests still retire in iteration order and before the next atomic is issued, so exact z ties
// keep resolving to the earlier iteration.
//
// TILE = 1 — the north star's "per-block count tiles in shared memory before a global atomicAdd reduction", for
// images that FIT a tile (W*H*8 bytes of shared memory, i.e. up to 25 600 pixels: thumbnails and previews; one block of
// 896 lanes per SM).  Every block keeps a private (count u32, z max u32) pair per pixel in shared memory, seeded with
// the pixel's global depth hint; a hit is two shared-memory atomics (ATOMS.ADD + ATOMS.MAX) instead of one L2 atomic
// — on such small images the L2 path serialises on a few thousand addresses (64x64: 30 G it/s) —, a hit that reaches
// the block's running max is a depth-test candidate and goes through the same exact record path (128-bit CAS on
// (zkey, ~job)), and at the end the block adds its counts and max-es its hints into the global arrays.  Exact for the
// same reasons as the L2 path: counts commute; a hit that is the global winner of its pixel is >= every earlier hit
// of its own block, so it always reaches the record path, which alone decides.
template <int NT, int MODE, int PIPE, int AK, int TILE>
__global__ void __launch_bounds__(TILE ? TILE_BLOCK : 128 / NT, TILE ? 1 : ((NT == 1 && AK == 0) ? (PIPE ? 5 : 7) : (NT == 1 ? 4 : 8)))   // register budgets: 7 x 128 / 5 x 128 / 8 x 64 / 8 x 32 threads per SM
iterate_kernel(const __grid_constant__ IterParams P)
{
    extern __shared__ unsigned int s_tile[];                                  // TILE: count[npix] | zmax[npix]
    const unsigned int npix = P.W * P.H;
    if (TILE) {
        for (unsigned int p = threadIdx.x; p < npix; p += blockDim.x) {
            s_tile[p] = 0u;
            s_tile[npix + p] = (uint32_t)(P.fast[slot_of(p, P.slots)] >> 32);   // the pixel's current hint (canonical key)
        }
        __syncthreads();
    }
    const unsigned long long nthreads = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t cmax = 0u;                                                       // greatest count this thread's hits produced (lib.rs:813-815)
    for (unsigned long long job0 = tid; job0 < P.n_jobs; job0 += nthreads * NT) {
        double x[NT], y[NT], z[NT];
        uint32_t job_inv[NT];
        bool live[NT];
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            const unsigned long long job = job0 + (unsigned long long)k * nthreads;
            live[k] = job < P.n_jobs;
            x[k] = y[k] = z[k] = 0.;
            if (live[k]) {
                if (P.init != nullptr) {
                    x[k] = P.init[3 * job + 0]; y[k] = P.init[3 * job + 1]; z[k] = P.init[3 * job + 2];
                } else {
                    const unsigned long long g = 3ull * (P.first_job + job);
                    x[k] = seed_coord(P.seed, g); y[k] = seed_coord(P.seed, g + 1); z[k] = seed_coord(P.seed, g + 2);
                }
            }
            // order key of this job: earlier jobs win z ties (see sar_device.cuh); the host
            // guarantees job_key0 + n_jobs <= 2^32
            job_inv[k] = 0xFFFFFFFFu - (P.job_key0 + (uint32_t)job);
        }
        for (unsigned int w = 0; w < P.warmup; ++w) {                         // lib.rs:750-752
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                double nx, ny, nz;
                next_point<AK>(P, x[k], y[k], z[k], nx, ny, nz);
                x[k] = nx; y[k] = ny; z[k] = nz;
            }
        }

        // One iteration in three pieces (lambdas, all inlined): arith (branch-free arithmetic of the NT
        // lanes -> candidate set), scatter (the count atomic), test (depth test on what the atomic
        // returned).  PIPE 0: arith(A) scatter(A) test(A).  PIPE 1 ping-pongs two candidate sets:
        //     scatter(A) | arith(B) test(A) | scatter(B) | arith(A) test(B) | ...
        // so the L2 round trip of an atomic is covered by the next iteration's arithmetic, each atomic
        // result is produced and consumed inside one loop trip (nothing is copied behind an atomic,
        // which would wait for it), and within a lane tests retire in iteration order.
        bool any = true;
        auto arith = [&](Cand (&c)[NT]) {
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                double nx, ny, nz;
                next_point<AK>(P, x[k], y[k], z[k], nx, ny, nz);               // lib.rs:770
                // screen_space = rotation_matrix.mul_right(current_point), lib.rs:773 / 208-215
                const double sx = __dadd_rn(__dadd_rn(__dmul_rn(P.m[0][0], nx), __dmul_rn(P.m[0][1], ny)), __dmul_rn(P.m[0][2], nz));
                const double sy = __dadd_rn(__dadd_rn(__dmul_rn(P.m[1][0], nx), __dmul_rn(P.m[1][1], ny)), __dmul_rn(P.m[1][2], nz));
                const double sz = __dadd_rn(__dadd_rn(__dmul_rn(P.m[2][0], nx), __dmul_rn(P.m[2][1], ny)), __dmul_rn(P.m[2][2], nz));
                // rotate around center_camera, lib.rs:776-779 (center_camera.y pairs with screen_space.z)
                const double a = __dadd_rn(sx, P.ccx);
                const double b = __dadd_rn(sz, P.ccy);
                const double x2 = __dadd_rn(__dmul_rn(a, P.cv), __dmul_rn(b, P.sv));
                const double z2 = __dsub_rn(__dmul_rn(a, P.sv), __dmul_rn(b, P.cv));
                const double fi = __dmul_rn(__dsub_rn(P.sam, x2), P.ws);                       // lib.rs:783
                const double fj = __dsub_rn(P.half_h, __dmul_rn(__dadd_rn(sy, P.ccz), P.ws));  // lib.rs:786
                // Bounds test + `as u32` (lib.rs:789-802) for the common case in one step: floor-convert
                // (saturating) and compare unsigned.  i in [0,W) <=> 0 <= floor(i) < W, and floor == trunc
                // there.  Anything else — out of view, NaN, pixel 0 — is marked IDX_RARE and goes through
                // classify_rare(), which applies the reference's comparisons literally.
                const unsigned int ii = (unsigned int)__double2int_rd(fi);
                const unsigned int jj = (unsigned int)__double2int_rd(fj);
                const unsigned int idx = jj * P.W + ii;
                c[k].idx = ((ii < P.W && jj < P.H) && idx != 0u) ? idx : IDX_RARE;
                const float zf = __double2float_rn(z2) + 0.0f;                // `z2 as f32`; -0 folded onto +0 (f32 `>` sees them equal)
                c[k].key = zkey_of(zf);
                if (c[k].key > ZKEY_POS_INF) c[k].key = 0u;                   // NaN never passes `>` (lib.rs:821)
                c[k].sx = sx; c[k].sy = sy; c[k].sz = sz;
                c[k].dx = __dsub_rn(nx, x[k]); c[k].dy = __dsub_rn(ny, y[k]); c[k].dz = __dsub_rn(nz, z[k]);   // delta, lib.rs:822
                x[k] = nx; y[k] = ny; z[k] = nz;                              // previous_point = current_point, lib.rs:793/836
            }
        };
        // count += 1 (lib.rs:811) + fetch of the depth hint: one L2 atomic per recorded lane.  Runs right
        // after arith() of the same candidate set, so (x,y,z) is still that iteration's current point.
        auto scatter = [&](Cand (&c)[NT], unsigned long long it) {
            any = false;
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                bool hit = live[k];
                if (live[k] && c[k].idx == IDX_RARE) {
                    const unsigned long long r = classify_rare(&P, c[k].sx, c[k].sy, c[k].sz, x[k], y[k], z[k]);
                    const unsigned int act = (unsigned int)(r >> 32);
                    c[k].idx = (unsigned int)r;
                    hit = act == 1u;
                    if (act == 2u) {                                          // NaN state: pay the whole debt at once, the job is over
                        atomicAdd(&P.scal->nan_sink, P.iterations - it);
                        live[k] = false;
                    }
                }
                unsigned long long *slot = P.fast + slot_of(c[k].idx, P.slots);
                if (TILE) {
                    if (!hit) { c[k].key = 0u; c[k].idx = IDX_RARE; c[k].old = ~0ull; }
                    else {
                        const unsigned int before = atomicAdd(&s_tile[c[k].idx], 1u);                 // count += 1, lib.rs:811
                        const unsigned int zbefore = atomicMax(&s_tile[npix + c[k].idx], c[k].key);   // the block's running z max
                        c[k].old = ((unsigned long long)zbefore << 32) | before;
                    }
                } else
#ifdef SAR_DIAGNOSTICS
                if (MODE == 5) {
                    // cost model of a per-SM shared-memory table for hot pixels: a pseudo-random P.diag_hot / 1024 of the
                    // hits pay a tag load + two shared-memory atomics instead of the L2 atomic (results are meaningless)
                    extern __shared__ unsigned int s_tab[];
                    const unsigned int h = c[k].idx * 0x9E3779B1u;
                    const unsigned int e = (h >> 8) % (unsigned int)(P.diag_tab_entries);
                    const unsigned int tag = *((volatile unsigned int *)&s_tab[3 * e]);
                    const bool hot = hit && ((h >> 22) < P.diag_hot) && tag != 0xDEADBEEFu;
                    if (!hit) { c[k].key = 0u; c[k].idx = IDX_RARE; }
                    if (hot) {
                        const unsigned int cnt = atomicAdd(&s_tab[3 * e + 1], 1u);
                        const unsigned int zm = atomicMax(&s_tab[3 * e + 2], c[k].key);
                        c[k].old = ((unsigned long long)(zm | 0x80000000u) << 32) | cnt;
                        c[k].key = 0u;
                    } else {
                        c[k].old = atom_inc_if(hit, slot);
                    }
                } else
#endif
                if (MODE == 0 || MODE == 4) {
                    if (!hit) { c[k].key = 0u; c[k].idx = IDX_RARE; }         // nothing recorded: nothing to test, `old` stays undefined
                    c[k].old = atom_inc_if(hit, slot);
                } else {
                    c[k].old = ~0ull;
                    if (hit) {
                        if (MODE == 2) asm volatile("red.global.add.u64 [%0], 1;" ::"l"(slot) : "memory");
                        else c[k].old = ~0ull ^ (unsigned long long)(c[k].idx == 0xFFFFFFFEu);
                    }
                }
                any |= live[k];
            }
        };
        // depth test (lib.rs:821) on the returned hints; the winning branch inline (in the loop) or as a call (loop exits)
        auto test = [&](Cand (&c)[NT], auto inl) {
#pragma unroll
            for (int k = 0; k < NT; ++k) {
                if ((MODE == 0 || MODE == 5) && !TILE && c[k].idx != IDX_RARE) {   // count after this hit; the running max of lib.rs:813-815
                    const uint32_t now = (uint32_t)c[k].old + 1u;
                    cmax = now > cmax ? now : cmax;
                }
                if (c[k].key >= (uint32_t)(c[k].old >> 32) && c[k].key != 0u) {      // may beat zbuf
                    if (MODE != 0 && MODE != 5) { if (c[k].idx == 0xFFFFFFF0u) P.scal->pad = c[k].key; }   // diagnostics: keep the returned value live
                    else if (decltype(inl)::value) record_win(P, c[k].idx, c[k].key, job_inv[k], c[k].old, c[k].dx, c[k].dy, c[k].dz, c[k].sx, c[k].sy, c[k].sz, !TILE);
                    else record_win_call(&P, c[k].idx, c[k].key, job_inv[k], c[k].old, c[k].dx, c[k].dy, c[k].dz, c[k].sx, c[k].sy, c[k].sz, !TILE);
                }
            }
        };
        using inl_t = std::integral_constant<bool, true>;
        using call_t = std::integral_constant<bool, false>;

        Cand A[NT];
        unsigned long long it = 0;
        if (PIPE) {
            if (P.iterations > 0) {
                Cand B[NT];
                arith(A);
                for (;;) {                                                    // lib.rs:769, two iterations per trip
                    scatter(A, it);
                    ++it;
                    if (!(it < P.iterations && any)) { test(A, call_t{}); break; }
                    arith(B);
                    test(A, inl_t{});
                    scatter(B, it);
                    ++it;
                    if (!(it < P.iterations && any)) { test(B, call_t{}); break; }
                    arith(A);
                    test(B, inl_t{});
                }
            }
        } else {
#pragma unroll 2
            for (; it < P.iterations && any; ++it) {                          // lib.rs:769
                arith(A); scatter(A, it); test(A, inl_t{});
            }
        }
    }
    if (TILE) {
        // the block's tile -> the global arrays: one 32-bit add on the count half and one 32-bit max on the hint half of
        // every pixel this block touched (the hint stays <= the recorded key: every hit that raised the block's max went
        // through record_win first)
        __syncthreads();
        for (unsigned int p = threadIdx.x; p < npix; p += blockDim.x) {
            const unsigned int c = s_tile[p];
            if (c) {
                unsigned int *w = reinterpret_cast<unsigned int *>(P.fast + slot_of(p, P.slots));
                atomicAdd(w, c);
                atomicMax(w + 1, s_tile[npix + p]);
            }
        }
        return;                                       // Runtime.max: the host runs the full reduction after a tile launch
    }
    // Runtime.max, kept current by the render itself: one reduction per warp (lanes leave the job loop together,
    // except NaN jobs, hence the active mask), not 132 608 same-address atomics
    {
        const unsigned int m = __activemask();
        const uint32_t wmax = __reduce_max_sync(m, cmax);
        if (wmax && (threadIdx.x & 31u) == (unsigned int)(__ffs(m) - 1)) atomicMax(&P.scal->max, wmax);
    }
}

// The warm-up of render() on its own (lib.rs:748-752): used when a frame sweep shares one list of
// start points, so that the 1000 unrecorded steps run once instead of once per frame.
__global__ void __launch_bounds__(128)
warm_kernel(const __grid_constant__ IterParams P, double *__restrict__ out)
{
    const unsigned long long job = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (job >= P.n_jobs) return;
    double x, y, z;
    if (P.init != nullptr) {
        x = P.init[3 * job + 0]; y = P.init[3 * job + 1]; z = P.init[3 * job + 2];
    } else {
        const unsigned long long g = 3ull * (P.first_job + job);
        x = seed_coord(P.seed, g); y = seed_coord(P.seed, g + 1); z = seed_coord(P.seed, g + 2);
    }
    for (unsigned int w = 0; w < P.warmup; ++w) {
        double nx, ny, nz;
        next_point_any(P, x, y, z, nx, ny, nz);
        x = nx; y = ny; z = nz;
    }
    out[3 * job + 0] = x; out[3 * job + 1] = y; out[3 * job + 2] = z;
}
void launch_warm(const IterParams &p, double *out, cudaStream_t s)
{
    if (p.n_jobs == 0) return;
    warm_kernel<<<(unsigned int)((p.n_jobs + 127) / 128), 128, 0, s>>>(p, out);
    ++g_launches;
}

// Lanes per thread (NT): a tuning knob that never changes results.
// ---------------------------------------------------------------------------------------------
// Auto-framing first pass — the reference author's TODO at lib.rs:326-334: "Add option to make
// first-pass to get these values [the attractor's extent in screen space], to then compute
// center_camera".  One lane per trajectory: start point, warm-up, then `iterations` steps whose
// screen_space = R·p (lib.rs:773) is folded into a per-lane bounding box.  A trajectory that
// diverged (non-finite state — absorbing for this map) is left out and counted.
// The fold is exact (min/max only), so the box equals the oracle's bit for bit.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long dkey_of(double v)      // order-preserving f64 -> u64
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__global__ void __launch_bounds__(128)
bbox_kernel(const __grid_constant__ IterParams P, BBoxAccum *acc)
{
    const unsigned long long job = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    double lo[3] = {CUDART_INF, CUDART_INF, CUDART_INF}, hi[3] = {-CUDART_INF, -CUDART_INF, -CUDART_INF};
    bool ok = false;
    if (job < P.n_jobs) {
        double x, y, z;
        if (P.init != nullptr) {
            x = P.init[3 * job + 0]; y = P.init[3 * job + 1]; z = P.init[3 * job + 2];
        } else {
            const unsigned long long g = 3ull * (P.first_job + job);
            x = seed_coord(P.seed, g); y = seed_coord(P.seed, g + 1); z = seed_coord(P.seed, g + 2);
        }
        for (unsigned int w = 0; w < P.warmup; ++w) {                         // lib.rs:750-752
            double nx, ny, nz;
            next_point_any(P, x, y, z, nx, ny, nz);
            x = nx; y = ny; z = nz;
        }
        for (unsigned long long it = 0; it < P.iterations; ++it) {
            double nx, ny, nz;
            next_point_any(P, x, y, z, nx, ny, nz);
            x = nx; y = ny; z = nz;
            const double s[3] = {
                __dadd_rn(__dadd_rn(__dmul_rn(P.m[0][0], nx), __dmul_rn(P.m[0][1], ny)), __dmul_rn(P.m[0][2], nz)),
                __dadd_rn(__dadd_rn(__dmul_rn(P.m[1][0], nx), __dmul_rn(P.m[1][1], ny)), __dmul_rn(P.m[1][2], nz)),
                __dadd_rn(__dadd_rn(__dmul_rn(P.m[2][0], nx), __dmul_rn(P.m[2][1], ny)), __dmul_rn(P.m[2][2], nz))};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                if (s[c] < lo[c]) lo[c] = s[c];
                if (s[c] > hi[c]) hi[c] = s[c];
            }
        }
        ok = isfinite(x) && isfinite(y) && isfinite(z) && isfinite(lo[0]) && isfinite(hi[0]) && isfinite(lo[1]) && isfinite(hi[1])
             && isfinite(lo[2]) && isfinite(hi[2]);
        if (!ok) atomicAdd(&acc->diverged, 1ull);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        unsigned long long kl = ok ? dkey_of(lo[c]) : ~0ull, kh = ok ? dkey_of(hi[c]) : 0ull;
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long a = __shfl_xor_sync(0xffffffffu, kl, o), b = __shfl_xor_sync(0xffffffffu, kh, o);
            kl = a < kl ? a : kl; kh = b > kh ? b : kh;
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(&acc->lo[c], kl); atomicMax(&acc->hi[c], kh); }
    }
}
void launch_bbox(const IterParams &p, BBoxAccum *acc, cudaStream_t s)
{
    if (p.n_jobs == 0) return;
    bbox_kernel<<<(unsigned int)((p.n_jobs + 127) / 128), 128, 0, s>>>(p, acc);
    ++g_launches;
}

static std::atomic<int> g_nt{SAR_DEFAULT_NT};
bool set_traj_per_thread(int nt)
{
    if (nt != 1 && nt != 2 && nt != 4) return false;
    g_nt = nt;
    return true;
}
static std::atomic<int> g_pipe{SAR_DEFAULT_PIPE};
bool set_pipeline(int on)
{
    if (on != 0 && on != 1) return false;
    g_pipe = on;
    return true;
}
#ifdef SAR_DIAGNOSTICS
static std::atomic<int> g_diag_hot{0};
void set_diag_hot(int per_1024) { g_diag_hot = per_1024; }
int get_diag_hot() { return g_diag_hot.load(); }
static std::atomic<int> g_mode{0};
bool set_mode(int m)
{
    if (m != 0 && m != 1 && m != 2 && m != 4 && m != 5) return false;
    g_mode = m;
    return true;
}
#else
bool set_mode(int m) { return m == 0; }
#endif

static std::atomic<int> g_tile{1};
bool set_tile_scatter(int on)
{
    if (on != 0 && on != 1) return false;
    g_tile = on;
    return true;
}
// a tile launch needs the whole image in one block's shared memory and at least one full block of lanes
static bool tile_fits(const IterParams &p, unsigned long long want)
{
    return g_tile.load() && (size_t)p.W * p.H * 8 <= TILE_SMEM_MAX && want >= TILE_BLOCK;
}

template <int MODE>
static bool launch_iterate_mode(const IterParams &p, unsigned long long want, cudaStream_t s)
{
#ifdef SAR_DIAGNOSTICS
    const size_t smem = MODE == 5 ? (size_t)p.diag_tab_entries * 12 : 0;
#else
    const size_t smem = 0;
#endif
    if (MODE == 0 && tile_fits(p, want)) {       // the image fits a shared-memory tile: per-block private histograms
        const size_t tsm = (size_t)p.W * p.H * 8;
        const unsigned int grid = (unsigned int)((want + TILE_BLOCK - 1) / TILE_BLOCK);
        if (p.attractor_kind == 1u) {
            cudaFuncSetAttribute(iterate_kernel<1, 0, 0, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM_MAX);
            iterate_kernel<1, 0, 0, 1, 1><<<grid, TILE_BLOCK, tsm, s>>>(p);
        } else {
            cudaFuncSetAttribute(iterate_kernel<1, 0, 0, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_SMEM_MAX);
            iterate_kernel<1, 0, 0, 0, 1><<<grid, TILE_BLOCK, tsm, s>>>(p);
        }
        return true;
    }
    // `want` lanes; NT lanes per thread; narrow blocks so that the grid stays a multiple of the SM
    // count at the default 896 lanes per SM (7 blocks per SM) and small launches spread over the SMs
    const int nt = g_nt.load();
    const unsigned long long threads = (want + nt - 1) / nt;
    const unsigned int block = threads >= 148ull * 128ull ? 128u / (unsigned int)nt : 32u;
    const unsigned int grid = (unsigned int)((threads + block - 1) / block);
    if (p.attractor_kind == 1u) {            // the cubic family: one instantiation (the knobs above are tuned for AK 0)
        const unsigned int blk = want >= 148ull * 128ull ? 128u : 32u;
        iterate_kernel<1, MODE, 0, 1, 0><<<(unsigned int)((want + blk - 1) / blk), blk, smem, s>>>(p);
        return false;
    }
    if (g_pipe.load()) {
        switch (nt) {
        case 1: iterate_kernel<1, MODE, 1, 0, 0><<<grid, block, smem, s>>>(p); break;
        case 2: iterate_kernel<2, MODE, 1, 0, 0><<<grid, block, smem, s>>>(p); break;
        default: iterate_kernel<4, MODE, 1, 0, 0><<<grid, block, smem, s>>>(p); break;
        }
    } else {
        switch (nt) {
        case 1: iterate_kernel<1, MODE, 0, 0, 0><<<grid, block, smem, s>>>(p); break;
        case 2: iterate_kernel<2, MODE, 0, 0, 0><<<grid, block, smem, s>>>(p); break;
        default: iterate_kernel<4, MODE, 0, 0, 0><<<grid, block, smem, s>>>(p); break;
        }
    }
    return false;
}

// returns true when the launch used the shared-memory tile path (Runtime.max is then not tracked by the kernel)
bool launch_iterate(const IterParams &p, unsigned int lanes, cudaStream_t s)
{
    if (p.n_jobs == 0) return false;
    const unsigned long long want = p.n_jobs < lanes ? p.n_jobs : lanes;
    bool tile = false;
#ifdef SAR_DIAGNOSTICS
    switch (g_mode.load()) {
    case 1: tile = launch_iterate_mode<1>(p, want, s); break;
    case 2: tile = launch_iterate_mode<2>(p, want, s); break;
    case 4: tile = launch_iterate_mode<4>(p, want, s); break;
    case 5: tile = launch_iterate_mode<5>(p, want, s); break;
    default: tile = launch_iterate_mode<0>(p, want, s); break;
    }
#else
    tile = launch_iterate_mode<0>(p, want, s);
#endif
    ++g_launches;
    return tile;
}

// ---------------------------------------------------------------------------------------------
// Runtime::reset (lib.rs:682-699)
// ---------------------------------------------------------------------------------------------
__global__ void reset_kernel(unsigned long long *fast, ulonglong2 *rec, Scalars *scal, size_t npix, size_t nslots)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nslots; i += stride) {
        fast[i] = FAST_RESET;                                   // count 0 (lib.rs:687); every slot, whatever pixel it belongs to
        if (i < npix) rec[i] = make_ulonglong2(0ull, REC_HI_RESET);   // steps 0.0 (lib.rs:690), zbuf -1.0 (lib.rs:693)
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scal->nan_sink = 0ull; scal->max = 0u;                  // lib.rs:694
        scal->zmax_key = ZKEY_ZERO; scal->zmin_key = ZKEY_FLT_MAX; scal->pad = 0u;
    }
}
void launch_reset(unsigned long long *fast, ulonglong2 *rec, Scalars *scal, size_t npix, size_t nslots, cudaStream_t s)
{
    const unsigned int block = 256;
    size_t g = (nslots + block - 1) / block;
    const unsigned int grid = (unsigned int)(g > 148u * 16u ? 148u * 16u : (g ? g : 1));
    reset_kernel<<<grid, block, 0, s>>>(fast, rec, scal, npix, nslots);
    ++g_launches;
}

// ---------------------------------------------------------------------------------------------
// Runtime.max (lib.rs:643, 813-815) as a reduction: counts only grow, so the running max the
// reference tracks equals the max over the final counts.  Also folds the Depth min/max
// (lib.rs:877-882).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pixel_count(const unsigned long long *fast, const Scalars *scal, size_t i, SlotMap slots)
{
    uint32_t c = (uint32_t)fast[slot_of((uint32_t)i, slots)];
    if (i == 0) c += (uint32_t)scal->nan_sink;                  // wrapping u32, like lib.rs:811 in release
    return c;
}
__global__ void max_kernel(const unsigned long long *fast, const ulonglong2 *rec, Scalars *scal, size_t pix0, size_t npix, SlotMap slots)
{
    uint32_t m = 0, zmx = ZKEY_ZERO, zmn = ZKEY_FLT_MAX;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        const size_t p = pix0 + i;
        const uint32_t c = pixel_count(fast, scal, p, slots);
        m = c > m ? c : m;
        const uint32_t k = canon_key((uint32_t)(rec[p].y >> 32));
        if (k != ZKEY_SENTINEL) { zmx = k > zmx ? k : zmx; zmn = k < zmn ? k : zmn; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t m2 = __shfl_xor_sync(0xffffffffu, m, o);
        const uint32_t a2 = __shfl_xor_sync(0xffffffffu, zmx, o);
        const uint32_t b2 = __shfl_xor_sync(0xffffffffu, zmn, o);
        m = m2 > m ? m2 : m; zmx = a2 > zmx ? a2 : zmx; zmn = b2 < zmn ? b2 : zmn;
    }
    if ((threadIdx.x & 31) == 0) {
        if (m) atomicMax(&scal->max, m);
        atomicMax(&scal->zmax_key, zmx);
        atomicMin(&scal->zmin_key, zmn);
    }
}
// When the iterate kernel has kept scal->max current (every count change since the last reset came from it), all
// that is left of the reduction is the NaN debt owed to pixel (0,0): one thread.
__global__ void fold_max_kernel(const unsigned long long *fast, Scalars *scal, SlotMap slots)
{
    const uint32_t c0 = pixel_count(fast, scal, 0, slots);
    if (c0 > scal->max) scal->max = c0;
}
void launch_fold_max(const unsigned long long *fast, Scalars *scal, SlotMap slots, cudaStream_t s)
{
    fold_max_kernel<<<1, 1, 0, s>>>(fast, scal, slots);
    ++g_launches;
}
void launch_max(const unsigned long long *fast, const ulonglong2 *rec, Scalars *scal, size_t pix0, size_t npix, SlotMap slots, cudaStream_t s)
{
    if (npix == 0) return;
    const unsigned int block = 256;
    size_t g = (npix + block - 1) / block;
    const unsigned int grid = (unsigned int)(g > 148u * 8u ? 148u * 8u : g);
    max_kernel<<<grid, block, 0, s>>>(fast, rec, scal, pix0, npix, slots);
    ++g_launches;
}

// ---------------------------------------------------------------------------------------------
// Cross-GPU frame protocol without the host (one process per GPU).  Every rank keeps,
// inside its exported allocation, one flag per (kind, source rank).  A source rank announces an
// event by storing the frame epoch into that flag on every target (a remote store over NVLink,
// release at system scope); a target polls its OWN memory.  The waits and signals are the prologues
// and epilogues of the kernels that need them — reset, export, merge, colourise — not launches of
// their own:
//   prologue  every block polls the local flags it depends on (frame_wait);
//   epilogue  every thread fences its writes to system scope, the block counts itself in, and the
//             last block to do so stores the flags (frame_last_block).
// A wait that sees no progress for sync_timeout (default ~10 s) records sync_error; every later
// kernel of the protocol then does nothing, so a dead peer costs a timeout, never a GPU hang or a
// frame computed from half-delivered data.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

static std::atomic<long long> g_sync_timeout_cycles{20000000000ll};   // ~10 s at 1.965 GHz
void set_sync_timeout_ms(long long ms) { g_sync_timeout_cycles = ms * 1965000ll; }

// All threads of the block call this.  Threads [0, n) poll mine->flag[kind][first + t] until it
// reaches `epoch`.  Returns false (for the whole block) if the runtime already carries a sync error
// or a wait times out.
__device__ bool frame_wait(Scalars *mine, int kind, int first, int n, unsigned int epoch, long long timeout)
{
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = *((volatile unsigned int *)&mine->sync_error) != 0u;
    __syncthreads();
    if ((int)threadIdx.x < n && !s_bad) {
        const unsigned int *f = &mine->flag[kind][first + threadIdx.x];
        const long long t0 = clock64();
        while ((int)(ld_acquire_sys(f) - epoch) < 0) {              // epochs only grow; wrap-safe compare
            __nanosleep(64);
            if (clock64() - t0 > timeout) {                          // a peer is gone: give up, do not hang the GPU
                atomicCAS(&mine->sync_error, 0u, 1u + (unsigned int)kind);
                s_bad = 1;
                break;
            }
        }
    }
    __syncthreads();
    __threadfence_system();
    return !s_bad;
}
// All threads call this after their last write of the kernel.  True in exactly one block, and only
// after every block's writes are visible system-wide.  counter: one word per kernel kind, left at 0.
__device__ bool frame_last_block(unsigned int *counter)
{
    __shared__ int s_last;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(counter, 1u);
        s_last = t == gridDim.x - 1u;
        if (s_last) *counter = 0u;
    }
    __syncthreads();
    if (s_last) __threadfence_system();
    return s_last;
}
__device__ __forceinline__ void frame_signal(const FrameSync &S, int kind, bool to_all, int only_rank)
{
    if ((int)threadIdx.x < S.n_ranks && (to_all || (int)threadIdx.x == only_rank))
        st_release_sys(&S.scal[threadIdx.x]->flag[kind][S.my_rank], S.epoch);
}

// ---------------------------------------------------------------------------------------------
// colorize() (lib.rs:841-904)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t sat_u16(double v)          // Rust `as u16`: saturating, NaN -> 0
{
    if (!(v > 0.)) return 0;                                    // NaN, negative, zero
    const unsigned int u = __double2uint_rz(v);
    return (uint16_t)(u > 65535u ? 65535u : u);
}
__device__ __forceinline__ uint16_t sat_u16f(float v)
{
    if (!(v > 0.f)) return 0;
    const unsigned int u = __float2uint_rz(v);
    return (uint16_t)(u > 65535u ? 65535u : u);
}

// One pixel of colorize(): Gas (lib.rs:853-874) or Depth (lib.rs:875-900).
__device__ __forceinline__ void color_pixel(const ColorParams &C, const ulonglong2 r, uint32_t cnt, uint32_t max, double lnmax,
                                            float zmax, float zmin, ushort4 &px, float4 &fx)
{
    if (C.render_kind == 0u) {                              // RenderKind::Gas
        // Palette::interpolate(steps), lib.rs:442-472
        double v = __longlong_as_double((long long)r.x);
        if (v < 0.) v = 0.; else if (v >= 1.) v = 0.999999;
        v = __dmul_rn(v, C.pal_len);
        const double fl = floor(v);
        unsigned int n = (fl > 0.) ? __double2uint_rz(fl) : 0u;   // `as usize`: NaN -> 0
        if (n > C.palette_len - 1u) n = C.palette_len - 1u;
        const double t = fmod(v, 1.);                       // `value % 1.`, lib.rs:454
        const double t1 = __dsub_rn(1.0, t);
        const double cr = __dsqrt_rn(__dadd_rn(__dmul_rn(C.pal[n + 1][0], t), __dmul_rn(C.pal[n][0], t1)));
        const double cg = __dsqrt_rn(__dadd_rn(__dmul_rn(C.pal[n + 1][1], t), __dmul_rn(C.pal[n][1], t1)));
        const double cb = __dsqrt_rn(__dadd_rn(__dmul_rn(C.pal[n + 1][2], t), __dmul_rn(C.pal[n][2], t1)));
        const uint32_t c1 = cnt + 1u;
        const double lnc = c1 < C.lnlut_len ? __ldg(C.lnlut + c1) : (cnt == max ? lnmax : log((double)c1));
        const double factor = __ddiv_rn(lnc, lnmax);        // lib.rs:860
        const double vr = __dmul_rn(__dadd_rn(__dmul_rn(cr, factor), C.bright_offset), C.bright_factor);
        const double vg = __dmul_rn(__dadd_rn(__dmul_rn(cg, factor), C.bright_offset), C.bright_factor);
        const double vb = __dmul_rn(__dadd_rn(__dmul_rn(cb, factor), C.bright_offset), C.bright_factor);
        px.x = sat_u16(__dmul_rn(vr, 65535.));              // lib.rs:862-864
        px.y = sat_u16(__dmul_rn(vg, 65535.));
        px.z = sat_u16(__dmul_rn(vb, 65535.));
        px.w = C.transparent ? sat_u16(__dmul_rn(factor, 65535.)) : (uint16_t)65535u;  // lib.rs:865-869
        fx = make_float4((float)vr, (float)vg, (float)vb, C.transparent ? (float)factor : 1.0f);
    } else {                                                // RenderKind::Depth
        const uint32_t k = (uint32_t)(r.y >> 32);
        float zz;
        if (k == ZKEY_SENTINEL) zz = 0.0f;                  // z == -1.0, lib.rs:889-890
        else zz = __fdiv_rn(__fsub_rn(__uint_as_float(zbits_from_key(k)), zmin), __fsub_rn(zmax, zmin));
        const uint16_t g = sat_u16f(__fmul_rn(zz, 65535.0f));                          // lib.rs:895
        px.x = g; px.y = g; px.z = g; px.w = 65535u;
        fx = make_float4(zz, zz, zz, 1.0f);
    }
}
// ln(max + 1): from the host when it has read max back (exact: the platform libm the reference's f64::ln
// resolves to), else from the host-built table, else (max + 1 >= 2^20 on a device-resident path) the device log
__device__ __forceinline__ double ln_max1(const ColorParams &C, uint32_t max)
{
    const uint32_t m1 = max + 1u;                           // f64::from(runtime.max + 1), lib.rs:860
    return C.host_lnmax_valid ? C.ln_max1_host : (m1 < C.lnlut_len ? C.lnlut[m1] : log((double)m1));
}

__global__ void __launch_bounds__(256)
colorize_kernel(const __grid_constant__ ColorParams C, const unsigned long long *__restrict__ fast,
                const ulonglong2 *__restrict__ rec, const Scalars *__restrict__ scal,
                uint16_t *__restrict__ out16, float *__restrict__ out32)
{
    __shared__ double s_lnmax;
    __shared__ float s_zmax, s_zmin;
    __shared__ uint32_t s_max;
    __shared__ ushort4 s_px0;                                       // the colour of an untouched pixel (count 0, steps +0.0, z -1.0):
    __shared__ float4 s_fx0;                                        // ~80 % of a frame — computed once per block by the same function
    if (threadIdx.x == 0) {
        s_max = scal->max;
        s_lnmax = ln_max1(C, s_max);
        s_zmax = __uint_as_float(zbits_from_key(scal->zmax_key));
        s_zmin = __uint_as_float(zbits_from_key(scal->zmin_key));
        ushort4 px; float4 fx;
        color_pixel(C, make_ulonglong2(0ull, REC_HI_RESET), 0u, s_max, s_lnmax, s_zmax, s_zmin, px, fx);
        s_px0 = px; s_fx0 = fx;
    }
    __syncthreads();
    const size_t pix0 = (size_t)C.row0 * C.W, npix = (size_t)C.rows * C.W;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        const size_t p = pix0 + i;
        ushort4 px = s_px0;
        float4 fx = s_fx0;
        const ulonglong2 r = rec[p];
        const uint32_t cnt = pixel_count(fast, scal, p, C.slots);
        if (cnt != 0u || r.x != 0ull || (uint32_t)(r.y >> 32) != ZKEY_SENTINEL)
            color_pixel(C, r, cnt, s_max, s_lnmax, s_zmax, s_zmin, px, fx);
        if (out16) reinterpret_cast<ushort4 *>(out16)[p] = px;
        if (out32) reinterpret_cast<float4 *>(out32)[p] = fx;
    }
}
void launch_colorize(const ColorParams &cp, const unsigned long long *fast, const ulonglong2 *rec, const Scalars *scal,
                     uint16_t *rgba_u16, float *rgba_f32, cudaStream_t s)
{
    const size_t npix = (size_t)cp.rows * cp.W;
    if (npix == 0) return;
    const unsigned int block = 256;
    size_t g = (npix + block - 1) / block;
    const unsigned int grid = (unsigned int)(g > 148u * 32u ? 148u * 32u : g);
    colorize_kernel<<<grid, block, 0, s>>>(cp, fast, rec, scal, rgba_u16, rgba_f32);
    ++g_launches;
}

// colorize() of this rank's stripe inside the cross-GPU frame protocol: wait for every rank's stripe
// maximum (MAX_READY) and for the image owner to be done with the previous frame (IMAGE_FREE), fold
// the stripe maxima — Runtime.max (the log base of lib.rs:860) and the Depth min/max (lib.rs:877-882)
// are global —, colourise from the merged pixel-order counts straight into the owner's image (remote
// stores over NVLink), and raise IMAGE_DONE at the owner.
__global__ void __launch_bounds__(256)
frame_colorize_kernel(const __grid_constant__ ColorParams C, const uint32_t *__restrict__ cnt,
                      const ulonglong2 *__restrict__ rec, Scalars *scal, uint16_t *out16,
                      const __grid_constant__ FrameSync S, int owner, long long timeout)
{
    __shared__ double s_lnmax;
    __shared__ float s_zmax, s_zmin;
    __shared__ uint32_t s_max;
    __shared__ ushort4 s_px0;
    if (!frame_wait(scal, SYNC_MAX_READY, 0, S.n_ranks, S.epoch, timeout)) return;
    if (!frame_wait(scal, SYNC_IMAGE_FREE, owner, 1, S.epoch - 1u, timeout)) return;
    if (threadIdx.x == 0) {
        uint32_t m = 0, zmx = ZKEY_ZERO, zmn = ZKEY_FLT_MAX;
        for (int r = 0; r < S.n_ranks; ++r) {
            const uint32_t a = *((volatile unsigned int *)&scal->stripe_max[r]);
            const uint32_t b = *((volatile unsigned int *)&scal->stripe_zmax[r]);
            const uint32_t c = *((volatile unsigned int *)&scal->stripe_zmin[r]);
            m = a > m ? a : m; zmx = b > zmx ? b : zmx; zmn = c < zmn ? c : zmn;
        }
        s_max = m;
        s_lnmax = ln_max1(C, m);
        s_zmax = __uint_as_float(zbits_from_key(zmx));
        s_zmin = __uint_as_float(zbits_from_key(zmn));
        if (blockIdx.x == 0) { scal->max = m; scal->zmax_key = zmx; scal->zmin_key = zmn; }   // Runtime.max of the whole frame
        ushort4 px; float4 fx;
        color_pixel(C, make_ulonglong2(0ull, REC_HI_RESET), 0u, s_max, s_lnmax, s_zmax, s_zmin, px, fx);
        s_px0 = px;
    }
    __syncthreads();
    const size_t pix0 = (size_t)C.row0 * C.W, npix = (size_t)C.rows * C.W;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        const size_t p = pix0 + i;
        ushort4 px = s_px0;                                         // untouched pixel: the block's precomputed colour
        float4 fx;
        const ulonglong2 r = rec[p];
        const uint32_t c = cnt[p];
        if (c != 0u || r.x != 0ull || (uint32_t)(r.y >> 32) != ZKEY_SENTINEL)
            color_pixel(C, r, c, s_max, s_lnmax, s_zmax, s_zmin, px, fx);
        reinterpret_cast<ushort4 *>(out16)[p] = px;
    }
    if (frame_last_block(&scal->done_counter[2])) frame_signal(S, SYNC_IMAGE_DONE, false, owner);
}
void launch_frame_colorize(const ColorParams &cp, const uint32_t *cnt, const ulonglong2 *rec, Scalars *scal, uint16_t *rgba_u16,
                           const FrameSync &S, int owner, cudaStream_t s)
{
    const size_t npix = (size_t)cp.rows * cp.W;
    const unsigned int block = 256;
    size_t g = (npix + block - 1) / block;
    const unsigned int grid = (unsigned int)(g > 148u * 16u ? 148u * 16u : (g ? g : 1));
    frame_colorize_kernel<<<grid, block, 0, s>>>(cp, cnt, rec, scal, rgba_u16, S, owner, g_sync_timeout_cycles.load());
    ++g_launches;
}

// ---------------------------------------------------------------------------------------------
// Output conversion (src/bin/main.rs:52-57): FinalImage (RGBA u16) -> what the encoders are handed.
//   to_rgb16  drops alpha; to_rgba8 / to_rgb8 narrow every sample with the `image` crate's
//   u16 -> u8 rule (image 0.25, color.rs, FromPrimitive<u16> for u8): (c + 128) / 257, i.e.
//   round(c * 255 / 65535) — third-party code absent from /root/reference, restated from its
//   published source ("parity unpinned (third-party)").
// Sample order: native little-endian (what DynamicImage::as_bytes() holds), big-endian 16-bit
// samples (what the PNM/PAM and PNG encoders write, main.rs:62-68), or BMP order (B,G,R[,A], rows
// bottom-up, each row padded to 4 bytes; 8-bit only, main.rs:70-76).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t narrow_u16(uint32_t c) { return (c + 128u) / 257u; }

__global__ void convert_kernel(const ushort4 *__restrict__ img, uint8_t *__restrict__ out, unsigned int W, unsigned int H,
                               unsigned int fmt, unsigned int order, size_t row_stride)
{
    const size_t npix = (size_t)W * H;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
        const ushort4 v = img[p];
        const unsigned int y = (unsigned int)(p / W), x = (unsigned int)(p - (size_t)y * W);
        const bool alpha = fmt == PIX_RGBA16 || fmt == PIX_RGBA8;
        if (fmt == PIX_RGBA16 || fmt == PIX_RGB16) {
            uint16_t c[4] = {v.x, v.y, v.z, v.w};
            if (order == ORDER_BIG_ENDIAN)
                for (int k = 0; k < 4; ++k) c[k] = (uint16_t)((c[k] >> 8) | (c[k] << 8));
            uint16_t *o = reinterpret_cast<uint16_t *>(out + (size_t)y * row_stride) + (size_t)x * (alpha ? 4 : 3);
            if (alpha) *reinterpret_cast<ushort4 *>(o) = make_ushort4(c[0], c[1], c[2], c[3]);
            else { o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; }
        } else {
            const uint32_t r = narrow_u16(v.x), g = narrow_u16(v.y), b = narrow_u16(v.z), a = narrow_u16(v.w);
            const unsigned int yy = order == ORDER_BMP ? H - 1u - y : y;
            uint8_t *o = out + (size_t)yy * row_stride + (size_t)x * (alpha ? 4 : 3);
            if (order == ORDER_BMP) {
                if (alpha) *reinterpret_cast<uint32_t *>(o) = b | (g << 8) | (r << 16) | (a << 24);
                else { o[0] = (uint8_t)b; o[1] = (uint8_t)g; o[2] = (uint8_t)r; }
                if (x == W - 1u) for (size_t k = (size_t)W * (alpha ? 4 : 3); k < row_stride; ++k) out[(size_t)yy * row_stride + k] = 0;   // row padding
            } else {
                if (alpha) *reinterpret_cast<uint32_t *>(o) = r | (g << 8) | (b << 16) | (a << 24);
                else { o[0] = (uint8_t)r; o[1] = (uint8_t)g; o[2] = (uint8_t)b; }
            }
        }
    }
}
void launch_convert(const uint16_t *rgba, uint8_t *out, unsigned int W, unsigned int H, unsigned int fmt, unsigned int order,
                    size_t row_stride, cudaStream_t s)
{
    const size_t npix = (size_t)W * H;
    if (npix == 0) return;
    const unsigned int block = 256;
    size_t g = (npix + block - 1) / block;
    convert_kernel<<<(unsigned int)(g > 148u * 16u ? 148u * 16u : g), block, 0, s>>>(reinterpret_cast<const ushort4 *>(rgba), out, W, H, fmt, order, row_stride);
    ++g_launches;
}

// ---------------------------------------------------------------------------------------------
// PNG (main.rs:78-89) without the compressor: the converted image as filter-type-0 scanlines inside
// zlib "stored" deflate blocks — a valid PNG any decoder reads back to the same pixels; the
// reference's encoder additionally deflates (png::CompressionType::Default), which stays on the host.
// The device writes the IDAT payload (block headers + scanlines, 16-bit samples big-endian) and the
// per-chunk partial checksums the host folds into the zlib Adler-32 and the chunk CRC-32.
//   raw stream  : for each row  [0x00 filter byte][W * bpp sample bytes]
//   payload     : raw stream cut into blocks of <= 65535 bytes, each behind [BFINAL, LEN lo, hi, ~LEN lo, hi]
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline size_t png_payload_offset(size_t k) { return k + 5 * (k / 65535 + 1); }

__global__ void png_pack_kernel(const ushort4 *__restrict__ img, uint8_t *__restrict__ out, unsigned int W, unsigned int H,
                                unsigned int fmt, size_t raw_row, size_t raw_len, size_t n_blocks)
{
    const size_t npix = (size_t)W * H;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const bool wide = fmt == PIX_RGBA16 || fmt == PIX_RGB16, alpha = fmt == PIX_RGBA16 || fmt == PIX_RGBA8;
    const unsigned int nch = alpha ? 4u : 3u, bpp = nch * (wide ? 2u : 1u);
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
        const ushort4 v = img[p];
        const unsigned int y = (unsigned int)(p / W), x = (unsigned int)(p - (size_t)y * W);
        const uint16_t c[4] = {v.x, v.y, v.z, v.w};
        size_t k = (size_t)y * raw_row;
        if (x == 0) out[png_payload_offset(k)] = 0;                          // filter type 0 (None)
        k += 1 + (size_t)x * bpp;
        for (unsigned int ch = 0; ch < nch; ++ch) {
            if (wide) {
                out[png_payload_offset(k)] = (uint8_t)(c[ch] >> 8); ++k;    // most significant byte first
                out[png_payload_offset(k)] = (uint8_t)c[ch]; ++k;
            } else {
                out[png_payload_offset(k)] = (uint8_t)narrow_u16(c[ch]); ++k;
            }
        }
        if (p < n_blocks) {                                                  // the stored-block headers
            const size_t first = p * 65535;
            const size_t len = raw_len - first < 65535 ? raw_len - first : 65535;
            uint8_t *hd = out + first + 5 * p;
            hd[0] = p + 1 == n_blocks ? 1 : 0;
            hd[1] = (uint8_t)len; hd[2] = (uint8_t)(len >> 8); hd[3] = (uint8_t)~len; hd[4] = (uint8_t)(~len >> 8);
        }
    }
}
// Partial checksums over chunks of PNG_CHUNK bytes: crc[i] = CRC-32 register after chunk i of the PAYLOAD starting
// from register 0 (no pre/post conditioning: the linear part, combined on the host); adler[i] = (sum d, sum (len - j) d_j)
// over chunk i of the RAW stream.
__global__ void png_sums_kernel(const uint8_t *__restrict__ payload, size_t payload_len, size_t raw_len,
                                uint32_t *__restrict__ crc, unsigned long long *__restrict__ adler, size_t n_crc, size_t n_adler)
{
    __shared__ uint32_t table[256];
    for (unsigned int n = threadIdx.x; n < 256; n += blockDim.x) {
        uint32_t c = n;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        table[n] = c;
    }
    __syncthreads();
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_crc) {
        const size_t lo = t * PNG_CHUNK, hi = lo + PNG_CHUNK < payload_len ? lo + PNG_CHUNK : payload_len;
        uint32_t c = 0;
        for (size_t i = lo; i < hi; ++i) c = table[(c ^ payload[i]) & 0xFFu] ^ (c >> 8);
        crc[t] = c;
    }
    if (t < n_adler) {
        const size_t lo = t * PNG_CHUNK, hi = lo + PNG_CHUNK < raw_len ? lo + PNG_CHUNK : raw_len;
        unsigned long long a = 0, b = 0;
        for (size_t k = lo; k < hi; ++k) { a += payload[png_payload_offset(k)]; b += a; }   // b = sum over j of (hi - lo - j) d_j
        adler[2 * t] = a; adler[2 * t + 1] = b;
    }
}
void launch_png_pack(const uint16_t *rgba, uint8_t *out, unsigned int W, unsigned int H, unsigned int fmt, size_t raw_row,
                     size_t raw_len, size_t n_blocks, cudaStream_t s)
{
    const size_t npix = (size_t)W * H;
    if (npix == 0) return;
    const unsigned int block = 256;
    size_t g = (npix + block - 1) / block;
    png_pack_kernel<<<(unsigned int)(g > 148u * 16u ? 148u * 16u : g), block, 0, s>>>(reinterpret_cast<const ushort4 *>(rgba), out, W, H, fmt, raw_row, raw_len, n_blocks);
    ++g_launches;
}
void launch_png_sums(const uint8_t *payload, size_t payload_len, size_t raw_len, uint32_t *crc, unsigned long long *adler,
                     size_t n_crc, size_t n_adler, cudaStream_t s)
{
    const size_t n = n_crc > n_adler ? n_crc : n_adler;
    if (n == 0) return;
    png_sums_kernel<<<(unsigned int)((n + 127) / 128), 128, 0, s>>>(payload, payload_len, raw_len, crc, adler, n_crc, n_adler);
    ++g_launches;
}

// ---------------------------------------------------------------------------------------------
// layout conversion to / from the reference's three textures (lib.rs:633-639)
// ---------------------------------------------------------------------------------------------
__global__ void unpack_kernel(const unsigned long long *fast, const ulonglong2 *rec, const Scalars *scal, size_t npix, SlotMap slots,
                              uint32_t *count, double *steps, float *zbuf)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        if (count) count[i] = pixel_count(fast, scal, i, slots);
        const ulonglong2 r = rec[i];
        if (steps) steps[i] = __longlong_as_double((long long)r.x);
        if (zbuf) zbuf[i] = __uint_as_float(zbits_from_key((uint32_t)(r.y >> 32)));
    }
}
void launch_unpack(const unsigned long long *fast, const ulonglong2 *rec, const Scalars *scal, size_t npix, SlotMap slots,
                   uint32_t *count, double *steps, float *zbuf, cudaStream_t s)
{
    if (npix == 0) return;
    const unsigned int block = 256;
    size_t g = (npix + block - 1) / block;
    unpack_kernel<<<(unsigned int)(g > 148u * 16u ? 148u * 16u : g), block, 0, s>>>(fast, rec, scal, npix, slots, count, steps, zbuf);
    ++g_launches;
}

__global__ void pack_kernel(unsigned long long *fast, ulonglong2 *rec, Scalars *scal, size_t npix, SlotMap slots,
                            const uint32_t *count, const double *steps, const float *zbuf)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        const float z = zbuf[i];
        uint32_t k = zkey_of(z);                                // keeps the sign of a zero
        if (!(z > -1.0f)) k = ZKEY_SENTINEL;                    // untouched (or invalid) pixels
        const uint32_t hint = k == ZKEY_SENTINEL ? ZKEY_SENTINEL + 1u : canon_key(k);
        fast[slot_of((uint32_t)i, slots)] = ((unsigned long long)hint << 32) | count[i];
        // uploaded records predate every future job: they keep all z ties (job key 0)
        rec[i] = make_ulonglong2((unsigned long long)__double_as_longlong(steps[i]), ((unsigned long long)k << 32) | 0xFFFFFFFFull);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scal->nan_sink = 0ull; scal->max = 0u; scal->zmax_key = ZKEY_ZERO; scal->zmin_key = ZKEY_FLT_MAX; scal->pad = 0u;
    }
}
void launch_pack(unsigned long long *fast, ulonglong2 *rec, Scalars *scal, size_t npix, SlotMap slots,
                 const uint32_t *count, const double *steps, const float *zbuf, cudaStream_t s)
{
    const unsigned int block = 256;
    size_t g = (npix + block - 1) / block;
    pack_kernel<<<(unsigned int)(g > 148u * 16u ? 148u * 16u : (g ? g : 1)), block, 0, s>>>(fast, rec, scal, npix, slots, count, steps, zbuf);
    ++g_launches;
}

// ---------------------------------------------------------------------------------------------
// Runtime::merge (lib.rs:708-738): count +=, `other` wins iff its z is strictly greater.
// ---------------------------------------------------------------------------------------------
__global__ void merge_kernel(unsigned long long *dfast, ulonglong2 *drec, Scalars *dscal,
                             const unsigned long long *sfast, const ulonglong2 *srec, const Scalars *sscal, size_t npix, SlotMap slots)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        const uint32_t sl = slot_of((uint32_t)i, slots);
        const uint32_t c = (uint32_t)dfast[sl] + (uint32_t)sfast[sl];               // lib.rs:719
        ulonglong2 d = drec[i];
        const ulonglong2 o = srec[i];
        if (canon_key((uint32_t)(o.y >> 32)) > canon_key((uint32_t)(d.y >> 32))) { d = o; drec[i] = d; }  // lib.rs:728-735
        const uint32_t k = canon_key((uint32_t)(d.y >> 32));
        const uint32_t hint = k == ZKEY_SENTINEL ? ZKEY_SENTINEL + 1u : k;
        dfast[sl] = ((unsigned long long)hint << 32) | c;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) dscal->nan_sink += sscal->nan_sink;
}
void launch_merge(unsigned long long *dfast, ulonglong2 *drec, Scalars *dscal,
                  const unsigned long long *sfast, const ulonglong2 *srec, const Scalars *sscal,
                  size_t npix, SlotMap slots, cudaStream_t s)
{
    const unsigned int block = 256;
    size_t g = (npix + block - 1) / block;
    merge_kernel<<<(unsigned int)(g > 148u * 16u ? 148u * 16u : (g ? g : 1)), block, 0, s>>>(dfast, drec, dscal, sfast, srec, sscal, npix, slots);
    ++g_launches;
}

// ---------------------------------------------------------------------------------------------
// All-ranks merge of one row stripe, reading every peer's accumulators directly over
// NVLink (CUDA IPC mappings).  Deterministic form of Runtime::merge: counts add, the
// record with the greatest (zkey, ~job) wins — identical to rendering every job on one
// Runtime, whatever the number of ranks.
// ---------------------------------------------------------------------------------------------
__global__ void merge_peers_kernel(unsigned long long *dfast, ulonglong2 *drec, Scalars *dscal,
                                   const __grid_constant__ PeerList peers, size_t pix0, size_t npix, SlotMap slots)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        const size_t p = pix0 + i;
        const uint32_t sl = slot_of((uint32_t)p, slots);
        uint32_t c = (uint32_t)dfast[sl];
        ulonglong2 d = drec[p];
        for (int r = 0; r < peers.n; ++r) {
            const uint32_t pc = (uint32_t)__ldcv(peers.fast[r] + sl);
            if (pc == 0u) continue;                                      // never hit there: its record is the reset value
            c += pc;
            const unsigned long long oy = __ldcv(&peers.rec[r][p].y);
            if (rec_order(oy) > rec_order(d.y)) { d.y = oy; d.x = __ldcv(&peers.rec[r][p].x); }
        }
        const uint32_t k = canon_key((uint32_t)(d.y >> 32));
        const uint32_t hint = k == ZKEY_SENTINEL ? ZKEY_SENTINEL + 1u : k;
        drec[p] = d;
        dfast[sl] = ((unsigned long long)hint << 32) | c;
    }
    if (pix0 == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long s = dscal->nan_sink;
        for (int r = 0; r < peers.n; ++r) s += *((volatile const unsigned long long *)&peers.scal[r]->nan_sink);
        dscal->nan_sink = s;
    }
}
void launch_merge_peers(unsigned long long *dfast, ulonglong2 *drec, Scalars *dscal, const PeerList &peers,
                        size_t pix0, size_t npix, SlotMap slots, cudaStream_t s)
{
    if (npix == 0) return;
    const unsigned int block = 256;
    size_t g = (npix + block - 1) / block;
    merge_peers_kernel<<<(unsigned int)(g > 148u * 16u ? 148u * 16u : g), block, 0, s>>>(dfast, drec, dscal, peers, pix0, npix, slots);
    ++g_launches;
}

// ---------------------------------------------------------------------------------------------
// Cross-GPU frame protocol: kernels (helpers: frame_wait / frame_last_block / frame_signal above)
// ---------------------------------------------------------------------------------------------
// standalone forms (image hand-over on the owner rank; tests)
__global__ void signal_kernel(const __grid_constant__ FrameSync S, int kind)
{
    __threadfence_system();
    frame_signal(S, kind, true, -1);
}
void launch_signal(const FrameSync &S, int kind, cudaStream_t s)
{
    signal_kernel<<<1, 32, 0, s>>>(S, kind);
    ++g_launches;
}
__global__ void wait_kernel(Scalars *mine, int kind, int n_ranks, unsigned int epoch, long long timeout)
{
    (void)frame_wait(mine, kind, 0, n_ranks, epoch, timeout);
}
void launch_wait(Scalars *mine, int kind, int n_ranks, unsigned int epoch, cudaStream_t s)
{
    wait_kernel<<<1, 32, 0, s>>>(mine, kind, n_ranks, epoch, g_sync_timeout_cycles.load());
    ++g_launches;
}

// Runtime::reset for a frame of the protocol: first wait until every peer has finished reading this
// rank's accumulators of the previous frame (MERGE_DONE, epoch - 1).
__global__ void frame_reset_kernel(unsigned long long *fast, ulonglong2 *rec, Scalars *scal, size_t npix, size_t nslots,
                                   int n_ranks, unsigned int epoch, long long timeout)
{
    if (!frame_wait(scal, SYNC_MERGE_DONE, 0, n_ranks, epoch - 1u, timeout)) return;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nslots; i += stride) {
        fast[i] = FAST_RESET;
        if (i < npix) rec[i] = make_ulonglong2(0ull, REC_HI_RESET);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scal->nan_sink = 0ull; scal->max = 0u;
        scal->zmax_key = ZKEY_ZERO; scal->zmin_key = ZKEY_FLT_MAX; scal->pad = 0u;
    }
}
void launch_frame_reset(unsigned long long *fast, ulonglong2 *rec, Scalars *scal, size_t npix, size_t nslots,
                        int n_ranks, unsigned int epoch, cudaStream_t s)
{
    const unsigned int block = 256;
    size_t g = (nslots + block - 1) / block;
    const unsigned int grid = (unsigned int)(g > 148u * 16u ? 148u * 16u : (g ? g : 1));
    frame_reset_kernel<<<grid, block, 0, s>>>(fast, rec, scal, npix, nslots, n_ranks, epoch, g_sync_timeout_cycles.load());
    ++g_launches;
}

// After the trajectories: the counts in PIXEL order (cnt[p], u32, NaN debt folded into pixel 0) so that
// the stripe owners read them from this rank with coalesced 16-byte loads instead of one scattered
// 8-byte `fast` word (= one 32-byte sector over NVLink) per pixel; then RENDER_DONE to every rank.
__global__ void frame_export_kernel(const unsigned long long *fast, Scalars *scal, uint32_t *cnt, size_t npix, SlotMap slots,
                                    const __grid_constant__ FrameSync S)
{
    if (*((volatile unsigned int *)&scal->sync_error) != 0u) return;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride)
        cnt[i] = pixel_count(fast, scal, i, slots);
    // the iterate kernel left this rank's own running max here; the merge reduces the max of the MERGED stripe
    // into the same word, so start it from 0 (the local value is a lower bound of the merged one, but be exact by construction)
    if (blockIdx.x == 0 && threadIdx.x == 0) scal->max = 0u;
    if (frame_last_block(&scal->done_counter[0])) frame_signal(S, SYNC_RENDER_DONE, true, -1);
}
void launch_frame_export(const unsigned long long *fast, Scalars *scal, uint32_t *cnt, size_t npix, SlotMap slots,
                         const FrameSync &S, cudaStream_t s)
{
    const unsigned int block = 256;
    size_t g = (npix + block - 1) / block;
    frame_export_kernel<<<(unsigned int)(g > 148u * 16u ? 148u * 16u : (g ? g : 1)), block, 0, s>>>(fast, scal, cnt, npix, slots, S);
    ++g_launches;
}

// All-ranks merge of one pixel stripe, reading every peer's counts and records directly over NVLink
// (CUDA IPC mappings).  Deterministic form of Runtime::merge (lib.rs:708-738): counts add, the record
// with the greatest (z, earlier job) wins — identical to rendering every job on one Runtime, whatever
// the number of ranks.  A peer's 16-byte record is only fetched where that peer counted a hit
// (~19 % of the pixels).  The stripe's share of Runtime.max (lib.rs:721-723) and of the Depth fold
// (lib.rs:877-882) is reduced on the way; the last block publishes it to every rank and raises
// MAX_READY and MERGE_DONE.
__global__ void __launch_bounds__(256)
frame_merge_kernel(unsigned long long *dfast, ulonglong2 *drec, uint32_t *dcnt, Scalars *dscal,
                   const __grid_constant__ PeerList peers, size_t pix0, size_t npix, SlotMap slots,
                   const __grid_constant__ FrameSync S, long long timeout)
{
    if (!frame_wait(dscal, SYNC_RENDER_DONE, 0, S.n_ranks, S.epoch, timeout)) return;
    uint32_t m = 0, zmx = ZKEY_ZERO, zmn = ZKEY_FLT_MAX;
    // groups of 4 pixels, aligned to 16 bytes of cnt[]; the stripe's ends may cut a group
    const size_t g0 = pix0 / 4, g1 = (pix0 + npix + 3) / 4;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t g = g0 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < g1; g += stride) {
        const uint4 own = reinterpret_cast<const uint4 *>(dcnt)[g];
        uint32_t c[4] = {own.x, own.y, own.z, own.w};
        bool in[4], changed[4];
        ulonglong2 d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const size_t p = 4 * g + j;
            in[j] = p >= pix0 && p < pix0 + npix;
            changed[j] = false;
            d[j] = in[j] ? drec[p] : make_ulonglong2(0ull, REC_HI_RESET);
        }
        for (int r = 0; r < peers.n; ++r) {
            const uint4 v4 = __ldcv(reinterpret_cast<const uint4 *>(peers.cnt[r]) + g);
            const uint32_t v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!in[j] || v[j] == 0u) continue;                      // this peer never hit the pixel: its record is the reset value
                c[j] += v[j];                                            // lib.rs:719 (wrapping)
                const ulonglong2 *pr = &peers.rec[r][4 * g + j];
                const unsigned long long oy = __ldcv(&pr->y);
                if (rec_order(oy) > rec_order(d[j].y)) { d[j].y = oy; d[j].x = __ldcv(&pr->x); changed[j] = true; }
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (!in[j]) { c[j] = j == 0 ? own.x : j == 1 ? own.y : j == 2 ? own.z : own.w; continue; }
            const size_t p = 4 * g + j;
            const uint32_t k = canon_key((uint32_t)(d[j].y >> 32));
            if (changed[j]) drec[p] = d[j];
            dfast[slot_of((uint32_t)p, slots)] = ((unsigned long long)(k == ZKEY_SENTINEL ? ZKEY_SENTINEL + 1u : k) << 32) | c[j];
            m = c[j] > m ? c[j] : m;
            if (k != ZKEY_SENTINEL) { zmx = k > zmx ? k : zmx; zmn = k < zmn ? k : zmn; }
        }
        // merged counts back in pixel order for the colourise pass (pixels outside the stripe keep their value)
        reinterpret_cast<uint4 *>(dcnt)[g] = make_uint4(c[0], c[1], c[2], c[3]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t m2 = __shfl_xor_sync(0xffffffffu, m, o);
        const uint32_t a2 = __shfl_xor_sync(0xffffffffu, zmx, o);
        const uint32_t b2 = __shfl_xor_sync(0xffffffffu, zmn, o);
        m = m2 > m ? m2 : m; zmx = a2 > zmx ? a2 : zmx; zmn = b2 < zmn ? b2 : zmn;
    }
    if ((threadIdx.x & 31) == 0) {
        if (m) atomicMax(&dscal->max, m);
        atomicMax(&dscal->zmax_key, zmx);
        atomicMin(&dscal->zmin_key, zmn);
    }
    if (frame_last_block(&dscal->done_counter[1])) {
        if ((int)threadIdx.x < S.n_ranks) {
            Scalars *t = S.scal[threadIdx.x];
            *((volatile unsigned int *)&t->stripe_max[S.my_rank]) = *((volatile unsigned int *)&dscal->max);
            *((volatile unsigned int *)&t->stripe_zmax[S.my_rank]) = *((volatile unsigned int *)&dscal->zmax_key);
            *((volatile unsigned int *)&t->stripe_zmin[S.my_rank]) = *((volatile unsigned int *)&dscal->zmin_key);
            __threadfence_system();
        }
        if (pix0 == 0 && threadIdx.x == 0) dscal->nan_sink = 0ull;      // the debts of all ranks are inside the merged count of pixel 0 now
        frame_signal(S, SYNC_MAX_READY, true, -1);
        frame_signal(S, SYNC_MERGE_DONE, true, -1);
    }
}
void launch_frame_merge(unsigned long long *dfast, ulonglong2 *drec, uint32_t *dcnt, Scalars *dscal, const PeerList &peers,
                        size_t pix0, size_t npix, SlotMap slots, const FrameSync &S, cudaStream_t s)
{
    const unsigned int block = 256;
    size_t g = (npix / 4 + block) / block;
    frame_merge_kernel<<<(unsigned int)(g > 148u * 8u ? 148u * 8u : (g ? g : 1)), block, 0, s>>>(
        dfast, drec, dcnt, dscal, peers, pix0, npix, slots, S, g_sync_timeout_cycles.load());
    ++g_launches;
}

__global__ void seed_points_kernel(unsigned long long seed, unsigned long long first, unsigned long long n, double *out)
{
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 3ull * n) out[i] = seed_coord(seed, 3ull * first + i);
}
void launch_seed_points(unsigned long long seed, unsigned long long first, unsigned long long n, double *out, cudaStream_t s)
{
    if (n == 0) return;
    const unsigned long long total = 3ull * n;
    seed_points_kernel<<<(unsigned int)((total + 255) / 256), 256, 0, s>>>(seed, first, n, out);
    ++g_launches;
}

}  // namespace sar
