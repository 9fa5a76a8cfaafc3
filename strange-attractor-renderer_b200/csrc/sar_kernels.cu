// sar_kernels.cu — hand-written sm_100a kernels for the iterate → project → scatter →
// tone-map/colourise path of Icelk/strange-attractor-renderer (src/lib.rs:747-904).
//
// Arithmetic contract: every f64 operation on the trajectory path is an explicit
// round-to-nearest __dmul_rn/__dadd_rn/__dsub_rn (never contracted to FMA, whatever the
// compiler flags), in the reference's exact association order, so that the chaotic map
// reproduces the CPU trajectory bit for bit.
#include "sar_device.cuh"

#include <atomic>

namespace sar {

static std::atomic<unsigned long long> g_launches{0};
unsigned long long launch_count() { return g_launches.load(); }

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
#define SAR_NEXT_POINT(P, x, y, z, nx, ny, nz)                                       \
    {                                                                                \
        const double xx_ = __dmul_rn(x, x), xy_ = __dmul_rn(x, y), xz_ = __dmul_rn(x, z); \
        const double yy_ = __dmul_rn(y, y), yz_ = __dmul_rn(y, z), zz_ = __dmul_rn(z, z); \
        nx = sprott_sum(P.c[0], x, y, z, xx_, xy_, xz_, yy_, yz_, zz_);              \
        ny = sprott_sum(P.c[1], x, y, z, xx_, xy_, xz_, yy_, yz_, zz_);              \
        nz = sprott_sum(P.c[2], x, y, z, xx_, xy_, xz_, yy_, yz_, zz_);              \
    }

// Vec3::magnitude (lib.rs:129-131)
__device__ __forceinline__ double magnitude(double x, double y, double z)
{
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
}

// The winning branch of the depth test (lib.rs:821-833), off the hot loop.
//   value = color_transform.transform(delta, screen_space, view)   lib.rs:826-828
//   steps[idx] = value; zbuf[idx] = z2 as f32                       lib.rs:830-832
// made atomic and order-independent: the record is replaced iff (zkey, ~job) is strictly
// greater than the stored one.
__device__ __noinline__ void record_win(unsigned long long *fast, ulonglong2 *rec, unsigned int idx,
                                        uint32_t key, uint32_t job_inv, unsigned int ct_kind,
                                        double ct_offset, double ct_factor, double vccx, double vccy,
                                        double dx, double dy, double dz, double sx, double sy, double sz)
{
    double value;
    const double mag = magnitude(dx, dy, dz);
    if (ct_kind == 1u) {
        value = __dmul_rn(__dadd_rn(mag, ct_offset), ct_factor);             // AdjustedVelocity, lib.rs:514
    } else {
        // color_transforms::poisson_saturne, lib.rs:520-558 (COS/SIN literals lib.rs:529-536)
        const double COS = 0.7009092642998508981833083453238941729068756103515625;
        const double SIN = 0.7132504491541815649924274111981503665447235107421875;
        const double x2 = __dadd_rn(__dmul_rn(__dadd_rn(sx, vccx), COS), __dmul_rn(__dadd_rn(sz, vccy), SIN));
        const bool out = (x2 < -0.0839) ||
                         (__dadd_rn(__dmul_rn(10.55, x2), sy) < (0.46 - 1.0941)) ||
                         (__dadd_rn(__dmul_rn(1.0426, x2), sy) < (0.179 - 0.1576)) ||
                         (__dsub_rn(__dmul_rn(0.5139, x2), sy) > (-0.04 - 0.04092));
        const double part = out ? 0. : 1.;
        const double color = __ddiv_rn(__dadd_rn(part, mag), 2.);            // lib.rs:556
        value = __ddiv_rn(__dsub_rn(color, 0.1), 0.9);                       // lib.rs:557
    }
    const unsigned long long hi = ((unsigned long long)key << 32) | job_inv;
    ulonglong2 *r = rec + idx;
    ulonglong2 cur;
    cur.y = __ldcg(&r->y);
    cur.x = __ldcg(&r->x);            // may be torn against a concurrent writer; the CAS validates it
    while (hi > cur.y) {
        const ulonglong2 want = make_ulonglong2((unsigned long long)__double_as_longlong(value), hi);
        const ulonglong2 old = cas128(r, cur, want);
        if (old.x == cur.x && old.y == cur.y) break;
        cur = old;
    }
    // raise the hint so later candidates below this z skip the slow path
    unsigned long long f = __ldcg(fast + idx);
    while ((uint32_t)(f >> 32) < key) {
        const unsigned long long nf = ((unsigned long long)key << 32) | (f & 0xFFFFFFFFull);
        const unsigned long long old = atomicCAS(fast + idx, f, nf);
        if (old == f) break;
        f = old;
    }
}

// ---------------------------------------------------------------------------------------------
// iterate → project → scatter: render() (lib.rs:747-838), one lane per trajectory.
// Lane L runs jobs L, L+lanes, L+2*lanes, ...; each job is one reference render() call:
// start point, 1000 warm-up steps (lib.rs:750-752), `iterations` recorded steps.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
iterate_kernel(const __grid_constant__ IterParams P)
{
    const unsigned long long lanes = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long job = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; job < P.n_jobs; job += lanes) {
        double x, y, z;
        if (P.init != nullptr) {
            x = P.init[3 * job + 0]; y = P.init[3 * job + 1]; z = P.init[3 * job + 2];
        } else {
            const unsigned long long g = 3ull * (P.first_job + job);
            x = seed_coord(P.seed, g); y = seed_coord(P.seed, g + 1); z = seed_coord(P.seed, g + 2);
        }
        for (int w = 0; w < 1000; ++w) {                                      // lib.rs:750-752
            double nx, ny, nz;
            SAR_NEXT_POINT(P, x, y, z, nx, ny, nz);
            x = nx; y = ny; z = nz;
        }
        // order key of this job: earlier jobs win z ties (see sar_device.cuh)
        const unsigned long long jk = (unsigned long long)P.job_key0 + job;
        const uint32_t job_inv = 0xFFFFFFFFu - (uint32_t)(jk > 0xFFFFFFFFull ? 0xFFFFFFFFull : jk);

        for (unsigned long long it = 0; it < P.iterations; ++it) {            // lib.rs:769
            double nx, ny, nz;
            SAR_NEXT_POINT(P, x, y, z, nx, ny, nz);                           // lib.rs:770
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
            // Bounds test + `as u32` (lib.rs:789-802) in one step: floor-convert (saturating,
            // NaN -> 0) and compare unsigned.  i in [0,W) <=> 0 <= floor(i) < W; floor == trunc
            // there; -0.0 -> 0 and NaN -> 0 pass exactly as in the reference.
            const unsigned int ii = (unsigned int)__double2int_rd(fi);
            const unsigned int jj = (unsigned int)__double2int_rd(fj);
            if (ii < P.W && jj < P.H) {
                const unsigned int idx = jj * P.W + ii;
                if (idx == 0u && (nx != nx || ny != ny || nz != nz)) {
                    // NaN is absorbing: this and every remaining iteration lands on count[(0,0)]
                    // and can never win the z test.  Pay the debt in one atomic.
                    atomicAdd(&P.scal->nan_sink, P.iterations - it);
                    break;
                }
                // count += 1 (lib.rs:811) and fetch the depth hint in one L2 atomic
                const unsigned long long old = atomicAdd(P.fast + idx, 1ull);
                const float zf = __double2float_rn(z2) + 0.0f;                // `z2 as f32`; -0 -> +0
                const uint32_t key = zkey_of(zf);
                if (key >= (uint32_t)(old >> 32) && key <= ZKEY_POS_INF) {    // may beat zbuf (lib.rs:821); NaN never does
                    record_win(P.fast, P.rec, idx, key, job_inv, P.ct_kind, P.ct_offset, P.ct_factor, P.ccx, P.ccy,
                               __dsub_rn(nx, x), __dsub_rn(ny, y), __dsub_rn(nz, z), sx, sy, sz);   // delta, lib.rs:822
                }
            }
            x = nx; y = ny; z = nz;                                           // previous_point = current_point, lib.rs:793/836
        }
    }
}

void launch_iterate(const IterParams &p, unsigned int lanes, cudaStream_t s)
{
    if (p.n_jobs == 0) return;
    unsigned long long want = p.n_jobs < lanes ? p.n_jobs : lanes;
    // small launches: narrow blocks so the jobs spread over the SMs
    const unsigned int block = want >= 148ull * 128ull ? 128u : 32u;
    const unsigned int grid = (unsigned int)((want + block - 1) / block);
    iterate_kernel<<<grid, block, 0, s>>>(p);
    ++g_launches;
}

// ---------------------------------------------------------------------------------------------
// Runtime::reset (lib.rs:682-699)
// ---------------------------------------------------------------------------------------------
__global__ void reset_kernel(unsigned long long *fast, ulonglong2 *rec, Scalars *scal, size_t npix)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        fast[i] = FAST_RESET;                                   // count 0 (lib.rs:687)
        rec[i] = make_ulonglong2(0ull, REC_HI_RESET);           // steps 0.0 (lib.rs:690), zbuf -1.0 (lib.rs:693)
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scal->nan_sink = 0ull; scal->max = 0u;                  // lib.rs:694
        scal->zmax_key = ZKEY_ZERO; scal->zmin_key = ZKEY_FLT_MAX; scal->pad = 0u;
    }
}
void launch_reset(unsigned long long *fast, ulonglong2 *rec, Scalars *scal, size_t npix, cudaStream_t s)
{
    const unsigned int block = 256;
    size_t g = (npix + block - 1) / block;
    const unsigned int grid = (unsigned int)(g > 148u * 16u ? 148u * 16u : (g ? g : 1));
    reset_kernel<<<grid, block, 0, s>>>(fast, rec, scal, npix);
    ++g_launches;
}

// ---------------------------------------------------------------------------------------------
// Runtime.max (lib.rs:643, 813-815) as a reduction: counts only grow, so the running max the
// reference tracks equals the max over the final counts.  Also folds the Depth min/max
// (lib.rs:877-882).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pixel_count(const unsigned long long *fast, const Scalars *scal, size_t i)
{
    uint32_t c = (uint32_t)fast[i];
    if (i == 0) c += (uint32_t)scal->nan_sink;                  // wrapping u32, like lib.rs:811 in release
    return c;
}
__global__ void max_kernel(const unsigned long long *fast, const ulonglong2 *rec, Scalars *scal, size_t pix0, size_t npix)
{
    uint32_t m = 0, zmx = ZKEY_ZERO, zmn = ZKEY_FLT_MAX;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        const size_t p = pix0 + i;
        const uint32_t c = pixel_count(fast, scal, p);
        m = c > m ? c : m;
        const uint32_t k = (uint32_t)(rec[p].y >> 32);
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
void launch_max(const unsigned long long *fast, const ulonglong2 *rec, Scalars *scal, size_t pix0, size_t npix, cudaStream_t s)
{
    if (npix == 0) return;
    const unsigned int block = 256;
    size_t g = (npix + block - 1) / block;
    const unsigned int grid = (unsigned int)(g > 148u * 8u ? 148u * 8u : g);
    max_kernel<<<grid, block, 0, s>>>(fast, rec, scal, pix0, npix);
    ++g_launches;
}

// ---------------------------------------------------------------------------------------------
// colorize() (lib.rs:841-904)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint16_t sat_u16(double v)          // Rust `as u16`: saturating, NaN -> 0
{
    const unsigned int u = __double2uint_rz(v);
    return (uint16_t)(u > 65535u ? 65535u : u);
}
__device__ __forceinline__ uint16_t sat_u16f(float v)
{
    const unsigned int u = __float2uint_rz(v);
    return (uint16_t)(u > 65535u ? 65535u : u);
}

__global__ void __launch_bounds__(256)
colorize_kernel(const __grid_constant__ ColorParams C, const unsigned long long *__restrict__ fast,
                const ulonglong2 *__restrict__ rec, const Scalars *__restrict__ scal,
                uint16_t *__restrict__ out16, float *__restrict__ out32)
{
    __shared__ double s_lnmax;
    __shared__ float s_zmax, s_zmin;
    if (threadIdx.x == 0) {
        s_lnmax = log((double)(uint32_t)(scal->max + 1u));      // f64::from(runtime.max + 1), lib.rs:860
        s_zmax = __uint_as_float(zbits_from_key(scal->zmax_key));
        s_zmin = __uint_as_float(zbits_from_key(scal->zmin_key));
    }
    __syncthreads();
    const size_t pix0 = (size_t)C.row0 * C.W, npix = (size_t)C.rows * C.W;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        const size_t p = pix0 + i;
        ushort4 px;
        float4 fx;
        if (C.render_kind == 0u) {                              // RenderKind::Gas, lib.rs:853-874
            // Palette::interpolate(steps), lib.rs:442-472
            double v = __longlong_as_double((long long)rec[p].x);
            if (v < 0.) v = 0.; else if (v >= 1.) v = 0.999999;
            v = __dmul_rn(v, C.pal_len);
            const double fl = floor(v);
            unsigned int n = __double2uint_rz(fl);              // `as usize`, NaN -> 0
            if (n > C.palette_len - 1u) n = C.palette_len - 1u;
            const double t = fmod(v, 1.);                       // `value % 1.`, lib.rs:454
            const double t1 = __dsub_rn(1.0, t);
            const double r = __dsqrt_rn(__dadd_rn(__dmul_rn(C.pal[n + 1][0], t), __dmul_rn(C.pal[n][0], t1)));
            const double g = __dsqrt_rn(__dadd_rn(__dmul_rn(C.pal[n + 1][1], t), __dmul_rn(C.pal[n][1], t1)));
            const double b = __dsqrt_rn(__dadd_rn(__dmul_rn(C.pal[n + 1][2], t), __dmul_rn(C.pal[n][2], t1)));
            const uint32_t cnt = pixel_count(fast, scal, p);
            const double factor = __ddiv_rn(log((double)(uint32_t)(cnt + 1u)), s_lnmax);   // lib.rs:860
            const double vr = __dmul_rn(__dadd_rn(__dmul_rn(r, factor), C.bright_offset), C.bright_factor);
            const double vg = __dmul_rn(__dadd_rn(__dmul_rn(g, factor), C.bright_offset), C.bright_factor);
            const double vb = __dmul_rn(__dadd_rn(__dmul_rn(b, factor), C.bright_offset), C.bright_factor);
            px.x = sat_u16(__dmul_rn(vr, 65535.));              // lib.rs:862-864
            px.y = sat_u16(__dmul_rn(vg, 65535.));
            px.z = sat_u16(__dmul_rn(vb, 65535.));
            px.w = C.transparent ? sat_u16(__dmul_rn(factor, 65535.)) : (uint16_t)65535u;  // lib.rs:865-869
            fx = make_float4((float)vr, (float)vg, (float)vb, C.transparent ? (float)factor : 1.0f);
        } else {                                                // RenderKind::Depth, lib.rs:875-900
            const uint32_t k = (uint32_t)(rec[p].y >> 32);
            float zz;
            if (k == ZKEY_SENTINEL) zz = 0.0f;                  // z == -1.0, lib.rs:889-890
            else zz = __fdiv_rn(__fsub_rn(__uint_as_float(zbits_from_key(k)), s_zmin), __fsub_rn(s_zmax, s_zmin));
            const uint16_t g = sat_u16f(__fmul_rn(zz, 65535.0f));                          // lib.rs:895
            px.x = g; px.y = g; px.z = g; px.w = 65535u;
            fx = make_float4(zz, zz, zz, 1.0f);
        }
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

// ---------------------------------------------------------------------------------------------
// layout conversion to / from the reference's three textures (lib.rs:633-639)
// ---------------------------------------------------------------------------------------------
__global__ void unpack_kernel(const unsigned long long *fast, const ulonglong2 *rec, const Scalars *scal, size_t npix,
                              uint32_t *count, double *steps, float *zbuf)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        if (count) count[i] = pixel_count(fast, scal, i);
        const ulonglong2 r = rec[i];
        if (steps) steps[i] = __longlong_as_double((long long)r.x);
        if (zbuf) zbuf[i] = __uint_as_float(zbits_from_key((uint32_t)(r.y >> 32)));
    }
}
void launch_unpack(const unsigned long long *fast, const ulonglong2 *rec, const Scalars *scal, size_t npix,
                   uint32_t *count, double *steps, float *zbuf, cudaStream_t s)
{
    if (npix == 0) return;
    const unsigned int block = 256;
    size_t g = (npix + block - 1) / block;
    unpack_kernel<<<(unsigned int)(g > 148u * 16u ? 148u * 16u : g), block, 0, s>>>(fast, rec, scal, npix, count, steps, zbuf);
    ++g_launches;
}

__global__ void pack_kernel(unsigned long long *fast, ulonglong2 *rec, Scalars *scal, size_t npix,
                            const uint32_t *count, const double *steps, const float *zbuf)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        const float z = zbuf[i] + 0.0f;
        uint32_t k = zkey_of(z);
        if (!(z > -1.0f)) k = ZKEY_SENTINEL;                    // untouched (or invalid) pixels
        const uint32_t hint = k == ZKEY_SENTINEL ? ZKEY_SENTINEL + 1u : k;
        fast[i] = ((unsigned long long)hint << 32) | count[i];
        // uploaded records predate every future job: they keep all z ties (job key 0)
        rec[i] = make_ulonglong2((unsigned long long)__double_as_longlong(steps[i]), ((unsigned long long)k << 32) | 0xFFFFFFFFull);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scal->nan_sink = 0ull; scal->max = 0u; scal->zmax_key = ZKEY_ZERO; scal->zmin_key = ZKEY_FLT_MAX; scal->pad = 0u;
    }
}
void launch_pack(unsigned long long *fast, ulonglong2 *rec, Scalars *scal, size_t npix,
                 const uint32_t *count, const double *steps, const float *zbuf, cudaStream_t s)
{
    const unsigned int block = 256;
    size_t g = (npix + block - 1) / block;
    pack_kernel<<<(unsigned int)(g > 148u * 16u ? 148u * 16u : (g ? g : 1)), block, 0, s>>>(fast, rec, scal, npix, count, steps, zbuf);
    ++g_launches;
}

// ---------------------------------------------------------------------------------------------
// Runtime::merge (lib.rs:708-738): count +=, `other` wins iff its z is strictly greater.
// ---------------------------------------------------------------------------------------------
__global__ void merge_kernel(unsigned long long *dfast, ulonglong2 *drec, Scalars *dscal,
                             const unsigned long long *sfast, const ulonglong2 *srec, const Scalars *sscal, size_t npix)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        const uint32_t c = (uint32_t)dfast[i] + (uint32_t)sfast[i];                 // lib.rs:719
        ulonglong2 d = drec[i];
        const ulonglong2 o = srec[i];
        if ((uint32_t)(o.y >> 32) > (uint32_t)(d.y >> 32)) { d = o; drec[i] = d; }  // lib.rs:728-735
        const uint32_t k = (uint32_t)(d.y >> 32);
        const uint32_t hint = k == ZKEY_SENTINEL ? ZKEY_SENTINEL + 1u : k;
        dfast[i] = ((unsigned long long)hint << 32) | c;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) dscal->nan_sink += sscal->nan_sink;
}
void launch_merge(unsigned long long *dfast, ulonglong2 *drec, Scalars *dscal,
                  const unsigned long long *sfast, const ulonglong2 *srec, const Scalars *sscal,
                  size_t npix, cudaStream_t s)
{
    const unsigned int block = 256;
    size_t g = (npix + block - 1) / block;
    merge_kernel<<<(unsigned int)(g > 148u * 16u ? 148u * 16u : (g ? g : 1)), block, 0, s>>>(dfast, drec, dscal, sfast, srec, sscal, npix);
    ++g_launches;
}

// ---------------------------------------------------------------------------------------------
// All-ranks merge of one row stripe, reading every peer's accumulators directly over
// NVLink (CUDA IPC mappings).  Deterministic form of Runtime::merge: counts add, the
// record with the greatest (zkey, ~job) wins — identical to rendering every job on one
// Runtime, whatever the number of ranks.
// ---------------------------------------------------------------------------------------------
__global__ void merge_peers_kernel(unsigned long long *dfast, ulonglong2 *drec, Scalars *dscal,
                                   const __grid_constant__ PeerList peers, size_t pix0, size_t npix)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += stride) {
        const size_t p = pix0 + i;
        uint32_t c = (uint32_t)dfast[p];
        ulonglong2 d = drec[p];
        for (int r = 0; r < peers.n; ++r) {
            c += (uint32_t)__ldcv(peers.fast[r] + p);
            const unsigned long long oy = __ldcv(&peers.rec[r][p].y);
            if (oy > d.y) { d.y = oy; d.x = __ldcv(&peers.rec[r][p].x); }
        }
        const uint32_t k = (uint32_t)(d.y >> 32);
        const uint32_t hint = k == ZKEY_SENTINEL ? ZKEY_SENTINEL + 1u : k;
        drec[p] = d;
        dfast[p] = ((unsigned long long)hint << 32) | c;
    }
    if (pix0 == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long s = dscal->nan_sink;
        for (int r = 0; r < peers.n; ++r) s += *((volatile const unsigned long long *)&peers.scal[r]->nan_sink);
        dscal->nan_sink = s;
    }
}
void launch_merge_peers(unsigned long long *dfast, ulonglong2 *drec, Scalars *dscal, const PeerList &peers,
                        size_t pix0, size_t npix, cudaStream_t s)
{
    if (npix == 0) return;
    const unsigned int block = 256;
    size_t g = (npix + block - 1) / block;
    merge_peers_kernel<<<(unsigned int)(g > 148u * 16u ? 148u * 16u : g), block, 0, s>>>(dfast, drec, dscal, peers, pix0, npix);
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
