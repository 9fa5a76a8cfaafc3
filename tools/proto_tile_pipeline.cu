// PROTOTYPE v2 (not part of the product): the "bin-then-tile" scatter of the north star / VERDICT r1 item 3,
// with the two fixes round 1's prototype (tools/proto_binning.cu) showed to be necessary:
//   producer : real trajectory arithmetic (poisson-saturne, 2048x2048), TWO trajectories per thread (the
//              arithmetic-bound form, profiles/r2_iterate_variants.md), no L2 atomic per hit.  Every WIN iterations
//              a block counting-sorts its hit records by image tile IN SHARED MEMORY (two shared-memory atomics per
//              record) and appends each tile's run to that tile's global queue as one COALESCED copy (one global
//              atomic per non-empty tile per window to reserve the space).
//   consumer : blocks take (tile, segment) work items and accumulate, per pixel of the tile, count (u32 ATOMS.ADD)
//              and z max (u32 ATOMS.MAX, seeded with the pixel's current hint) in shared memory, reading the queue
//              with 16-byte loads, 4 in flight per thread; records that raise the max are the depth-test candidates
//              (counted here; the product would replay them from checkpoints).  Touched pixels are flushed with one
//              count add + one hint max each — "privatised count tiles in shared memory before a global atomicAdd".
// Measured: producer alone, consumer alone, and both overlapped chunk by chunk on two streams.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/proto_tile_pipeline.bin tools/proto_tile_pipeline.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Params {
    double c[3][10], m[3][3], ccx, ccy, ccz, cv, sv, sam, ws, half_h;
    unsigned int W, H;
};

constexpr int THREADS = 256, NTJ = 2, LANES = THREADS * NTJ;   // 512 trajectories per block
constexpr int WIN = 8;                                          // iterations per window -> 4096 records per block-window
constexpr int RECS = LANES * WIN;
constexpr int TSHIFT = 13, TPIX = 1 << TSHIFT;                  // tile = 8192 consecutive pixels (4 rows of 2048): 512 tiles
constexpr int NTILES = 512;

__device__ __forceinline__ double sum10(const double (&c)[10], double x, double y, double z, double xx, double xy, double xz, double yy, double yz, double zz)
{
    double s = c[0];
    s = __dadd_rn(s, __dmul_rn(x, c[1])); s = __dadd_rn(s, __dmul_rn(xx, c[2])); s = __dadd_rn(s, __dmul_rn(xy, c[3]));
    s = __dadd_rn(s, __dmul_rn(xz, c[4])); s = __dadd_rn(s, __dmul_rn(y, c[5])); s = __dadd_rn(s, __dmul_rn(yy, c[6]));
    s = __dadd_rn(s, __dmul_rn(yz, c[7])); s = __dadd_rn(s, __dmul_rn(z, c[8])); s = __dadd_rn(s, __dmul_rn(zz, c[9]));
    return s;
}
__device__ __forceinline__ void next_point(const Params &P, double &x, double &y, double &z)
{
    const double xx = __dmul_rn(x, x), xy = __dmul_rn(x, y), xz = __dmul_rn(x, z), yy = __dmul_rn(y, y), yz = __dmul_rn(y, z), zz = __dmul_rn(z, z);
    const double nx = sum10(P.c[0], x, y, z, xx, xy, xz, yy, yz, zz), ny = sum10(P.c[1], x, y, z, xx, xy, xz, yy, yz, zz),
                 nz = sum10(P.c[2], x, y, z, xx, xy, xz, yy, yz, zz);
    x = nx; y = ny; z = nz;
}
__device__ __forceinline__ unsigned long long splitmix(unsigned long long seed, unsigned long long n)
{
    unsigned long long z = seed + (n + 1ull) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct Queues {
    unsigned long long *rec;      // NTILES x cap records: [63:32] zkey | [31:23] tile | [22:10] pixel in tile | [9:0] spare
    unsigned short *id;           // NTILES x cap: lane (9 b) << 3 | iteration in window (3 b) — who produced the record (for the replay)
    unsigned long long *tail;     // per tile: records appended so far
    unsigned long long cap;       // capacity of one tile's queue
};

// mode 0: arithmetic + staging only; 1: + counting sort in shared memory; 2: + coalesced append to the tile queues
__global__ void __maxnreg__(96)      // 2 blocks/SM = 48 K registers: leaves room for a consumer block
producer(const __grid_constant__ Params P, Queues Q, double *state, unsigned int first_iter, unsigned int n_iter, int mode)
{
    extern __shared__ unsigned long long sm[];
    unsigned long long *stage = sm;                                   // RECS records in production order
    unsigned long long *sorted = stage + RECS;                        // RECS records in tile order
    // (the product would carry a 2-byte producer id per record in a parallel queue for the replay: +25 % of queue
    //  traffic, 8 KB more shared memory; left out here so that two producer blocks and a consumer block share an SM)
    unsigned int *hist = (unsigned int *)(sorted + RECS);             // NTILES counts
    unsigned int *offs = hist + NTILES;                               // NTILES running offsets (block-local)
    unsigned long long *gbase = (unsigned long long *)(offs + NTILES);   // NTILES reserved positions in the global queues
    __shared__ unsigned int s_warp_tot[THREADS / 32];
    const unsigned int tid = threadIdx.x;
    double x[NTJ], y[NTJ], z[NTJ];
    for (int k = 0; k < NTJ; ++k) {
        const unsigned long long job = ((unsigned long long)blockIdx.x * THREADS + tid) * NTJ + k;
        if (first_iter == 0) {
            x[k] = __dmul_rn(__dmul_rn((double)(splitmix(1234, 3 * job) >> 11), 0x1.0p-53), 0.1);
            y[k] = __dmul_rn(__dmul_rn((double)(splitmix(1234, 3 * job + 1) >> 11), 0x1.0p-53), 0.1);
            z[k] = __dmul_rn(__dmul_rn((double)(splitmix(1234, 3 * job + 2) >> 11), 0x1.0p-53), 0.1);
            for (int w = 0; w < 1000; ++w) next_point(P, x[k], y[k], z[k]);
        } else {
            x[k] = state[3 * job]; y[k] = state[3 * job + 1]; z[k] = state[3 * job + 2];
        }
    }
    for (unsigned int win = 0; win < n_iter / WIN; ++win) {
        for (int t = tid; t < NTILES; t += THREADS) hist[t] = 0;
        __syncthreads();
        for (int i = 0; i < WIN; ++i) {
#pragma unroll
            for (int k = 0; k < NTJ; ++k) {
                next_point(P, x[k], y[k], z[k]);
                const double sx = __dadd_rn(__dadd_rn(__dmul_rn(P.m[0][0], x[k]), __dmul_rn(P.m[0][1], y[k])), __dmul_rn(P.m[0][2], z[k]));
                const double sy = __dadd_rn(__dadd_rn(__dmul_rn(P.m[1][0], x[k]), __dmul_rn(P.m[1][1], y[k])), __dmul_rn(P.m[1][2], z[k]));
                const double sz = __dadd_rn(__dadd_rn(__dmul_rn(P.m[2][0], x[k]), __dmul_rn(P.m[2][1], y[k])), __dmul_rn(P.m[2][2], z[k]));
                const double a = __dadd_rn(sx, P.ccx), b = __dadd_rn(sz, P.ccy);
                const double x2 = __dadd_rn(__dmul_rn(a, P.cv), __dmul_rn(b, P.sv));
                const double z2 = __dsub_rn(__dmul_rn(a, P.sv), __dmul_rn(b, P.cv));
                const double fi = __dmul_rn(__dsub_rn(P.sam, x2), P.ws);
                const double fj = __dsub_rn(P.half_h, __dmul_rn(__dadd_rn(sy, P.ccz), P.ws));
                const unsigned int ii = (unsigned int)__double2int_rd(fi), jj = (unsigned int)__double2int_rd(fj);
                unsigned long long rec = 0ull;                      // 0 = no hit (zkey is never 0)
                if (ii < P.W && jj < P.H) {
                    const unsigned int bits = __float_as_uint(__double2float_rn(z2) + 0.0f);
                    const unsigned int key = (bits & 0x80000000u) ? ~bits : (bits | 0x80000000u);
                    const unsigned int idx = jj * P.W + ii;         // tile = idx >> 13, pixel in tile = idx & 8191: bits [31:10] of the low word
                    rec = ((unsigned long long)key << 32) | ((unsigned long long)idx << 10);
                    if (mode >= 1) atomicAdd(&hist[idx >> TSHIFT], 1u);
                }
                stage[(i * NTJ + k) * THREADS + tid] = rec;
            }
        }
        if (mode == 0) continue;
        __syncthreads();
        // exclusive scan of hist[512] -> offs (2 tiles per thread), and the reservation in the global queues
        {
            const unsigned int a = hist[2 * tid], b = hist[2 * tid + 1];
            unsigned int v = a + b;
            const unsigned int lane = tid & 31, warp = tid >> 5;
            unsigned int inc = v;
            for (int o = 1; o < 32; o <<= 1) { const unsigned int n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += n; }
            if (lane == 31) s_warp_tot[warp] = inc;
            __syncthreads();
            unsigned int wbase = 0;
            for (int w = 0; w < (int)warp; ++w) wbase += s_warp_tot[w];
            const unsigned int ex = wbase + inc - v;
            offs[2 * tid] = ex; offs[2 * tid + 1] = ex + a;
            if (mode >= 2) {
                gbase[2 * tid] = a ? atomicAdd(&Q.tail[2 * tid], (unsigned long long)a) - ex : 0ull;                // so that global pos = gbase[t] + sorted pos
                gbase[2 * tid + 1] = b ? atomicAdd(&Q.tail[2 * tid + 1], (unsigned long long)b) - (ex + a) : 0ull;
            }
        }
        __syncthreads();
        // scatter into tile order (second shared-memory atomic per record: its final position)
        for (int r = tid; r < RECS; r += THREADS) {
            const unsigned long long rec = stage[r];
            if (rec) {
                const unsigned int t = (unsigned int)(rec >> 10) >> TSHIFT & (NTILES - 1);
                const unsigned int pos = atomicAdd(&offs[t], 1u);
                sorted[pos] = rec;
            }
        }
        __syncthreads();
        if (mode >= 2) {
            // coalesced copy-out: consecutive threads write consecutive records of (mostly) the same tile's run
            const unsigned int total = offs[NTILES - 1];            // after the scatter, offs[t] = end of tile t's run
            for (unsigned int p = tid; p < total; p += THREADS) {
                const unsigned long long rec = sorted[p];
                const unsigned int t = (unsigned int)(rec >> 10) >> TSHIFT & (NTILES - 1);
                const unsigned long long g = gbase[t] + p;
                if (g < Q.cap) Q.rec[(size_t)t * Q.cap + g] = rec;
            }
        }
        __syncthreads();
    }
    for (int k = 0; k < NTJ; ++k) {
        const unsigned long long job = ((unsigned long long)blockIdx.x * THREADS + tid) * NTJ + k;
        state[3 * job] = x[k]; state[3 * job + 1] = y[k]; state[3 * job + 2] = z[k];
    }
}

// consumer: one block per (tile, segment of the tile's queue)
constexpr unsigned int SEG = 1u << 17;      // records per work item
constexpr int CTHREADS = 256;
__global__ void __maxnreg__(48)
consumer(Queues Q, const unsigned long long *first_of_tile, const unsigned int *work_tile, const unsigned int *work_seg,
         unsigned int *gcount, unsigned int *ghint, unsigned long long *n_candidates)
{
    extern __shared__ unsigned int csm[];
    unsigned int *cnt = csm, *zmax = csm + TPIX;
    const unsigned int t = work_tile[blockIdx.x];
    const unsigned long long lo = first_of_tile[t] + (unsigned long long)work_seg[blockIdx.x] * SEG;
    unsigned long long hi = lo + SEG;
    const unsigned long long end = Q.tail[t] < Q.cap ? Q.tail[t] : Q.cap;
    if (hi > end) hi = end;
    if (lo >= hi) return;                                            // empty work item (the list is sized for the worst case)
    for (int p = threadIdx.x; p < TPIX; p += blockDim.x) { cnt[p] = 0u; zmax[p] = ghint[(size_t)t * TPIX + p]; }
    __syncthreads();
    const ulonglong2 *q = reinterpret_cast<const ulonglong2 *>(Q.rec + (size_t)t * Q.cap);
    unsigned int cand = 0;
    auto one = [&](unsigned long long rec) {
        if (!rec) return;
        const unsigned int pix = (unsigned int)(rec >> 10) & (TPIX - 1), key = (unsigned int)(rec >> 32);
        atomicAdd(&cnt[pix], 1u);
        if (key > zmax[pix]) { const unsigned int old = atomicMax(&zmax[pix], key); cand += key > old; }   // filtered by a plain read first
    };
    const unsigned long long p0 = (lo + 1) / 2, p1 = hi / 2;        // 16-byte pairs fully inside [lo, hi)
    if ((lo & 1) && threadIdx.x == 0 && lo < hi) one(Q.rec[(size_t)t * Q.cap + lo]);
    if ((hi & 1) && threadIdx.x == 0 && hi > lo) one(Q.rec[(size_t)t * Q.cap + hi - 1]);
    unsigned long long p = p0 + threadIdx.x;
    for (; p + 3ull * blockDim.x < p1; p += 4ull * blockDim.x) {     // 4 independent 16-byte loads in flight per thread
        const ulonglong2 a = __ldcs(q + p), b = __ldcs(q + p + blockDim.x), c = __ldcs(q + p + 2 * blockDim.x), d = __ldcs(q + p + 3 * blockDim.x);
        one(a.x); one(a.y); one(b.x); one(b.y); one(c.x); one(c.y); one(d.x); one(d.y);
    }
    for (; p < p1; p += blockDim.x) { const ulonglong2 a = __ldcs(q + p); one(a.x); one(a.y); }
    __syncthreads();
    for (int px = threadIdx.x; px < TPIX; px += blockDim.x) {
        if (cnt[px]) {
            atomicAdd(&gcount[(size_t)t * TPIX + px], cnt[px]);
            atomicMax(&ghint[(size_t)t * TPIX + px], zmax[px]);
        }
    }
    for (int o = 16; o > 0; o >>= 1) cand += __shfl_xor_sync(0xffffffffu, cand, o);
    if ((threadIdx.x & 31) == 0 && cand) atomicAdd(n_candidates, (unsigned long long)cand);
}

int main(int argc, char **argv)
{
    const unsigned int W = 2048, H = 2048;
    const int blocks_per_sm = 2;
    const int grid = 148 * blocks_per_sm;
    const unsigned long long njobs = (unsigned long long)grid * LANES;
    const unsigned int chunk = argc > 1 ? atoi(argv[1]) : 256;                 // iterations per chunk (multiple of WIN)
    const unsigned int iters = (unsigned int)(1000000000ull / njobs) / chunk * chunk;
    const unsigned int n_chunks = iters / chunk;
    Params P;
    const double cx[10] = {0.021, 1.182, -1.183, 0.128, -1.12, -0.641, -1.152, -0.834, -0.97, 0.722};
    const double cy[10] = {0.243038, -0.825, -1.2, -0.835443, -0.835443, -0.364557, 0.458, 0.622785, -0.394937, -1.032911};
    const double cz[10] = {-0.455696, 0.673, 0.915, -0.258228, -0.495, -0.264, -0.432, -0.416, -0.877, -0.3};
    for (int i = 0; i < 10; ++i) { P.c[0][i] = cx[i]; P.c[1][i] = cy[i]; P.c[2][i] = cz[i]; }
    const double ax = 0.304289493528802, ay = 0.760492682863655, az = 0.573636455813981, rot = 1.78268191887446;
    const double c = cos(rot), c1 = 1. - c, s = sin(rot);
    P.m[0][0] = c + ax * ax * c1; P.m[0][1] = ax * ay * c1 - az * s; P.m[0][2] = ax * az * c1 + ay * s;
    P.m[1][0] = ay * ax * c1 + az * s; P.m[1][1] = c + ay * ay * c1; P.m[1][2] = ay * az * c1 - ax * s;
    P.m[2][0] = az * ax * c1 - ay * s; P.m[2][1] = az * ay * c1 + ax * s; P.m[2][2] = c + az * az * c1;
    P.ccx = -0.005; P.ccy = 0.262; P.ccz = -0.366 + 0.12; P.cv = 1.0; P.sv = 0.0; P.sam = 0.5; P.ws = 2048.0; P.half_h = 1024.0;
    P.W = W; P.H = H;

    const unsigned long long per_chunk = njobs * chunk;
    // the densest 8192-pixel tile gets < 3 % of the hits; 16x the mean = 3.1 % of a chunk per tile, two chunk buffers
    Queues Q[2];
    const unsigned long long cap = per_chunk / NTILES * 16;
    for (int b = 0; b < 2; ++b) {
        Q[b].cap = cap;
        CK(cudaMalloc(&Q[b].rec, (size_t)NTILES * cap * 8)); CK(cudaMalloc(&Q[b].id, (size_t)NTILES * cap * 2)); CK(cudaMalloc(&Q[b].tail, NTILES * 8));
    }
    double *state; CK(cudaMalloc(&state, njobs * 24));
    unsigned int *gcount, *ghint; unsigned long long *ncand;
    CK(cudaMalloc(&gcount, (size_t)W * H * 4)); CK(cudaMalloc(&ghint, (size_t)W * H * 4)); CK(cudaMalloc(&ncand, 8));
    const size_t psm = (size_t)RECS * 8 * 2 + NTILES * 4 * 2 + NTILES * 8;
    CK(cudaFuncSetAttribute(producer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
    const size_t csmb = (size_t)TPIX * 8;
    CK(cudaFuncSetAttribute(consumer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmb));
    printf("jobs %llu (2 per thread, %d threads/block, %d blocks/SM), iterations/job %u in %u chunks of %u, window %d; producer smem %zu B/block, consumer %zu B/block; queue memory 2 x %.2f GB\n",
           njobs, THREADS, blocks_per_sm, iters, n_chunks, chunk, WIN, psm, csmb, (double)NTILES * cap * 10 / 1e9);
    cudaStream_t sp, sc; CK(cudaStreamCreate(&sp)); CK(cudaStreamCreate(&sc));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<cudaEvent_t> produced(n_chunks), consumed(n_chunks);
    for (auto &e : produced) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto &e : consumed) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    float ms;
    // ---- producer alone, three depths of the scatter ----
    for (int mode = 0; mode <= 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0, sp);
            for (unsigned int cidx = 0; cidx < n_chunks; ++cidx) {
                CK(cudaMemsetAsync(Q[0].tail, 0, NTILES * 8, sp));
                producer<<<grid, THREADS, psm, sp>>>(P, Q[0], state, cidx * chunk, chunk, mode);
            }
            cudaEventRecord(e1, sp);
            CK(cudaDeviceSynchronize());
            cudaEventElapsedTime(&ms, e0, e1);
        }
        const char *what[3] = {"arithmetic + staging in shared memory", "+ counting sort by tile in shared memory", "+ coalesced append to the tile queues"};
        printf("producer, %s: %.3f ms  %.2f G recorded it/s\n", what[mode], ms, (double)njobs * iters / ms / 1e6);
    }
    // ---- consumer alone on the last chunk's queues ----
    std::vector<unsigned long long> tails(NTILES), firsts(NTILES, 0);
    CK(cudaMemcpy(tails.data(), Q[0].tail, NTILES * 8, cudaMemcpyDeviceToHost));
    unsigned long long sum = 0, mx = 0, dropped = 0;
    std::vector<unsigned int> wt, ws;
    for (int t = 0; t < NTILES; ++t) {
        sum += tails[t]; if (tails[t] > mx) mx = tails[t];
        const unsigned long long have = tails[t] < cap ? tails[t] : cap;
        dropped += tails[t] - have;
        for (unsigned long long f = 0, sgi = 0; f < have; f += SEG, ++sgi) { wt.push_back(t); ws.push_back((unsigned int)sgi); }
    }
    unsigned long long *dfirst; unsigned int *dwt, *dws;
    CK(cudaMalloc(&dfirst, NTILES * 8)); CK(cudaMemcpy(dfirst, firsts.data(), NTILES * 8, cudaMemcpyHostToDevice));
    // work list for the worst case: every tile with its maximum number of segments (empty items return at once)
    std::vector<unsigned int> awt, aws;
    for (int t = 0; t < NTILES; ++t) for (unsigned long long f = 0, sgi = 0; f < cap; f += SEG, ++sgi) { awt.push_back(t); aws.push_back((unsigned int)sgi); }
    CK(cudaMalloc(&dwt, awt.size() * 4)); CK(cudaMalloc(&dws, aws.size() * 4));
    CK(cudaMemcpy(dwt, wt.data(), wt.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dws, ws.data(), ws.size() * 4, cudaMemcpyHostToDevice));
    printf("last chunk: records %llu (%.4f of its iterations), densest tile %.4f of them, over capacity %.5f, consumer work items %zu\n",
           sum, (double)sum / per_chunk, (double)mx / sum, (double)dropped / sum, wt.size());
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(gcount, 0, (size_t)W * H * 4)); CK(cudaMemset(ghint, 0, (size_t)W * H * 4)); CK(cudaMemset(ncand, 0, 8));
        cudaEventRecord(e0, sc);
        consumer<<<(unsigned int)wt.size(), CTHREADS, csmb, sc>>>(Q[0], dfirst, dwt, dws, gcount, ghint, ncand);
        cudaEventRecord(e1, sc);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
    }
    unsigned long long hc = 0; CK(cudaMemcpy(&hc, ncand, 8, cudaMemcpyDeviceToHost));
    printf("consumer alone (one chunk, empty image): %.3f ms for %llu records  (%.2f G records/s, %.1f GB/s of queue reads), depth-test candidates %llu (%.4f)\n",
           ms, sum - dropped, (double)(sum - dropped) / ms / 1e6, (double)(sum - dropped) * 8 / ms / 1e6, hc, (double)hc / (sum - dropped));
    // ---- the whole frame: producer of chunk c+1 overlapped with the consumer of chunk c (two streams, two queue buffers) ----
    CK(cudaMemcpy(dwt, awt.data(), awt.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dws, aws.data(), aws.size() * 4, cudaMemcpyHostToDevice));
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaMemset(gcount, 0, (size_t)W * H * 4)); CK(cudaMemset(ghint, 0, (size_t)W * H * 4)); CK(cudaMemset(ncand, 0, 8));
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0, sp);
        for (unsigned int cidx = 0; cidx < n_chunks; ++cidx) {
            const int b = cidx & 1;
            if (cidx >= 2) CK(cudaStreamWaitEvent(sp, consumed[cidx - 2], 0));      // this buffer's previous chunk has been consumed
            CK(cudaMemsetAsync(Q[b].tail, 0, NTILES * 8, sp));
            producer<<<grid, THREADS, psm, sp>>>(P, Q[b], state, cidx * chunk, chunk, 2);
            CK(cudaEventRecord(produced[cidx], sp));
            CK(cudaStreamWaitEvent(sc, produced[cidx], 0));
            consumer<<<(unsigned int)awt.size(), CTHREADS, csmb, sc>>>(Q[b], dfirst, dwt, dws, gcount, ghint, ncand);
            CK(cudaEventRecord(consumed[cidx], sc));
        }
        CK(cudaStreamWaitEvent(sp, consumed[n_chunks - 1], 0));
        cudaEventRecord(e1, sp);
        CK(cudaDeviceSynchronize());
        cudaEventElapsedTime(&ms, e0, e1);
    }
    CK(cudaMemcpy(&hc, ncand, 8, cudaMemcpyDeviceToHost));
    std::vector<unsigned int> hcount((size_t)W * H);
    CK(cudaMemcpy(hcount.data(), gcount, (size_t)W * H * 4, cudaMemcpyDeviceToHost));
    unsigned long long tot = 0; for (unsigned int v : hcount) tot += v;
    printf("whole frame, producer || consumer: %.3f ms  %.2f G recorded it/s; counted %llu of %llu iterations; depth-test candidates %llu (%.4f of the hits)\n",
           ms, (double)njobs * iters / ms / 1e6, tot, njobs * iters, hc, (double)hc / tot);
    return 0;
}
