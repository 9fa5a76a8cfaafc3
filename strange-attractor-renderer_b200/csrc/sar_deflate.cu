// sar_deflate.cu — device side of the compressed PNG writer (src/bin/main.rs:78-89): scanline filter, one deflate block
// per warp (sar_deflate.cuh holds everything that decides bits), compaction of the blocks into one zlib stream, and
// the partial CRC-32 / Adler-32 sums the host folds into the IDAT chunk.
#include "sar_deflate.cuh"
#include "sar_device.cuh"

namespace sar {

__device__ __forceinline__ uint32_t narrow_u16_d(uint32_t c) { return (c + 128u) / 257u; }   // image 0.25: u16 -> u8

// sample bytes of one pixel as the PNG holds them (16-bit samples most significant byte first)
__device__ __forceinline__ void pixel_bytes(const ushort4 v, bool wide, unsigned int nch, uint8_t *b)
{
    const uint16_t c[4] = {v.x, v.y, v.z, v.w};
    for (unsigned int ch = 0; ch < nch; ++ch) {
        if (wide) { b[2 * ch] = (uint8_t)(c[ch] >> 8); b[2 * ch + 1] = (uint8_t)c[ch]; }
        else b[ch] = (uint8_t)narrow_u16_d(c[ch]);
    }
}

// Filtered scanline stream: every row is [1][Sub(x)], Sub(x) = Raw(x) - Raw(x - bpp) mod 256 (PNG spec §9.2, filter type 1;
// the reference's encoder picks a filter per row adaptively — any choice decodes to the same pixels).  With Sub the
// untouched part of a frame (80 % of it, one constant colour) becomes zeros, which the run-length matcher eats.
__global__ void png_filter_kernel(const ushort4 *__restrict__ img, uint8_t *__restrict__ raw, unsigned int W, unsigned int H,
                                  unsigned int fmt, size_t raw_row)
{
    const size_t npix = (size_t)W * H;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const bool wide = fmt == PIX_RGBA16 || fmt == PIX_RGB16, alpha = fmt == PIX_RGBA16 || fmt == PIX_RGBA8;
    const unsigned int nch = alpha ? 4u : 3u, bpp = nch * (wide ? 2u : 1u);
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += stride) {
        const unsigned int y = (unsigned int)(p / W), x = (unsigned int)(p - (size_t)y * W);
        uint8_t cur[8], left[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        pixel_bytes(img[p], wide, nch, cur);
        if (x > 0) pixel_bytes(img[p - 1], wide, nch, left);
        uint8_t *o = raw + (size_t)y * raw_row;
        if (x == 0) o[0] = 1;
        o += 1 + (size_t)x * bpp;
        for (unsigned int k = 0; k < bpp; ++k) o[k] = (uint8_t)(cur[k] - left[k]);
    }
}

struct AtomicOr { __device__ __forceinline__ void operator()(uint32_t *w, uint32_t v) const { if (v) atomicOr(w, v); } };

// One thread block of dfl::LANES lanes per CHUNK bytes of the scanline stream -> one deflate block in
// out[chunk * CHUNK_CAP ...] (zeroed by the caller), its byte size in sizes[chunk].  The block's bytes are staged in shared
// memory, lane l's SUB-byte range at l * (SUB + 4) so that the lanes of a warp walk 32 different banks; the symbol sort, the
// canonical codes and the header are spread over the lanes, only the two-queue Huffman construction itself runs on one.
// A deflate block is one chain of dependent work (ncu: 115 K warp instructions when one warp did it all, §profiles), so
// its latency is what a frame's encode costs: hence several warps per block of the stream, not several blocks per warp.
constexpr unsigned int DFL_PITCH = dfl::SUB + 4;                           // bytes between two lanes' ranges in shared memory
constexpr unsigned int DFL_MASK_PITCH = dfl::SUB / 32u + 1u;               // words of match-start bits per lane (+1: bank spread)
constexpr unsigned int DFL_SUB_SHIFT = 8;
static_assert(dfl::SUB == (1u << DFL_SUB_SHIFT) && dfl::LANES % 32u == 0 && dfl::LANES * dfl::SUB == dfl::CHUNK, "lane geometry");
struct DflShared {
    uint8_t bytes[dfl::LANES * DFL_PITCH];
    uint32_t mask[dfl::LANES * DFL_MASK_PITCH];
    uint32_t freq[dfl::NSYM + 2];
    uint16_t code[dfl::NSYM + 2];
    uint8_t len[dfl::NSYM + 2];
    dfl::CodeScratch scratch;
    uint32_t warp_bits[dfl::LANES / 32u];
    uint32_t n_used;
};
__device__ __forceinline__ uint32_t dfl_slot(uint32_t r) { return (r >> DFL_SUB_SHIFT) * DFL_PITCH + (r & (dfl::SUB - 1u)); }

__global__ void __launch_bounds__(dfl::LANES) deflate_chunks_kernel(const uint8_t *__restrict__ raw, size_t raw_len,
                                                                   uint8_t *__restrict__ out, uint32_t *__restrict__ sizes,
                                                                   unsigned int n_chunks)
{
    __shared__ __align__(16) DflShared S;
    const unsigned int lane = threadIdx.x, wlane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t *freq = S.freq;
    uint16_t *code = S.code;
    uint8_t *len = S.len;
    for (unsigned int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const size_t g0 = (size_t)chunk * dfl::CHUNK;
        const size_t g1 = raw_len - g0 < dfl::CHUNK ? raw_len : g0 + dfl::CHUNK;
        const uint32_t n_bytes = (uint32_t)(g1 - g0);
        const bool final = chunk + 1u == n_chunks;
        size_t a = g0 + (size_t)lane * dfl::SUB, b = a + dfl::SUB;
        if (a > g1) a = g1;
        if (b > g1) b = g1;
        // stage the block: 16-byte loads (g0 is a multiple of CHUNK, raw is 256-byte aligned), the tail byte by byte
        for (uint32_t i = lane; i < dfl::CHUNK / 16u; i += dfl::LANES) {
            const uint32_t r = i * 16u;
            if (r >= n_bytes) break;
            if (r + 16u <= n_bytes) {
                uint32_t *dst = reinterpret_cast<uint32_t *>(S.bytes + dfl_slot(r));
                const uint4 v = *reinterpret_cast<const uint4 *>(raw + g0 + r);
                dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
            } else {
                for (uint32_t k = r; k < n_bytes; ++k) S.bytes[dfl_slot(k)] = raw[g0 + k];
            }
        }
        for (unsigned int s = lane; s < dfl::NSYM + 2u; s += dfl::LANES) { freq[s] = s == dfl::EOB ? 1u : 0u; len[s] = 0; }
        if (lane == 0) S.n_used = 0u;
        const uint32_t before = g0 > 0 ? (uint32_t)raw[g0 - 1] : 0u;
        __syncthreads();
        // pass 1: the parse of dfl::parse(), done once: histogram of the tokens, and the tokens themselves left in place for
        // the two later passes — a match overwrites its first byte with (length - 3) and sets that position's bit in the
        // lane's mask; its other bytes are skipped from then on.  Runs are scanned a 32-bit word at a time.  (The last byte of
        // a range is never overwritten — a match is at least 3 long — so the next lane's predecessor byte stays intact.)
        uint8_t *my = S.bytes + lane * DFL_PITCH;
        uint32_t *mask = S.mask + lane * DFL_MASK_PITCH;
        const uint32_t n_my = (uint32_t)(b - a);
        for (unsigned int w = 0; w < dfl::SUB / 32u; ++w) mask[w] = 0u;
        {
            uint32_t prev = a == 0 ? 256u : (lane == 0 ? before : (uint32_t)S.bytes[(lane - 1u) * DFL_PITCH + dfl::SUB - 1u]);
            uint32_t i = 0;
            while (i < n_my) {
                const uint32_t bv = my[i];
                if (bv == prev) {
                    const uint32_t lim = n_my - i < 258u ? n_my : i + 258u;
                    uint32_t e = i + 1u;
                    bool stop = false;
                    while (e < lim && (e & 3u)) { if (my[e] != bv) { stop = true; break; } ++e; }
                    if (!stop) {
                        const uint32_t pat = bv * 0x01010101u;
                        while (e + 4u <= lim) {
                            const uint32_t x = *reinterpret_cast<const uint32_t *>(my + e) ^ pat;
                            if (x) { e += (uint32_t)(__ffs((int)x) - 1) >> 3; stop = true; break; }
                            e += 4u;
                        }
                        if (!stop) while (e < lim && my[e] == bv) ++e;
                    }
                    const uint32_t l = e - i;
                    if (l >= 3u) {
                        uint32_t sy, eb, ev;
                        dfl::length_symbol(l, sy, eb, ev);
                        atomicAdd(&freq[sy], 1u);
                        my[i] = (uint8_t)(l - 3u);
                        mask[i >> 5] |= 1u << (i & 31u);
                        i = e;
                        continue;
                    }
                }
                atomicAdd(&freq[bv], 1u);
                prev = bv;
                ++i;
            }
        }
        __syncthreads();
        // used symbols in ascending (frequency, symbol) order: every lane ranks its symbols against all of them
        {
            uint32_t used = 0;
            for (unsigned int s = lane; s < dfl::NSYM; s += dfl::LANES) used += freq[s] ? 1u : 0u;
            for (unsigned int d = 16; d > 0; d >>= 1) used += __shfl_xor_sync(0xFFFFFFFFu, used, d);
            if (wlane == 0) atomicAdd(&S.n_used, used);
        }
        for (unsigned int s = lane; s < dfl::NSYM; s += dfl::LANES) {
            const uint32_t f = freq[s];
            if (f == 0u) continue;
            uint32_t rank = 0;
            for (unsigned int j = 0; j < dfl::NSYM; ++j) {
                const uint32_t fj = freq[j];
                rank += (fj != 0u && (fj < f || (fj == f && j < s))) ? 1u : 0u;
            }
            S.scratch.key[rank] = f;
            S.scratch.sym[rank] = (uint16_t)s;
        }
        __syncthreads();
        if (lane == 0) dfl::lengths_from_sorted(S.scratch, S.n_used, len);
        __syncthreads();
        // canonical codes (RFC 1951 §3.2.2): lane b of the first warp owns the codes of length b
        if (warp == 0) {
            uint32_t cnt = 0;
            if (wlane >= 1u && wlane <= dfl::MAX_BITS)
                for (unsigned int s = 0; s < dfl::NSYM; ++s) cnt += len[s] == wlane ? 1u : 0u;
            uint32_t c = 0, first = 0;
            for (unsigned int bits = 1; bits <= dfl::MAX_BITS; ++bits) {
                c = (c + __shfl_sync(0xFFFFFFFFu, cnt, bits - 1u)) << 1;
                if (bits == wlane) first = c;
            }
            if (wlane >= 1u && wlane <= dfl::MAX_BITS)
                for (unsigned int s = 0; s < dfl::NSYM; ++s)
                    if (len[s] == wlane) code[s] = (uint16_t)dfl::bit_reverse(first++, wlane);
        }
        __syncthreads();
        // pass 2: what each lane's range costs; prefix sum over the block -> where it starts
        uint32_t mine = 0;
        for (uint32_t i = 0; i < n_my;) {
            if ((mask[i >> 5] >> (i & 31u)) & 1u) {
                const uint32_t l = (uint32_t)my[i] + 3u;
                uint32_t sy, eb, ev;
                dfl::length_symbol(l, sy, eb, ev);
                mine += len[sy] + eb + 1u;
                i += l;
            } else { mine += len[my[i]]; ++i; }
        }
        uint32_t incl = mine;
        for (unsigned int d = 1; d < 32u; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (wlane >= d) incl += up;
        }
        if (wlane == 31u) S.warp_bits[warp] = incl;
        __syncthreads();
        uint32_t total = 0, start = incl - mine;
        for (unsigned int w = 0; w < dfl::LANES / 32u; ++w) {
            if (w < warp) start += S.warp_bits[w];
            total += S.warp_bits[w];
        }
        const size_t dyn_bits = (size_t)dfl::HEADER_BITS + total + len[dfl::EOB];
        const size_t dyn_bytes = dfl::dynamic_block_bytes(dyn_bits, final), st_bytes = dfl::stored_block_bytes(n_bytes);
        uint8_t *dst = out + (size_t)chunk * dfl::CHUNK_CAP;
        uint32_t *words = reinterpret_cast<uint32_t *>(dst);
        if (dyn_bytes < st_bytes) {
            // header: lane 0 the fixed part, then each lane of the first warp 9 of the 288 code lengths (4 bits each)
            if (lane == 0) {
                dfl::BitSink<AtomicOr> hs(words, 0, AtomicOr());
                dfl::put_header_fixed(hs, final);
                hs.flush();
            }
            if (warp == 0) {
                dfl::BitSink<AtomicOr> ls(words, (size_t)dfl::HEADER_FIXED_BITS + 36u * wlane, AtomicOr());
                for (unsigned int k = wlane * 9u; k < wlane * 9u + 9u; ++k) ls.put(dfl::bit_reverse(k < dfl::NSYM ? len[k] : 1u, 4u), 4u);
                ls.flush();
            }
            // pass 3: the tokens
            dfl::BitSink<AtomicOr> bs(words, (size_t)dfl::HEADER_BITS + start, AtomicOr());
            for (uint32_t i = 0; i < n_my;) {
                if ((mask[i >> 5] >> (i & 31u)) & 1u) {
                    const uint32_t l = (uint32_t)my[i] + 3u;
                    uint32_t sy, eb, ev;
                    dfl::length_symbol(l, sy, eb, ev);
                    bs.put((uint32_t)code[sy] | (ev << len[sy]), len[sy] + eb + 1u);      // symbol, extra bits, distance code "0"
                    i += l;
                } else { const uint32_t v = my[i]; bs.put(code[v], len[v]); ++i; }
            }
            bs.flush();
            if (lane == dfl::LANES - 1u) {
                dfl::BitSink<AtomicOr> ts(words, (size_t)dfl::HEADER_BITS + total, AtomicOr());
                ts.put(code[dfl::EOB], len[dfl::EOB]);
                ts.flush();
                if (!final) {                                  // empty stored block: LEN = 0x0000, NLEN = 0xFFFF
                    const size_t nlen = (dyn_bits + 3u + 7u) / 8u + 2u;
                    for (size_t k = nlen; k < nlen + 2u; ++k) atomicOr(words + (k >> 2), 0xFFu << (8u * (unsigned int)(k & 3u)));
                }
            }
            if (lane == 0) sizes[chunk] = (uint32_t)dyn_bytes;
        } else {
            if (lane == 0) {
                dst[0] = final ? 1 : 0;
                dst[1] = (uint8_t)n_bytes; dst[2] = (uint8_t)(n_bytes >> 8); dst[3] = (uint8_t)~n_bytes; dst[4] = (uint8_t)(~n_bytes >> 8);
                sizes[chunk] = (uint32_t)st_bytes;
            }
            for (uint32_t k = lane; k < n_bytes; k += dfl::LANES) dst[5 + k] = raw[g0 + k];
        }
        __syncthreads();
    }
}

// offsets[c] = sum of sizes[0..c), offsets[n] = total; one block
__global__ void deflate_scan_kernel(const uint32_t *__restrict__ sizes, unsigned long long *__restrict__ offsets, unsigned int n)
{
    __shared__ unsigned long long s_part[1024];
    const unsigned int t = threadIdx.x, per = (n + blockDim.x - 1) / blockDim.x;
    const unsigned int lo = t * per < n ? t * per : n, hi = lo + per < n ? lo + per : n;
    unsigned long long sum = 0;
    for (unsigned int i = lo; i < hi; ++i) sum += sizes[i];
    s_part[t] = sum;
    __syncthreads();
    if (t == 0) {
        unsigned long long run = 0;
        for (unsigned int k = 0; k < blockDim.x; ++k) { const unsigned long long v = s_part[k]; s_part[k] = run; run += v; }
        offsets[n] = run;
    }
    __syncthreads();
    unsigned long long run = s_part[t];
    for (unsigned int i = lo; i < hi; ++i) { offsets[i] = run; run += sizes[i]; }
}

__global__ void deflate_gather_kernel(const uint8_t *__restrict__ chunks, const uint32_t *__restrict__ sizes,
                                      const unsigned long long *__restrict__ offsets, uint8_t *__restrict__ pay, unsigned int n_chunks)
{
    for (unsigned int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const uint8_t *src = chunks + (size_t)c * dfl::CHUNK_CAP;
        uint8_t *dst = pay + offsets[c];
        const uint32_t n = sizes[c];
        for (uint32_t k = threadIdx.x; k < n; k += blockDim.x) dst[k] = src[k];
    }
}

// crc[t]: CRC-32 register (from 0, no conditioning) over bytes [t * PNG_CHUNK, ...) of the payload, whose length is the
// device-side total; adler[2t], adler[2t+1]: (sum d, sum (len - j) d_j) over piece t of the scanline stream.  A piece is one
// serial chain per thread, so the chain is kept short: CRC four bytes per step (slicing-by-4: four independent look-ups),
// and the Adler pieces go to threads of their own (threads [n_crc_max, n_crc_max + n_adler)).
__global__ void deflate_sums_kernel(const uint8_t *__restrict__ pay, const unsigned long long *__restrict__ total,
                                    const uint8_t *__restrict__ raw, size_t raw_len, uint32_t *__restrict__ crc,
                                    unsigned long long *__restrict__ adler, size_t n_crc_max, size_t n_adler)
{
    __shared__ uint32_t table[4][256];
    for (unsigned int n = threadIdx.x; n < 256; n += blockDim.x) {
        uint32_t c = n;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        table[0][n] = c;
    }
    __syncthreads();
    for (int k = 1; k < 4; ++k) {
        for (unsigned int n = threadIdx.x; n < 256; n += blockDim.x) {
            const uint32_t p = table[k - 1][n];
            table[k][n] = (p >> 8) ^ table[0][p & 0xFFu];
        }
        __syncthreads();
    }
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t pay_len = (size_t)*total;
    // pieces start at multiples of PNG_CHUNK of 256-byte aligned buffers: whole pieces are read 16 bytes at a time
    if (t < n_crc_max) {
        const size_t lo = t * PNG_CHUNK, hi = lo + PNG_CHUNK < pay_len ? lo + PNG_CHUNK : pay_len;
        uint32_t c = 0;
        size_t i = lo;
        for (; i + 16 <= hi; i += 16) {
            const uint4 v = *reinterpret_cast<const uint4 *>(pay + i);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                c ^= w[q];
                c = table[3][c & 0xFFu] ^ table[2][(c >> 8) & 0xFFu] ^ table[1][(c >> 16) & 0xFFu] ^ table[0][c >> 24];
            }
        }
        for (; i < hi; ++i) c = table[0][(c ^ pay[i]) & 0xFFu] ^ (c >> 8);
        crc[t] = c;
    } else if (t < n_crc_max + n_adler) {
        const size_t u = t - n_crc_max;
        const size_t lo = u * PNG_CHUNK, hi = lo + PNG_CHUNK < raw_len ? lo + PNG_CHUNK : raw_len;
        unsigned long long a = 0, b = 0;
        size_t k = lo;
        for (; k + 16 <= hi; k += 16) {
            const uint4 v = *reinterpret_cast<const uint4 *>(raw + k);
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int e = 0; e < 4; ++e) { a += (w[q] >> (8 * e)) & 0xFFu; b += a; }
        }
        for (; k < hi; ++k) { a += raw[k]; b += a; }
        adler[2 * u] = a; adler[2 * u + 1] = b;
    }
}

void launch_png_deflate(const uint16_t *rgba, unsigned int W, unsigned int H, unsigned int fmt, size_t raw_row, size_t raw_len,
                        uint8_t *raw, uint8_t *chunks, uint32_t *sizes, unsigned long long *offsets, uint8_t *pay,
                        uint32_t *crc, unsigned long long *adler, size_t n_crc_max, size_t n_adler, cudaStream_t s)
{
    const size_t npix = (size_t)W * H;
    if (npix == 0) return;
    const unsigned int n_chunks = (unsigned int)((raw_len + dfl::CHUNK - 1) / dfl::CHUNK);
    size_t g = (npix + 255) / 256;
    png_filter_kernel<<<(unsigned int)(g > 148u * 16u ? 148u * 16u : g), 256, 0, s>>>(reinterpret_cast<const ushort4 *>(rgba), raw, W, H, fmt, raw_row);
    cudaMemsetAsync(chunks, 0, (size_t)n_chunks * dfl::CHUNK_CAP, s);
    static_assert(sizeof(DflShared) <= 48 * 1024, "static shared memory");
    deflate_chunks_kernel<<<n_chunks > 148u * 9u ? 148u * 9u : n_chunks, dfl::LANES, 0, s>>>(raw, raw_len, chunks, sizes, n_chunks);
    deflate_scan_kernel<<<1, 1024, 0, s>>>(sizes, offsets, n_chunks);
    deflate_gather_kernel<<<n_chunks > 148u * 8u ? 148u * 8u : n_chunks, 256, 0, s>>>(chunks, sizes, offsets, pay, n_chunks);
    const size_t n = n_crc_max + n_adler;
    // one piece per thread is a long serial chain (4 096 dependent table look-ups): few threads per block, so that the
    // ~3 000 pieces of a frame spread over all SMs
    deflate_sums_kernel<<<(unsigned int)((n + 31) / 32), 32, 0, s>>>(pay, offsets + n_chunks, raw, raw_len, crc, adler, n_crc_max, n_adler);
    bump_launches(5);
}

}  // namespace sar
