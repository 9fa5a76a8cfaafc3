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

// One warp per CHUNK bytes of the scanline stream -> one deflate block in out[chunk * CHUNK_CAP ...] (zeroed by the caller),
// its byte size in sizes[chunk].
constexpr unsigned int DFL_WARPS = 4;
__global__ void __launch_bounds__(DFL_WARPS * 32) deflate_chunks_kernel(const uint8_t *__restrict__ raw, size_t raw_len,
                                                                        uint8_t *__restrict__ out, uint32_t *__restrict__ sizes,
                                                                        unsigned int n_chunks)
{
    __shared__ uint32_t s_freq[DFL_WARPS][dfl::NSYM + 2];
    __shared__ uint16_t s_code[DFL_WARPS][dfl::NSYM + 2];
    __shared__ uint8_t s_len[DFL_WARPS][dfl::NSYM + 2];
    __shared__ dfl::CodeScratch s_scratch[DFL_WARPS];
    const unsigned int warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint32_t *freq = s_freq[warp];
    uint16_t *code = s_code[warp];
    uint8_t *len = s_len[warp];
    for (unsigned int chunk = blockIdx.x * DFL_WARPS + warp; chunk < n_chunks; chunk += gridDim.x * DFL_WARPS) {
        const size_t g0 = (size_t)chunk * dfl::CHUNK;
        const size_t g1 = raw_len - g0 < dfl::CHUNK ? raw_len : g0 + dfl::CHUNK;
        const bool final = chunk + 1u == n_chunks;
        size_t a = g0 + (size_t)lane * dfl::SUB, b = a + dfl::SUB;
        if (a > g1) a = g1;
        if (b > g1) b = g1;
        for (unsigned int s = lane; s < dfl::NSYM; s += 32u) freq[s] = s == dfl::EOB ? 1u : 0u;
        __syncwarp();
        dfl::parse(raw, a, b, [&](uint32_t v) { atomicAdd(&freq[v], 1u); },
                   [&](uint32_t l) { uint32_t sy, eb, ev; dfl::length_symbol(l, sy, eb, ev); atomicAdd(&freq[sy], 1u); });
        __syncwarp();
        if (lane == 0) {
            dfl::code_lengths(freq, len, s_scratch[warp]);
            dfl::canonical_codes(len, code);
        }
        __syncwarp();
        const uint32_t mine = dfl::range_bits(raw, a, b, len);
        uint32_t incl = mine;
        for (unsigned int d = 1; d < 32u; d <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += up;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        const size_t dyn_bits = (size_t)dfl::HEADER_BITS + total + len[dfl::EOB];
        const size_t dyn_bytes = dfl::dynamic_block_bytes(dyn_bits, final), st_bytes = dfl::stored_block_bytes(g1 - g0);
        uint8_t *dst = out + (size_t)chunk * dfl::CHUNK_CAP;
        uint32_t *words = reinterpret_cast<uint32_t *>(dst);
        if (dyn_bytes < st_bytes) {
            if (lane == 0) {
                dfl::BitSink<AtomicOr> hs(words, 0, AtomicOr());
                dfl::put_header(hs, len, final);
                hs.flush();
            }
            dfl::BitSink<AtomicOr> bs(words, (size_t)dfl::HEADER_BITS + (incl - mine), AtomicOr());
            dfl::range_emit(bs, raw, a, b, len, code);
            bs.flush();
            if (lane == 31u) {
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
            const size_t n = g1 - g0;
            if (lane == 0) {
                dst[0] = final ? 1 : 0;
                dst[1] = (uint8_t)n; dst[2] = (uint8_t)(n >> 8); dst[3] = (uint8_t)~n; dst[4] = (uint8_t)(~n >> 8);
                sizes[chunk] = (uint32_t)st_bytes;
            }
            for (size_t k = lane; k < n; k += 32u) dst[5 + k] = raw[g0 + k];
        }
        __syncwarp();
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
// device-side total; adler[2t], adler[2t+1]: (sum d, sum (len - j) d_j) over piece t of the scanline stream
__global__ void deflate_sums_kernel(const uint8_t *__restrict__ pay, const unsigned long long *__restrict__ total,
                                    const uint8_t *__restrict__ raw, size_t raw_len, uint32_t *__restrict__ crc,
                                    unsigned long long *__restrict__ adler, size_t n_crc_max, size_t n_adler)
{
    __shared__ uint32_t table[256];
    for (unsigned int n = threadIdx.x; n < 256; n += blockDim.x) {
        uint32_t c = n;
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        table[n] = c;
    }
    __syncthreads();
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t pay_len = (size_t)*total;
    if (t < n_crc_max) {
        const size_t lo = t * PNG_CHUNK, hi = lo + PNG_CHUNK < pay_len ? lo + PNG_CHUNK : pay_len;
        uint32_t c = 0;
        for (size_t i = lo; i < hi; ++i) c = table[(c ^ pay[i]) & 0xFFu] ^ (c >> 8);
        crc[t] = c;
    }
    if (t < n_adler) {
        const size_t lo = t * PNG_CHUNK, hi = lo + PNG_CHUNK < raw_len ? lo + PNG_CHUNK : raw_len;
        unsigned long long a = 0, b = 0;
        for (size_t k = lo; k < hi; ++k) { a += raw[k]; b += a; }
        adler[2 * t] = a; adler[2 * t + 1] = b;
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
    const unsigned int blocks = (n_chunks + DFL_WARPS - 1) / DFL_WARPS;
    deflate_chunks_kernel<<<blocks > 148u * 8u ? 148u * 8u : blocks, DFL_WARPS * 32, 0, s>>>(raw, raw_len, chunks, sizes, n_chunks);
    deflate_scan_kernel<<<1, 1024, 0, s>>>(sizes, offsets, n_chunks);
    deflate_gather_kernel<<<n_chunks > 148u * 8u ? 148u * 8u : n_chunks, 256, 0, s>>>(chunks, sizes, offsets, pay, n_chunks);
    const size_t n = n_crc_max > n_adler ? n_crc_max : n_adler;
    deflate_sums_kernel<<<(unsigned int)((n + 127) / 128), 128, 0, s>>>(pay, offsets + n_chunks, raw, raw_len, crc, adler, n_crc_max, n_adler);
    bump_launches(5);
}

}  // namespace sar
