// Host harness for strange-attractor-renderer_b200/csrc/sar_deflate.cuh: the same inline functions the CUDA kernel
// (sar_deflate.cu: deflate_chunks_kernel) calls, driven with the lanes of a block emulated by a loop, so that the
// bit-level logic (run-length parse, length symbols, Huffman code construction and its 15-bit cap, canonical codes, block
// header, bit packing at arbitrary offsets, sync flush, stored fallback) is checked on the CPU against zlib's inflate
// (tests/test_deflate_host.py).  Built by the test with g++; not part of the product library.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../strange-attractor-renderer_b200/csrc/sar_deflate.cuh"

using namespace sar;

struct HostOr { void operator()(uint32_t *w, uint32_t v) const { *w |= v; } };

static size_t chunk_host(const uint8_t *raw, size_t raw_len, unsigned chunk, unsigned n_chunks, uint8_t *dst)
{
    static thread_local uint32_t freq[dfl::NSYM + 2];
    static thread_local uint16_t code[dfl::NSYM + 2];
    static thread_local uint8_t len[dfl::NSYM + 2];
    static thread_local dfl::CodeScratch scratch;
    const size_t g0 = (size_t)chunk * dfl::CHUNK;
    const size_t g1 = raw_len - g0 < dfl::CHUNK ? raw_len : g0 + dfl::CHUNK;
    const bool final = chunk + 1u == n_chunks;
    size_t a[dfl::LANES], b[dfl::LANES];
    for (unsigned lane = 0; lane < dfl::LANES; ++lane) {
        a[lane] = g0 + (size_t)lane * dfl::SUB; b[lane] = a[lane] + dfl::SUB;
        if (a[lane] > g1) a[lane] = g1;
        if (b[lane] > g1) b[lane] = g1;
    }
    for (unsigned s = 0; s < dfl::NSYM; ++s) freq[s] = s == dfl::EOB ? 1u : 0u;
    for (unsigned lane = 0; lane < dfl::LANES; ++lane)
        dfl::parse(dfl::PtrAt{raw}, a[lane], b[lane], [&](uint32_t v) { ++freq[v]; },
                   [&](uint32_t l) { uint32_t sy, eb, ev; dfl::length_symbol(l, sy, eb, ev); ++freq[sy]; });
    dfl::code_lengths(freq, len, scratch);
    dfl::canonical_codes(len, code);
    uint32_t mine[dfl::LANES], excl[dfl::LANES], total = 0;
    for (unsigned lane = 0; lane < dfl::LANES; ++lane) { mine[lane] = dfl::range_bits(dfl::PtrAt{raw}, a[lane], b[lane], len); excl[lane] = total; total += mine[lane]; }
    const size_t dyn_bits = (size_t)dfl::HEADER_BITS + total + len[dfl::EOB];
    const size_t dyn_bytes = dfl::dynamic_block_bytes(dyn_bits, final), st_bytes = dfl::stored_block_bytes(g1 - g0);
    uint32_t *words = reinterpret_cast<uint32_t *>(dst);
    if (dyn_bytes < st_bytes) {
        dfl::BitSink<HostOr> hs(words, 0, HostOr());              // as the kernel: lane 0 the fixed part, every lane 9 lengths
        dfl::put_header_fixed(hs, final);
        hs.flush();
        for (unsigned lane = 0; lane < 32; ++lane) {
            dfl::BitSink<HostOr> ls(words, (size_t)dfl::HEADER_FIXED_BITS + 36u * lane, HostOr());
            for (unsigned k = lane * 9u; k < lane * 9u + 9u; ++k) ls.put(dfl::bit_reverse(k < dfl::NSYM ? len[k] : 1u, 4u), 4u);
            ls.flush();
        }
        for (unsigned lane = 0; lane < dfl::LANES; ++lane) {
            dfl::BitSink<HostOr> bs(words, (size_t)dfl::HEADER_BITS + excl[lane], HostOr());
            dfl::range_emit(bs, dfl::PtrAt{raw}, a[lane], b[lane], len, code);
            bs.flush();
        }
        dfl::BitSink<HostOr> ts(words, (size_t)dfl::HEADER_BITS + total, HostOr());
        ts.put(code[dfl::EOB], len[dfl::EOB]);
        ts.flush();
        if (!final) {
            const size_t nlen = (dyn_bits + 3u + 7u) / 8u + 2u;
            for (size_t k = nlen; k < nlen + 2u; ++k) words[k >> 2] |= 0xFFu << (8u * (unsigned)(k & 3u));
        }
        return dyn_bytes;
    }
    const size_t n = g1 - g0;
    dst[0] = final ? 1 : 0;
    dst[1] = (uint8_t)n; dst[2] = (uint8_t)(n >> 8); dst[3] = (uint8_t)~n; dst[4] = (uint8_t)(~n >> 8);
    memcpy(dst + 5, raw + g0, n);
    return st_bytes;
}

extern "C" {

// raw deflate stream (no zlib wrapper) of raw[0..n); returns its size, or 0 if `cap` is too small.  kinds (optional,
// one byte per block): 1 = dynamic Huffman, 0 = stored.
size_t dfl_compress_host(const uint8_t *raw, size_t n, uint8_t *out, size_t cap, uint8_t *kinds)
{
    const unsigned n_chunks = (unsigned)((n + dfl::CHUNK - 1) / dfl::CHUNK);
    alignas(8) static thread_local uint8_t tmp[dfl::CHUNK_CAP];
    size_t pos = 0;
    for (unsigned c = 0; c < n_chunks; ++c) {
        memset(tmp, 0, sizeof tmp);
        const size_t sz = chunk_host(raw, n, c, n_chunks, tmp);
        if (pos + sz > cap) return 0;
        memcpy(out + pos, tmp, sz);
        if (kinds) kinds[c] = (tmp[0] & 6u) == 4u ? 1 : 0;
        pos += sz;
    }
    return pos;
}

// code lengths alone, for the 15-bit cap test: freq[286] -> len[286]
void dfl_code_lengths_host(const uint32_t *freq, uint8_t *len)
{
    static thread_local dfl::CodeScratch scratch;
    dfl::code_lengths(freq, len, scratch);
}

uint32_t dfl_chunk_bytes(void) { return dfl::CHUNK; }

}  // extern "C"
