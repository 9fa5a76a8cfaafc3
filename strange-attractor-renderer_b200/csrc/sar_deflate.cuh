// sar_deflate.cuh — the compressor of the PNG writer (src/bin/main.rs:78-89: PngEncoder with CompressionType::Default).
// The reference hands its image to the `png` crate, which deflates the filtered scanlines on one host thread.  A deflate
// stream is free to choose its own blocks and matches — only what it DECODES to is specified (RFC 1951) — so this is not a
// restatement of that crate but a compressor shaped for a GPU:
//   * the scanline stream (Sub-filtered, so the untouched 80 % of a frame is zeros) is cut into CHUNK-byte pieces, one
//     deflate block each, two warps per block; every block ends on a byte boundary (an empty stored block, the "sync
//     flush" of zlib), so blocks are produced independently and concatenated;
//   * matches are run-length only (distance 1, length 3..258): finding them needs no hash table, each lane parses its
//     own 1/64 of the block, and on these images it is what zlib's Z_RLE strategy does — 3.60 MB for the reference's
//     poisson-saturne frame against the 3.63 MB of the file the reference published;
//   * each block carries its own dynamic Huffman code (literal/length alphabet from the block's histogram, built by
//     the warp — the two-queue construction itself on one lane —, length-limited to 15 bits); a block that would not
//     shrink is stored instead.
// Everything that decides bits is in this header as host+device inline functions, so the very same code is exercised
// on the CPU by tests/cpp/deflate_host.cpp (lanes emulated by a loop) against zlib's inflate.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __CUDACC__
#define SAR_HD __host__ __device__ __forceinline__
#else
#define SAR_HD inline
#endif

namespace sar {
namespace dfl {

constexpr uint32_t CHUNK = 16384;              // scanline bytes per deflate block
constexpr uint32_t LANES = 64;                 // lanes (two warps) that share one block of the stream
constexpr uint32_t SUB = CHUNK / LANES;        // bytes parsed by one lane
constexpr uint32_t NSYM = 286;                 // literal/length alphabet, RFC 1951 §3.2.5
constexpr uint32_t EOB = 256;
constexpr uint32_t MAX_BITS = 15;
constexpr uint32_t CHUNK_CAP = CHUNK + 64;     // bytes reserved per block: the stored fallback needs 5 + CHUNK
// block header: BFINAL + BTYPE (3), HLIT (5), HDIST (5), HCLEN (4), 19 code-length-code lengths (3 each), then the
// 286 + 2 code lengths themselves in a flat 4-bit code (symbols 0..15 all of length 4: complete, no repeat codes)
constexpr uint32_t HEADER_FIXED_BITS = 3 + 5 + 5 + 4 + 19 * 3;
constexpr uint32_t HEADER_BITS = HEADER_FIXED_BITS + (NSYM + 2) * 4;

// length 3..258 -> (symbol 257..285, extra bits, extra value), RFC 1951 §3.2.5
SAR_HD void length_symbol(uint32_t len, uint32_t &sym, uint32_t &ebits, uint32_t &eval)
{
    if (len == 258u) { sym = 285u; ebits = 0u; eval = 0u; return; }
    const uint32_t l = len - 3u;
    if (l < 8u) { sym = 257u + l; ebits = 0u; eval = 0u; return; }
    uint32_t e = 1u;                            // e = floor(log2 l) - 2
    while ((l >> (e + 3u)) != 0u) ++e;
    sym = 257u + 4u * e + (l >> e);
    ebits = e;
    eval = l & ((1u << e) - 1u);
}

SAR_HD uint32_t bit_reverse(uint32_t v, uint32_t n)
{
#ifdef __CUDA_ARCH__
    return n ? __brev(v) >> (32u - n) : 0u;
#endif
    uint32_t r = 0u;
    for (uint32_t i = 0; i < n; ++i) { r = (r << 1) | (v & 1u); v >>= 1; }
    return r;
}

// Greedy run-length parse of positions [g0, g1) of the scanline stream, read through `at(g)`: a position whose byte equals
// its predecessor opens a match of distance 1 over the rest of the run (at most 258 bytes, never past g1, at least 3),
// anything else is a literal.  The predecessor of g0 may lie in an earlier block: the 32 KB window of a deflate stream
// runs across blocks.  (`at` is a plain pointer read on the host and a shared-memory read on the device.)
template <class At, class Lit, class Match>
SAR_HD void parse(At &&at, size_t g0, size_t g1, Lit &&lit, Match &&match)
{
    size_t g = g0;
    uint32_t prev = g0 > 0 ? (uint32_t)at(g0 - 1) : 256u;     // 256: no predecessor (first byte of the stream)
    while (g < g1) {
        const uint32_t b = (uint32_t)at(g);
        if (prev == b) {
            const size_t lim = g1 - g < 258u ? g1 : g + 258u;
            size_t e = g + 1;
            while (e < lim && (uint32_t)at(e) == b) ++e;
            if (e - g >= 3u) { match((uint32_t)(e - g)); g = e; continue; }
        }
        lit(b);
        prev = b;
        ++g;
    }
}
struct PtrAt { const uint8_t *p; SAR_HD uint8_t operator()(size_t g) const { return p[g]; } };

// Code lengths of an optimal prefix code for freq[0..NSYM), limited to MAX_BITS.  Serial (one lane): sort the used
// symbols by frequency, run the in-place two-queue Huffman construction (Moffat & Katajainen: after the first pass a
// node holds its parent's index, after the second its depth), then cap the depth by moving leaves down until the Kraft
// sum is 1 again.  At least two symbols get a code, so the code is complete (zlib rejects incomplete literal codes).
struct CodeScratch { uint32_t key[NSYM]; uint16_t sym[NSYM]; };
// the used symbols in ascending (frequency, symbol) order -> w.key / w.sym; returns how many.  Serial form; the kernel
// ranks the symbols with the whole warp instead (the order is unique, so both give the same arrays).
SAR_HD uint32_t sort_symbols(const uint32_t *freq, CodeScratch &w)
{
    uint32_t n = 0;
    for (uint32_t s = 0; s < NSYM; ++s) if (freq[s]) { w.key[n] = freq[s]; w.sym[n] = (uint16_t)s; ++n; }
    for (uint32_t gap = n / 2u; gap > 0u; gap /= 2u)           // shell sort
        for (uint32_t i = gap; i < n; ++i) {
            const uint32_t k = w.key[i]; const uint16_t s = w.sym[i];
            uint32_t j = i;
            while (j >= gap && (w.key[j - gap] > k || (w.key[j - gap] == k && w.sym[j - gap] > s))) {
                w.key[j] = w.key[j - gap]; w.sym[j] = w.sym[j - gap]; j -= gap;
            }
            w.key[j] = k; w.sym[j] = s;
        }
    return n;
}
// w holds n >= 2 sorted symbols (a block always has its end-of-block symbol and at least one token); len[] must be zeroed
SAR_HD void lengths_from_sorted(CodeScratch &w, uint32_t n, uint8_t *len)
{
    uint32_t *a = w.key;
    if (n == 2u) { a[0] = 1u; a[1] = 1u; }
    else {
        // pass 1: internal node `next` = the two smallest of (unused leaves from `leaf`, unused internal nodes from `root`)
        a[0] += a[1];
        uint32_t root = 0u, leaf = 2u;
        for (uint32_t next = 1u; next < n - 1u; ++next) {
            if (leaf >= n || a[root] < a[leaf]) { a[next] = a[root]; a[root++] = next; } else a[next] = a[leaf++];
            if (leaf >= n || (root < next && a[root] < a[leaf])) { a[next] += a[root]; a[root++] = next; } else a[next] += a[leaf++];
        }
        // pass 2: parent indices -> depths of the internal nodes
        a[n - 2u] = 0u;
        for (int next = (int)n - 3; next >= 0; --next) a[next] = a[a[next]] + 1u;
        // pass 3: depths of the leaves, deepest first
        int avail = 1, used = 0, depth = 0, r = (int)n - 2, next = (int)n - 1;
        while (avail > 0) {
            while (r >= 0 && (int)a[r] == depth) { ++used; --r; }
            while (avail > used) { a[next--] = (uint32_t)depth; --avail; }
            avail = 2 * used; ++depth; used = 0;
        }
    }
    // a[i] = code length of the i-th least frequent symbol (non-increasing in i); cap at MAX_BITS
    uint32_t count[33];
    for (uint32_t i = 0; i <= 32u; ++i) count[i] = 0u;
    for (uint32_t i = 0; i < n; ++i) ++count[a[i] > 32u ? 32u : a[i]];
    bool over = false;
    for (uint32_t i = MAX_BITS + 1u; i <= 32u; ++i) if (count[i]) { over = true; count[MAX_BITS] += count[i]; count[i] = 0u; }
    if (over) {
        uint32_t total = 0u;
        for (uint32_t i = MAX_BITS; i > 0u; --i) total += count[i] << (MAX_BITS - i);
        while (total != (1u << MAX_BITS)) {     // over-subscribed: push one leaf of the deepest level out, split a shallower one
            --count[MAX_BITS];
            for (uint32_t i = MAX_BITS - 1u; i > 0u; --i)
                if (count[i]) { --count[i]; count[i + 1u] += 2u; break; }
            --total;
        }
    }
    uint32_t j = n;                              // shortest codes to the most frequent symbols
    for (uint32_t bits = 1u; bits <= MAX_BITS; ++bits)
        for (uint32_t c = count[bits]; c > 0u; --c) len[w.sym[--j]] = (uint8_t)bits;
}
SAR_HD void code_lengths(const uint32_t *freq, uint8_t *len, CodeScratch &w)
{
    uint32_t fixed[NSYM];
    bool patched = false;
    uint32_t used = 0;
    for (uint32_t s = 0; s < NSYM; ++s) { len[s] = 0; used += freq[s] ? 1u : 0u; }
    if (used < 2u) {                             // cannot happen for a block, kept so that the function is total
        for (uint32_t s = 0; s < NSYM; ++s) fixed[s] = freq[s];
        fixed[EOB] = fixed[EOB] ? fixed[EOB] : 1u;
        if (used == 0u || freq[EOB]) fixed[0] = fixed[0] ? fixed[0] : 1u;
        patched = true;
    }
    const uint32_t n = sort_symbols(patched ? fixed : freq, w);
    lengths_from_sorted(w, n, len);
}

// canonical codes (RFC 1951 §3.2.2), stored bit-reversed: deflate packs Huffman codes most significant bit first into a
// stream that is otherwise filled from the least significant bit
SAR_HD void canonical_codes(const uint8_t *len, uint16_t *code)
{
    uint32_t count[MAX_BITS + 1u], next[MAX_BITS + 2u];
    for (uint32_t i = 0; i <= MAX_BITS; ++i) count[i] = 0u;
    for (uint32_t s = 0; s < NSYM; ++s) ++count[len[s]];
    count[0] = 0u;
    uint32_t c = 0u;
    for (uint32_t b = 1u; b <= MAX_BITS; ++b) { c = (c + count[b - 1u]) << 1; next[b] = c; }
    for (uint32_t s = 0; s < NSYM; ++s) code[s] = len[s] ? (uint16_t)bit_reverse(next[len[s]]++, len[s]) : (uint16_t)0;
}

// Bits are ORed into 32-bit little-endian words of a zeroed buffer; `Or` is atomicOr on the device (lanes share the
// words at the seams of their ranges) and |= on the host.
template <class Or>
struct BitSink {
    uint32_t *words; uint64_t acc; uint32_t fill; size_t word; Or orw;
    SAR_HD BitSink(uint32_t *w, size_t bit0, Or o) : words(w), acc(0), fill((uint32_t)(bit0 & 31u)), word(bit0 >> 5), orw(o) {}
    SAR_HD void put(uint32_t value, uint32_t nbits)           // nbits <= 25
    {
        acc |= (uint64_t)value << fill;
        fill += nbits;
        if (fill >= 32u) { orw(words + word, (uint32_t)acc); acc >>= 32; fill -= 32u; ++word; }
    }
    SAR_HD void flush() { if (fill) orw(words + word, (uint32_t)acc); acc = 0; }
};

// the block header for a dynamic block; `final` sets BFINAL.  The first HEADER_FIXED_BITS do not depend on the code.
template <class Or>
SAR_HD void put_header_fixed(BitSink<Or> &s, bool final)
{
    s.put(final ? 1u : 0u, 1u);
    s.put(2u, 2u);                                            // BTYPE = 10, dynamic Huffman
    s.put(NSYM - 257u, 5u);                                   // HLIT
    s.put(1u, 5u);                                            // HDIST: two distance codes, 1 bit each (only "distance 1" is used)
    s.put(15u, 4u);                                           // HCLEN: all 19 code-length-code lengths follow
    // order 16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15: the repeat codes unused (0), the lengths 0..15 in 4 bits each
    for (uint32_t i = 0; i < 19u; ++i) s.put(i < 3u ? 0u : 4u, 3u);
}
template <class Or>
SAR_HD void put_header(BitSink<Or> &s, const uint8_t *len, bool final)
{
    put_header_fixed(s, final);
    for (uint32_t k = 0; k < NSYM; ++k) s.put(bit_reverse(len[k], 4u), 4u);
    s.put(bit_reverse(1u, 4u), 4u);
    s.put(bit_reverse(1u, 4u), 4u);
}

// bits one lane's range costs under the code `len` (a match adds its extra bits and the 1-bit distance code)
template <class At>
SAR_HD uint32_t range_bits(At &&at, size_t g0, size_t g1, const uint8_t *len)
{
    uint32_t bits = 0u;
    parse(at, g0, g1, [&](uint32_t b) { bits += len[b]; },
          [&](uint32_t l) { uint32_t sy, eb, ev; length_symbol(l, sy, eb, ev); bits += len[sy] + eb + 1u; });
    return bits;
}

template <class Or, class At>
SAR_HD void range_emit(BitSink<Or> &s, At &&at, size_t g0, size_t g1, const uint8_t *len, const uint16_t *code)
{
    parse(at, g0, g1, [&](uint32_t b) { s.put(code[b], len[b]); },
          [&](uint32_t l) {
              uint32_t sy, eb, ev;
              length_symbol(l, sy, eb, ev);
              s.put((uint32_t)code[sy] | (ev << len[sy]), len[sy] + eb + 1u);    // symbol, extra bits, distance code "0"
          });
}

// size in bytes of a block whose dynamic part (header + tokens + EOB) is `bits` long: a non-final block is followed by
// an empty stored block (3 header bits, pad to a byte, LEN = 0, NLEN = 0xFFFF) so that the next block starts on a byte
SAR_HD size_t dynamic_block_bytes(size_t bits, bool final) { return final ? (bits + 7u) / 8u : (bits + 3u + 7u) / 8u + 4u; }
SAR_HD size_t stored_block_bytes(size_t n) { return n + 5u; }

}  // namespace dfl
}  // namespace sar
