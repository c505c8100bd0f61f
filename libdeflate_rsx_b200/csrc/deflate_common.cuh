// deflate_common.cuh — pieces shared by the compression kernels (sm_100a):
// framing, the warp bit sink (prefix-sum bit packer), cooperative match length,
// RFC 1951 symbol geometry, and the length-limited Huffman code builder.
#pragma once
#include "checksum.cuh"
#include "common.cuh"

namespace bdf {


struct DeflateArgs {
    const uint8_t *in;
    const uint64_t *in_off;
    uint8_t *out;
    const uint64_t *out_off;
    uint64_t *out_size;
    int32_t *status;
    unsigned long long *work_counter;
    void *scratch;           // per-CTA scratch slabs (levels 2..9)
    uint64_t scratch_stride; // bytes per CTA
    uint32_t n;
    int level;
    int format;
    // Units of a chunked call (Compressor::compress cuts inputs above 256 KiB into 256 KiB chunks,
    // each through a fresh compressor, src/compress/mod.rs:699-772): bit 0 = this unit ends the
    // stream (FlushMode::Finish), bit 1 = it is followed by a sync flush (:662-681).  NULL = every
    // entry is a whole stream (finish, no sync).
    const uint8_t *unit_flags;
    // Size estimation (Compressor::compress_to_size, src/compress/mod.rs:792-1094): the same kernels
    // run with out == NULL, count bits instead of storing them and follow the estimator's block
    // rules (level 1 always runs the split statistics, levels 10..12 seed their costs from a second
    // greedy parse, an empty input costs nothing).  final_block only matters at level 0 (:1073-1082).
    int size_only = 0;
    int final_block = 1;
    // Levels 2..9, streams of at most 64 KiB: two kernels share a batch (deflate_hc.cuh for streams
    // that are runs / short periods, deflate_hcs.cuh for the rest).  klass[i] is written by
    // deflate_classify_kernel; a kernel takes the streams whose class equals `want` (NULL: all).
    const uint8_t *klass = nullptr;
    uint8_t want = 0;
    int nos_depth = 0;       // levels 10..12: chain depth override for experiments (BDF_NOS_DEPTH), 0 = the level's own
    int l1_window = 11;      // level 1: bit 0 = whole-window rounds, bits 1 / 2 = next window's buckets prefetched into L1 / L2, bit 3 = also in the block-split path (deflate_l1.cuh; BDF_L1_WINDOW)
};
constexpr unsigned UNIT_FINISH = 1u, UNIT_SYNC = 2u, UNIT_CAP5 = 4u;   // CAP5: 5 more bytes of room (DeflateEncoder, src/stream.rs:66-69)
__host__ __device__ inline uint64_t unit_cap(uint64_t len, unsigned flags) { return len + (len / 65535 + 1) * 5 + 10 + ((flags & 4u) ? 5 : 0); }
__device__ __forceinline__ unsigned unit_flags_of(const DeflateArgs &a, uint64_t idx)
{
    return a.unit_flags ? a.unit_flags[idx] : UNIT_FINISH;
}

__host__ __device__ inline uint64_t deflate_bound(uint64_t len) { return len + (len / 65535 + 1) * 5 + 10; }

// ---- framing (compress_zlib / compress_gzip, src/compress/mod.rs:2248-2357); warp-uniform
__device__ __forceinline__ unsigned frame_header(int format, int level, uint8_t *out, unsigned lane)
{
    if (format == BDF_ZLIB) {
        unsigned hint = level < 2 ? 0 : level < 6 ? 1 : level < 8 ? 2 : 3;
        unsigned hdr = (8u << 8) | (7u << 12) | (hint << 6);
        hdr |= 31 - (hdr % 31);
        if (lane == 0) { out[0] = (uint8_t)(hdr >> 8); out[1] = (uint8_t)hdr; }
        return 2;
    }
    if (format == BDF_GZIP) {
        if (lane < 10) {
            uint8_t b = 0;
            if (lane == 0) b = 0x1F;
            else if (lane == 1) b = 0x8B;
            else if (lane == 2) b = 8;
            else if (lane == 8) b = level < 2 ? 4 : level >= 8 ? 2 : 0;
            else if (lane == 9) b = 255;
            out[lane] = b;
        }
        return 10;
    }
    return 0;
}
// Writes the footer after `at` bytes; returns the framed size.  One warp.
__device__ __forceinline__ uint64_t frame_footer(int format, const uint8_t *in, uint64_t len, uint8_t *out,
                                                 uint64_t at, const uint32_t (*s_crc)[256],
                                                 const uint32_t *s_x2n, unsigned lane)
{
    if (format == BDF_ZLIB) {
        uint32_t a = warp_adler32(in, len, lane);
        if (lane < 4) out[at + lane] = (uint8_t)(a >> (24 - 8 * lane));     // big-endian
        return at + 4;
    }
    if (format == BDF_GZIP) {
        uint32_t c = warp_crc32(in, len, s_crc, s_x2n, lane);
        uint32_t isz = (uint32_t)len;
        if (lane < 4) out[at + lane] = (uint8_t)(c >> (8 * lane));
        else if (lane < 8) out[at + lane] = (uint8_t)(isz >> (8 * (lane - 4)));
        return at + 8;
    }
    return at;
}
__device__ __forceinline__ void load_crc_tables_to_smem(uint32_t (*s_crc)[256], uint32_t *s_x2n)
{
    for (unsigned i = threadIdx.x; i < 1024; i += blockDim.x) s_crc[i >> 8][i & 255] = g_crc_tables.slice[i >> 8][i & 255];
    if (threadIdx.x < 32) s_x2n[threadIdx.x] = g_crc_tables.x2n[threadIdx.x];
    __syncthreads();
}

// ---- symbol geometry
__device__ __forceinline__ unsigned length_slot_of(unsigned len)      // LENGTH_WRITE_TABLE[len] >> 24
{
    if (len <= 10) return len - 3;
    if (len == 258) return 28;
    unsigned v = len - 3;
    unsigned l = 31u - __clz(v);           // >= 3
    return 4 * (l - 1) + ((v >> (l - 2)) & 3u);
}
__device__ __forceinline__ unsigned offset_slot_of(unsigned off)      // get_offset_slot, src/compress/mod.rs:2197-2207
{
    if (off <= 2) return off - 1;
    unsigned v = off - 1;
    unsigned l = 31u - __clz(v);
    return 2 * l + ((v >> (l - 1)) & 1u);
}
__device__ __forceinline__ uint32_t ld24(const uint8_t *p)
{
    return p[0] | (uint32_t)p[1] << 8 | (uint32_t)p[2] << 16;
}
__device__ __forceinline__ uint32_t hash3(uint32_t v24) { return (v24 * 0x1E35A7BDu) >> 17; }

// ---- warp bit sink: LSB-first bit stream (Bitstream, src/compress/bitstream.rs).
// Every lane contributes (bits, n <= 32) in lane order; bit offsets come from a
// warp inclusive scan, bits are OR-ed into a zeroed shared staging buffer, and
// whole words are flushed to global memory.  Capacity is enforced the way the
// reference's checked writes do: the stream fails as soon as a completed byte
// does not fit (bitstream.rs:143-189,194-222).
constexpr uint32_t SINK_WORDS = 512;                 // 2 KiB staging per sink
constexpr uint32_t SINK_FLUSH_BITS = 8192;           // flush once 1 KiB is pending

// COUNT = true: size estimation, nothing is stored and nothing can overflow.
template <bool COUNT>
struct BitSinkT {
    uint32_t *buf;      // shared, SINK_WORDS, zero-filled outside [0, nbits)
    uint8_t *out;       // global destination
    uint64_t cap;       // bytes available at out
    uint64_t flushed;   // bytes already stored
    uint32_t nbits;     // bits pending in buf
    bool overflow;

    __device__ __forceinline__ void init(uint32_t *staging, uint8_t *dst, uint64_t capacity, unsigned lane)
    {
        buf = staging; out = dst; cap = capacity; flushed = 0; nbits = 0; overflow = false;
        for (unsigned i = lane; i < SINK_WORDS; i += 32) buf[i] = 0;
        __syncwarp();
    }
    __device__ __forceinline__ void flush_words(unsigned lane)
    {
        const uint32_t nfull = nbits >> 5;
        const uint64_t bytes = (uint64_t)nfull * 4;
        __syncwarp();
        if (flushed + bytes > cap) overflow = true;
        if (!COUNT && !overflow) {
            for (uint32_t w = lane; w < nfull; w += 32) {
                uint32_t v = buf[w];
                uint8_t *d = out + flushed + 4ull * w;
                d[0] = (uint8_t)v; d[1] = (uint8_t)(v >> 8); d[2] = (uint8_t)(v >> 16); d[3] = (uint8_t)(v >> 24);
            }
        }
        const uint32_t rem = buf[nfull];
        __syncwarp();
        for (uint32_t w = lane; w <= nfull; w += 32) buf[w] = 0;
        __syncwarp();
        if (lane == 0) buf[0] = rem;
        __syncwarp();
        flushed += bytes;
        nbits &= 31u;
    }
    // all 32 lanes call; lanes with n == 0 contribute nothing
    __device__ __forceinline__ void put(uint32_t bits, uint32_t n, unsigned lane)
    {
        uint32_t incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(BDF_FULL_MASK, incl, d);
            if (lane >= (unsigned)d) incl += t;
        }
        const uint32_t total = __shfl_sync(BDF_FULL_MASK, incl, 31);
        if (n) {
            const uint32_t at = nbits + incl - n;
            const uint32_t w = at >> 5, sh = at & 31u;
            atomicOr(&buf[w], bits << sh);
            if (sh + n > 32) atomicOr(&buf[w + 1], bits >> (32 - sh));
        }
        nbits += total;
        __syncwarp();
        if (nbits >= SINK_FLUSH_BITS) flush_words(lane);
    }
    // single (bits, n) item from a warp-uniform caller
    __device__ __forceinline__ void put1(uint32_t bits, uint32_t n, unsigned lane)
    {
        if (n == 0) return;
        if (lane == 0) {
            const uint32_t w = nbits >> 5, sh = nbits & 31u;
            buf[w] |= bits << sh;
            if (sh + n > 32) buf[w + 1] |= bits >> (32 - sh);
        }
        nbits += n;
        __syncwarp();
        if (nbits >= SINK_FLUSH_BITS) flush_words(lane);
    }
    // bit position of the next bit, counted from the start of the stream
    __device__ __forceinline__ uint64_t bitpos() const { return flushed * 8 + nbits; }
    // sets one bit that has already been put (a BFINAL bit that is only known at the end of its
    // block); the byte is either still in the staging buffer or already in global memory
    __device__ __forceinline__ void set_bit(uint64_t at, unsigned lane)
    {
        __syncwarp();
        if (lane == 0 && !overflow && (!COUNT || at >= flushed * 8)) {
            if (at >= flushed * 8) {
                const uint32_t r = (uint32_t)(at - flushed * 8);
                buf[r >> 5] |= 1u << (r & 31u);
            } else {
                out[at >> 3] |= (uint8_t)(1u << (at & 7u));
            }
        }
        __syncwarp();
    }
    // FlushMode::Sync, src/compress/mod.rs:662-681: an empty stored block (000, pad to a byte,
    // LEN = 0x0000, NLEN = 0xFFFF)
    __device__ __forceinline__ void sync_marker(unsigned lane)
    {
        put1(0, 3, lane);
        nbits = (nbits + 7u) & ~7u;            // the staging buffer is zero outside [0, nbits)
        put1(0xFFFF0000u, 32, lane);
    }
    // Bitstream::flush: zero-pad to a byte and store everything.  Returns total bytes or ~0 on overflow.
    __device__ __forceinline__ uint64_t finish(unsigned lane)
    {
        __syncwarp();
        const uint32_t bytes = (nbits + 7) >> 3;
        if (flushed + bytes > cap) overflow = true;
        if (!COUNT && !overflow) {
            for (uint32_t b = lane; b < bytes; b += 32) out[flushed + b] = (uint8_t)(buf[b >> 2] >> (8 * (b & 3)));
        }
        __syncwarp();
        return overflow ? ~0ull : flushed + bytes;
    }
};

using BitSink = BitSinkT<false>;

// 8 bytes at any alignment from the aligned words that contain them (a word without a requested
// byte is never touched)
__device__ __forceinline__ uint64_t ld64_any(const uint8_t *p)
{
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p - sh);
    const uint32_t w0 = w[0], w1 = w[1];
    if (sh == 0) return (uint64_t)w1 << 32 | w0;
    const uint32_t w2 = w[2];
    return (uint64_t)__funnelshift_r(w1, w2, 8 * sh) << 32 | __funnelshift_r(w0, w1, 8 * sh);
}

// Common-prefix length of a[0..maxlen) and b[0..maxlen), maxlen <= 258, all lanes cooperate
// (every match_len_* variant of the reference, src/compress/matchfinder.rs:245-694).  Each lane
// compares 8 bytes with one 64-bit XOR (byte-wise only in the last, partial group of eight).
__device__ __forceinline__ unsigned warp_match_len(const uint8_t *a, const uint8_t *b, unsigned maxlen, unsigned lane)
{
    for (unsigned base = 0; base < maxlen; base += 256) {
        const unsigned i0 = base + lane * 8;
        unsigned cnt = 0;
        if (i0 + 8 <= maxlen) {
            const uint64_t x = ld64_any(a + i0) ^ ld64_any(b + i0);
            cnt = x ? (unsigned)(__ffsll((long long)x) - 1) >> 3 : 8u;
        } else {
#pragma unroll
            for (unsigned k = 0; k < 7; k++) {
                unsigned idx = i0 + k;
                if (cnt == k && idx < maxlen && a[idx] == b[idx]) cnt++;
            }
        }
        const unsigned stop = __ballot_sync(BDF_FULL_MASK, cnt != 8);
        if (stop) {
            const unsigned first = __ffs(stop) - 1;
            return base + first * 8 + __shfl_sync(BDF_FULL_MASK, cnt, first);
        }
    }
    return maxlen;
}

// ---- length-limited canonical Huffman code (make_huffman_code and helpers,
// src/compress/huffman_comp.rs:8-155).  Serial, run by one thread on shared
// arrays: a[] (num_syms words, doubles as the codeword output), lens[], and
// cnt[] (num_syms words of scratch).  Ties are broken exactly as in the
// reference: counting sort on min(freq, n-1), last bucket ordered by the
// packed (freq << 10 | sym) key.
__device__ void make_huffman_code_serial(unsigned num_syms, unsigned max_len, const uint32_t *freqs,
                                         uint8_t *lens, uint32_t *a, uint32_t *cnt)
{
    const uint32_t SYM_MASK = 1023u, FREQ_MASK = ~1023u;
    for (unsigned i = 0; i < num_syms; i++) cnt[i] = 0;
    for (unsigned s = 0; s < num_syms; s++) {
        uint32_t f = freqs[s];
        cnt[f < num_syms - 1 ? f : num_syms - 1]++;
    }
    uint32_t run = 0;
    for (unsigned i = 1; i < num_syms; i++) {
        uint32_t c = cnt[i];
        cnt[i] = run;
        run += c;
    }
    const unsigned used = run;
    for (unsigned s = 0; s < num_syms; s++) {
        uint32_t f = freqs[s];
        if (f) {
            unsigned b = f < num_syms - 1 ? f : num_syms - 1;
            a[cnt[b]++] = s | (f << 10);
        } else {
            lens[s] = 0;
        }
    }
    {   // order the overflow bucket by the packed key (insertion sort; keys are distinct)
        const unsigned lo = cnt[num_syms - 2], hi = cnt[num_syms - 1];
        for (unsigned i = lo + 1; i < hi; i++) {
            uint32_t key = a[i];
            unsigned j = i;
            while (j > lo && a[j - 1] > key) { a[j] = a[j - 1]; j--; }
            a[j] = key;
        }
    }
    if (used < 2) {
        unsigned sym = used ? (a[0] & SYM_MASK) : 0;
        unsigned nz = sym ? sym : 1;
        a[0] = 0; lens[0] = 1;
        a[nz] = 1; lens[nz] = 1;
        return;
    }
    {
        const unsigned last = used - 1;
        unsigned i = 0, b = 0, e = 0;
        while (e < last) {
            uint32_t nf;
            if (i < last && (b == e || (a[i + 1] & FREQ_MASK) <= (a[b] & FREQ_MASK))) {
                nf = (a[i] & FREQ_MASK) + (a[i + 1] & FREQ_MASK);
                i += 2;
            } else if (b + 2 <= e && (i > last || (a[b + 1] & FREQ_MASK) < (a[i] & FREQ_MASK))) {
                nf = (a[b] & FREQ_MASK) + (a[b + 1] & FREQ_MASK);
                a[b] = (e << 10) | (a[b] & SYM_MASK);
                a[b + 1] = (e << 10) | (a[b + 1] & SYM_MASK);
                b += 2;
            } else {
                nf = (a[i] & FREQ_MASK) + (a[b] & FREQ_MASK);
                a[b] = (e << 10) | (a[b] & SYM_MASK);
                i += 1;
                b += 1;
            }
            a[e] = nf | (a[e] & SYM_MASK);
            e++;
        }
    }
    uint32_t len_counts[16];
#pragma unroll
    for (int l = 0; l < 16; l++) len_counts[l] = 0;
    {
        const unsigned root = used - 2;
        len_counts[1] = 2;
        a[root] &= SYM_MASK;
        for (int node = (int)root - 1; node >= 0; node--) {
            unsigned parent = a[node] >> 10;
            unsigned depth = (a[parent] >> 10) + 1;
            a[node] = (a[node] & SYM_MASK) | (depth << 10);
            if (depth >= max_len) {
                depth = max_len - 1;
                while (len_counts[depth] == 0) depth--;
            }
            len_counts[depth]--;
            len_counts[depth + 1] += 2;
        }
    }
    {
        unsigned i = 0;
        for (unsigned len = max_len; len >= 1; len--)
            for (uint32_t c = len_counts[len]; c > 0; c--) lens[a[i++] & SYM_MASK] = (uint8_t)len;
        uint32_t next[16];
        next[0] = 0; next[1] = 0;
        for (unsigned len = 2; len <= max_len; len++) next[len] = (next[len - 1] + len_counts[len - 1]) << 1;
        for (unsigned s = 0; s < num_syms; s++) {
            unsigned l = lens[s];
            if (l) a[s] = __brev(next[l]++) >> (32 - l);
        }
    }
}


// The same code construction by a whole CTA (all threads call it; it contains barriers).  The
// order the reference's counting sort + overflow-bucket sort produces is simply ascending
// (frequency, symbol) over the used symbols, so every symbol finds its slot by counting the
// symbols that come before it; the in-place tree build and the depth pass stay with one thread
// (they are chains of dependent steps), code lengths and codewords are assigned by all threads
// again.  Same results as make_huffman_code_serial, bit for bit.
//   a[] / lens[] as above; ctl: 20 words of shared scratch.
__device__ void make_huffman_code_cta(unsigned num_syms, unsigned max_len, const uint32_t *freqs,
                                      uint8_t *lens, uint32_t *a, uint32_t *ctl)
{
    const uint32_t SYM_MASK = 1023u, FREQ_MASK = ~1023u;
    const unsigned tid = threadIdx.x;
    if (tid == 0) ctl[16] = 0;
    __syncthreads();
    uint32_t f = 0;
    if (tid < num_syms) {
        f = freqs[tid];
        if (f) {
            unsigned rank = 0;
            for (unsigned t = 0; t < num_syms; t++) {
                const uint32_t g = freqs[t];
                rank += (g != 0 && (g < f || (g == f && t < tid))) ? 1u : 0u;
            }
            a[rank] = tid | (f << 10);
            atomicAdd(&ctl[16], 1u);
        } else {
            lens[tid] = 0;
        }
    }
    __syncthreads();
    const unsigned used = ctl[16];
    if (used < 2) {
        if (tid == 0) {
            unsigned sym = used ? (a[0] & SYM_MASK) : 0;
            unsigned nz = sym ? sym : 1;
            a[0] = 0; lens[0] = 1;
            a[nz] = 1; lens[nz] = 1;
        }
        __syncthreads();
        return;
    }
    if (tid == 0) {
        {
            const unsigned last = used - 1;
            unsigned i = 0, b = 0, e = 0;
            while (e < last) {
                uint32_t nf;
                if (i < last && (b == e || (a[i + 1] & FREQ_MASK) <= (a[b] & FREQ_MASK))) {
                    nf = (a[i] & FREQ_MASK) + (a[i + 1] & FREQ_MASK);
                    i += 2;
                } else if (b + 2 <= e && (i > last || (a[b + 1] & FREQ_MASK) < (a[i] & FREQ_MASK))) {
                    nf = (a[b] & FREQ_MASK) + (a[b + 1] & FREQ_MASK);
                    a[b] = (e << 10) | (a[b] & SYM_MASK);
                    a[b + 1] = (e << 10) | (a[b + 1] & SYM_MASK);
                    b += 2;
                } else {
                    nf = (a[i] & FREQ_MASK) + (a[b] & FREQ_MASK);
                    a[b] = (e << 10) | (a[b] & SYM_MASK);
                    i += 1;
                    b += 1;
                }
                a[e] = nf | (a[e] & SYM_MASK);
                e++;
            }
        }
        uint32_t len_counts[16];
#pragma unroll
        for (int l = 0; l < 16; l++) len_counts[l] = 0;
        const unsigned root = used - 2;
        len_counts[1] = 2;
        a[root] &= SYM_MASK;
        for (int node = (int)root - 1; node >= 0; node--) {
            unsigned parent = a[node] >> 10;
            unsigned depth = (a[parent] >> 10) + 1;
            a[node] = (a[node] & SYM_MASK) | (depth << 10);
            if (depth >= max_len) {
                depth = max_len - 1;
                while (len_counts[depth] == 0) depth--;
            }
            len_counts[depth]--;
            len_counts[depth + 1] += 2;
        }
#pragma unroll
        for (int l = 0; l < 16; l++) ctl[l] = len_counts[l];
    }
    __syncthreads();
    // sorted position i gets the longest lengths first
    if (tid < used) {
        unsigned acc = 0, L = 0;
        for (unsigned l = max_len; l >= 1; l--) {
            acc += ctl[l];
            if (L == 0 && tid < acc) L = l;
        }
        lens[a[tid] & SYM_MASK] = (uint8_t)L;
    }
    __syncthreads();
    if (tid < num_syms) {
        const unsigned l = lens[tid];
        if (l) {
            uint32_t nx = 0;                                 // first codeword of length l
            for (unsigned k = 2; k <= l; k++) nx = (nx + ctl[k - 1]) << 1;
            unsigned rank = 0;
            for (unsigned t = 0; t < tid; t++) rank += lens[t] == l ? 1u : 0u;
            a[tid] = __brev(nx + rank) >> (32 - l);
        }
    }
    __syncthreads();
}

}  // namespace bdf
