// inflate.cuh — batch DEFLATE decompression, one LANE GROUP per stream (sm_100a).
//
// Replaces Decompressor::decompress / decompress_zlib / decompress_gzip as
// called from BatchDecompressor::decompress_batch (reference src/batch.rs:74-101,
// src/decompress/mod.rs:164-202,1074-1240).  Accept/reject rules for Huffman
// code sets follow build_decode_table (src/decompress/mod.rs:1365-1383); the
// table SHAPE is our own: a 9-bit direct litlen table and a 7-bit offset table
// in shared memory, with codewords longer than the table resolved by a
// canonical first-code search (no sub-tables), all built lane-parallel.
//
// Execution model.  Profiling the first version (one warp per stream) showed
// the kernel bound by instruction issue, not by memory: every instruction
// served one stream.  A warp is therefore cut into 32/G groups of G lanes and
// each group owns a stream: the G lanes hold identical bit-reader state and
// decode the same symbol (table entries are shared-memory broadcasts, input
// words are uniform loads), and the groups of a warp run the same code in
// lock-step whenever their streams agree on the kind of the next symbol, so
// one warp instruction advances up to 32/G streams.  Literals are parked one
// per lane and flushed as one store; a match is copied in 16-byte,
// destination-aligned chunks (one per lane per round) whose source run is
// fetched as aligned words and realigned with PRMT.  Adler-32 is accumulated
// from the bytes as they are written (position-weighted partial sums per
// lane, dp4a), so the zlib path never re-reads its output.
#pragma once
#include <cstddef>
#include "common.cuh"

namespace bdf {

constexpr int LT_BITS = 9;    // litlen direct-table bits
constexpr int OT_BITS = 7;    // offset direct-table bits (the precode table, 7 bits, overlays it)
constexpr int PT_BITS = 7;    // precode direct-table bits

// First-block headers decoded ahead of the inflate kernels (inflate_prehdr.cuh): per stream a row
// of code lengths and a meta word: [15:0] bits of DEFLATE data up to the end of the header,
// [20:16] nlit - 257, [25:21] noff - 1, [26] final-block bit, [31] valid.
constexpr int PREHDR_ROW_BYTES = 320, PREHDR_ROW_WORDS = PREHDR_ROW_BYTES / 4;
constexpr uint32_t PREHDR_VALID = 1u << 31;

// Table entry (u16 — 32-bit entries cost 1.3 KiB more shared memory per stream, i.e. two CTAs per
// SM): [3:0] codeword bits (0 = longer than the table), [5:4] kind, [6] literal,
// [15:7] literal byte / length slot / offset slot / precode symbol.  Base value and extra-bit
// count of a length or offset slot are computed where a match is decoded.
constexpr uint32_t K_LIT = 0u << 4, K_BASE = 1u << 4, K_EOB = 2u << 4, K_MASK = 3u << 4;
constexpr uint32_t LITFLAG = 1u << 6;     // set in literal entries only: one test selects the literal fast path
constexpr uint32_t E_LEN = 15u;           // mask of the codeword length
constexpr int E_VAL = 7;                  // shift of the value

struct HuffCode {            // canonical description used for build + long codes
    uint16_t first[16];      // first codeword of each length (MSB-first value)
    uint16_t count[16];
    uint16_t offs[16];       // index of the first symbol of each length in sorted[]
};

// Scratch of the table builder.  rows: one row of 16 per-length counters per lane (u16, padded to a
// 9-word stride so that the rows of a group fall into different banks); big_*: direct-table fills
// of 32 slots or more, deferred so that the whole group can do them together.
constexpr int ROW_STRIDE = 18;
template <int G>
struct BuildScratch {
    uint32_t cnt[16];                  // also the 32-byte period buffer of copy_match_part (mode 3)
    uint32_t big_e[16];
    uint16_t big_r[16];
    uint32_t nbig;
    uint16_t rows[G * ROW_STRIDE];
};

template <int G>
struct __align__(16) InflateSmem {     // one per lane group
    uint16_t lit_tab[1 << LT_BITS];
    uint16_t off_tab[1 << OT_BITS];
    uint16_t lit_sorted[288];
    uint16_t off_sorted[32];
    HuffCode lit_code, off_code;
    BuildScratch<G> bs;
    uint8_t lens[320 + 8];
};

// Block buffer of copy_match (see there): the builder scratch + lens, or everything from
// lit_sorted on when the block has no long codewords.  Both end at the end of lens.
template <int G> constexpr uint32_t BLOCK_BUF_SCRATCH = (uint32_t)(offsetof(InflateSmem<G>, lens) + 328 - offsetof(InflateSmem<G>, bs)) & ~15u;
template <int G> constexpr uint32_t BLOCK_BUF_ALL = (uint32_t)(offsetof(InflateSmem<G>, lens) + 328 - offsetof(InflateSmem<G>, lit_sorted)) & ~15u;

// ------------------------------------------------------------------ bit reader
// Group-uniform LSB-first reader over [p, p+len) using aligned 32-bit loads.
struct BitReader {
    const uint8_t *p;      // stream start
    uint32_t len;          // stream length in bytes
    uint32_t mis;          // p & 3
    uint32_t nwords;       // aligned words covering the stream
    uint32_t widx;         // next word to append to buf
    uint32_t ahead;        // word widx, already loaded (its latency hides behind decoding)
    uint64_t buf;
    int32_t left;          // valid bits in buf (may include zero fill past the end)

    __device__ __forceinline__ uint32_t load_word(uint32_t w) const
    {
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(p - mis) + w;
        if (w - 1u < nwords - 2u) return __ldg(wp);     // interior word (0 < w < nwords-1)
        if (w >= nwords) return 0;                     // zero fill past the end
        // first / last word: only touch bytes that belong to the stream
        uint32_t v = 0;
        int64_t b0 = (int64_t)w * 4 - mis;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int64_t bi = b0 + k;
            if (bi >= 0 && bi < (int64_t)len) v |= (uint32_t)__ldg(p + bi) << (8 * k);
        }
        return v;
    }
    __device__ __forceinline__ void init(const uint8_t *ptr, uint32_t n)
    {
        p = ptr; len = n;
        mis = (uint32_t)(reinterpret_cast<uintptr_t>(ptr) & 3u);
        nwords = (mis + n + 3) >> 2;
        buf = (uint64_t)load_word(0) >> (8 * mis);
        left = 32 - 8 * (int32_t)mis;
        widx = 1;
        ahead = load_word(1);
    }
    // the stream without loading anything: seek_bits() / seek() start the reader
    __device__ __forceinline__ void attach(const uint8_t *ptr, uint32_t n)
    {
        p = ptr; len = n;
        mis = (uint32_t)(reinterpret_cast<uintptr_t>(ptr) & 3u);
        nwords = (mis + n + 3) >> 2;
        buf = 0; left = 0; widx = 0; ahead = 0;
    }
    // keep at least 33 valid bits
    __device__ __forceinline__ void refill()
    {
        if (left <= 32) {
            buf |= (uint64_t)ahead << left;
            widx++;
            left += 32;
            ahead = load_word(widx);
        }
    }
    __device__ __forceinline__ uint32_t peek(unsigned n) const { return (uint32_t)buf & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(unsigned n) { buf >>= n; left -= (int32_t)n; }
    __device__ __forceinline__ uint32_t take(unsigned n) { uint32_t v = peek(n); drop(n); return v; }
    // bits consumed from the stream so far (can exceed 8*len when reading zero fill)
    __device__ __forceinline__ int64_t consumed_bits() const
    {
        return (int64_t)widx * 32 - 8 * (int64_t)mis - left;
    }
    __device__ __forceinline__ bool overrun() const { return consumed_bits() > (int64_t)len * 8; }
    // restart at bit `bit` of the stream (consumed_bits() == bit afterwards)
    __device__ __forceinline__ void seek_bits(int64_t bit)
    {
        const uint64_t a = 8ull * mis + (uint64_t)bit;
        widx = (uint32_t)(a >> 5);
        const uint32_t sh = (uint32_t)a & 31u;
        buf = (uint64_t)load_word(widx) >> sh;
        left = 32 - (int32_t)sh;
        widx++;
        ahead = load_word(widx);
    }
    // restart at byte offset `at` from the stream start
    __device__ __forceinline__ void seek(uint32_t at)
    {
        uint32_t a = mis + at;
        widx = a >> 2;
        uint32_t sh = 8 * (a & 3);
        buf = (uint64_t)load_word(widx) >> sh;
        left = 32 - (int32_t)sh;
        widx++;
        ahead = load_word(widx);
    }
};

// ------------------------------------------------------------ table building
__device__ __forceinline__ uint32_t make_litlen_entry(unsigned sym, unsigned l)
{
    if (sym < 256) return (sym << E_VAL) | LITFLAG | K_LIT | l;
    if (sym == 256) return K_EOB | l;
    return ((sym - 257) << E_VAL) | K_BASE | l;         // slots 29 / 30 (symbols 286 / 287) decode as 258
}
__device__ __forceinline__ uint32_t make_offset_entry(unsigned sym, unsigned l) { return (sym << E_VAL) | K_BASE | l; }
__device__ __forceinline__ uint32_t make_precode_entry(unsigned sym, unsigned l) { return (sym << E_VAL) | l; }
// match length / offset of a decoded slot entry: base + the extra bits that follow the codeword
__device__ __forceinline__ unsigned take_length(BitReader &br, uint32_t e)
{
    unsigned base, extra;
    length_slot_info(e >> E_VAL, base, extra);
    return base + br.take(extra);
}
__device__ __forceinline__ unsigned take_offset(BitReader &br, uint32_t f)
{
    unsigned base, extra;
    offset_slot_info(f >> E_VAL, base, extra);
    return base + br.take(extra);
}

enum { CODE_LITLEN = 0, CODE_OFFSET = 1, CODE_PRECODE = 2 };

template <int KIND>
__device__ __forceinline__ uint32_t make_entry(unsigned sym, unsigned l)
{
    if (KIND == CODE_LITLEN) return make_litlen_entry(sym, l);
    if (KIND == CODE_OFFSET) return make_offset_entry(sym, l);
    return make_precode_entry(sym, l);
}

// Builds the direct table + canonical description for `nsyms` code lengths.
// Returns false (uniformly over the group) for code sets the reference
// rejects: over-subscribed, or incomplete other than "no codes" / "one 1-bit
// code" (src/decompress/mod.rs:1365-1383).
//
// Lane j owns the symbols [j*per, (j+1)*per) and, as "length owner", the code length j.  Pass 1:
// every lane counts the lengths of its symbols into its row.  The length owners then turn their
// column into exclusive prefixes (= canonical rank of a lane's first symbol of that length) and
// derive first codeword / sorted offset of their length with two shuffle scans.  Pass 2: every
// lane walks its symbols again; rank -> codeword -> the symbol writes its own table slots.  No
// group synchronisation inside either pass.
template <int KIND, int TBITS, int G>
__device__ bool build_code(const Grp<G> &g, const uint8_t *lens, unsigned nsyms, uint16_t *tab,
                           uint16_t *sorted, HuffCode &hc, BuildScratch<G> &bs, uint32_t *nlong = nullptr)
{
    static_assert(G >= 16, "one length owner per code length");
    constexpr unsigned BIG = TBITS >= 5 ? TBITS - 5 : 0;      // codewords of <= BIG bits fill >= 32 slots
    uint16_t *row = bs.rows + g.lane * ROW_STRIDE;
    {
        uint32_t *rw = reinterpret_cast<uint32_t *>(row);      // 36-byte stride: 4-byte aligned
#pragma unroll
        for (int k = 0; k < 8; k++) rw[k] = 0;
    }
    const unsigned per = (nsyms + G - 1) / G;
    const unsigned s0 = g.lane * per;
    const unsigned s1 = s0 + per < nsyms ? s0 + per : nsyms;
    for (unsigned s = s0; s < s1; s++) row[lens[s]]++;
    if (g.lane == 0) bs.nbig = 0;
    g.sync();
    const unsigned L = g.lane;
    const bool owner = L >= 1 && L < 16;
    uint32_t tot = 0;
    if (owner) {
#pragma unroll 4
        for (int j = 0; j < G; j++) {
            const uint32_t c = bs.rows[j * ROW_STRIDE + L];
            bs.rows[j * ROW_STRIDE + L] = (uint16_t)tot;
            tot += c;
        }
    }
    // inclusive scans over the lengths: Kraft sum in units of 2^-15, and symbol count
    const uint32_t v = owner ? tot << (15 - L) : 0;
    uint32_t iv = v, ic = tot;
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
        const uint32_t x = g.shfl_up(iv, d), y = g.shfl_up(ic, d);
        if (g.lane >= (unsigned)d) { iv += x; ic += y; }
    }
    const uint32_t used = g.shfl(iv, 15), total = g.shfl(ic, 15), c1 = g.shfl(tot, 1);
    if (nlong) *nlong = total - g.shfl(ic, TBITS);       // codewords longer than the direct table
    if (owner) {
        hc.first[L] = (uint16_t)((iv - v) >> (15 - L));
        hc.count[L] = (uint16_t)tot;
        hc.offs[L] = (uint16_t)(ic - tot);
    }
    g.sync();
    if (used > (1u << 15)) return false;
    if (used < (1u << 15)) {
        if (!(total == 0 || (total == 1 && c1 == 1))) return false;
        // every lookup yields symbol 0 (or the lone symbol) with a 1-bit codeword
        unsigned sym = 0;
        if (total == 1) {
            for (unsigned base = 0; base < nsyms; base += G) {
                unsigned s = base + g.lane;
                unsigned hit = g.ballot(s < nsyms && lens[s] == 1);
                if (hit) sym = base + (__ffs(hit) - 1);
            }
        }
        const uint32_t e = make_entry<KIND>(sym, 1);
        for (unsigned i = g.lane; i < (1u << TBITS); i += G) tab[i] = (uint16_t)e;
        for (unsigned i = g.lane; i < 16; i += G) hc.count[i] = 0;   // no long codes
        g.sync();
        return true;
    }
    // A codeword of l <= TBITS bits owns the 2^(TBITS-l) slots whose low l bits are its reversed
    // bits; a longer codeword marks its TBITS-bit prefix slot with 0 (-> canonical search).
    for (unsigned s = s0; s < s1; s++) {
        const unsigned l = lens[s];
        if (l == 0) continue;
        const unsigned rank = row[l];
        row[l] = (uint16_t)(rank + 1);
        sorted[hc.offs[l] + rank] = (uint16_t)s;
        const unsigned rev = __brev((unsigned)hc.first[l] + rank) >> (32 - l);
        if (l <= TBITS) {
            const uint32_t e = make_entry<KIND>(s, l);
            if (l <= BIG) {
                const unsigned k = atomicAdd(&bs.nbig, 1u);    // at most 2^BIG <= 16 such codewords
                bs.big_e[k] = e;
                bs.big_r[k] = (uint16_t)(rev | l << 12);
            } else {
                for (unsigned i = rev; i < (1u << TBITS); i += 1u << l) tab[i] = (uint16_t)e;
            }
        } else {
            tab[rev & ((1u << TBITS) - 1u)] = 0;
        }
    }
    g.sync();
    const unsigned nbig = bs.nbig;
    for (unsigned k = 0; k < nbig; k++) {
        const uint32_t e = bs.big_e[k];
        const unsigned r = bs.big_r[k], l = r >> 12;
        for (unsigned i = (r & 0xFFFu) + (g.lane << l); i < (1u << TBITS); i += (unsigned)G << l) tab[i] = (uint16_t)e;
    }
    g.sync();
    return true;
}

// Codeword longer than the direct table: canonical search on the next 15 bits.
template <int KIND, int TBITS>
__device__ __forceinline__ uint32_t decode_long(uint32_t bits15, const uint16_t *sorted, const HuffCode &hc)
{
    unsigned x = __brev(bits15) >> 17;
#pragma unroll 1
    for (unsigned l = TBITS + 1; l <= 15; l++) {
        unsigned d = (x >> (15 - l)) - hc.first[l];
        if (d < hc.count[l]) return make_entry<KIND>(sorted[hc.offs[l] + d], l);
    }
    return 0;   // not a codeword (only reachable with an incomplete code)
}

// ------------------------------------------------------------------- output
struct OutState {
    uint8_t *out;
    uint32_t pos, cap;
    uint32_t npend;        // literals parked in lanes 0..npend-1 of the group
    uint32_t mylit;
    uint32_t sumA;         // Σ b           over bytes written by this lane
    uint64_t sumB;         // Σ i·b  (i = 0-based output index)
    uint32_t next_fold;    // output position at which the sums are folded mod 65521
    uint32_t zfill;        // output sectors below this offset are fully valid in L2 (see make_valid)
    uint32_t zlimit;       // last offset up to which whole 32-byte sectors belong to this stream
    uint32_t blkcap;       // bytes of shared memory copy_match may use as its block buffer (see InflateSmem)
};

// Reading a sector of which only some bytes have been written makes L2 fetch the
// rest from DRAM; a match source always ends at the write position, so without
// care every match waits for a DRAM round trip (measured: ~18 KiB of DRAM reads
// per 64 KiB stream and a third of all stall samples on the source loads).
// The group therefore zero-fills whole sectors ahead of the write position:
// full-sector writes need no fill, later partial writes merge in L2, and the
// source loads become L2 hits.  DRAM still sees each line written once.
constexpr uint32_t ZFILL_AHEAD = 1024;
template <int G>
__device__ __forceinline__ void make_valid(const Grp<G> &g, OutState &o, uint32_t upto)
{
    if (upto > o.zfill && o.zfill < o.zlimit) {
        // zfill and zlimit are offsets x with (out + x) 32-byte aligned.  Never touch a sector
        // that already holds output: start at the first boundary at or after the write position
        // (stored blocks and end-of-block flushes can move pos past zfill).
        uint32_t zbeg = o.zfill;
        if (o.pos > zbeg) zbeg += (o.pos - zbeg + 31u) & ~31u;
        uint32_t zend = zbeg + ((upto + ZFILL_AHEAD - zbeg + 31u) & ~31u);
        if (upto + ZFILL_AHEAD <= zbeg) zend = zbeg;
        if (zend > o.zlimit) zend = o.zlimit;
        for (uint32_t x = zbeg + 16 * g.lane; x < zend; x += 16 * G)
            *reinterpret_cast<uint4 *>(o.out + x) = make_uint4(0, 0, 0, 0);
        o.zfill = zend > o.zfill ? zend : o.zfill;
        g.sync();
    }
}

// One lane can end up writing every byte (short matches always land on the low
// lanes), so fold well before 255 * bytes overflows 32 bits / 2^40 * bytes 64.
constexpr uint32_t ADLER_FOLD_INTERVAL = 1u << 22;
__device__ __forceinline__ void adler_fold(OutState &o)
{
    if (o.pos >= o.next_fold) {
        o.sumA %= 65521u;
        o.sumB %= 65521u;
        o.next_fold = o.pos + ADLER_FOLD_INTERVAL;
    }
}

template <bool ADLER>
__device__ __forceinline__ void adler_acc1(OutState &o, uint32_t idx, uint32_t b)
{
    if (ADLER) { o.sumA += b; o.sumB += (uint64_t)idx * b; }
}
template <bool ADLER>
__device__ __forceinline__ void adler_acc16(OutState &o, uint32_t idx, const uint4 &v)
{
    if (ADLER) {
        uint32_t s = __dp4a(v.x, 0x01010101u, __dp4a(v.y, 0x01010101u, __dp4a(v.z, 0x01010101u, __dp4a(v.w, 0x01010101u, 0u))));
        uint32_t w = __dp4a(v.x, 0x03020100u, __dp4a(v.y, 0x07060504u, __dp4a(v.z, 0x0B0A0908u, __dp4a(v.w, 0x0F0E0D0Cu, 0u))));
        o.sumA += s;
        o.sumB += (uint64_t)idx * s + w;
    }
}

template <bool ADLER>
__device__ __forceinline__ void flush_literals(OutState &o, unsigned lane)
{
    if (o.npend) {
        if (lane < o.npend) {
            BDF_ASSERT(o.pos + lane < o.cap);
            o.out[o.pos + lane] = (uint8_t)o.mylit;
            adler_acc1<ADLER>(o, o.pos + lane, o.mylit);
        }
        o.pos += o.npend;
        o.npend = 0;
    }
}

// ---- match copy ------------------------------------------------------------
// out[pos+i] = out[pos-offset+i], i < length, with the period-replication rule
// for offset < length (src/decompress/mod.rs:1259-1317, x86.rs copy_match_bmi2).
// 16 contiguous bytes at an arbitrarily aligned address (reads the aligned words that cover them)
__device__ __forceinline__ uint4 load16_unaligned(const uint8_t *sp)
{
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(sp) & 3u);
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(sp - sh);
    const uint32_t w0 = sw[0], w1 = sw[1], w2 = sw[2], w3 = sw[3], w4 = sh ? sw[4] : 0u;
    const uint32_t sel = 0x3210u + 0x1111u * sh;
    return make_uint4(__byte_perm(w0, w1, sel), __byte_perm(w1, w2, sel), __byte_perm(w2, w3, sel),
                      __byte_perm(w3, w4, sel));
}
// bytes [0, t) of word w from a, the rest from b, for t relative to the word (may be <= 0 or >= 4)
__device__ __forceinline__ uint32_t merge_word(uint32_t a, uint32_t b, int t)
{
    const int c = t < 0 ? 0 : t > 4 ? 4 : t;
    const uint32_t sel = 0x7654u ^ (0x4444u & ((1u << (4 * c)) - 1u));
    return __byte_perm(a, b, sel);
}

// x mod offset through a float reciprocal (x < 2^17): the quotient estimate is off by at most one
__device__ __forceinline__ uint32_t mod_period(uint32_t x, uint32_t offset, float inv)
{
    uint32_t j = x - (uint32_t)((float)x * inv) * offset;
    if ((int32_t)j < 0) j += offset;
    if (j >= offset) j -= offset;
    return j;
}

// rare path (period shorter than a chunk, or a match at the very start of the output), kept
// out of line so that the hot decode loop stays small
__device__ __noinline__ uint4 chunk_gather(const uint8_t *pat, uint32_t j, uint32_t offset)
{
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int k = 0; k < 16; k++) {
        w[k >> 2] |= (uint32_t)pat[j] << (8 * (k & 3));
        j = j + 1 == offset ? 0 : j + 1;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// Source bytes of one 16-byte chunk at match-relative offset rel (see copy_match).
// mode 0: offset >= length (plain run); 1: periodic, period >= 16 and the 16 bytes in front of the
// period are addressable; 2: periodic at the very start of the output -> byte gather.
// mode 3: period < 16: the period has been unrolled to 32 bytes in shared memory (ext), so a chunk
// is the 16 bytes at its phase.
__device__ __forceinline__ uint4 chunk_source(const uint8_t *pat, uint32_t rel, uint32_t offset, float inv, int mode,
                                              const uint8_t *ext)
{
    if (mode == 0) return load16_unaligned(pat + rel);
    uint32_t j = mod_period(rel, offset, inv);
    if (mode == 3) return load16_unaligned(ext + j);
    if (mode == 1) {
        // a chunk is one contiguous run of the period (A), or its end followed by its start (A then B)
        const int t = (int)(offset - j);              // bytes left in this period, >= 1
        const uint4 a = load16_unaligned(pat + j);
        if (t >= 16) return a;
        const uint4 b = load16_unaligned(pat - t);
        return make_uint4(merge_word(a.x, b.x, t), merge_word(a.y, b.y, t - 4), merge_word(a.z, b.z, t - 8),
                          merge_word(a.w, b.w, t - 12));
    }
    return chunk_gather(pat, j, offset);
}

// One pass; every source byte lies in front of the match, so ALL loads (ragged
// edge bytes and 16-byte chunks) are issued before the first store and the
// match costs a single memory round trip.
template <bool ADLER, int G>
__device__ __forceinline__ void copy_match_part(const Grp<G> &g, OutState &o, unsigned length, unsigned offset,
                                                uint8_t *ext)
{
    constexpr int ER = (30 + G - 1) / G;     // rounds for up to 15 head + 15 tail bytes
    constexpr int BR = (16 + G - 1) / G;     // rounds for up to 16 chunks
    g.sync();   // earlier stores by other lanes of the group may be our source
    uint8_t *base = o.out;
    const uint32_t dpos = o.pos;
    BDF_ASSERT(offset >= 1 && offset <= dpos && dpos + length <= o.cap);
    const uint8_t *pat = base + dpos - offset;       // one period of the match source
    if (length <= (unsigned)G) {
        // short match (the common case in text): one byte per lane
        if (g.lane < length) {
            uint32_t j = g.lane;
            if (offset < length) j = g.lane % offset;
            const uint32_t b = pat[j];
            base[dpos + g.lane] = (uint8_t)b;
            adler_acc1<ADLER>(o, dpos + g.lane, b);
        }
        o.pos += length;
        return;
    }
    uint32_t head = (uint32_t)((16u - (uint32_t)(reinterpret_cast<uintptr_t>(base + dpos) & 15u)) & 15u);
    if (head > length) head = length;
    const uint32_t nbody = (length - head) >> 4;      // <= 16 chunks of 16 bytes
    const uint32_t nedge = head + ((length - head) & 15u);
    const bool wrap = offset < length;
    const float inv = wrap ? __fdividef(1.f, (float)offset) : 0.f;
    const int mode = !wrap ? 0 : offset < 16 ? 3 : dpos - offset >= 20 ? 1 : 2;
    if (mode == 3) {
        // unroll the short period to 32 bytes in shared memory (one byte per lane per round)
        for (uint32_t k = g.lane; k < 32; k += G) ext[k] = pat[k % offset];
        g.sync();
    }
    uint32_t eb[ER];
    uint4 cv[BR];
#pragma unroll
    for (int r = 0; r < ER; r++) {
        const uint32_t e = g.lane + r * G;
        eb[r] = 0;
        if (e < nedge) {
            const uint32_t ei = e < head ? e : nbody * 16 + e;    // tail bytes follow the body
            eb[r] = pat[wrap ? mod_period(ei, offset, inv) : ei];
        }
    }
#pragma unroll
    for (int r = 0; r < BR; r++) {
        const uint32_t c = g.lane + r * G;
        if (c < nbody) cv[r] = chunk_source(pat, head + 16 * c, offset, inv, mode, ext);
    }
#pragma unroll
    for (int r = 0; r < ER; r++) {
        const uint32_t e = g.lane + r * G;
        if (e < nedge) {
            const uint32_t ei = e < head ? e : nbody * 16 + e;
            base[dpos + ei] = (uint8_t)eb[r];
            adler_acc1<ADLER>(o, dpos + ei, eb[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < BR; r++) {
        const uint32_t c = g.lane + r * G;
        if (c < nbody) {
            const uint32_t rel = head + 16 * c;
            *reinterpret_cast<uint4 *>(base + dpos + rel) = cv[r];
            adler_acc16<ADLER>(o, dpos + rel, cv[r]);
        }
    }
    // coalesced runs of matches can be much longer than 258 bytes: remaining chunks, one per lane
    // per step (the source never overlaps what this call writes: it is the period in front of
    // dpos, or for offset >= length the bytes in front of it)
#pragma unroll 1
    for (uint32_t c = g.lane + BR * G; c < nbody; c += G) {
        const uint32_t rel = head + 16 * c;
        const uint4 v = chunk_source(pat, rel, offset, inv, mode, ext);
        *reinterpret_cast<uint4 *>(base + dpos + rel) = v;
        adler_acc16<ADLER>(o, dpos + rel, v);
    }
    o.pos += length;
}


// Whole match.  Three things happen here:
//  * Parts.  The generic copy works on parts of at most COPY_PART bytes, each preceded by the
//    zero-fill of its own range only: filling all of a 64 KiB coalesced match ahead of the copy
//    would push the zeroes out to DRAM before they are overwritten (measured: DRAM writes 2x).
//    A long periodic match that starts right behind its first period (a stream that is one
//    repeated record) cannot look back in front of the period; its first period(s) are their own part.
//  * Long periodic matches (corpus A: one 65 KiB match of period 100 per stream; a run of zeroes
//    is period 1).  After a lead-in that ends on a 16-byte boundary and covers D = lcm(period, 16)
//    bytes, the rest of the match is the D bytes in front of it over and over, chunk c being
//    chunk (c mod D/16) of that block.  The block never changes while the match is written, so
//    the loop has no dependences at all: one L1-resident LDG.128 and one STG.128 per chunk, no
//    realignment, no barrier, whole sectors written (no zero-fill needed).
//  * The Adler-32 contribution of that part follows from the sums over ONE block (and over the
//    partial last one): out[p2 + m*D + k] = blk[k], so sum i*b is a polynomial in the block sums.
constexpr uint32_t COPY_PART = 4096;
constexpr uint32_t LONG_D_MAX = 4096;
constexpr uint32_t LONG_MIN_BODY = 512;
#ifndef BDF_TMA_REPLAY
#define BDF_TMA_REPLAY 1
#endif

// bytes [0, n) of w (n may be <= 0 or >= 4)
__device__ __forceinline__ uint32_t keep_low_bytes(uint32_t w, int n)
{
    return n <= 0 ? 0u : n >= 4 ? w : w & ((1u << (8 * n)) - 1u);
}

// d16[c] = src[c mod nblk] for c < nch, chunk c handled by lane (c mod G).  The loads of four
// steps are issued before their stores (the compiler cannot know that the stores never hit src).
template <int G>
__device__ __forceinline__ void replay_block(const Grp<G> &g, const uint4 *src, uint4 *d16, uint32_t nch, uint32_t nblk)
{
    uint32_t j = g.lane % nblk, c = g.lane;
    const uint32_t gs = (uint32_t)G % nblk;
#pragma unroll 1
    for (; c + 3 * G < nch; c += 4 * G) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            v[u] = src[j];
            j += gs;
            if (j >= nblk) j -= nblk;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) d16[c + u * G] = v[u];
    }
#pragma unroll 1
    for (; c < nch; c += G) {
        d16[c] = src[j];
        j += gs;
        if (j >= nblk) j -= nblk;
    }
}

// Block sums for the closed-form Adler-32 of a replayed block: SD = sum blk[k], WD = sum k*blk[k]
// over the whole block, SR / WR the same over k < r; reduced over the group, all mod 65521.
template <int G>
__device__ __forceinline__ void block_sums(const Grp<G> &g, const uint4 *blk, uint32_t nblk, uint32_t r,
                                           uint32_t &SD, uint32_t &WD, uint32_t &SR, uint32_t &WR)
{
    SD = 0; WD = 0; SR = 0; WR = 0;
    for (uint32_t k = g.lane; k < nblk; k += G) {
        const uint4 x = blk[k];
        const uint32_t sx = __dp4a(x.x, 0x01010101u, __dp4a(x.y, 0x01010101u, __dp4a(x.z, 0x01010101u, __dp4a(x.w, 0x01010101u, 0u))));
        const uint32_t wx = 16u * k * sx + __dp4a(x.x, 0x03020100u, __dp4a(x.y, 0x07060504u, __dp4a(x.z, 0x0B0A0908u, __dp4a(x.w, 0x0F0E0D0Cu, 0u))));
        SD += sx;
        WD += wx;
        if (k < (r >> 4)) { SR += sx; WR += wx; }
    }
    if ((r & 15u) && g.lane == 0) {           // the chunk the remainder ends in
        const uint32_t k = r >> 4;
        const int nb = (int)(r & 15u);
        uint4 x = blk[k];
        x.x = keep_low_bytes(x.x, nb); x.y = keep_low_bytes(x.y, nb - 4);
        x.z = keep_low_bytes(x.z, nb - 8); x.w = keep_low_bytes(x.w, nb - 12);
        const uint32_t sx = __dp4a(x.x, 0x01010101u, __dp4a(x.y, 0x01010101u, __dp4a(x.z, 0x01010101u, __dp4a(x.w, 0x01010101u, 0u))));
        SR += sx;
        WR += 16u * k * sx + __dp4a(x.x, 0x03020100u, __dp4a(x.y, 0x07060504u, __dp4a(x.z, 0x0B0A0908u, __dp4a(x.w, 0x0F0E0D0Cu, 0u))));
    }
    SD %= 65521u; WD %= 65521u; SR %= 65521u; WR %= 65521u;
#pragma unroll
    for (int sft = G / 2; sft > 0; sft >>= 1) {
        SD += g.shfl_xor(SD, sft); WD += g.shfl_xor(WD, sft);
        SR += g.shfl_xor(SR, sft); WR += g.shfl_xor(WR, sft);
    }
}

template <bool ADLER, int G>
__device__ __forceinline__ void copy_match(const Grp<G> &g, OutState &o, unsigned length, unsigned offset,
                                           uint8_t *ext)
{
    uint32_t lead = length, dist = 0;
    if (length >= 2 * LONG_MIN_BODY && offset < length) {
        const uint32_t tz = __ffs(offset) - 1;
        const uint32_t d = offset << (4u - (tz < 4u ? tz : 4u));          // lcm(offset, 16)
        const uint32_t l1 = d + ((16u - (uint32_t)(reinterpret_cast<uintptr_t>(o.out + o.pos + d) & 15u)) & 15u);
        if (d <= LONG_D_MAX && length >= l1 + LONG_MIN_BODY) { lead = l1; dist = d; }
    }
    uint32_t rem = lead;
#pragma unroll 1
    while (rem) {
        uint32_t part = rem < COPY_PART ? rem : COPY_PART;
        if (offset >= 16 && o.pos - offset < 20 && rem > 2 * offset + 64) part = offset >= 20 ? offset : 2 * offset;
        make_valid<G>(g, o, o.pos + part + G);        // + G: also covers the next literal flush
        copy_match_part<ADLER, G>(g, o, part, offset, ext);
        rem -= part;
    }
    if (dist) {
        g.sync();                                     // the block was written by all lanes of the group
        const uint32_t p2 = o.pos, n2 = length - lead;
        const uint32_t nch = n2 >> 4, nblk = dist >> 4;
        uint4 *d16 = reinterpret_cast<uint4 *>(o.out + p2);
        const uint4 *blk = reinterpret_cast<const uint4 *>(o.out + p2 - dist);
        const uint32_t q = n2 / dist, r = n2 - q * dist;
        uint32_t SD = 0, WD = 0, SR = 0, WR = 0;
        if (dist <= o.blkcap) {
            // Small blocks are staged in shared memory (the table builder's scratch is idle while
            // symbols are decoded), repeated until the buffer is full, and then written by the TMA
            // unit: one lane issues cp.async.bulk shared -> global copies of the whole buffer
            // (UBLKCP), so the 64 KiB of a config-2 stream cost ~80 issue slots instead of 4096
            // LDS/STG pairs and no load-to-store latency sits on the warp.  The block sums for
            // Adler-32 are taken from the same buffer while the copies are in flight.
            uint4 *sb = reinterpret_cast<uint4 *>(ext);
            const uint32_t rch = (o.blkcap / dist) * nblk;                // chunks in the repeated buffer
            {
                uint32_t j = g.lane % nblk;
                const uint32_t gs = (uint32_t)G % nblk;
                for (uint32_t k = g.lane; k < rch; k += G) {
                    sb[k] = blk[j];
                    j += gs;
                    if (j >= nblk) j -= nblk;
                }
            }
#if BDF_TMA_REPLAY
            // the generic-proxy writes above (and the zero-fill / lead-in stores to the output) come
            // before the async-proxy accesses below
            asm volatile("fence.proxy.async;" ::: "memory");
            g.sync();
            if (g.lane == 0) {
                const uint32_t rbytes = rch * 16u, total = nch * 16u;
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(sb);
                uint8_t *dst = o.out + p2;
                uint32_t off = 0;
                for (; off + rbytes <= total; off += rbytes)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 ::"l"(dst + off), "r"(src), "r"(rbytes) : "memory");
                if (off < total)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                 ::"l"(dst + off), "r"(src), "r"(total - off) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (ADLER) block_sums<G>(g, sb, nblk, r, SD, WD, SR, WR);
            // complete (not only read) before anyone continues: a later match may copy from these bytes
            if (g.lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            g.sync();
#else
            g.sync();
            replay_block<G>(g, sb, d16, nch, rch);
            if (ADLER) block_sums<G>(g, sb, nblk, r, SD, WD, SR, WR);
            g.sync();                                 // ext is reused by the next match
#endif
        } else {
            replay_block<G>(g, blk, d16, nch, nblk);
            if (ADLER) block_sums<G>(g, blk, nblk, r, SD, WD, SR, WR);
        }
        const uint32_t tail = n2 & 15u;
        if (g.lane < tail)
            o.out[p2 + 16 * nch + g.lane] = reinterpret_cast<const uint8_t *>(blk)[16 * (nch % nblk) + g.lane];
        if (ADLER && g.lane == 0) {
            // sum over m < q, k < D of (p2 + m*D + k) blk[k]  +  sum over k < r of (p2 + q*D + k) blk[k]
            const uint64_t M = 65521u;
            const uint64_t sd = SD % M, wd = WD % M, sr = SR % M, wr = WR % M;
            const uint64_t qm = q % M, pm = p2 % M, dm = dist % M;
            const uint64_t tri = ((uint64_t)q * (q - (q ? 1u : 0u)) / 2u) % M;
            const uint64_t db = qm * pm % M * sd + dm * sd % M * tri + qm * wd + (pm + qm * dm) % M * sr + wr;
            o.sumA += (uint32_t)((qm * sd + sr) % M);
            o.sumB += db % M;
        }
        o.pos += n2;
    }
}

// ------------------------------------------------------------ block decoding
// Consecutive matches with the same offset are one longer match with that offset (each one
// continues the same period).  After a long match the decoder therefore looks ahead: while the
// next symbol is another match with the same offset it is folded into the current one, so that
// periodic / run-length data costs one long, fully populated copy instead of hundreds of
// 258-byte ones.  The look-ahead works on a copy of the bit reader and is simply dropped when
// the next symbol turns out to be something else; short matches (text) never pay for it.
#ifndef BDF_COALESCE_MAX
#define BDF_COALESCE_MAX (1u << 16)
#endif
constexpr uint32_t COALESCE_MAX = BDF_COALESCE_MAX;
constexpr uint32_t COALESCE_MIN_LEN = 64;

// One decoding step of a Huffman block: up to two literals, or one match (with everything that
// coalesces into it), or the end of the block.  Returns STEP_MORE, or the status that ends the
// block (BDF_OK at end-of-block).  The callers' loops are WARP-uniform (see inflate_stream): every
// lane group of the warp enters each step together, which is what keeps groups with similar
// streams on one instruction stream.
constexpr int STEP_MORE = -1;
template <bool ADLER, int G>
__device__ __forceinline__ int decode_step(const Grp<G> &g, BitReader &br, OutState &o, InflateSmem<G> &sm)
{
    // A literal: parked in lane npend of the group, flushed as one store when every lane has one.
#define BDF_PARK_LITERAL(e)                                                          \
    do {                                                                             \
        if (o.pos + o.npend >= o.cap) return BDF_INSUFFICIENT_SPACE;                 \
        br.drop((e) & E_LEN);                                                        \
        if (g.lane == o.npend) o.mylit = (e) >> E_VAL;                               \
        if (++o.npend == G) {                                                        \
            make_valid<G>(g, o, o.pos + 2 * G);                                      \
            flush_literals<ADLER>(o, g.lane);                                        \
        }                                                                            \
    } while (0)
    {
        if (br.left <= 32) {
            // more than two zero-fill words loaded: the stream ended inside this block
            if (br.widx > br.nwords + 2) return BDF_SHORT_INPUT;
            br.refill();
        }
        uint32_t e = sm.lit_tab[br.peek(LT_BITS)];
        if (e & LITFLAG) {
            // literals come in runs: a second one is looked up without another refill
            // (a direct literal is at most LT_BITS bits, so >= 24 valid bits remain)
            BDF_PARK_LITERAL(e);
            e = sm.lit_tab[br.peek(LT_BITS)];
            if (e & LITFLAG) {
                BDF_PARK_LITERAL(e);
                return STEP_MORE;
            }
            br.refill();                               // the general path starts from >= 33 valid bits
        }
        const uint32_t tok_lo = (uint32_t)br.buf;      // >= 33 valid bits: the next 32 bits of the stream
        const int32_t tok_left = br.left;
        const uint32_t tok_widx = br.widx;
        if ((e & E_LEN) == 0) {
            e = decode_long<CODE_LITLEN, LT_BITS>(br.peek(15), sm.lit_sorted, sm.lit_code);
            if (e == 0) return BDF_BAD_DATA;
        }
        const uint32_t kind = e & K_MASK;
        if (kind == K_LIT) {                           // a literal with a codeword longer than the table
            BDF_PARK_LITERAL(e);
            return STEP_MORE;
        }
        br.drop(e & E_LEN);
        if (kind == K_EOB) {
            return br.overrun() ? BDF_SHORT_INPUT : BDF_OK;
        }
        unsigned length = take_length(br, e);
        br.refill();                                   // appends above `left`: consumed-bit accounting is unchanged
        uint32_t f = sm.off_tab[br.peek(OT_BITS)];
        if ((f & E_LEN) == 0) {
            f = decode_long<CODE_OFFSET, OT_BITS>(br.peek(15), sm.off_sorted, sm.off_code);
            if (f == 0) return BDF_BAD_DATA;
        }
        br.drop(f & E_LEN);
        const unsigned offset = take_offset(br, f);
        flush_literals<ADLER>(o, g.lane);
        if (offset > o.pos) return BDF_BAD_DATA;
        if (o.pos + length > o.cap) return BDF_INSUFFICIENT_SPACE;
        if (length >= COALESCE_MIN_LEN && COALESCE_MAX) {
            // (a) The same match again: Huffman decoding is a function of the bits, so if the next
            // bits equal the bits of the token just decoded they decode to the same (length,
            // offset).  Run-length / periodic data is hundreds of identical tokens in a row; each
            // costs one compare here instead of two table look-ups.
            const uint32_t tok_bits = (uint32_t)(tok_left - br.left) + 32u * (br.widx - tok_widx);
            if (tok_bits <= 32u) {
                const uint32_t tmask = 0xFFFFFFFFu >> (32u - tok_bits);
                const uint32_t tpat = tok_lo & tmask;
                const unsigned len1 = length;
                // Long runs: the G lanes check one input word each against the periodic bit
                // pattern (the token repeated, shifted to the word's phase), a ballot gives the
                // number of words that continue the run, and the reader jumps over all whole
                // tokens at once — G words per round instead of one refill per 32 bits, each of
                // which used to wait for its prefetched word.
                {
                    uint64_t rep64 = tpat;
                    for (uint32_t b = tok_bits; b < 64u; b += tok_bits) rep64 |= (uint64_t)tpat << b;
                    const uint32_t *wbase = reinterpret_cast<const uint32_t *>(br.p - br.mis);
#pragma unroll 1
                    for (;;) {
                        const uint32_t have = (uint32_t)br.left;           // bits in buf: they continue the pattern too
                        const uint64_t hmask = have >= 64u ? ~0ull : ((1ull << have) - 1ull);
                        if ((br.buf ^ rep64) & hmask) break;
                        const uint32_t w = br.widx + g.lane;
                        const bool interior = w - 1u < br.nwords - 2u;     // words that lie wholly inside the stream
                        const uint32_t word = interior ? __ldg(wbase + w) : 0u;
                        const uint32_t q = have + 32u * g.lane;            // bit offset of the word from the token boundary
                        const uint32_t phi = q % tok_bits;
                        const bool ok = interior && word == (uint32_t)(rep64 >> phi);
                        const unsigned bad = g.ballot(!ok);
                        const uint32_t m = bad ? (uint32_t)__ffs(bad) - 1u : (uint32_t)G;
                        uint32_t tokens = (have + 32u * m) / tok_bits;
                        const uint32_t lim = o.cap - o.pos < COALESCE_MAX ? o.cap - o.pos : COALESCE_MAX;
                        const uint32_t max_tok = (lim - length) / len1;    // o.pos + length <= cap was checked
                        const bool capped = tokens >= max_tok;
                        if (capped) tokens = max_tok;
                        if (tokens == 0) break;
                        length += tokens * len1;
                        br.seek_bits(br.consumed_bits() + (int64_t)tokens * tok_bits);
                        if (m < (uint32_t)G || capped) break;
                    }
                }
                // several tokens per compare while they fit into the 32 bits a refill guarantees
                const uint32_t nrep = 32u / tok_bits;
                if (nrep >= 2) {
                    uint32_t rpat = tpat;
                    for (uint32_t k = 1; k < nrep; k++) rpat |= tpat << (k * tok_bits);
                    const uint32_t rbits = nrep * tok_bits, rmask = 0xFFFFFFFFu >> (32u - rbits);
                    const uint32_t rlen = nrep * len1;
#pragma unroll 1
                    for (;;) {
                        if (br.widx > br.nwords + 2) break;
                        br.refill();
                        if ((((uint32_t)br.buf ^ rpat) & rmask) != 0) break;
                        if (length + rlen > COALESCE_MAX || o.pos + length + rlen > o.cap) break;
                        br.drop(rbits);
                        length += rlen;
                    }
                }
#pragma unroll 1
                for (;;) {
                    if (br.widx > br.nwords + 2) break;
                    br.refill();
                    if ((((uint32_t)br.buf ^ tpat) & tmask) != 0) break;
                    if (length + len1 > COALESCE_MAX || o.pos + length + len1 > o.cap) break;
                    br.drop(tok_bits);
                    length += len1;
                }
            }
            // (b) another match with the same offset but a different length (the last one of a run)
#pragma unroll 1
            for (;;) {
                BitReader t = br;
                if (t.widx > t.nwords + 2) break;
                t.refill();
                uint32_t e2 = sm.lit_tab[t.peek(LT_BITS)];
                if ((e2 & E_LEN) == 0) e2 = decode_long<CODE_LITLEN, LT_BITS>(t.peek(15), sm.lit_sorted, sm.lit_code);
                if ((e2 & E_LEN) == 0 || (e2 & K_MASK) != K_BASE) break;
                t.drop(e2 & E_LEN);
                const unsigned len2 = take_length(t, e2);
                t.refill();
                uint32_t f2 = sm.off_tab[t.peek(OT_BITS)];
                if ((f2 & E_LEN) == 0) f2 = decode_long<CODE_OFFSET, OT_BITS>(t.peek(15), sm.off_sorted, sm.off_code);
                if ((f2 & E_LEN) == 0) break;
                t.drop(f2 & E_LEN);
                const unsigned off2 = take_offset(t, f2);
                if (off2 != offset || length + len2 > COALESCE_MAX || o.pos + length + len2 > o.cap) break;
                br = t;
                length += len2;
            }
        }
        copy_match<ADLER, G>(g, o, length, offset,
                             reinterpret_cast<uint8_t *>(o.blkcap == BLOCK_BUF_ALL<G> ? (void *)sm.lit_sorted : (void *)&sm.bs));
        if (ADLER) adler_fold(o);
    }
    return STEP_MORE;
#undef BDF_PARK_LITERAL
}

// read_dynamic_huffman_header, src/decompress/mod.rs:403-507
// SM: anything with the members of InflateSmem (lit_tab, off_tab, lit_sorted, off_sorted, lit_code,
// off_code, bs, lens) — the lane kernel passes a view whose tables sit in its own slot layout.
// where the precode table lives while a header is read (128 entries): on top of the offset table
// unless the holder says otherwise
template <class SM> __device__ __forceinline__ uint16_t *precode_table(SM &sm) { return sm.off_tab; }

template <int G, class SM, int LTB = LT_BITS, int OTB = OT_BITS>
__device__ int read_code_lengths(const Grp<G> &g, BitReader &br, SM &sm, unsigned &nlit, unsigned &noff)
{
    br.refill();
    nlit = 257 + br.take(5);
    noff = 1 + br.take(5);
    const unsigned npre = 4 + br.take(4);
    uint8_t *pre_lens = sm.lens + 300;   // only needed until the precode table is built
    {
        // up to 19 * 3 = 57 bits of precode lengths, read in two steps; lane k mod G stores entry k
        br.refill();
        const unsigned n_lo = npre < 10 ? npre : 10;
        const uint32_t lo = br.take(3 * n_lo);
        br.refill();
        const uint32_t hi = npre > 10 ? br.take(3 * (npre - 10)) : 0;
        // order 16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15 packed 5 bits each
        const uint64_t perm_lo = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 |
                                 9ull << 30 | 6ull << 35 | 10ull << 40 | 5ull << 45;
        const uint64_t perm_hi = 11ull | 4ull << 5 | 12ull << 10 | 3ull << 15 | 13ull << 20 | 2ull << 25 |
                                 14ull << 30 | 1ull << 35 | 15ull << 40;
        for (unsigned k = g.lane; k < 19; k += G) {
            const unsigned sym = k < 10 ? (unsigned)(perm_lo >> (5 * k)) & 31u
                                        : (unsigned)(perm_hi >> (5 * (k - 10))) & 31u;
            const unsigned v = k < 10 ? (lo >> (3 * k)) & 7u : (hi >> (3 * (k - 10))) & 7u;
            pre_lens[sym] = (uint8_t)(k < npre ? v : 0);
        }
        g.sync();
    }
    if (br.overrun()) return BDF_SHORT_INPUT;
    uint16_t *pre_tab = precode_table(sm);
    if (!build_code<CODE_PRECODE, PT_BITS, G>(g, pre_lens, 19, pre_tab, sm.off_sorted, sm.off_code, sm.bs))
        return BDF_BAD_DATA;
    // run-length decode of the litlen + offset code lengths (group-uniform)
    const unsigned total = nlit + noff;
    unsigned i = 0, prev = 0;
    while (i < total) {
        // >= 33 valid bits after a refill: up to three plain lengths (<= 7 bits each), or two and
        // one repeat symbol (<= 7 + 7 bits), are decoded per refill
        br.refill();
        uint32_t e = pre_tab[br.peek(PT_BITS)];
        unsigned k = 0;
        bool again = false;
        while ((e >> E_VAL) < 16) {
            if (g.lane == 0) sm.lens[i] = (uint8_t)(e >> E_VAL);
            prev = e >> E_VAL;
            br.drop(e & E_LEN);
            i++;
            if (++k == 3 || i >= total) { again = true; break; }
            e = pre_tab[br.peek(PT_BITS)];
        }
        if (again) continue;
        br.drop(e & E_LEN);
        const unsigned sym = e >> E_VAL;
        unsigned rep, val;
        if (sym == 16) {
            if (i == 0) return BDF_BAD_DATA;
            rep = 3 + br.take(2);
            val = prev;
        } else if (sym == 17) {
            rep = 3 + br.take(3);
            val = 0;
        } else {
            rep = 11 + br.take(7);
            val = 0;
        }
        if (rep > total - i) rep = total - i;      // overruns are clamped (:462-493)
        for (unsigned q = g.lane; q < rep; q += G) sm.lens[i + q] = (uint8_t)val;
        prev = val;
        i += rep;
    }
    if (br.overrun()) return BDF_SHORT_INPUT;
    g.sync();
    return BDF_OK;
}

// The same lengths from the row the pre-pass left (inflate_prehdr.cuh); the reader restarts behind
// the header.
template <int G, class SM>
__device__ __forceinline__ void load_code_lengths(const Grp<G> &g, BitReader &br, SM &sm, uint32_t meta, const uint32_t *row,
                                                  unsigned &nlit, unsigned &noff)
{
    nlit = 257 + ((meta >> 16) & 31u);
    noff = 1 + ((meta >> 21) & 31u);
    uint32_t *lw = reinterpret_cast<uint32_t *>(sm.lens);
    const unsigned nw = (nlit + noff + 3) >> 2;
    BDF_ASSERT(nw <= (unsigned)PREHDR_ROW_WORDS && (meta & PREHDR_VALID));
    for (unsigned j = g.lane; j < nw; j += G) lw[j] = __ldg(row + j);
    br.seek_bits((int64_t)(meta & 0xFFFFu));
    g.sync();
}

// decode tables of a dynamic block from the code lengths in sm.lens
template <int G, class SM, int LTB = LT_BITS, int OTB = OT_BITS>
__device__ int build_dynamic_codes(const Grp<G> &g, SM &sm, unsigned nlit, unsigned noff, uint32_t &nlong)
{
    nlong = 1;
    uint32_t long_off = 0, long_lit = 0;
    if (!build_code<CODE_OFFSET, OTB, G>(g, sm.lens + nlit, noff, sm.off_tab, sm.off_sorted, sm.off_code, sm.bs, &long_off))
        return BDF_BAD_DATA;
    if (!build_code<CODE_LITLEN, LTB, G>(g, sm.lens, nlit, sm.lit_tab, sm.lit_sorted, sm.lit_code, sm.bs, &long_lit))
        return BDF_BAD_DATA;
    nlong = long_off + long_lit;
    return BDF_OK;
}

template <int G, class SM, int LTB = LT_BITS, int OTB = OT_BITS>
__device__ void load_static_codes(const Grp<G> &g, SM &sm)
{
    for (unsigned s = g.lane; s < 320; s += G)
        sm.lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : s < 288 ? 8 : 5;
    g.sync();
    build_code<CODE_OFFSET, OTB, G>(g, sm.lens + 288, 32, sm.off_tab, sm.off_sorted, sm.off_code, sm.bs);
    build_code<CODE_LITLEN, LTB, G>(g, sm.lens, 288, sm.lit_tab, sm.lit_sorted, sm.lit_code, sm.bs);
}

// Raw DEFLATE stream [p, p+len) -> o; returns status, *used = bytes consumed.
// Called by ALL lanes of the warp together; `have` says whether this group has a stream.  Both
// loops run until no group of the warp has work left (a group that is done idles), so the groups
// meet again at every block header and at every decoding step.
// hdr_meta / hdr_rows + idx: the stream's first header as the pre-pass left it (inflate_prehdr.cuh).  A reader
// that was attached instead of started (widx == 0) marks a stream whose first header comes from there, so
// the state costs no register once the first block is under way.
template <bool ADLER, int G>
__device__ int inflate_stream(const Grp<G> &g, const uint8_t *p, uint32_t len, OutState &o, InflateSmem<G> &sm,
                              uint32_t *used, bool have, const uint32_t *hdr_meta, const uint32_t *hdr_rows,
                              unsigned long long idx)
{
    BitReader br;
    br.p = p; br.len = 0; br.mis = 0; br.nwords = 0; br.widx = 1; br.ahead = 0; br.buf = 0; br.left = 0;
    if (have) {
        if (hdr_meta && (__ldg(hdr_meta + idx) & PREHDR_VALID)) br.attach(p, len);
        else br.init(p, len);
    }
    int st = BDF_OK;
    bool live = have;
    while (__any_sync(BDF_FULL_MASK, live)) {
        bool inblock = false;
        unsigned final = 0;
        const bool pre = live && br.widx == 0;
        if (live && !pre) {
            br.refill();
            if (br.consumed_bits() + 3 > (int64_t)len * 8) { st = BDF_SHORT_INPUT; live = false; }
        }
        if (live) {
            unsigned type = 2;
            if (!pre) {
                final = br.take(1);
                type = br.take(2);
            }
            if (type == 0) {
                // stored block (src/decompress/mod.rs:282-346, x86.rs:2216-2246)
                uint32_t at = (uint32_t)((br.consumed_bits() + 7) >> 3);
                unsigned blen = 0;
                if (at + 4 > len) st = BDF_SHORT_INPUT;
                else {
                    blen = p[at] | (unsigned)p[at + 1] << 8;
                    const unsigned nlen = p[at + 2] | (unsigned)p[at + 3] << 8;
                    at += 4;
                    if (blen != (~nlen & 0xFFFFu)) st = BDF_BAD_DATA;
                    else if (o.pos + blen > o.cap) st = BDF_INSUFFICIENT_SPACE;
                    else if (at + blen > len) st = BDF_SHORT_INPUT;
                }
                if (st == BDF_OK) {
                    for (unsigned i = g.lane; i < blen; i += G) {
                        const uint8_t b = p[at + i];
                        o.out[o.pos + i] = b;
                        adler_acc1<ADLER>(o, o.pos + i, b);
                    }
                    o.pos += blen;
                    if (ADLER) adler_fold(o);
                    br.seek(at + blen);
                    if (final) live = false;
                } else {
                    live = false;
                }
            } else if (type == 3) {
                st = BDF_BAD_DATA;
                live = false;
            } else {
                uint32_t nlong = 0;                    // the static codes fit the direct tables
                if (type == 1) {
                    load_static_codes<G>(g, sm);
                } else {
                    unsigned nlit = 0, noff = 0;
                    if (pre) {
                        const uint32_t meta = __ldg(hdr_meta + idx);
                        final = (meta >> 26) & 1u;
                        load_code_lengths<G>(g, br, sm, meta, hdr_rows + idx * PREHDR_ROW_WORDS, nlit, noff);
                    } else st = read_code_lengths<G>(g, br, sm, nlit, noff);
                    if (st == BDF_OK) st = build_dynamic_codes<G>(g, sm, nlit, noff, nlong);
                    if (st != BDF_OK) live = false;
                }
                // A block without codewords longer than the direct tables never looks at the sorted
                // symbol lists / first-code arrays again: copy_match may use them, together with the
                // builder scratch, as its block buffer (bigger TMA copies).
                o.blkcap = nlong == 0 ? BLOCK_BUF_ALL<G> : BLOCK_BUF_SCRATCH<G>;
                g.sync();
                inblock = live;
            }
        }
        while (__any_sync(BDF_FULL_MASK, inblock)) {
            if (inblock) {
                const int r = decode_step<ADLER, G>(g, br, o, sm);
                if (r != STEP_MORE) {
                    inblock = false;
                    flush_literals<ADLER>(o, g.lane);
                    if (r != BDF_OK) { st = r; live = false; }
                    else if (final) live = false;
                }
            }
        }
    }
    int64_t cb = have ? br.consumed_bits() : 0;
    if (cb < 0) cb = 0;
    *used = (uint32_t)((cb + 7) >> 3);
    g.sync();
    return st;
}

// Group-wide CRC-32 of d[0..n): G contiguous slices, slice-by-4 per lane,
// partial CRCs shifted to the end with x^(8k) mod P and XOR-reduced.
template <int G>
__device__ uint32_t grp_crc32(const Grp<G> &g, const uint8_t *d, uint64_t n, const uint32_t (*slice)[256],
                              const uint32_t *x2n)
{
    uint64_t chunk = (n + G - 1) / G;
    uint64_t beg = chunk * g.lane, end = beg + chunk;
    if (beg > n) beg = n;
    if (end > n) end = n;
    uint32_t c = 0xFFFFFFFFu;
    uint64_t i = beg;
    while (i < end && ((uintptr_t)(d + i) & 3)) { c = (c >> 8) ^ slice[0][(c ^ d[i]) & 0xFF]; i++; }
    for (; i + 4 <= end; i += 4) {
        uint32_t w = c ^ *reinterpret_cast<const uint32_t *>(d + i);
        c = slice[3][w & 0xFF] ^ slice[2][(w >> 8) & 0xFF] ^ slice[1][(w >> 16) & 0xFF] ^ slice[0][w >> 24];
    }
    for (; i < end; i++) c = (c >> 8) ^ slice[0][(c ^ d[i]) & 0xFF];
    c = ~c;
    if (end == beg) c = 0;
    uint32_t part = (n - end) ? gf2_mulmod(gf2_xpow8n(n - end, x2n), c) : c;
#pragma unroll
    for (int s = G / 2; s > 0; s >>= 1) part ^= g.shfl_xor(part, s);
    return part;
}
__device__ __forceinline__ uint32_t warp_crc32(const uint8_t *d, uint64_t n, const uint32_t (*slice)[256],
                                               const uint32_t *x2n, unsigned)
{
    return grp_crc32<32>(Grp<32>(), d, n, slice, x2n);
}

// (Σb, Σ i·b) partials -> Adler-32 of n bytes with seed 1.
template <int G>
__device__ uint32_t grp_adler_finish(const Grp<G> &g, uint32_t sumA, uint64_t sumB, uint64_t n)
{
    uint32_t a = sumA % 65521u, b = (uint32_t)(sumB % 65521u);
#pragma unroll
    for (int s = G / 2; s > 0; s >>= 1) {
        a += g.shfl_xor(a, s);
        b += g.shfl_xor(b, s);
    }
    a %= 65521u;
    b %= 65521u;
    const uint64_t nm = n % 65521u;
    const uint64_t s1 = (1 + a) % 65521u;
    const uint64_t s2 = (nm + nm * a + 65521u - b) % 65521u;
    return (uint32_t)(s2 << 16 | s1);
}
__device__ __forceinline__ uint32_t warp_adler_finish(uint32_t sumA, uint64_t sumB, uint64_t n)
{
    return grp_adler_finish<32>(Grp<32>(), sumA, sumB, n);
}

struct InflateArgs {
    const uint8_t *in;
    const uint64_t *in_off;
    uint8_t *out;
    const uint64_t *out_off;
    const uint64_t *max_out;
    uint64_t *out_size;
    uint32_t *checksum;
    int32_t *status;
    unsigned long long *work_counter;    // queue head of inflate_kernel (lane groups)
    unsigned long long *work_counter2;   // queue head of inflate_lane_kernel
    uint8_t *lane_scratch;               // inflate_lane_kernel: LANE_SORTED_BYTES per lane of the grid
    uint32_t n;
    // first-block headers decoded by inflate_prehdr_kernel (null: the kernels read every header themselves)
    const uint32_t *hdr_rows;            // n x PREHDR_ROW_WORDS
    const uint32_t *hdr_meta;            // n
    // Two engines share a batch: a stream whose capacity is at least split_ratio times its
    // compressed length ("heavy": a few long matches, run-length / periodic data) goes to the
    // lane-group kernel, every other stream to the lane-per-stream kernel (inflate_lane.cuh).
    // 0 = no split: the kernel that is launched takes every stream.
    uint32_t split_ratio;
};

// capacities are clamped so that pos + length (<= 65536 + 258 after coalescing) cannot wrap
constexpr uint32_t INFLATE_CAP_MAX = 0xFFFE0000u;
__device__ __forceinline__ bool inflate_is_heavy(uint32_t ratio, uint64_t len, uint64_t cap)
{
    return ratio != 0 && cap >= (uint64_t)ratio * (len ? len : 1);
}

// Framing in front of the DEFLATE data (decompress_zlib_uninit / decompress_gzip_uninit,
// src/decompress/mod.rs:1074-1127,1144-1240): offset and length of the DEFLATE data, or the
// status that ends the stream.
template <int FORMAT>
__device__ __forceinline__ int inflate_frame_header(const uint8_t *p, uint32_t len, uint32_t &at, uint32_t &dlen)
{
    at = 0; dlen = 0;
    if (FORMAT == BDF_RAW) { dlen = len; return BDF_OK; }
    if (FORMAT == BDF_ZLIB) {
        if (len < 6) return BDF_SHORT_INPUT;
        const unsigned hdr = (unsigned)p[0] << 8 | p[1];
        at = 2;
        dlen = len - 6;
        if (hdr % 31 != 0 || ((hdr >> 8) & 0xF) != 8 || ((hdr >> 12) & 0xF) > 7 || ((hdr >> 5) & 1)) return BDF_BAD_DATA;
        return BDF_OK;
    }
    if (len < 18) return BDF_SHORT_INPUT;
    if (p[0] != 0x1F || p[1] != 0x8B || p[2] != 8 || (p[3] & 0xE0)) return BDF_BAD_DATA;
    const unsigned flg = p[3];
    uint64_t h = 10;
    if (flg & 0x04) {
        if (h + 2 > len) return BDF_SHORT_INPUT;
        h += 2 + (p[h] | (uint64_t)p[h + 1] << 8);
    }
    if (flg & 0x08) { while (h < len && p[h]) h++; h++; }
    if (flg & 0x10) { while (h < len && p[h]) h++; h++; }
    if (flg & 0x02) h += 2;
    if (h + 8 > len) return BDF_SHORT_INPUT;
    at = (uint32_t)h;
    dlen = (uint32_t)(len - 8 - h);
    return BDF_OK;
}

constexpr int INF_THREADS = 64;     // threads per CTA; 64 / G lane groups = streams in flight per CTA

template <int FORMAT, int G>
__global__ void __launch_bounds__(INF_THREADS, 14)
inflate_kernel(InflateArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_crc[FORMAT == BDF_GZIP ? 4 : 1][256];
    __shared__ uint32_t s_x2n[32];
    InflateSmem<G> &sm = reinterpret_cast<InflateSmem<G> *>(smem_raw)[threadIdx.x / G];
    static_assert(offsetof(InflateSmem<G>, bs) % 16 == 0 && offsetof(InflateSmem<G>, lit_sorted) % 16 == 0 &&
                      sizeof(InflateSmem<G>) % 16 == 0 && BLOCK_BUF_SCRATCH<G> >= 1024 && sizeof(sm.lens) == 328 &&
                      offsetof(InflateSmem<G>, lens) % 4 == 0,
                  "builder scratch + lens (and the long-code lists in front of them) double as the 16-byte "
                  "aligned block buffer of copy_match");
    const Grp<G> g;
    if (FORMAT == BDF_GZIP) {
        for (unsigned i = threadIdx.x; i < 1024; i += blockDim.x) s_crc[i >> 8][i & 255] = g_crc_tables.slice[i >> 8][i & 255];
        if (threadIdx.x < 32) s_x2n[threadIdx.x] = g_crc_tables.x2n[threadIdx.x];
        __syncthreads();
    }
    // The groups of a warp take their streams TOGETHER: streams of one batch tend to look alike, and
    // groups that start together stay in lock-step (one instruction stream for 32/G streams);
    // groups that fetched on their own would drift apart for good and execute one after the other.
    constexpr unsigned GPW = 32 / G;
    bool q_empty = false;
    for (;;) {
        // Every group of the warp takes the next stream of this kernel's class (see
        // InflateArgs::split_ratio); without a split that is GPW consecutive streams in one round.
        bool have = false;
        unsigned long long idx = 0;
        uint64_t len64 = 0, cap64 = 0;
        while (!q_empty) {
            const unsigned need = __ballot_sync(BDF_FULL_MASK, !have && g.lane == 0);
            if (!need) break;
            unsigned long long base = 0;
            if (lane_id() == 0) base = atomicAdd(a.work_counter, (unsigned long long)__popc(need));
            base = __shfl_sync(BDF_FULL_MASK, base, 0);
            if (base + __popc(need) >= a.n) q_empty = true;
            if (!have) {
                const unsigned long long my = base + __popc(need & ((1u << g.shift) - 1u));
                if (my < a.n) {
                    len64 = a.in_off[my + 1] - a.in_off[my];
                    cap64 = a.max_out[my];
                    if (a.split_ratio == 0 || inflate_is_heavy(a.split_ratio, len64, cap64)) { have = true; idx = my; }
                }
            }
        }
        if (!__any_sync(BDF_FULL_MASK, have)) break;
        const uint8_t *p = a.in;
        OutState o;
        o.out = a.out;
        if (have) {
            p = a.in + a.in_off[idx];
            o.out = a.out + a.out_off[idx];
        }
        o.pos = 0;
        o.cap = cap64 > INFLATE_CAP_MAX ? INFLATE_CAP_MAX : (uint32_t)cap64;
        o.npend = 0; o.mylit = 0; o.sumA = 0; o.sumB = 0; o.next_fold = ADLER_FOLD_INTERVAL;
        o.blkcap = BLOCK_BUF_SCRATCH<G>;
        {
            // whole sectors owned by this stream: [first 32-byte boundary at or after out, last one at or before out + cap)
            const uint32_t a0 = (uint32_t)(reinterpret_cast<uintptr_t>(o.out) & 31u);
            const uint32_t first = (32u - a0) & 31u;
            const int64_t last = (int64_t)(((uint64_t)a0 + o.cap) & ~31ull) - (int64_t)a0;
            o.zfill = first;
            o.zlimit = last > (int64_t)first ? (uint32_t)last : first;
        }
        // framing in front of the DEFLATE data: decides whether (and where) this group inflates
        int st = BDF_OK;
        uint32_t sum = 0, used = 0, at = 0, dlen = 0;
        const uint32_t len = (uint32_t)len64;
        if (!have) st = BDF_BAD_DATA;
        else if (len64 > 0xFFFFFFF0ull) st = BDF_BAD_DATA;      // single streams above 4 GiB are outside this engine's range
        else st = inflate_frame_header<FORMAT>(p, len, at, dlen);
        const bool go = st == BDF_OK;
        // all lanes of the warp call this together (see inflate_stream)
        const int ist = inflate_stream<FORMAT == BDF_ZLIB, G>(g, p + at, dlen, o, sm, &used, go, a.hdr_meta, a.hdr_rows, idx);
        if (go) {
            st = ist;
            if (st == BDF_OK && FORMAT == BDF_ZLIB) {
                sum = grp_adler_finish<G>(g, o.sumA, o.sumB, o.pos);
                const uint8_t *f = p + at + used;
                const uint32_t want = (uint32_t)f[0] << 24 | (uint32_t)f[1] << 16 | (uint32_t)f[2] << 8 | f[3];
                if (want != sum) st = BDF_BAD_DATA;
            }
            if (st == BDF_OK && FORMAT == BDF_GZIP) {
                sum = grp_crc32<G>(g, o.out, o.pos, s_crc, s_x2n);
                const uint8_t *f = p + at + used;
                const uint32_t want = (uint32_t)f[3] << 24 | (uint32_t)f[2] << 16 | (uint32_t)f[1] << 8 | f[0];
                const uint32_t isz = (uint32_t)f[7] << 24 | (uint32_t)f[6] << 16 | (uint32_t)f[5] << 8 | f[4];
                if (want != sum || isz != o.pos) st = BDF_BAD_DATA;
            }
        }
        if (have && g.lane == 0) {
            a.status[idx] = st;
            a.out_size[idx] = st == BDF_OK ? o.pos : 0;
            if (a.checksum) a.checksum[idx] = st == BDF_OK ? sum : 0;
        }
        __syncwarp();
    }
}

}  // namespace bdf
