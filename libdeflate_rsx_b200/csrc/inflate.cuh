// inflate.cuh — batch DEFLATE decompression, one warp per stream (sm_100a).
//
// Replaces Decompressor::decompress / decompress_zlib / decompress_gzip as
// called from BatchDecompressor::decompress_batch (reference src/batch.rs:74-101,
// src/decompress/mod.rs:164-202,1074-1240).  Accept/reject rules for Huffman
// code sets follow build_decode_table (src/decompress/mod.rs:1365-1383); the
// table SHAPE is our own: a 10-bit direct litlen table and an 8-bit offset
// table in shared memory, with codewords longer than the table resolved by a
// canonical first-code search (no sub-tables), all built lane-parallel.
//
// Execution model: the 32 lanes of a warp hold identical bit-reader state and
// decode the same symbol (table entries are shared-memory broadcasts, input
// words are uniform global loads), so a symbol costs one warp-wide dependent
// chain and no shuffles.  Literals are parked one per lane and flushed as a
// single 32-byte store; a match is copied by all lanes at once with its
// source loads hoisted ahead of the stores.  Adler-32 is accumulated from the
// bytes as they are written (position-weighted partial sums per lane), so the
// zlib path never re-reads its output.
#pragma once
#include "common.cuh"

namespace bdf {

constexpr int INF_WARPS_PER_BLOCK = 4;
constexpr int LT_BITS = 10;   // litlen direct-table bits
constexpr int OT_BITS = 8;    // offset direct-table bits
constexpr int PT_BITS = 7;    // precode direct-table bits

// Table entry (u32): [4:0] codeword bits (0 = longer than the table),
// [8:5] extra bits, [10:9] kind, [31:16] literal / base value.
constexpr uint32_t K_LIT = 0u << 9, K_BASE = 1u << 9, K_EOB = 2u << 9, K_MASK = 3u << 9;

struct HuffCode {            // canonical description used for build + long codes
    uint16_t first[16];      // first codeword of each length (MSB-first value)
    uint16_t count[16];
    uint16_t offs[16];       // index of the first symbol of each length in sorted[]
};

struct __align__(16) InflateWarpSmem {
    uint32_t lit_tab[1 << LT_BITS];
    uint32_t off_tab[1 << OT_BITS];   // the precode table overlays the front of this
    uint16_t lit_sorted[288];
    uint16_t off_sorted[32];
    HuffCode lit_code, off_code;
    uint32_t cnt[16];                  // scratch: length histogram / running ranks
    uint8_t lens[320 + 8];
};

// ------------------------------------------------------------------ bit reader
// Warp-uniform LSB-first reader over [p, p+len) using aligned 32-bit loads.
struct BitReader {
    const uint8_t *p;      // stream start
    uint32_t len;          // stream length in bytes
    uint32_t mis;          // p & 3
    uint32_t nwords;       // aligned words covering the stream
    uint32_t widx;         // next word to load
    uint64_t buf;
    int32_t left;          // valid bits in buf (may include zero fill past the end)

    __device__ __forceinline__ uint32_t load_word(uint32_t w) const
    {
        if (w >= nwords) return 0;                     // zero fill past the end
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(p - mis) + w;
        if (w != 0 && w + 1 != nwords) return __ldg(wp);
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
    }
    // keep at least 33 valid bits
    __device__ __forceinline__ void refill()
    {
        if (left <= 32) {
            buf |= (uint64_t)load_word(widx) << left;
            widx++;
            left += 32;
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
    // restart at byte offset `at` from the stream start
    __device__ __forceinline__ void seek(uint32_t at)
    {
        uint32_t a = mis + at;
        widx = a >> 2;
        uint32_t sh = 8 * (a & 3);
        buf = (uint64_t)load_word(widx) >> sh;
        left = 32 - (int32_t)sh;
        widx++;
    }
};

// ------------------------------------------------------------ table building
__device__ __forceinline__ uint32_t make_litlen_entry(unsigned sym, unsigned l)
{
    if (sym < 256) return (sym << 16) | K_LIT | l;
    if (sym == 256) return K_EOB | l;
    unsigned base, extra;
    length_slot_info(sym - 257, base, extra);
    return (base << 16) | K_BASE | (extra << 5) | l;
}
__device__ __forceinline__ uint32_t make_offset_entry(unsigned sym, unsigned l)
{
    unsigned base, extra;
    offset_slot_info(sym, base, extra);
    return (base << 16) | K_BASE | (extra << 5) | l;
}
__device__ __forceinline__ uint32_t make_precode_entry(unsigned sym, unsigned l) { return (sym << 16) | l; }

enum { CODE_LITLEN = 0, CODE_OFFSET = 1, CODE_PRECODE = 2 };

template <int KIND>
__device__ __forceinline__ uint32_t make_entry(unsigned sym, unsigned l)
{
    if (KIND == CODE_LITLEN) return make_litlen_entry(sym, l);
    if (KIND == CODE_OFFSET) return make_offset_entry(sym, l);
    return make_precode_entry(sym, l);
}

// Builds the direct table + canonical description for `nsyms` code lengths.
// Returns false (uniformly) for code sets the reference rejects:
// over-subscribed, or incomplete other than "no codes" / "one 1-bit code"
// (src/decompress/mod.rs:1365-1383).
template <int KIND, int TBITS>
__device__ bool build_code(const uint8_t *lens, unsigned nsyms, uint32_t *tab, uint16_t *sorted,
                           HuffCode &hc, uint32_t *cnt)
{
    const unsigned lane = lane_id();
    if (lane < 16) cnt[lane] = 0;
    __syncwarp();
    for (unsigned s = lane; s < nsyms; s += 32) {
        unsigned l = lens[s];
        if (l) atomicAdd(&cnt[l], 1u);
    }
    __syncwarp();
    // every lane derives the same canonical description
    uint32_t used = 0, code = 0, run = 0, total = 0;
    uint32_t my_first = 0, my_cnt = 0, my_off = 0;
#pragma unroll
    for (unsigned l = 1; l <= 15; l++) {
        uint32_t c = cnt[l];
        used += c << (15 - l);
        if (lane == l) { my_first = code; my_cnt = c; my_off = run; }
        code = (code + c) << 1;
        run += c;
        total += c;
    }
    const uint32_t c1 = cnt[1];
    __syncwarp();
    if (lane < 16) {
        hc.first[lane] = (uint16_t)my_first;
        hc.count[lane] = (uint16_t)my_cnt;
        hc.offs[lane] = (uint16_t)my_off;
        cnt[lane] = 0;      // becomes the running rank per length
    }
    __syncwarp();
    if (used > (1u << 15)) return false;
    if (used < (1u << 15)) {
        if (!(total == 0 || (total == 1 && c1 == 1))) return false;
        // every lookup yields symbol 0 (or the lone symbol) with a 1-bit codeword
        unsigned sym = 0;
        if (total == 1) {
            for (unsigned base = 0; base < nsyms; base += 32) {
                unsigned s = base + lane;
                unsigned hit = __ballot_sync(BDF_FULL_MASK, s < nsyms && lens[s] == 1);
                if (hit) sym = base + (__ffs(hit) - 1);
            }
        }
        uint32_t e = make_entry<KIND>(sym, 1);
        for (unsigned i = lane; i < (1u << TBITS); i += 32) tab[i] = e;
        if (lane < 16) { hc.count[lane] = 0; }   // no long codes
        __syncwarp();
        return true;
    }
    // sorted[] = symbols in (length, symbol) order
    for (unsigned base = 0; base < nsyms; base += 32) {
        unsigned s = base + lane;
        unsigned l = s < nsyms ? lens[s] : 0;
        unsigned peers = __match_any_sync(BDF_FULL_MASK, l);
        unsigned rank = __popc(peers & lanemask_lt());
        if (l) sorted[hc.offs[l] + cnt[l] + rank] = (uint16_t)s;
        __syncwarp();
        if (l && (peers >> lane) == 1u) cnt[l] += __popc(peers);   // highest lane of the group
        __syncwarp();
    }
    // direct table: each slot decodes its own index canonically
    for (unsigned i = lane; i < (1u << TBITS); i += 32) {
        unsigned x = __brev(i) >> (32 - TBITS);      // first TBITS bits of the codeword, MSB first
        uint32_t e = 0;                              // 0 = codeword longer than the table
#pragma unroll 1
        for (unsigned l = 1; l <= TBITS; l++) {
            unsigned d = (x >> (TBITS - l)) - hc.first[l];
            if (d < hc.count[l]) {
                e = make_entry<KIND>(sorted[hc.offs[l] + d], l);
                break;
            }
        }
        tab[i] = e;
    }
    __syncwarp();
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
    uint32_t npend;        // literals parked in lanes 0..npend-1
    uint32_t mylit;
    uint32_t sumA;         // Σ b           over bytes written by this lane
    uint64_t sumB;         // Σ i·b  (i = 0-based output index)
    uint32_t next_fold;    // output position at which the sums are folded mod 65521
};

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
__device__ __forceinline__ void flush_literals(OutState &o, unsigned lane)
{
    if (o.npend) {
        if (lane < o.npend) {
            o.out[o.pos + lane] = (uint8_t)o.mylit;
            if (ADLER) { o.sumA += o.mylit; o.sumB += (uint64_t)(o.pos + lane) * o.mylit; }
        }
        o.pos += o.npend;
        o.npend = 0;
    }
}

// ---- match copy ------------------------------------------------------------
// out[pos+i] = out[pos-offset+i], i < length, with the period-replication rule
// for offset < length (src/decompress/mod.rs:1259-1317, x86.rs copy_match_bmi2).
// The warp moves 8 destination-aligned bytes per lane per round: the source run
// is fetched as aligned 32-bit words and realigned with PRMT, so a 258-byte
// match is two rounds of LDG/PRMT/STG.64 instead of 258 byte moves, and the
// Adler-32 partial sums advance 8 bytes at a time with dp4a.
template <bool ADLER>
__device__ __forceinline__ void adler_acc8(OutState &o, uint32_t idx, uint32_t x, uint32_t y)
{
    if (ADLER) {
        uint32_t s = __dp4a(x, 0x01010101u, __dp4a(y, 0x01010101u, 0u));
        uint32_t w = __dp4a(x, 0x03020100u, __dp4a(y, 0x07060504u, 0u));
        o.sumA += s;
        o.sumB += (uint64_t)idx * s + w;
    }
}
template <bool ADLER>
__device__ __forceinline__ void adler_acc1(OutState &o, uint32_t idx, uint32_t b)
{
    if (ADLER) { o.sumA += b; o.sumB += (uint64_t)idx * b; }
}

// Splits [dpos, dpos+n) into head bytes (to 8-byte alignment), 8-byte words, tail bytes.
struct CopySplit { uint32_t head, nbody, tail; };
__device__ __forceinline__ CopySplit split_dst(const uint8_t *base, uint32_t dpos, uint32_t n)
{
    CopySplit c;
    c.head = (uint32_t)((8u - (uint32_t)(reinterpret_cast<uintptr_t>(base + dpos) & 7u)) & 7u);
    if (c.head > n) c.head = n;
    c.nbody = (n - c.head) >> 3;
    c.tail = (n - c.head) & 7u;
    return c;
}
// byte index (relative to dpos) this lane moves in the head/tail round, or ~0u
__device__ __forceinline__ uint32_t edge_index(const CopySplit &c, unsigned lane)
{
    if (lane < c.head) return lane;
    if (lane >= 8 && lane - 8 < c.tail) return c.head + c.nbody * 8 + (lane - 8);
    return 0xFFFFFFFFu;
}

// Non-overlapping forward copy inside the output buffer: out[dpos..dpos+n) = out[spos..spos+n),
// spos + n <= dpos.
template <bool ADLER>
__device__ __forceinline__ void warp_copy_fwd(OutState &o, unsigned lane, uint32_t dpos, uint32_t spos, uint32_t n)
{
    uint8_t *base = o.out;
    const CopySplit c = split_dst(base, dpos, n);
    const uint32_t ei = edge_index(c, lane);
    if (ei != 0xFFFFFFFFu) {
        uint32_t b = base[spos + ei];
        base[dpos + ei] = (uint8_t)b;
        adler_acc1<ADLER>(o, dpos + ei, b);
    }
    for (uint32_t w = lane; w < c.nbody; w += 32) {
        const uint32_t rel = c.head + 8 * w;
        const uint8_t *sp = base + spos + rel;
        const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(sp) & 3u);
        const uint32_t *sw = reinterpret_cast<const uint32_t *>(sp - sh);
        uint32_t w0 = sw[0], w1 = sw[1], w2 = sh ? sw[2] : 0u;
        const uint32_t sel = 0x3210u + 0x1111u * sh;
        uint32_t x = __byte_perm(w0, w1, sel), y = __byte_perm(w1, w2, sel);
        *reinterpret_cast<uint2 *>(base + dpos + rel) = make_uint2(x, y);
        adler_acc8<ADLER>(o, dpos + rel, x, y);
    }
}

// Overlapping copy with a short period (2 <= offset < 32, offset < length): every lane
// builds its 8 bytes from the period directly.
template <bool ADLER>
__device__ __forceinline__ void warp_copy_period(OutState &o, unsigned lane, uint32_t dpos, uint32_t offset,
                                                 uint32_t n)
{
    uint8_t *base = o.out;
    const uint8_t *pat = base + dpos - offset;
    const CopySplit c = split_dst(base, dpos, n);
    const uint32_t ei = edge_index(c, lane);
    if (ei != 0xFFFFFFFFu) {
        uint32_t b = pat[ei % offset];
        base[dpos + ei] = (uint8_t)b;
        adler_acc1<ADLER>(o, dpos + ei, b);
    }
    for (uint32_t w = lane; w < c.nbody; w += 32) {
        const uint32_t rel = c.head + 8 * w;
        uint32_t j = rel % offset;
        uint32_t v[2] = {0, 0};
#pragma unroll
        for (int k = 0; k < 8; k++) {
            v[k >> 2] |= (uint32_t)pat[j] << (8 * (k & 3));
            j = j + 1 == offset ? 0 : j + 1;
        }
        *reinterpret_cast<uint2 *>(base + dpos + rel) = make_uint2(v[0], v[1]);
        adler_acc8<ADLER>(o, dpos + rel, v[0], v[1]);
    }
}

// Run of one byte (offset == 1).
template <bool ADLER>
__device__ __forceinline__ void warp_fill(OutState &o, unsigned lane, uint32_t dpos, uint32_t n)
{
    uint8_t *base = o.out;
    const uint32_t b = base[dpos - 1];
    const uint32_t word = b * 0x01010101u;
    const CopySplit c = split_dst(base, dpos, n);
    const uint32_t ei = edge_index(c, lane);
    if (ei != 0xFFFFFFFFu) {
        base[dpos + ei] = (uint8_t)b;
        adler_acc1<ADLER>(o, dpos + ei, b);
    }
    for (uint32_t w = lane; w < c.nbody; w += 32) {
        const uint32_t rel = c.head + 8 * w;
        *reinterpret_cast<uint2 *>(base + dpos + rel) = make_uint2(word, word);
        adler_acc8<ADLER>(o, dpos + rel, word, word);
    }
}

template <bool ADLER>
__device__ __forceinline__ void copy_match(OutState &o, unsigned lane, unsigned length, unsigned offset)
{
    __syncwarp();   // earlier stores by other lanes may be our source
    if (offset >= length) {
        warp_copy_fwd<ADLER>(o, lane, o.pos, o.pos - offset, length);
    } else if (offset == 1) {
        warp_fill<ADLER>(o, lane, o.pos, length);
    } else if (offset < 32) {
        warp_copy_period<ADLER>(o, lane, o.pos, offset, length);
    } else {
        // period >= 32: each pass copies everything that is already periodic
        // ([pos-offset, pos+done) holds 1 + done/offset periods), doubling per pass
        uint32_t done = 0;
        for (;;) {
            uint32_t n = offset + done;
            if (n > length - done) n = length - done;
            warp_copy_fwd<ADLER>(o, lane, o.pos + done, o.pos - offset, n);
            done += n;
            if (done >= length) break;
            __syncwarp();
        }
    }
    o.pos += length;
}

// ------------------------------------------------------------ block decoding
template <bool ADLER>
__device__ int decode_huffman_block(BitReader &br, OutState &o, InflateWarpSmem &sm, unsigned lane)
{
    for (;;) {
        // more than two zero-fill words loaded: the stream ended inside this block
        if (br.widx > br.nwords + 2) return BDF_SHORT_INPUT;
        br.refill();
        uint32_t e = sm.lit_tab[br.peek(LT_BITS)];
        if ((e & 31u) == 0) {
            e = decode_long<CODE_LITLEN, LT_BITS>(br.peek(15), sm.lit_sorted, sm.lit_code);
            if (e == 0) return BDF_BAD_DATA;
        }
        br.drop(e & 31u);
        const uint32_t kind = e & K_MASK;
        if (kind == K_LIT) {
            if (o.pos + o.npend >= o.cap) return BDF_INSUFFICIENT_SPACE;
            if (lane == o.npend) o.mylit = e >> 16;
            if (++o.npend == 32) flush_literals<ADLER>(o, lane);
            continue;
        }
        if (kind == K_EOB) {
            return br.overrun() ? BDF_SHORT_INPUT : BDF_OK;
        }
        unsigned length = (e >> 16) + br.take((e >> 5) & 15u);
        br.refill();
        uint32_t f = sm.off_tab[br.peek(OT_BITS)];
        if ((f & 31u) == 0) {
            f = decode_long<CODE_OFFSET, OT_BITS>(br.peek(15), sm.off_sorted, sm.off_code);
            if (f == 0) return BDF_BAD_DATA;
        }
        br.drop(f & 31u);
        unsigned offset = (f >> 16) + br.take((f >> 5) & 15u);
        if (br.overrun()) return BDF_SHORT_INPUT;
        flush_literals<ADLER>(o, lane);
        if (offset > o.pos) return BDF_BAD_DATA;
        if (o.pos + length > o.cap) return BDF_INSUFFICIENT_SPACE;
        copy_match<ADLER>(o, lane, length, offset);
        if (ADLER) adler_fold(o);
    }
}

// read_dynamic_huffman_header, src/decompress/mod.rs:403-507
__device__ int read_dynamic_header(BitReader &br, InflateWarpSmem &sm, unsigned lane)
{
    br.refill();
    const unsigned nlit = 257 + br.take(5);
    const unsigned noff = 1 + br.take(5);
    const unsigned npre = 4 + br.take(4);
    // precode lengths in permutation order; lane k handles entry k
    uint8_t *pre_lens = sm.lens + 300;   // only needed until the precode table is built
    {
        // 19 * 3 = 57 bits: read in two steps
        br.refill();
        uint32_t lo = 0, hi = 0;
        unsigned n_lo = npre < 10 ? npre : 10;
        lo = br.take(3 * n_lo);
        br.refill();
        if (npre > 10) hi = br.take(3 * (npre - 10));
        // order 16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15 packed 5 bits each
        const uint64_t perm_lo = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 |
                                 9ull << 30 | 6ull << 35 | 10ull << 40 | 5ull << 45;
        const uint64_t perm_hi = 11ull | 4ull << 5 | 12ull << 10 | 3ull << 15 | 13ull << 20 | 2ull << 25 |
                                 14ull << 30 | 1ull << 35 | 15ull << 40;
        if (lane < 19) {
            unsigned sym = lane < 10 ? (unsigned)(perm_lo >> (5 * lane)) & 31u
                                     : (unsigned)(perm_hi >> (5 * (lane - 10))) & 31u;
            unsigned v = lane < 10 ? (lo >> (3 * lane)) & 7u : (hi >> (3 * (lane - 10))) & 7u;
            pre_lens[sym] = (uint8_t)(lane < npre ? v : 0);
        }
        __syncwarp();
    }
    if (br.overrun()) return BDF_SHORT_INPUT;
    uint32_t *pre_tab = sm.off_tab;
    if (!build_code<CODE_PRECODE, PT_BITS>(pre_lens, 19, pre_tab, sm.off_sorted, sm.off_code, sm.cnt))
        return BDF_BAD_DATA;
    // run-length decode of the litlen + offset code lengths (uniform; lane 0 stores)
    const unsigned total = nlit + noff;
    unsigned i = 0, prev = 0;
    while (i < total) {
        br.refill();
        uint32_t e = pre_tab[br.peek(PT_BITS)];
        br.drop(e & 31u);
        unsigned sym = e >> 16;
        if (sym < 16) {
            if (lane == 0) sm.lens[i] = (uint8_t)sym;
            prev = sym;
            i++;
            continue;
        }
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
        for (unsigned k = lane; k < rep; k += 32) sm.lens[i + k] = (uint8_t)val;
        prev = val;
        i += rep;
    }
    if (br.overrun()) return BDF_SHORT_INPUT;
    __syncwarp();
    if (!build_code<CODE_OFFSET, OT_BITS>(sm.lens + nlit, noff, sm.off_tab, sm.off_sorted, sm.off_code, sm.cnt))
        return BDF_BAD_DATA;
    if (!build_code<CODE_LITLEN, LT_BITS>(sm.lens, nlit, sm.lit_tab, sm.lit_sorted, sm.lit_code, sm.cnt))
        return BDF_BAD_DATA;
    return BDF_OK;
}

__device__ void load_static_codes(InflateWarpSmem &sm, unsigned lane)
{
    for (unsigned s = lane; s < 320; s += 32)
        sm.lens[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : s < 288 ? 8 : 5;
    __syncwarp();
    build_code<CODE_OFFSET, OT_BITS>(sm.lens + 288, 32, sm.off_tab, sm.off_sorted, sm.off_code, sm.cnt);
    build_code<CODE_LITLEN, LT_BITS>(sm.lens, 288, sm.lit_tab, sm.lit_sorted, sm.lit_code, sm.cnt);
}

// Raw DEFLATE stream [p, p+len) -> o; returns status, *used = bytes consumed.
template <bool ADLER>
__device__ int inflate_stream(const uint8_t *p, uint32_t len, OutState &o, InflateWarpSmem &sm,
                              unsigned lane, uint32_t *used)
{
    BitReader br;
    br.init(p, len);
    int st;
    for (;;) {
        br.refill();
        if (br.consumed_bits() + 3 > (int64_t)len * 8) { st = BDF_SHORT_INPUT; break; }
        const unsigned final = br.take(1);
        const unsigned type = br.take(2);
        if (type == 0) {
            // stored block (src/decompress/mod.rs:282-346, x86.rs:2216-2246)
            uint32_t at = (uint32_t)((br.consumed_bits() + 7) >> 3);
            if (at + 4 > len) { st = BDF_SHORT_INPUT; break; }
            unsigned blen = p[at] | (unsigned)p[at + 1] << 8;
            unsigned nlen = p[at + 2] | (unsigned)p[at + 3] << 8;
            at += 4;
            if (blen != (~nlen & 0xFFFFu)) { st = BDF_BAD_DATA; break; }
            if (o.pos + blen > o.cap) { st = BDF_INSUFFICIENT_SPACE; break; }
            if (at + blen > len) { st = BDF_SHORT_INPUT; break; }
            for (unsigned i = lane; i < blen; i += 32) {
                uint8_t b = p[at + i];
                o.out[o.pos + i] = b;
                if (ADLER) { o.sumA += b; o.sumB += (uint64_t)(o.pos + i) * b; }
            }
            o.pos += blen;
            if (ADLER) adler_fold(o);
            br.seek(at + blen);
        } else if (type == 3) {
            st = BDF_BAD_DATA;
            break;
        } else {
            if (type == 1) {
                load_static_codes(sm, lane);
            } else {
                st = read_dynamic_header(br, sm, lane);
                if (st != BDF_OK) break;
            }
            st = decode_huffman_block<ADLER>(br, o, sm, lane);
            flush_literals<ADLER>(o, lane);
            if (st != BDF_OK) break;
        }
        if (final) { st = BDF_OK; break; }
    }
    int64_t cb = br.consumed_bits();
    if (cb < 0) cb = 0;
    *used = (uint32_t)((cb + 7) >> 3);
    __syncwarp();
    return st;
}

// Warp-wide CRC-32 of out[0..n): 32 contiguous slices, slice-by-4 per lane,
// partial CRCs shifted to the end with x^(8k) mod P and XOR-reduced.
__device__ uint32_t warp_crc32(const uint8_t *d, uint64_t n, const uint32_t (*slice)[256],
                               const uint32_t *x2n, unsigned lane)
{
    uint64_t chunk = (n + 31) / 32;
    uint64_t beg = chunk * lane, end = beg + chunk;
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
    for (int s = 16; s > 0; s >>= 1) part ^= __shfl_xor_sync(BDF_FULL_MASK, part, s);
    return part;
}

// (Σb, Σ i·b) partials -> Adler-32 of n bytes with seed 1.
__device__ uint32_t warp_adler_finish(uint32_t sumA, uint64_t sumB, uint64_t n)
{
    uint64_t a = sumA % 65521u, b = sumB % 65521u;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        a += __shfl_xor_sync(BDF_FULL_MASK, a, s);
        b += __shfl_xor_sync(BDF_FULL_MASK, b, s);
    }
    a %= 65521u;
    b %= 65521u;
    uint64_t nm = n % 65521u;
    uint64_t s1 = (1 + a) % 65521u;
    uint64_t s2 = (nm + nm * a + 65521u - b) % 65521u;
    return (uint32_t)(s2 << 16 | s1);
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
    unsigned long long *work_counter;
    uint32_t n;
};

template <int FORMAT>
__global__ void __launch_bounds__(INF_WARPS_PER_BLOCK * 32, 8)
inflate_kernel(InflateArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_crc[FORMAT == BDF_GZIP ? 4 : 1][256];
    __shared__ uint32_t s_x2n[32];
    InflateWarpSmem &sm = reinterpret_cast<InflateWarpSmem *>(smem_raw)[threadIdx.x >> 5];
    const unsigned lane = lane_id();
    if (FORMAT == BDF_GZIP) {
        for (unsigned i = threadIdx.x; i < 1024; i += blockDim.x) s_crc[i >> 8][i & 255] = g_crc_tables.slice[i >> 8][i & 255];
        if (threadIdx.x < 32) s_x2n[threadIdx.x] = g_crc_tables.x2n[threadIdx.x];
        __syncthreads();
    }
    for (;;) {
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(a.work_counter, 1ull);
        idx = __shfl_sync(BDF_FULL_MASK, idx, 0);
        if (idx >= a.n) break;
        const uint8_t *p = a.in + a.in_off[idx];
        uint64_t len64 = a.in_off[idx + 1] - a.in_off[idx];
        uint64_t cap64 = a.max_out[idx];
        OutState o;
        o.out = a.out + a.out_off[idx];
        o.pos = 0;
        o.cap = cap64 > 0xFFFFFFF0ull ? 0xFFFFFFF0u : (uint32_t)cap64;
        o.npend = 0; o.mylit = 0; o.sumA = 0; o.sumB = 0; o.next_fold = ADLER_FOLD_INTERVAL;
        int st = BDF_OK;
        uint32_t sum = 0, used = 0;
        if (len64 > 0xFFFFFFF0ull) {
            st = BDF_BAD_DATA;      // single streams above 4 GiB are outside this engine's range
        } else if (FORMAT == BDF_RAW) {
            st = inflate_stream<false>(p, (uint32_t)len64, o, sm, lane, &used);
        } else if (FORMAT == BDF_ZLIB) {
            // decompress_zlib_uninit, src/decompress/mod.rs:1074-1127
            uint32_t len = (uint32_t)len64;
            if (len < 6) st = BDF_SHORT_INPUT;
            else {
                unsigned hdr = (unsigned)p[0] << 8 | p[1];
                if (hdr % 31 != 0 || ((hdr >> 8) & 0xF) != 8 || ((hdr >> 12) & 0xF) > 7 || ((hdr >> 5) & 1))
                    st = BDF_BAD_DATA;
                else {
                    st = inflate_stream<true>(p + 2, len - 6, o, sm, lane, &used);
                    if (st == BDF_OK) {
                        sum = warp_adler_finish(o.sumA, o.sumB, o.pos);
                        const uint8_t *f = p + 2 + used;
                        uint32_t want = (uint32_t)f[0] << 24 | (uint32_t)f[1] << 16 | (uint32_t)f[2] << 8 | f[3];
                        if (want != sum) st = BDF_BAD_DATA;
                    }
                }
            }
        } else {
            // decompress_gzip_uninit, src/decompress/mod.rs:1144-1240
            uint32_t len = (uint32_t)len64;
            if (len < 18) st = BDF_SHORT_INPUT;
            else if (p[0] != 0x1F || p[1] != 0x8B || p[2] != 8 || (p[3] & 0xE0)) st = BDF_BAD_DATA;
            else {
                unsigned flg = p[3];
                uint64_t at = 10;
                if (flg & 0x04) {
                    if (at + 2 > len) st = BDF_SHORT_INPUT;
                    else at += 2 + (p[at] | (uint64_t)p[at + 1] << 8);
                }
                if (st == BDF_OK && (flg & 0x08)) { while (at < len && p[at]) at++; at++; }
                if (st == BDF_OK && (flg & 0x10)) { while (at < len && p[at]) at++; at++; }
                if (st == BDF_OK && (flg & 0x02)) at += 2;
                if (st == BDF_OK && at + 8 > len) st = BDF_SHORT_INPUT;
                if (st == BDF_OK) {
                    st = inflate_stream<false>(p + at, (uint32_t)(len - 8 - at), o, sm, lane, &used);
                    if (st == BDF_OK) {
                        sum = warp_crc32(o.out, o.pos, s_crc, s_x2n, lane);
                        const uint8_t *f = p + at + used;
                        uint32_t want = (uint32_t)f[3] << 24 | (uint32_t)f[2] << 16 | (uint32_t)f[1] << 8 | f[0];
                        uint32_t isz = (uint32_t)f[7] << 24 | (uint32_t)f[6] << 16 | (uint32_t)f[5] << 8 | f[4];
                        if (want != sum || isz != o.pos) st = BDF_BAD_DATA;
                    }
                }
            }
        }
        if (lane == 0) {
            a.status[idx] = st;
            a.out_size[idx] = st == BDF_OK ? o.pos : 0;
            if (a.checksum) a.checksum[idx] = st == BDF_OK ? sum : 0;
        }
        __syncwarp();
    }
}

}  // namespace bdf
