// inflate_lane.cuh — batch DEFLATE decompression, one LANE per stream (sm_100a).
//
// Second engine behind BatchDecompressor::decompress_batch (reference src/batch.rs:74-101,
// src/decompress/mod.rs:164-202,509-1072,1074-1240), for streams that are dominated by literals
// and short matches (text, records, low-entropy bytes).  The lane-GROUP kernel of inflate.cuh spends
// one warp instruction on at most two streams and is at its best when a stream is a handful of
// long matches; ncu on mixed data showed it issue-bound with 15 of 16 lanes repeating the same
// table look-up (profiles/r1_inflate_mixed_summary.md).  Here every lane decodes its OWN stream
// the way a CPU core would, so one warp instruction advances 32 streams.  The launcher sends a
// stream to this kernel or to the lane-group kernel by its expansion ratio
// (InflateArgs::split_ratio).
//
// What limits a lane-per-stream decoder is the number of streams an SM can hold, i.e. shared
// memory per stream (the first version of this file kept 2 KiB of tables and a 1 KiB window per
// lane: 2 warps per SM, 22 GB/s on text).  Per lane there is now
//   * an 8-bit litlen and a 7-bit offset direct table (768 B).  Longer codewords are resolved by a
//     canonical search whose first-code / count pairs live in REGISTERS (7 + 8 words, loaded when
//     the block starts); only the symbol they select is read from a per-lane list in global
//     memory;
//   * a 128-byte output ring, indexed by the low bits of the GLOBAL output address.  The lane
//     writes literals and match bytes into it and moves whole 32-byte sectors to global memory
//     with 16-byte stores as soon as they are complete (Adler-32 partial sums come from the same
//     words, dp4a).  A match source further back than the ring is read from the lane's own
//     flushed output through L2; the load is issued when the match is decoded and consumed one
//     iteration later.
// 31 KB per warp: 7 warps (224 streams) per SM.
//  * Loop body = three predicated sections and no per-symbol branch between lanes: decode one
//    symbol (up to two literals, or a length/offset pair) | copy up to 8 bytes of the pending
//    match | flush one sector.
//  * Block headers are read by the whole warp for one stream at a time with the lane-group
//    kernel's code (read_code_lengths / build_code with G = 32): the owner's bit reader is
//    broadcast, the tables are built into the owner's slot, the reader is handed back.
#pragma once
#include "inflate.cuh"

namespace bdf {

constexpr int LANE_RING = 128;      // bytes of output a lane keeps in shared memory
constexpr int LANE_SORTED_BYTES = (288 + 32) * 2;      // per-lane symbol lists in global memory

// Direct-table sizes are a trade between codewords that miss the table and streams per SM:
// (8, 7) bits = 772 B per lane, 7 warps per SM; (9, 6) bits = 1156 B, 5 warps.
template <int LTB, int OTB>
struct LaneTab {                    // one per lane (stream slot)
    uint16_t lit_tab[1 << LTB];     // the precode table (128 entries) overlays it while a header is read
    uint16_t off_tab[1 << OTB];
    uint32_t pad;                   // odd stride in words: equal indices of different lanes fall into different banks
};
// GT: the per-lane tables live in GLOBAL memory (lane_scratch, read through L1) instead of shared
// memory: what is left in shared memory is 6.4 KB per warp, so registers (16 warps per SM) and not
// shared memory (7) decide how many streams an SM holds.
template <int LTB, int OTB, bool GT = false>
struct LaneSmem {                   // one per warp
    LaneTab<LTB, OTB> tab[GT ? 1 : 32];
    uint4 ring[32][(LANE_RING + 16) / 16];  // + 16: a copy step may spill up to 16 bytes past the end
    HuffCode lit_code, off_code;    // of the block whose header was read last
    BuildScratch<32> bs;
    uint8_t lens[328];
};
// what read_code_lengths / build_dynamic_codes / load_static_codes see (member names of InflateSmem)
struct LaneView {
    uint16_t *lit_tab, *off_tab, *lit_sorted, *off_sorted;
    HuffCode &lit_code, &off_code;
    BuildScratch<32> &bs;
    uint8_t *lens;
};
__device__ __forceinline__ uint16_t *precode_table(LaneView &v) { return v.lit_tab; }

enum { LS_NEW = 0, LS_IDLE, LS_HDR, LS_RUN, LS_END };

__device__ __forceinline__ BitReader bcast_reader(const BitReader &b, unsigned src)
{
    BitReader r;
    r.p = reinterpret_cast<const uint8_t *>(__shfl_sync(BDF_FULL_MASK, reinterpret_cast<unsigned long long>(b.p), src));
    r.len = __shfl_sync(BDF_FULL_MASK, b.len, src);
    r.mis = __shfl_sync(BDF_FULL_MASK, b.mis, src);
    r.nwords = __shfl_sync(BDF_FULL_MASK, b.nwords, src);
    r.widx = __shfl_sync(BDF_FULL_MASK, b.widx, src);
    r.ahead = __shfl_sync(BDF_FULL_MASK, b.ahead, src);
    r.buf = __shfl_sync(BDF_FULL_MASK, (unsigned long long)b.buf, src);
    r.left = __shfl_sync(BDF_FULL_MASK, b.left, src);
    return r;
}

__device__ __forceinline__ uint32_t sum4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
    return __dp4a(w0, 0x01010101u, __dp4a(w1, 0x01010101u, __dp4a(w2, 0x01010101u, __dp4a(w3, 0x01010101u, 0u))));
}
__device__ __forceinline__ uint32_t wsum4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3)
{
    return __dp4a(w0, 0x03020100u, __dp4a(w1, 0x07060504u, __dp4a(w2, 0x0B0A0908u, __dp4a(w3, 0x0F0E0D0Cu, 0u))));
}

// first / last words of a stream and the zero fill behind it: out of line, the decode loop only
// carries the interior case
__device__ __noinline__ uint32_t lane_load_word_edge(const uint8_t *p, uint32_t len, uint32_t mis, uint32_t nwords, uint32_t w)
{
    BitReader t;
    t.p = p; t.len = len; t.mis = mis; t.nwords = nwords;
    return t.load_word(w);
}
__device__ __forceinline__ void lane_refill(BitReader &br)
{
    if (br.left <= 32) {
        br.buf |= (uint64_t)br.ahead << br.left;
        br.widx++;
        br.left += 32;
        const uint32_t w = br.widx;
        if (w - 1u < br.nwords - 2u) br.ahead = __ldg(reinterpret_cast<const uint32_t *>(br.p - br.mis) + w);
        else br.ahead = lane_load_word_edge(br.p, br.len, br.mis, br.nwords, w);
    }
}

// Canonical description of the codewords longer than the direct table, in registers:
// c[k] = first codeword << 16 | count for length FIRST + k; base = index of the first such symbol
// in the sorted list.
template <int FIRST>
struct LongCodes {
    static constexpr int N = 16 - FIRST;
    uint32_t c[N];
    uint32_t base;
    __device__ __forceinline__ void load(const HuffCode &hc)
    {
#pragma unroll
        for (int k = 0; k < N; k++) c[k] = (uint32_t)hc.first[FIRST + k] << 16 | hc.count[FIRST + k];
        base = hc.offs[FIRST];
    }
    // -> index into the sorted list and codeword length, or length 0 (not a codeword)
    __device__ __forceinline__ void find(uint32_t bits15, uint32_t &index, uint32_t &length) const
    {
        const uint32_t x = __brev(bits15) >> 17;
        uint32_t acc = base;
        index = 0; length = 0;
#pragma unroll
        for (int k = 0; k < N; k++) {
            const uint32_t l = FIRST + k;
            const uint32_t d = (x >> (15 - l)) - (c[k] >> 16), cnt = c[k] & 0xFFFFu;
            if (length == 0 && d < cnt) { length = l; index = acc + d; }
            acc += cnt;
        }
    }
};

// ADAPT (with GT): the litlen table of a block is 8 or 9 bits wide — LTB is the room a lane has, the
// width is chosen per block from the weight of the codewords an 8-bit table would miss (Kraft mass of
// the lengths above 8): binary records put ~40 % of their symbols there, text a few percent, and a
// 9-bit table that is not needed doubles what a lane keeps in L1.
template <int FORMAT, int LTB, int OTB, bool GT = false, bool ADAPT = false>
__global__ void __launch_bounds__(32, GT ? 16 : 1) inflate_lane_kernel(InflateArgs a)
{
    static_assert(!ADAPT || (GT && LTB == 9), "adaptive width: 8 or 9 bits in a 9-bit table in global memory");
    constexpr int RING = LANE_RING;
    constexpr uint32_t MASK = RING - 1, WMASK = RING / 4 - 1;
    // A copy step rewrites up to 20 bytes from the word pos lies in, i.e. it may clobber ring
    // positions further back than RING - 17.  Ring sources are therefore at most NEAR back, anything
    // older is read from global memory and must have been flushed: a lane copies / decodes only
    // while its unflushed bytes stay below COPY_MAX / DECODE_MAX (only the ragged head of a stream,
    // flushed byte by byte, ever gets there).
    constexpr uint32_t NEAR = RING - 24;
    constexpr uint32_t COPY_MAX = RING - 56, DECODE_MAX = RING - 40;
    constexpr uint32_t FLUSH_AT = 56;             // < COPY_MAX - 16: a lane that reaches it still has a round of room
    constexpr bool ADLER = FORMAT == BDF_ZLIB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    LaneSmem<LTB, OTB, GT> &sm = *reinterpret_cast<LaneSmem<LTB, OTB, GT> *>(smem_raw);
    using LaneSmemT = LaneSmem<LTB, OTB, GT>;
    static_assert(offsetof(LaneSmemT, lens) % 4 == 0, "load_code_lengths stores words");
    const unsigned lane = threadIdx.x;
    const Grp<32> g;
    // GT: the tables of the warp's 32 slots sit behind the symbol lists of the warp's scratch
    LaneTab<LTB, OTB> *const tabs = GT ? reinterpret_cast<LaneTab<LTB, OTB> *>(
                                             a.lane_scratch + (size_t)gridDim.x * 32 * LANE_SORTED_BYTES) + (size_t)blockIdx.x * 32
                                       : sm.tab;
    LaneTab<LTB, OTB> &T = tabs[lane];
    uint8_t *const ring = reinterpret_cast<uint8_t *>(sm.ring[lane]);
    uint32_t *const ringw = reinterpret_cast<uint32_t *>(ring);
    uint16_t *const my_sorted = reinterpret_cast<uint16_t *>(
        a.lane_scratch + ((size_t)blockIdx.x * 32 + lane) * LANE_SORTED_BYTES);

    // ---- per-lane stream state
    BitReader br;
    br.p = a.in; br.len = 0; br.mis = 0; br.nwords = 0; br.widx = 0; br.ahead = 0; br.buf = 0; br.left = 0;
    LongCodes<ADAPT ? 9 : LTB + 1> lcode;
    uint32_t lmask = (1u << LTB) - 1u;       // ADAPT: mask of the width the current block's table was built with
    LongCodes<OTB + 1> ocode;
    lcode.load(sm.lit_code); ocode.load(sm.off_code);      // (values are replaced before they are used)
    const uint8_t *sp = a.in;        // stream start (framing included)
    uint8_t *out = a.out;
    uint32_t at = 0;                 // offset of the DEFLATE data
    uint32_t pos = 0, cap = 0, flushed = 0, ring_lo = 0, rbias = 0;
    uint32_t copy_rem = 0, copy_src = 0;
    uint64_t pfa = 0, pfb = 0, pfc = 0;     // the 24 aligned bytes around a far source, requested one round ahead
    bool pf_valid = false, final_blk = false;
    uint32_t sumA = 0;
    uint64_t sumB = 0;
    uint32_t idx = 0;
    uint32_t pre_meta = 0;           // first header of the stream as the pre-pass left it (inflate_prehdr.cuh), 0 = none
    int st = LS_NEW, status = BDF_OK;
    bool q_empty = false;

    // the aligned 8-byte words that hold out[src .. src + min(n, 16)) (L2: the lane wrote them itself)
    auto prefetch = [&](uint32_t src, uint32_t n) {
        const uintptr_t A = reinterpret_cast<uintptr_t>(out + src);
        const unsigned long long *wp = reinterpret_cast<const unsigned long long *>(A & ~(uintptr_t)7);
        const uint32_t need = (uint32_t)(A & 7u) + (n < 16u ? n : 16u);
        pfa = __ldcg(wp);
        if (need > 8) pfb = __ldcg(wp + 1);
        if (need > 16) pfc = __ldcg(wp + 2);
        pf_valid = true;
    };
    // everything the ring still holds -> global memory, byte by byte where it has to be (end of a
    // stream, stored block ahead)
    auto flush_rest = [&]() {
        while (flushed < pos) {
            if ((reinterpret_cast<uintptr_t>(out + flushed) & 15u) == 0 && flushed + 16 <= pos) {
                const uint32_t w = ((flushed + rbias) & MASK) >> 2;
                const uint32_t w0 = ringw[w], w1 = ringw[w + 1], w2 = ringw[w + 2], w3 = ringw[w + 3];
                *reinterpret_cast<uint4 *>(out + flushed) = make_uint4(w0, w1, w2, w3);
                if (ADLER) {
                    const uint32_t s = sum4(w0, w1, w2, w3);
                    sumA += s;
                    sumB += (uint64_t)flushed * s + wsum4(w0, w1, w2, w3);
                }
                flushed += 16;
            } else {
                const uint32_t b = ring[(flushed + rbias) & MASK];
                out[flushed] = (uint8_t)b;
                if (ADLER) { sumA += b; sumB += (uint64_t)flushed * b; }
                flushed++;
            }
        }
    };

    for (;;) {
        // =============================================================== service
        for (;;) {
            // (1) finished streams: tail, checksum, results
            unsigned todo = __ballot_sync(BDF_FULL_MASK, st == LS_END);
            while (todo) {
                const unsigned s = __ffs(todo) - 1;
                todo &= todo - 1;
                if (lane == s && status == BDF_OK) flush_rest();
                __syncwarp();
                uint32_t crc = 0;
                if (FORMAT == BDF_GZIP) {
                    const int st_s = __shfl_sync(BDF_FULL_MASK, status, s);
                    uint8_t *out_s = reinterpret_cast<uint8_t *>(__shfl_sync(BDF_FULL_MASK, reinterpret_cast<unsigned long long>(out), s));
                    const uint32_t n_s = __shfl_sync(BDF_FULL_MASK, pos, s);
                    if (st_s == BDF_OK) crc = grp_crc32<32>(g, out_s, n_s, g_crc_tables.slice, g_crc_tables.x2n);
                }
                if (lane == s) {
                    uint32_t sum = 0;
                    if (status == BDF_OK && FORMAT != BDF_RAW) {
                        int64_t cb = br.consumed_bits();
                        if (cb < 0) cb = 0;
                        const uint8_t *f = sp + at + (uint32_t)((cb + 7) >> 3);
                        if (FORMAT == BDF_ZLIB) {
                            const uint64_t M = 65521u, nm = pos % M, sa = sumA % M, sb = sumB % M;
                            sum = (uint32_t)(((nm + nm * sa + M - sb) % M) << 16 | ((1 + sa) % M));
                            const uint32_t want = (uint32_t)f[0] << 24 | (uint32_t)f[1] << 16 | (uint32_t)f[2] << 8 | f[3];
                            if (want != sum) status = BDF_BAD_DATA;
                        } else {
                            sum = crc;
                            const uint32_t want = (uint32_t)f[3] << 24 | (uint32_t)f[2] << 16 | (uint32_t)f[1] << 8 | f[0];
                            const uint32_t isz = (uint32_t)f[7] << 24 | (uint32_t)f[6] << 16 | (uint32_t)f[5] << 8 | f[4];
                            if (want != sum || isz != pos) status = BDF_BAD_DATA;
                        }
                    }
                    a.status[idx] = status;
                    a.out_size[idx] = status == BDF_OK ? pos : 0;
                    if (a.checksum) a.checksum[idx] = status == BDF_OK ? sum : 0;
                    st = LS_NEW;
                }
                __syncwarp();
            }
            // (2) free slots take the next streams of this kernel's class
            for (;;) {
                const unsigned need = __ballot_sync(BDF_FULL_MASK, st == LS_NEW);
                if (!need || q_empty) break;
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(a.work_counter2, (unsigned long long)__popc(need));
                base = __shfl_sync(BDF_FULL_MASK, base, 0);
                if (base + __popc(need) >= a.n) q_empty = true;
                if (st == LS_NEW) {
                    const unsigned long long my = base + __popc(need & ((1u << lane) - 1u));
                    if (my < a.n) {
                        const uint64_t o0 = a.in_off[my], len64 = a.in_off[my + 1] - o0, cap64 = a.max_out[my];
                        if (!inflate_is_heavy(a.split_ratio, len64, cap64)) {
                            idx = (uint32_t)my;
                            sp = a.in + o0;
                            out = a.out + a.out_off[my];
                            cap = cap64 > INFLATE_CAP_MAX ? INFLATE_CAP_MAX : (uint32_t)cap64;
                            pos = 0; flushed = 0; ring_lo = 0; copy_rem = 0; pf_valid = false; final_blk = false;
                            sumA = 0; sumB = 0;
                            rbias = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 31u);
                            status = BDF_OK;
                            uint32_t dlen = 0;
                            at = 0;
                            if (len64 > 0xFFFFFFF0ull) status = BDF_BAD_DATA;    // outside this engine's range
                            else status = inflate_frame_header<FORMAT>(sp, (uint32_t)len64, at, dlen);
                            pre_meta = 0;
                            if (status == BDF_OK) {
                                if (a.hdr_meta) pre_meta = __ldg(a.hdr_meta + my);
                                if (pre_meta & PREHDR_VALID) br.attach(sp + at, dlen);
                                else br.init(sp + at, dlen);
                                st = LS_HDR;
                            } else st = LS_END;
                        }
                    }
                }
            }
            if (q_empty && st == LS_NEW) st = LS_IDLE;
            // (3) block headers, one stream at a time, all lanes
            unsigned hdr = __ballot_sync(BDF_FULL_MASK, st == LS_HDR);
            while (hdr) {
                const unsigned s = __ffs(hdr) - 1;
                hdr &= hdr - 1;
                BitReader b = bcast_reader(br, s);
                uint32_t pm = __shfl_sync(BDF_FULL_MASK, pre_meta, s);
                const uint32_t *prow = a.hdr_rows + (size_t)__shfl_sync(BDF_FULL_MASK, idx, s) * PREHDR_ROW_WORDS;
                int hst = BDF_OK;           // status that ends the stream (uniform)
                bool ended = false, fin = false, narrow = false;
                LaneTab<LTB, OTB> &Ts = tabs[s];
                uint16_t *sorted_s = reinterpret_cast<uint16_t *>(
                    a.lane_scratch + ((size_t)blockIdx.x * 32 + s) * LANE_SORTED_BYTES);
                LaneView v{Ts.lit_tab, Ts.off_tab, sorted_s, sorted_s + 288, sm.lit_code, sm.off_code, sm.bs, sm.lens};
                for (;;) {                  // stored blocks are consumed here, one after the other
                    unsigned type = 2;
                    if (pm & PREHDR_VALID) fin = (pm >> 26 & 1u) != 0;
                    else {
                        b.refill();
                        if (b.consumed_bits() + 3 > (int64_t)b.len * 8) { hst = BDF_SHORT_INPUT; ended = true; break; }
                        fin = b.take(1) != 0;
                        type = b.take(2);
                    }
                    if (type == 1) {
                        load_static_codes<32, LaneView, LTB, OTB>(g, v);       // 112 of its literals have 9-bit codes
                        break;
                    }
                    if (type == 2) {
                        uint32_t nlong;
                        unsigned nlit = 0, noff = 0;
                        if (pm & PREHDR_VALID) load_code_lengths<32, LaneView>(g, b, v, pm, prow, nlit, noff);
                        else hst = read_code_lengths<32, LaneView, LTB, OTB>(g, b, v, nlit, noff);
                        if (ADAPT && hst == BDF_OK) {
                            // Kraft mass (units of 2^-15) of the litlen codewords longer than 8 bits
                            uint32_t mass = 0;
                            for (unsigned q = lane; q < nlit; q += 32) {
                                const unsigned l = sm.lens[q];
                                if (l > 8) mass += 1u << (15 - l);
                            }
#pragma unroll
                            for (int d = 16; d > 0; d >>= 1) mass += __shfl_xor_sync(BDF_FULL_MASK, mass, d);
                            narrow = mass < (1u << 15) / 8;              // < 12.5 % of the symbols: 8 bits
                        }
                        if (hst == BDF_OK)
                            hst = (ADAPT && narrow) ? build_dynamic_codes<32, LaneView, 8, OTB>(g, v, nlit, noff, nlong)
                                                    : build_dynamic_codes<32, LaneView, LTB, OTB>(g, v, nlit, noff, nlong);
                        if (hst != BDF_OK) ended = true;
                        break;
                    }
                    if (type == 3) { hst = BDF_BAD_DATA; ended = true; break; }
                    // stored block (src/decompress/mod.rs:282-346): straight from the input to the output in
                    // global memory, after the owner has written out what its ring still holds
                    if (lane == s) flush_rest();
                    __syncwarp();
                    const uint32_t pos_s = __shfl_sync(BDF_FULL_MASK, pos, s), cap_s = __shfl_sync(BDF_FULL_MASK, cap, s);
                    uint8_t *out_s = reinterpret_cast<uint8_t *>(__shfl_sync(BDF_FULL_MASK, reinterpret_cast<unsigned long long>(out), s));
                    uint32_t sat = (uint32_t)((b.consumed_bits() + 7) >> 3);
                    unsigned blen = 0;
                    if (sat + 4 > b.len) hst = BDF_SHORT_INPUT;
                    else {
                        blen = b.p[sat] | (unsigned)b.p[sat + 1] << 8;
                        const unsigned nlen = b.p[sat + 2] | (unsigned)b.p[sat + 3] << 8;
                        sat += 4;
                        if (blen != (~nlen & 0xFFFFu)) hst = BDF_BAD_DATA;
                        else if (blen > cap_s - pos_s) hst = BDF_INSUFFICIENT_SPACE;
                        else if (sat + blen > b.len) hst = BDF_SHORT_INPUT;
                    }
                    if (hst != BDF_OK) { ended = true; break; }
                    uint32_t pa = 0;
                    uint64_t pb = 0;
                    for (unsigned i = lane; i < blen; i += 32) {
                        const uint32_t x = b.p[sat + i];
                        out_s[pos_s + i] = (uint8_t)x;
                        if (ADLER) { pa += x; pb += (uint64_t)(pos_s + i) * x; }
                    }
                    if (ADLER) {
                        pb %= 65521u;
#pragma unroll
                        for (int d = 16; d > 0; d >>= 1) {
                            pa += __shfl_xor_sync(BDF_FULL_MASK, pa, d);
                            pb += __shfl_xor_sync(BDF_FULL_MASK, pb, d);
                        }
                    }
                    __syncwarp();
                    if (lane == s) {
                        BDF_ASSERT(pos + blen <= cap);
                        pos += blen; flushed = pos; ring_lo = pos;
                        if (ADLER) { sumA = (sumA + pa) % 65521u; sumB = (sumB + pb) % 65521u; }
                    }
                    b.seek(sat + blen);
                    if (fin) { ended = true; break; }
                }
                __syncwarp();               // tables and code descriptions of slot s were written by all lanes
                if (lane == s) {
                    br = b;
                    pre_meta = 0;
                    final_blk = fin;
                    if (ended) { status = hst; st = LS_END; }
                    else {
                        st = LS_RUN;
                        lcode.load(sm.lit_code);
                        ocode.load(sm.off_code);
                        if (ADAPT) lmask = narrow ? 255u : 511u;
                    }
                }
                __syncwarp();               // before the next header overwrites the code descriptions
            }
            if (!__any_sync(BDF_FULL_MASK, st == LS_END || st == LS_HDR || (st == LS_NEW && !q_empty))) break;
        }
        if (!__any_sync(BDF_FULL_MASK, st == LS_RUN)) break;

        // ================================================================ decode
        for (;;) {
            if (__any_sync(BDF_FULL_MASK, st == LS_END || st == LS_HDR)) break;
            // ---- B: up to 16 bytes of the pending match.  It was decoded in section A of an earlier
            // round, so the loads of a far source (issued there) have had a whole round to arrive.
            if (st == LS_RUN && copy_rem != 0 && pos - flushed <= COPY_MAX) {
                const uint32_t avail = pos - copy_src;          // distance to the source, >= 1
                const uint32_t d = (pos + rbias) & MASK;
                uint32_t n = copy_rem < 16u ? copy_rem : 16u;
                if (avail < n) n = avail;                       // overlapping: one period at a time
                if ((uint32_t)RING - d < n) n = (uint32_t)RING - d;      // the bytes that count do not wrap
                uint32_t x0, x1, x2, x3;
                if (avail > NEAR || copy_src < ring_lo) {
                    if (copy_src < ring_lo && ring_lo - copy_src < n) n = ring_lo - copy_src;
                    if (!pf_valid) prefetch(copy_src, n);
                    const uintptr_t A = reinterpret_cast<uintptr_t>(out + copy_src);
                    const uint32_t sh = 8u * (uint32_t)(A & 3u);
                    const bool odd = (A & 4u) != 0;
                    const uint32_t g0 = (uint32_t)pfa, g1 = (uint32_t)(pfa >> 32), g2 = (uint32_t)pfb, g3 = (uint32_t)(pfb >> 32),
                                   g4 = (uint32_t)pfc, g5 = (uint32_t)(pfc >> 32);
                    const uint32_t h0 = odd ? g1 : g0, h1 = odd ? g2 : g1, h2 = odd ? g3 : g2, h3 = odd ? g4 : g3, h4 = odd ? g5 : g4;
                    x0 = __funnelshift_r(h0, h1, sh); x1 = __funnelshift_r(h1, h2, sh);
                    x2 = __funnelshift_r(h2, h3, sh); x3 = __funnelshift_r(h3, h4, sh);
                    pf_valid = false;
                } else {
                    const uint32_t i = (copy_src + rbias) & MASK, w = i >> 2, sh = 8u * (i & 3u);
                    const uint32_t a0 = ringw[w], a1 = ringw[(w + 1) & WMASK], a2 = ringw[(w + 2) & WMASK],
                                   a3 = ringw[(w + 3) & WMASK], a4 = ringw[(w + 4) & WMASK];
                    x0 = __funnelshift_r(a0, a1, sh); x1 = __funnelshift_r(a1, a2, sh);
                    x2 = __funnelshift_r(a2, a3, sh); x3 = __funnelshift_r(a3, a4, sh);
                }
                // Five aligned words are stored whatever n is: the bytes in front of pos are put back,
                // what lies beyond pos + n is overwritten before it is read (it aliases output that is
                // flushed and further back than any ring source, or the slack behind the ring).
                const uint32_t wd = d >> 2, ss = 8u * (d & 3u);
                BDF_ASSERT(wd + 4 < (RING + 16) / 4 && n >= 1 && pos + n <= cap && pos - flushed + n <= (uint32_t)RING - 24);
                BDF_ASSERT(!(avail > NEAR || copy_src < ring_lo) || copy_src + n <= flushed);      // a far source is flushed output
                const uint32_t old = ringw[wd];
                ringw[wd] = (old & ((1u << ss) - 1u)) | (x0 << ss);
                ringw[wd + 1] = __funnelshift_l(x0, x1, ss);
                ringw[wd + 2] = __funnelshift_l(x1, x2, ss);
                ringw[wd + 3] = __funnelshift_l(x2, x3, ss);
                ringw[wd + 4] = __funnelshift_l(x3, 0u, ss);
                pos += n;
                copy_rem -= n;
                // a whole period copied (n == avail < 16): the source stays where it is and the distance
                // doubles; otherwise source and destination advance together
                if (n < avail) copy_src += n;
                if (copy_rem != 0 && (avail > NEAR || copy_src < ring_lo)) prefetch(copy_src, copy_rem);
            }
            // ---- A: one symbol: up to two literals, or a length / offset pair
            if (st == LS_RUN && copy_rem == 0 && pos - flushed <= DECODE_MAX) {
                // more than two zero-fill words loaded: the stream ended inside this block
                if (br.left <= 32 && br.widx > br.nwords + 2) { status = BDF_SHORT_INPUT; st = LS_END; }
                else {
                    lane_refill(br);
                    uint32_t e = T.lit_tab[ADAPT ? ((uint32_t)br.buf & lmask) : br.peek(LTB)];
                    if (e & LITFLAG) {
                        if (pos >= cap) { status = BDF_INSUFFICIENT_SPACE; st = LS_END; }
                        else {
                            ring[(pos + rbias) & MASK] = (uint8_t)(e >> E_VAL);
                            pos++;
                            br.drop(e & E_LEN);
                            // literals come in runs: the second look-up needs no refill (>= 25 valid bits)
                            e = T.lit_tab[ADAPT ? ((uint32_t)br.buf & lmask) : br.peek(LTB)];
                            if ((e & LITFLAG) && pos < cap) {
                                ring[(pos + rbias) & MASK] = (uint8_t)(e >> E_VAL);
                                pos++;
                                br.drop(e & E_LEN);
                            }
                        }
                    } else {
                        if ((e & E_LEN) == 0) {
                            uint32_t li, ll;
                            lcode.find(br.peek(15), li, ll);
                            BDF_ASSERT(ll == 0 || li < 288);
                            e = ll ? make_litlen_entry(__ldcg(my_sorted + li), ll) : 0u;
                        }
                        const uint32_t kind = e & K_MASK;
                        if (e == 0) { status = BDF_BAD_DATA; st = LS_END; }
                        else if (kind == K_LIT) {                  // a literal with a codeword longer than the table
                            if (pos >= cap) { status = BDF_INSUFFICIENT_SPACE; st = LS_END; }
                            else { ring[(pos + rbias) & MASK] = (uint8_t)(e >> E_VAL); pos++; br.drop(e & E_LEN); }
                        } else if (kind == K_EOB) {
                            br.drop(e & E_LEN);
                            status = br.overrun() ? BDF_SHORT_INPUT : BDF_OK;
                            st = (status == BDF_OK && !final_blk) ? LS_HDR : LS_END;
                        } else {
                            br.drop(e & E_LEN);
                            const unsigned length = take_length(br, e);
                            lane_refill(br);
                            uint32_t f = T.off_tab[br.peek(OTB)];
                            if ((f & E_LEN) == 0) {
                                uint32_t oi, ol;
                                ocode.find(br.peek(15), oi, ol);
                                BDF_ASSERT(ol == 0 || oi < 32);
                                f = ol ? make_offset_entry(__ldcg(my_sorted + 288 + oi), ol) : 0u;
                            }
                            if (f == 0) { status = BDF_BAD_DATA; st = LS_END; }
                            else {
                                br.drop(f & E_LEN);
                                const unsigned offset = take_offset(br, f);
                                if (offset > pos) { status = BDF_BAD_DATA; st = LS_END; }
                                else if (length > cap - pos) { status = BDF_INSUFFICIENT_SPACE; st = LS_END; }
                                else {
                                    copy_rem = length;
                                    copy_src = pos - offset;
                                    pf_valid = false;
                                    if (offset > NEAR || copy_src < ring_lo) prefetch(copy_src, length);
                                }
                            }
                        }
                    }
                }
            }
            // ---- C: complete 32-byte sectors of the ring -> global memory.  The section runs for the whole
            // warp once a lane has FLUSH_AT bytes waiting (every second or third round) instead of
            // every round for the few lanes that happen to have a sector ready.
            if (__any_sync(BDF_FULL_MASK, st == LS_RUN && pos - flushed >= FLUSH_AT)) {
                if (st == LS_RUN) {
                    const uint32_t mis = (flushed + rbias) & 31u;   // == (out + flushed) & 31
                    if (mis != 0) {
                        // ragged head of a stream (or behind a stored block): byte by byte up to the boundary
                        uint32_t k = 32u - mis;
                        if (k > pos - flushed) k = pos - flushed;
                        for (; k; k--) {
                            const uint32_t b = ring[(flushed + rbias) & MASK];
                            out[flushed] = (uint8_t)b;
                            if (ADLER) { sumA += b; sumB += (uint64_t)flushed * b; }
                            flushed++;
                        }
                    } else if (pos - flushed >= 32u) {
                        const uint4 *rq = reinterpret_cast<const uint4 *>(ring + ((flushed + rbias) & MASK));
                        const uint4 v0 = rq[0], v1 = rq[1];
                        uint4 *dst = reinterpret_cast<uint4 *>(out + flushed);
                        BDF_ASSERT(flushed + 32 <= pos && pos <= cap && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
                        dst[0] = v0;
                        dst[1] = v1;
                        if (ADLER) {
                            const uint32_t s0 = sum4(v0.x, v0.y, v0.z, v0.w), s1 = sum4(v1.x, v1.y, v1.z, v1.w);
                            sumA += s0 + s1;
                            sumB += (uint64_t)flushed * (s0 + s1) +
                                    (16u * s1 + wsum4(v0.x, v0.y, v0.z, v0.w) + wsum4(v1.x, v1.y, v1.z, v1.w));
                        }
                        flushed += 32;
                    }
                }
            }
        }
        if (ADLER) { sumA %= 65521u; sumB %= 65521u; }     // between two services a lane adds far less than 2^31 / 2^63
    }
}

}  // namespace bdf
