// inflate_lane.cuh — batch DEFLATE decompression, one LANE per stream (sm_100a).
//
// Second engine behind BatchDecompressor::decompress_batch (reference src/batch.rs:74-101,
// src/decompress/mod.rs:164-202,509-1072,1074-1240), for streams that are dominated by literals
// and short matches (text, records, low-entropy bytes).  The lane-GROUP kernel of inflate.cuh spends
// one warp instruction on at most two streams and is at its best when a stream is a handful of
// long matches; ncu on mixed data showed it issue-bound with 15 of 16 lanes repeating the same
// table look-up (profiles/r1_inflate_mixed_summary.md).  Here every lane decodes its OWN stream
// the way a CPU core would — bit buffer in registers, Huffman tables of its stream in shared
// memory, LZ77 copies out of a private shared-memory ring — so one warp instruction advances 32
// streams.  The launcher sends a stream to this kernel or to the lane-group kernel by its
// expansion ratio (InflateArgs::split_ratio).
//
//  * Loop body = two predicated sections, no per-symbol branches between lanes: a lane is either
//    decoding a symbol (up to two literals, or one length/offset pair) or copying up to 8 bytes
//    of a pending match.  A match whose source lies further back than the ring is read from the
//    lane's own flushed output in global memory; the load is issued when the match is decoded
//    and used one iteration later, so its L2 latency hides behind the other lanes' work.
//  * The ring is indexed by the low bits of the GLOBAL output address, so a lane flushes its own
//    ring with 16-byte aligned stores (whole 32-byte sectors) and the Adler-32 partial sums are
//    taken from the same words (dp4a).  All lanes flush together when the first ring is full.
//  * Block headers are read by the whole warp for one stream at a time with the lane-group
//    kernel's code (read_dynamic_header / build_code with G = 32): the owner's bit reader is
//    broadcast, the tables are built into the owner's slot, the reader is handed back.
#pragma once
#include "inflate.cuh"

namespace bdf {

template <int LTB>
struct LaneTables {                 // one per lane (stream slot)
    uint16_t lit_tab[1 << LTB];
    uint16_t off_tab[1 << OT_BITS];
    uint16_t lit_sorted[288];
    uint16_t off_sorted[32];
    HuffCode lit_code, off_code;
    uint32_t pad;                   // odd stride in words: equal indices of different lanes fall into different banks
};
template <int LTB, int RING>
struct LaneSmem {                   // one per warp
    LaneTables<LTB> tab[32];
    uint8_t ring[32][RING + 4];     // + 4: odd word stride, 4-byte aligned rows
    BuildScratch<32> bs;
    uint8_t lens[328];
};
// what read_dynamic_header / load_static_codes see (member names of InflateSmem)
template <int LTB>
struct LaneView {
    uint16_t *lit_tab, *off_tab, *lit_sorted, *off_sorted;
    HuffCode &lit_code, &off_code;
    BuildScratch<32> &bs;
    uint8_t *lens;
};

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

template <int FORMAT, int LTB, int RING>
__global__ void __launch_bounds__(32) inflate_lane_kernel(InflateArgs a)
{
    static_assert(RING >= 256 && (RING & (RING - 1)) == 0, "ring size");
    constexpr uint32_t MASK = RING - 1, WMASK = RING / 4 - 1;
    constexpr uint32_t ROOM = RING - 32;          // a lane runs while its unflushed bytes are at most this
    constexpr uint32_t NEAR = RING - 16;          // a source at most this far back is read from the ring
    constexpr bool ADLER = FORMAT == BDF_ZLIB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_crc[FORMAT == BDF_GZIP ? 4 : 1][256];
    __shared__ uint32_t s_x2n[32];
    using SM = LaneSmem<LTB, RING>;
    SM &sm = *reinterpret_cast<SM *>(smem_raw);
    const unsigned lane = threadIdx.x;
    const Grp<32> g;
    if (FORMAT == BDF_GZIP) {
        for (unsigned i = lane; i < 1024; i += 32) s_crc[i >> 8][i & 255] = g_crc_tables.slice[i >> 8][i & 255];
        s_x2n[lane] = g_crc_tables.x2n[lane];
        __syncwarp();
    }
    LaneTables<LTB> &T = sm.tab[lane];
    uint8_t *const ring = sm.ring[lane];
    uint32_t *const ringw = reinterpret_cast<uint32_t *>(ring);

    // ---- per-lane stream state
    BitReader br;
    br.p = a.in; br.len = 0; br.mis = 0; br.nwords = 0; br.widx = 0; br.ahead = 0; br.buf = 0; br.left = 0;
    const uint8_t *sp = a.in;        // stream start (framing included)
    uint8_t *out = a.out;
    uint32_t slen = 0, at = 0;       // stream length, offset of the DEFLATE data
    uint32_t pos = 0, cap = 0, flushed = 0, ring_lo = 0, rbias = 0;
    uint32_t copy_rem = 0, copy_src = 0, pf0 = 0, pf1 = 0, pf2 = 0;
    bool pf_valid = false, final_blk = false;
    uint32_t sumA = 0;
    uint64_t sumB = 0;
    uint32_t idx = 0;
    int st = LS_NEW, status = BDF_OK;
    bool q_empty = false;

    auto ring_put = [&](uint32_t x, uint32_t b) { ring[(x + rbias) & MASK] = (uint8_t)b; };
    // issue the loads of the 8 source bytes at out[src ..) (aligned words, L2: the lane wrote them itself)
    auto prefetch = [&](uint32_t src) {
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(
            reinterpret_cast<uintptr_t>(out + src) & ~(uintptr_t)3);
        pf0 = __ldcg(wp); pf1 = __ldcg(wp + 1); pf2 = __ldcg(wp + 2);
        pf_valid = true;
    };
    // aligned part of the lane's ring -> global memory (16-byte stores, whole 32-byte sectors);
    // everything = true also writes the ragged tail (end of stream, stored block ahead)
    auto flush_local = [&](bool everything) {
        while (flushed < pos && (reinterpret_cast<uintptr_t>(out + flushed) & 15u)) {
            const uint32_t b = ring[(flushed + rbias) & MASK];
            out[flushed] = (uint8_t)b;
            if (ADLER) { sumA += b; sumB += (uint64_t)flushed * b; }
            flushed++;
        }
        const uint32_t tail = (uint32_t)(reinterpret_cast<uintptr_t>(out + pos) & 31u);
        const uint32_t end = pos >= tail ? pos - tail : 0;
        while (flushed + 16 <= end) {
            const uint32_t w = ((flushed + rbias) & MASK) >> 2;       // multiple of 4: (out + flushed) is 16-byte aligned
            const uint32_t w0 = ringw[w], w1 = ringw[w + 1], w2 = ringw[w + 2], w3 = ringw[w + 3];
            *reinterpret_cast<uint4 *>(out + flushed) = make_uint4(w0, w1, w2, w3);
            if (ADLER) {
                const uint32_t s = sum4(w0, w1, w2, w3);
                sumA += s;
                sumB += (uint64_t)flushed * s + wsum4(w0, w1, w2, w3);
            }
            flushed += 16;
        }
        if (everything) {
            while (flushed < pos) {
                const uint32_t b = ring[(flushed + rbias) & MASK];
                out[flushed] = (uint8_t)b;
                if (ADLER) { sumA += b; sumB += (uint64_t)flushed * b; }
                flushed++;
            }
        }
        if (ADLER) { sumA %= 65521u; sumB %= 65521u; }
    };

    for (;;) {
        // =============================================================== service
        // (1) everybody flushes when somebody's ring is full
        if (__any_sync(BDF_FULL_MASK, st == LS_RUN && pos - flushed > ROOM)) {
            if (st == LS_RUN) flush_local(false);
            __syncwarp();
        }
        for (;;) {
            // (2) finished streams: tail, checksum, results
            unsigned todo = __ballot_sync(BDF_FULL_MASK, st == LS_END);
            while (todo) {
                const unsigned s = __ffs(todo) - 1;
                todo &= todo - 1;
                if (lane == s && status == BDF_OK) flush_local(true);
                __syncwarp();
                uint32_t crc = 0;
                if (FORMAT == BDF_GZIP) {
                    const int st_s = __shfl_sync(BDF_FULL_MASK, status, s);
                    uint8_t *out_s = reinterpret_cast<uint8_t *>(__shfl_sync(BDF_FULL_MASK, reinterpret_cast<unsigned long long>(out), s));
                    const uint32_t n_s = __shfl_sync(BDF_FULL_MASK, pos, s);
                    if (st_s == BDF_OK) crc = grp_crc32<32>(g, out_s, n_s, s_crc, s_x2n);
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
            // (3) free slots take the next streams of this kernel's class
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
                            rbias = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u);
                            status = BDF_OK;
                            slen = (uint32_t)len64;
                            uint32_t dlen = 0;
                            at = 0;
                            if (len64 > 0xFFFFFFF0ull) status = BDF_BAD_DATA;    // outside this engine's range
                            else status = inflate_frame_header<FORMAT>(sp, slen, at, dlen);
                            if (status == BDF_OK) { br.init(sp + at, dlen); st = LS_HDR; }
                            else st = LS_END;
                        }
                    }
                }
            }
            if (q_empty && st == LS_NEW) st = LS_IDLE;
            // (4) block headers, one stream at a time, all lanes
            unsigned hdr = __ballot_sync(BDF_FULL_MASK, st == LS_HDR);
            while (hdr) {
                const unsigned s = __ffs(hdr) - 1;
                hdr &= hdr - 1;
                BitReader b = bcast_reader(br, s);
                int hst = BDF_OK;           // status that ends the stream (uniform)
                bool ended = false, fin = false;
                for (;;) {                  // stored blocks are consumed here, one after the other
                    b.refill();
                    if (b.consumed_bits() + 3 > (int64_t)b.len * 8) { hst = BDF_SHORT_INPUT; ended = true; break; }
                    fin = b.take(1) != 0;
                    const unsigned type = b.take(2);
                    if (type == 1) {
                        LaneTables<LTB> &Ts = sm.tab[s];
                        LaneView<LTB> v{Ts.lit_tab, Ts.off_tab, Ts.lit_sorted, Ts.off_sorted, Ts.lit_code, Ts.off_code, sm.bs, sm.lens};
                        load_static_codes<32, LaneView<LTB>, LTB>(g, v);
                        break;
                    }
                    if (type == 2) {
                        LaneTables<LTB> &Ts = sm.tab[s];
                        LaneView<LTB> v{Ts.lit_tab, Ts.off_tab, Ts.lit_sorted, Ts.off_sorted, Ts.lit_code, Ts.off_code, sm.bs, sm.lens};
                        uint32_t nlong;
                        hst = read_dynamic_header<32, LaneView<LTB>, LTB>(g, b, v, nlong);
                        if (hst != BDF_OK) ended = true;
                        break;
                    }
                    if (type == 3) { hst = BDF_BAD_DATA; ended = true; break; }
                    // stored block (src/decompress/mod.rs:282-346): straight from the input to the output in
                    // global memory, after the owner has written out what its ring still holds
                    if (lane == s) flush_local(true);
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
                        const uint32_t v = b.p[sat + i];
                        out_s[pos_s + i] = (uint8_t)v;
                        if (ADLER) { pa += v; pb += (uint64_t)(pos_s + i) * v; }
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
                        pos += blen; flushed = pos; ring_lo = pos;
                        if (ADLER) { sumA = (sumA + pa) % 65521u; sumB = (sumB + pb) % 65521u; }
                    }
                    b.seek(sat + blen);
                    if (fin) { ended = true; break; }
                }
                __syncwarp();               // the tables of slot s were written by all lanes
                if (lane == s) {
                    br = b;
                    final_blk = fin;
                    if (ended) { status = hst; st = LS_END; }
                    else st = LS_RUN;
                }
            }
            if (!__any_sync(BDF_FULL_MASK, st == LS_END || st == LS_HDR || (st == LS_NEW && !q_empty))) break;
        }
        if (!__any_sync(BDF_FULL_MASK, st == LS_RUN)) break;

        // ================================================================ decode
        for (;;) {
            const bool room = pos - flushed <= ROOM;
            if (__any_sync(BDF_FULL_MASK, st == LS_END || st == LS_HDR || (st == LS_RUN && !room))) break;
            bool fresh_far = false;
            // ---- one symbol: up to two literals, or a length / offset pair
            if (st == LS_RUN && copy_rem == 0) {
                bool ok = true;
                if (br.left <= 32) {
                    // more than two zero-fill words loaded: the stream ended inside this block
                    if (br.widx > br.nwords + 2) { status = BDF_SHORT_INPUT; st = LS_END; ok = false; }
                    else br.refill();
                }
                if (ok) {
                    uint32_t e = T.lit_tab[br.peek(LTB)];
                    if (e & LITFLAG) {
                        if (pos >= cap) { status = BDF_INSUFFICIENT_SPACE; st = LS_END; }
                        else {
                            ring_put(pos, e >> E_VAL);
                            pos++;
                            br.drop(e & E_LEN);
                            // literals come in runs: the second look-up needs no refill (>= 24 valid bits)
                            e = T.lit_tab[br.peek(LTB)];
                            if ((e & LITFLAG) && pos < cap) {
                                ring_put(pos, e >> E_VAL);
                                pos++;
                                br.drop(e & E_LEN);
                            }
                        }
                    } else {
                        if ((e & E_LEN) == 0) e = decode_long<CODE_LITLEN, LTB>(br.peek(15), T.lit_sorted, T.lit_code);
                        const uint32_t kind = e & K_MASK;
                        if (e == 0) { status = BDF_BAD_DATA; st = LS_END; }
                        else if (kind == K_LIT) {                  // a literal with a codeword longer than the table
                            if (pos >= cap) { status = BDF_INSUFFICIENT_SPACE; st = LS_END; }
                            else { ring_put(pos, e >> E_VAL); pos++; br.drop(e & E_LEN); }
                        } else if (kind == K_EOB) {
                            br.drop(e & E_LEN);
                            if (br.overrun()) { status = BDF_SHORT_INPUT; st = LS_END; }
                            else if (final_blk) { status = BDF_OK; st = LS_END; }
                            else st = LS_HDR;
                        } else {
                            br.drop(e & E_LEN);
                            const unsigned length = take_length(br, e);
                            br.refill();
                            uint32_t f = T.off_tab[br.peek(OT_BITS)];
                            if ((f & E_LEN) == 0) f = decode_long<CODE_OFFSET, OT_BITS>(br.peek(15), T.off_sorted, T.off_code);
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
                                    if (offset > NEAR || copy_src < ring_lo) { prefetch(copy_src); fresh_far = true; }
                                }
                            }
                        }
                    }
                }
            }
            // ---- up to 8 bytes of the pending match (a far source just requested waits one round)
            if (st == LS_RUN && copy_rem != 0 && !fresh_far) {
                const uint32_t avail = pos - copy_src;          // distance to the source, >= 1
                const bool far = avail > NEAR || copy_src < ring_lo;
                uint32_t n = copy_rem < 8u ? copy_rem : 8u;
                if (avail < n) n = avail;                       // overlapping: one period at a time
                uint32_t lo, hi;
                if (far) {
                    if (copy_src < ring_lo && ring_lo - copy_src < n) n = ring_lo - copy_src;
                    if (!pf_valid) prefetch(copy_src);
                    const uint32_t sh = 8u * (uint32_t)(reinterpret_cast<uintptr_t>(out + copy_src) & 3u);
                    lo = __funnelshift_r(pf0, pf1, sh);
                    hi = __funnelshift_r(pf1, pf2, sh);
                    pf_valid = false;
                } else {
                    const uint32_t i = (copy_src + rbias) & MASK, w = i >> 2, sh = 8u * (i & 3u);
                    const uint32_t a0 = ringw[w], a1 = ringw[(w + 1) & WMASK], a2 = ringw[(w + 2) & WMASK];
                    lo = __funnelshift_r(a0, a1, sh);
                    hi = __funnelshift_r(a1, a2, sh);
                }
                const uint32_t d = (pos + rbias) & MASK;
                if (d + 8 <= (uint32_t)RING) {
                    uint8_t *q = ring + d;
                    q[0] = (uint8_t)lo;
                    if (n > 1) q[1] = (uint8_t)(lo >> 8);
                    if (n > 2) q[2] = (uint8_t)(lo >> 16);
                    if (n > 3) q[3] = (uint8_t)(lo >> 24);
                    if (n > 4) q[4] = (uint8_t)hi;
                    if (n > 5) q[5] = (uint8_t)(hi >> 8);
                    if (n > 6) q[6] = (uint8_t)(hi >> 16);
                    if (n > 7) q[7] = (uint8_t)(hi >> 24);
                } else {
                    const uint64_t v = (uint64_t)hi << 32 | lo;
                    for (uint32_t k = 0; k < n; k++) ring[(d + k) & MASK] = (uint8_t)(v >> (8 * k));
                }
                pos += n;
                copy_rem -= n;
                // a source closer than 8 bytes stays where it is: the distance doubles with every step
                if (avail >= 8u) copy_src += n;
                if (copy_rem != 0 && (avail > NEAR || copy_src < ring_lo)) prefetch(copy_src);
            }
        }
    }
}

}  // namespace bdf
