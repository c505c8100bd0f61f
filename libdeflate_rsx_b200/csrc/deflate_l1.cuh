// deflate_l1.cuh — level 1: single-probe hash table + static Huffman block.
//
// Replaces compress_greedy_block's level-1 branch with HtMatchFinder
// (reference src/compress/mod.rs:1499-1583, src/compress/matchfinder.rs:1139-1231)
// Inputs of at most 65536 bytes are one block without split statistics (:1505-1529); longer
// ones (units of up to 256 KiB, the second kernel instance) run the block-split heuristic
// while staying static-Huffman (:1531-1564).
//
// The parse is inherently serial because only positions where find_match is
// CALLED enter the table (skip_positions is a no-op, :1231).  One warp owns a
// stream and resolves a WINDOW of 32 consecutive positions per round: every lane
// reads its bucket as it was before the round and probes as if every lower lane
// had been inserted (same-hash lower lanes are found with match.any), the warp
// walks the window from match to match on ballot masks, the lanes that were
// really probed commit their table writes, and all literals and matches of the
// window are bit-packed with one warp prefix sum.  The result is exactly the
// serial parse (lane-level model: tools/l1_window_sim.py, tests/test_l1_window.py).
// The round-1 form of the round — 32 positions speculated as "all literals", the
// first lane that really finds a match ends the round — is kept behind
// BDF_L1_WINDOW=0 (determinism test), and the run-of-maximum-length-matches round
// takes over after a 258-byte match.
#pragma once
#include "deflate_common.cuh"

namespace bdf {

constexpr int L1_WARPS = 4;                      // streams per CTA
constexpr int L1_RUN_BATCH = 8;                  // matches of a speculated run verified per memory round trip
constexpr uint32_t L1_SPEC_CAP = 16;             // bytes of a match every lane measures ahead of the window walk

// The hash table of each stream (last position per bucket, all-ones = empty; 16-bit positions
// for the 64 KiB instance, 32-bit for the 256 KiB one) lives in a per-warp global slab behind
// L1 / L2: in shared memory it would cap the SM at three streams, and this parse is a chain of
// dependent probes that only many streams in flight can hide (measured: 8 / 6 / 4 / 2 CTAs per SM
// = 35 / 27 / 22 / 14 GB/s on text, profiles/r4_l1_window_probe.txt).  With a full grid the tables
// (310 MB) and inputs exceed the L2, so the probes run at the memory system's random-sector rate.
template <bool BIG>
struct L1Cfg {
    using pos_t = typename std::conditional<BIG, uint32_t, uint16_t>::type;
    static constexpr uint32_t MAX_LEN = BIG ? 262144u : 65536u;
    static constexpr uint32_t EMPTY = BIG ? 0xFFFFFFFFu : 0xFFFFu;
    static constexpr size_t TABLE_BYTES = 32768 * sizeof(pos_t);
};

struct L1SplitStats {                            // BlockSplitStats, src/compress/mod.rs:271-416
    uint32_t new_obs[14], obs[14];
    uint32_t num_new, num_obs;
};

struct __align__(16) L1Smem {
    uint32_t sink[L1_WARPS][SINK_WORDS];
    uint32_t crc[4][256];
    uint32_t x2n[32];
    L1SplitStats st[L1_WARPS];
};

// static Huffman codes (RFC 1951 3.2.6), returned bit-reversed for the LSB-first stream
__device__ __forceinline__ void static_lit_code(unsigned b, uint32_t &bits, uint32_t &n)
{
    if (b < 144) { n = 8; bits = __brev(0x30u + b) >> 24; }
    else { n = 9; bits = __brev(0x190u + (b - 144)) >> 23; }
}
__device__ __forceinline__ void static_len_code(unsigned len, uint32_t &bits, uint32_t &n)
{
    unsigned slot = length_slot_of(len), base, extra;
    length_slot_info(slot, base, extra);
    uint32_t code, cl;
    if (slot < 23) { cl = 7; code = __brev(slot + 1) >> 25; }
    else { cl = 8; code = __brev(0xC0u + (slot - 23)) >> 24; }
    bits = code | ((len - base) << cl);
    n = cl + extra;
}
__device__ __forceinline__ void static_off_code(unsigned off, uint32_t &bits, uint32_t &n)
{
    unsigned slot = offset_slot_of(off), base, extra;
    offset_slot_info(slot, base, extra);
    bits = (__brev(slot) >> 27) | ((off - base) << 5);
    n = 5 + extra;
}

#ifndef BDF_L1_MIN_CTAS
#define BDF_L1_MIN_CTAS 8                        // resident CTAs per SM the register allocation aims at (64 registers)
#endif
template <bool BIG, bool SIZE = false>
__global__ void __launch_bounds__(L1_WARPS * 32, BDF_L1_MIN_CTAS) deflate_l1_kernel(DeflateArgs a)
{
    using CFG = L1Cfg<BIG>;
    using pos_t = typename CFG::pos_t;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    L1Smem &sm = *reinterpret_cast<L1Smem *>(smem_raw);
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    pos_t *table = reinterpret_cast<pos_t *>(static_cast<uint8_t *>(a.scratch) +
                                             a.scratch_stride * (blockIdx.x * L1_WARPS + warp));
    L1SplitStats &st = sm.st[warp];
    if (a.format == BDF_GZIP) load_crc_tables_to_smem(sm.crc, sm.x2n);
    for (;;) {
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(a.work_counter, 1ull);
        idx = __shfl_sync(BDF_FULL_MASK, idx, 0);
        if (idx >= a.n) break;
        const uint8_t *in = a.in + a.in_off[idx];
        const uint64_t len64 = a.in_off[idx + 1] - a.in_off[idx];
        uint8_t *out = SIZE ? nullptr : a.out + a.out_off[idx];
        if (SIZE && len64 == 0) {             // the estimator's block loop never runs (:808)
            if (lane == 0) { a.status[idx] = BDF_OK; a.out_size[idx] = 0; }
            continue;
        }
        if (len64 > CFG::MAX_LEN) {
            if (lane == 0) { a.status[idx] = BDF_STREAM_UNSUPPORTED; a.out_size[idx] = 0; }
            continue;
        }
        const uint32_t len = (uint32_t)len64;
        const unsigned uflags = unit_flags_of(a, idx);
        // :1505: at most 64 KiB is one block, no statistics; the size estimator always keeps them
        // (accumulate_greedy_frequencies, :1096-1140)
        const bool split = BIG && (len > 65536 || SIZE);
        for (unsigned i = lane; i < CFG::TABLE_BYTES / 16; i += 32) reinterpret_cast<uint4 *>(table)[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
        const unsigned hdr = SIZE ? 0 : frame_header(a.format, 1, out, lane);
        BitSinkT<SIZE> bs;
        bs.init(sm.sink[warp], SIZE ? nullptr : out + hdr, SIZE ? ~0ull : unit_cap(len, uflags), lane);
        // BTYPE = 01; BFINAL of a block that may still be split is written as 0 and set at the end
        uint64_t bfinal_at = bs.bitpos();
        bs.put1((!split && (uflags & UNIT_FINISH) ? 1u : 0u) | 2u, 3, lane);
        uint32_t block_start = 0;
        if (split) {
            if (lane < 14) { st.new_obs[lane] = 0; st.obs[lane] = 0; }
            if (lane == 0) { st.num_new = 0; st.num_obs = 0; }
            __syncwarp();
        }

        uint32_t pos = 0;
        bool run_mode = false;          // the last match had the maximum length: try a run of them
        while (pos < len) {
            if (run_mode) {
                // Run-length / periodic data is one maximum-length match after the other, and a round
                // of this (serial) parse is a chain of dependent memory round trips.  After a 258-byte
                // match the 32 lanes therefore speculate on a RUN of such matches: lane k probes
                // position pos + 258 k as if every lower lane had been inserted (same rule as the
                // literal speculation below), measures its own match, and the lanes up to the first
                // one that is not a full-length match are committed — exactly the serial result,
                // up to 32 matches per round.
                run_mode = false;
                const uint32_t q = pos + 258u * lane;
                const bool act = q + 3 <= len;
                uint32_t v = 0, h = 0x10000u + lane;
                if (act) { v = ld24(in + q); h = hash3(v); }
                const unsigned peers = __match_any_sync(BDF_FULL_MASK, h);
                const unsigned lower = peers & lanemask_lt();
                uint32_t cand = CFG::EMPTY;
                if (act) cand = lower ? pos + 258u * (31 - __clz(lower)) : (uint32_t)table[h];
                const bool found = act && cand != CFG::EMPTY && q - cand <= 32768u && ld24(in + cand) == v;
                // Which matches have the full 258 bytes?  The last two bytes are checked by each lane
                // for its own match; the first 256 are compared by the WHOLE warp, match after match
                // (one coalesced 256-byte access per side, L1_RUN_BATCH matches in flight) — 32 lanes each
                // walking their own 258 bytes touched 32 sectors per load and left the kernel waiting
                // on L2 / DRAM (long-scoreboard stall 51 per issue).
                bool maybe = found && len - q >= 258;
                if (maybe) maybe = in[cand + 256] == in[q + 256] && in[cand + 257] == in[q + 257];
                const unsigned cand_ok = __ballot_sync(BDF_FULL_MASK, maybe);
                unsigned full = 0;
                for (unsigned m0 = 0; m0 < 32; m0 += L1_RUN_BATCH) {
                    uint64_t x[L1_RUN_BATCH];
#pragma unroll
                    for (int t = 0; t < L1_RUN_BATCH; t++) {
                        const unsigned m = m0 + t;
                        const uint32_t cm = __shfl_sync(BDF_FULL_MASK, cand, m), qm = __shfl_sync(BDF_FULL_MASK, q, m);
                        x[t] = 1;
                        if ((cand_ok >> m) & 1u) x[t] = ld64_any(in + cm + 8u * lane) ^ ld64_any(in + qm + 8u * lane);
                    }
                    bool all4 = true;
#pragma unroll
                    for (int t = 0; t < L1_RUN_BATCH; t++) {
                        const bool eq = __ballot_sync(BDF_FULL_MASK, x[t] == 0) == BDF_FULL_MASK;
                        if (eq) full |= 1u << (m0 + t);
                        all4 = all4 && eq;
                    }
                    if (!all4) break;                                          // the run ends in this group
                }
                const unsigned j = ~full ? __ffs(~full) - 1 : 32;                 // first lane without a full-length match
                unsigned mylen = lane < j ? 258u : 0u;
                if (lane == j && found) mylen = prefix_len_bytes(in + cand, in + q, len - q < 258 ? len - q : 258);
                const unsigned jfound = __shfl_sync(BDF_FULL_MASK, (unsigned)found, j & 31);
                const unsigned nm = j < 32 ? j + (jfound ? 1u : 0u) : 32u;        // matches taken this round
                if (nm) {
                    // bucket writes of the probes that really happen (highest lane of a hash wins)
                    const unsigned committing = __ballot_sync(BDF_FULL_MASK, act && lane < nm);
                    const unsigned mine = peers & committing;
                    if (act && lane < nm && (mine >> lane) == 1u) table[h] = (pos_t)q;
                    uint32_t bits = 0, nb = 0;
                    if (lane < nm) {
                        uint32_t b0, n0, b1, n1;
                        static_len_code(mylen, b0, n0);
                        static_off_code(q - cand, b1, n1);
                        bits = b0 | b1 << n0;                                   // <= 13 + 18 bits
                        nb = n0 + n1;
                    }
                    bs.put(bits, nb, lane);
                    const unsigned last_len = __shfl_sync(BDF_FULL_MASK, mylen, nm - 1);
                    pos += 258u * (nm - 1) + last_len;
                    run_mode = last_len == 258;
                    __syncwarp();
                    continue;
                }
                __syncwarp();
                // the position is not a match at all: the literal speculation takes it from here
            }
            if ((a.l1_window & 1) && (!split || (a.l1_window & 8))) {
                // WHOLE-WINDOW ROUND (tools/l1_window_sim.py states it with plain lists and checks it
                // against the serial parse).  The 32 lanes read their buckets once, as they were
                // before the round; bucket writes of the round are stood in for by same-hash lower
                // lanes (match.any).  The warp then walks the window from match to match without
                // going back to memory: `inserted` collects the lanes that were really probed
                // (literals and match starts), lanes from `cur` on count as probed for the lanes
                // above them, lanes inside a match never enter the table (skip_positions is a
                // no-op, :1231).  Every lane has measured the match it would have if all lower lanes
                // were probed (the usual case) up to L1_SPEC_CAP bytes beforehand, so a step of the
                // walk touches memory only for long matches or for a candidate the speculation
                // did not foresee.  Up to 32 + 257 bytes per round instead of one match.
                const uint32_t p = pos + lane;
                const bool hashable = p + 3 <= len;
                uint32_t v = 0, h = 0x10000u + lane;
                if (hashable) { v = ld24(in + p); h = hash3(v); }
                const unsigned peers = __match_any_sync(BDF_FULL_MASK, h);
                const unsigned lower = peers & lanemask_lt();
                const unsigned hmask = __ballot_sync(BDF_FULL_MASK, hashable);
                uint32_t tc = CFG::EMPTY;
                if (hashable) tc = table[h];
                // The buckets of the NEXT window are prefetched while this one is worked on (bits 1 / 2 of
                // BDF_L1_WINDOW: into L1, the default, or L2; +1-4 %): its bytes are requested here and have
                // arrived once the bucket above has
                uint32_t vn = 0;
                const bool pn = (a.l1_window & 6) && p + 35 <= len;
                if (pn) vn = ld24(in + p + 32);
                const bool tfound = hashable && tc != CFG::EMPTY && p - tc <= 32768u && ld24(in + tc) == v;
                if (pn) {
                    const pos_t *nb = table + hash3(vn);
                    if (a.l1_window & 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(nb));
                    else asm volatile("prefetch.global.L2 [%0];" ::"l"(nb));
                }
                const unsigned jl = lower ? 31 - __clz(lower) : lane;
                const uint32_t vjl = __shfl_sync(BDF_FULL_MASK, v, jl);
                uint32_t scand = 0xFFFFFFFFu, slen = 0;
                if (lower) { if (vjl == v) scand = pos + jl; }
                else if (tfound) scand = tc;
                if (scand != 0xFFFFFFFFu) {
                    const uint32_t room = len - p < 258u ? len - p : 258u;
                    slen = prefix_len_bytes(in + scand, in + p, room < L1_SPEC_CAP ? room : L1_SPEC_CAP);
                }
                unsigned inserted = 0, lit = 0, mm = 0, cur = 0;
                uint32_t mylen = slen, mycand = scand, nxt = 0, lastlen = 0;
                // lanes whose candidate depends on what the round itself inserts, and the hits that do not
                const unsigned depmask = __ballot_sync(BDF_FULL_MASK, lower != 0);
                const unsigned smask = __ballot_sync(BDF_FULL_MASK, tfound && lower == 0);
                for (;;) {
                    const unsigned from_cur = 0xFFFFFFFFu << cur;             // cur < 32 here
                    const unsigned sh = smask & from_cur;
                    const unsigned ks = sh ? __ffs(sh) - 1 : 32u;
                    const unsigned below_ks = sh ? (sh & (0u - sh)) - 1u : 0xFFFFFFFFu;
                    unsigned k;
                    uint32_t ck = 0;
                    bool spec = true;                                         // the match is the one its lane measured
                    if ((depmask & from_cur & below_ks) == 0) {
                        // STATIC step: no lane between cur and the first plain hit looks at the round's own
                        // insertions, so that hit is the next match, with its bucket candidate: bit operations only
                        if (ks == 32) { lit |= from_cur; inserted |= from_cur; nxt = pos + 32; lastlen = 0; break; }
                        k = ks;
                    } else {
                        const unsigned eff = lower & (inserted | from_cur);
                        const unsigned j = eff ? 31 - __clz(eff) : lane;
                        const uint32_t vj = __shfl_sync(BDF_FULL_MASK, v, j);
                        const bool ok = eff ? vj == v : tfound;
                        const uint32_t c = eff ? pos + j : tc;
                        const unsigned fb = __ballot_sync(BDF_FULL_MASK, ok && hashable && lane >= cur);
                        if (!fb) { lit |= from_cur; inserted |= from_cur; nxt = pos + 32; lastlen = 0; break; }
                        k = __ffs(fb) - 1;
                        ck = __shfl_sync(BDF_FULL_MASK, c, k);
                        spec = ck == __shfl_sync(BDF_FULL_MASK, scand, k);
                        if (lane == k) mycand = ck;
                    }
                    const unsigned seg = from_cur & ((2u << k) - 1u);         // lanes cur..k are probed
                    lit |= seg & ~(1u << k);
                    inserted |= seg;
                    mm |= 1u << k;
                    const uint32_t slk = __shfl_sync(BDF_FULL_MASK, slen, k);
                    const uint32_t mp = pos + k;
                    const uint32_t rk = len - mp < 258u ? len - mp : 258u;
                    uint32_t mlen = slk;
                    if (!spec || (slk >= L1_SPEC_CAP && slk < rk)) {
                        if (spec) {
                            ck = __shfl_sync(BDF_FULL_MASK, scand, k);
                            mlen = L1_SPEC_CAP + warp_match_len(in + ck + L1_SPEC_CAP, in + mp + L1_SPEC_CAP, rk - L1_SPEC_CAP, lane);
                        } else {
                            mlen = warp_match_len(in + ck, in + mp, rk, lane);
                        }
                        if (lane == k) mylen = mlen;
                    }
                    BDF_ASSERT(mlen >= 3 && mlen <= rk && mp + mlen <= len);
                    BDF_ASSERT(lane != k || (mycand < mp && mp - mycand <= 32768u && mylen == mlen));
                    cur = k + (mlen ? mlen : 1u);                            // mlen >= 3 (the 3 hashed bytes are equal); never stall
                    lastlen = mlen;
                    if (cur >= 32) { nxt = pos + cur; break; }
                }
                // bucket writes of the probed lanes (the highest lane of a hash wins)
                inserted &= hmask;
                const unsigned mine = peers & inserted;
                BDF_ASSERT(!hashable || h < 32768u);
                BDF_ASSERT(nxt > pos && (mm & lit) == 0);
                if (((inserted >> lane) & 1u) && (mine >> lane) == 1u) table[h] = (pos_t)p;
                const bool is_match = (mm >> lane) & 1u;
                const bool is_lit = !is_match && ((lit >> lane) & 1u) && p < len;
                const uint32_t byte = hashable ? (v & 0xFFu) : (is_lit ? in[p] : 0u);
                // Block splitting (units above 64 KiB, size estimation): the statistics see the symbols of
                // the window in stream order = lane order.  should_end_block can only act once 2048
                // observations are pending, so a round that cannot reach that is counted in bulk; the others
                // are replayed symbol by symbol by lane 0 (same rule as the one-match rounds below).
                uint32_t cut = 0xFFFFFFFFu;                 // lane in front of whose symbol the block ends
                if (split) {
                    const unsigned litmask = __ballot_sync(BDF_FULL_MASK, is_lit);
                    const uint32_t nobs = __popc(litmask) + 2u * __popc(mm);
                    const unsigned oslot = is_match ? offset_slot_of(p - mycand) : 0u;
                    const unsigned ocls = oslot < 16 ? 0u : oslot < 24 ? 1u : oslot < 30 ? 2u : 0u;
                    const uint32_t pending = st.num_new;
                    __syncwarp();
                    if (pending + nobs < 2048) {
                        if (is_lit) atomicAdd(&st.new_obs[byte >> 5], 1u);
                        if (is_match) { atomicAdd(&st.new_obs[8 + (mylen >= 8)], 1u); atomicAdd(&st.new_obs[10 + ocls], 1u); }
                        __syncwarp();
                        if (lane == 0) st.num_new = pending + nobs;
                    } else {
                        const uint32_t info = is_match ? ((mylen >= 8 ? 1u : 0u) | ocls << 1) : byte;
                        for (unsigned rem = litmask | mm; rem; rem &= rem - 1u) {
                            const unsigned l = __ffs(rem) - 1;
                            const uint32_t x = __shfl_sync(BDF_FULL_MASK, info, l);
                            if (lane == 0) {
                                if (hc_should_end(st, pos + l - block_start, len - (pos + l))) {
                                    cut = l;                        // at most once per round: 2048 more are needed
                                    block_start = pos + l;
                                    for (int c = 0; c < 14; c++) { st.new_obs[c] = 0; st.obs[c] = 0; }
                                    st.num_new = 0; st.num_obs = 0;
                                }
                                if ((mm >> l) & 1u) { st.new_obs[8 + (x & 1u)]++; st.new_obs[10 + (x >> 1)]++; st.num_new += 2; }
                                else { st.new_obs[x >> 5]++; st.num_new++; }
                            }
                        }
                        cut = __shfl_sync(BDF_FULL_MASK, cut, 0);
                        block_start = __shfl_sync(BDF_FULL_MASK, block_start, 0);
                    }
                    __syncwarp();
                }
                uint32_t bits = 0, nb = 0;
                if (is_match) {
                    uint32_t b0, n0, b1, n1;
                    static_len_code(mylen, b0, n0);
                    static_off_code(p - mycand, b1, n1);
                    bits = b0 | b1 << n0;                                       // <= 13 + 18 bits
                    nb = n0 + n1;
                } else if (is_lit) {
                    static_lit_code(byte, bits, nb);
                }
                if (cut == 0xFFFFFFFFu) {
                    bs.put(bits, nb, lane);
                } else {
                    bs.put(lane < cut ? bits : 0, lane < cut ? nb : 0, lane);
                    bs.put1(0, 7, lane);                     // end of block (symbol 256 = 0000000)
                    bfinal_at = bs.bitpos();
                    bs.put1(2u, 3, lane);                    // BFINAL = 0 for now, BTYPE = 01
                    bs.put(lane >= cut ? bits : 0, lane >= cut ? nb : 0, lane);
                }
                pos = nxt;
                run_mode = !split && lastlen == 258;
                __syncwarp();
                continue;
            }
            const uint32_t p = pos + lane;
            const bool hashable = p + 3 <= len;          // the last two positions are never hashed (:1140)
            uint32_t v = 0, h = 0x10000u + lane;         // unique key for lanes that do not hash
            if (hashable) { v = ld24(in + p); h = hash3(v); }
            const unsigned peers = __match_any_sync(BDF_FULL_MASK, h);
            const unsigned lower = peers & lanemask_lt();
            uint32_t cand = CFG::EMPTY;
            if (hashable) cand = lower ? pos + (31 - __clz(lower)) : table[h];
            const bool found = hashable && cand != CFG::EMPTY && p - cand <= 32768u && ld24(in + cand) == v;
            const unsigned fb = __ballot_sync(BDF_FULL_MASK, found);
            const unsigned k = fb ? __ffs(fb) - 1 : 32;            // first lane whose probe hits
            // commit bucket writes of lanes 0..k (the match start is inserted too, :1162-1163)
            const unsigned committing = __ballot_sync(BDF_FULL_MASK, hashable && lane <= k);
            const unsigned mine = peers & committing;
            if (hashable && lane <= k && (mine >> lane) == 1u) table[h] = (pos_t)p;
            // this round's symbols: nl literals, then (k < 32) one match
            const uint32_t nl = len - pos < k ? len - pos : k;
            uint32_t mp = 0, mc = 0, mlen = 0;
            if (k < 32) {
                mp = pos + k;
                mc = __shfl_sync(BDF_FULL_MASK, cand, k);
                const unsigned room = len - mp < 258 ? len - mp : 258;
                mlen = warp_match_len(in + mc, in + mp, room, lane);
            }
            // Block splitting does not touch the parse (no lazy evaluation, the table carries
            // over), it only decides where end-of-block + a new header go into the bit stream.
            // should_end_block can only act once 2048 observations are pending, so rounds that
            // cannot reach that are counted in bulk; the others are replayed symbol by symbol.
            uint32_t cut = 0xFFFFFFFFu;                 // literal index in front of which the block ends
            if (split) {
                const uint32_t pending = st.num_new;
                __syncwarp();
                if (pending + nl < 2048) {
                    if (lane < nl) atomicAdd(&st.new_obs[in[p] >> 5], 1u);
                    __syncwarp();
                    if (lane == 0) {
                        st.num_new = pending + nl;
                        if (k < 32) {
                            const unsigned slot = offset_slot_of(mp - mc);
                            st.new_obs[8 + (mlen >= 8)]++;
                            st.new_obs[10 + (slot < 16 ? 0 : slot < 24 ? 1 : slot < 30 ? 2 : 0)]++;
                            st.num_new += 2;
                        }
                    }
                } else {
                    if (lane == 0) {
                        for (uint32_t j = 0; j <= nl; j++) {
                            if (j == nl && k == 32) break;      // no symbol follows in this round
                            if (hc_should_end(st, pos + j - block_start, len - (pos + j))) {
                                cut = j;                        // at most once per round: 2048 more are needed
                                block_start = pos + j;
                                for (int c = 0; c < 14; c++) { st.new_obs[c] = 0; st.obs[c] = 0; }
                                st.num_new = 0; st.num_obs = 0;
                            }
                            if (j < nl) { st.new_obs[in[pos + j] >> 5]++; st.num_new++; }
                        }
                        if (k < 32) {
                            const unsigned slot = offset_slot_of(mp - mc);
                            st.new_obs[8 + (mlen >= 8)]++;
                            st.new_obs[10 + (slot < 16 ? 0 : slot < 24 ? 1 : slot < 30 ? 2 : 0)]++;
                            st.num_new += 2;
                        }
                    }
                    cut = __shfl_sync(BDF_FULL_MASK, cut, 0);
                    block_start = __shfl_sync(BDF_FULL_MASK, block_start, 0);
                }
                __syncwarp();
            }
            // literals: lanes below nl
            uint32_t bits = 0, nb = 0;
            if (lane < nl) static_lit_code(in[p], bits, nb);
            if (cut == 0xFFFFFFFFu) {
                bs.put(bits, nb, lane);
            } else {
                bs.put(lane < cut ? bits : 0, lane < cut ? nb : 0, lane);
                bs.put1(0, 7, lane);                     // end of block (symbol 256 = 0000000)
                bfinal_at = bs.bitpos();
                bs.put1(2u, 3, lane);                    // BFINAL = 0 for now, BTYPE = 01
                bs.put(lane >= cut ? bits : 0, lane >= cut ? nb : 0, lane);
            }
            if (k == 32) { pos += 32; continue; }
            uint32_t b0, n0, b1, n1;
            static_len_code(mlen, b0, n0);
            static_off_code(mp - mc, b1, n1);
            bs.put(lane == 0 ? b0 : lane == 1 ? b1 : 0, lane == 0 ? n0 : lane == 1 ? n1 : 0, lane);
            pos = mp + mlen;
            run_mode = !split && mlen == 258;
            __syncwarp();
        }
        bs.put1(0, 7, lane);                     // end of block (symbol 256 = 0000000)
        if (split && (uflags & UNIT_FINISH)) bs.set_bit(bfinal_at, lane);
        if (uflags & UNIT_SYNC) bs.sync_marker(lane);
        uint64_t sz = bs.finish(lane);
        int st_code = BDF_OK;
        if (sz == ~0ull) { st_code = BDF_INSUFFICIENT_SPACE; sz = 0; }
        else if (!SIZE) sz = frame_footer(a.format, in, len, out, hdr + sz, sm.crc, sm.x2n, lane);
        if (lane == 0) { a.status[idx] = st_code; a.out_size[idx] = sz; }
        __syncwarp();
    }
}

}  // namespace bdf
