// deflate_hcs.cuh — levels 2..9 for streams of at most 64 KiB: hash chains, greedy / lazy /
// lazy2 parse and dynamic Huffman blocks with EVERYTHING a stream needs in shared memory.
// One 1024-thread CTA per stream, one CTA per SM.
//
// Replaces, for the batch path, the same reference code as deflate_hc.cuh (which stays as the
// instance for units up to 256 KiB):
//   MatchFinder::{find_match_impl, skip_match, skip_positions}  src/compress/matchfinder.rs:754-1106
//   decide_greedy_sequences + BlockSplitStats                    src/compress/mod.rs:1261-1373, 271-416
//   make_huffman_code, write_dynamic_huffman_header_impl, write_sequences
//
// Why a second kernel: ncu on the first one (profiles/r1_deflate_hc_summary.md) showed 41x DRAM
// amplification (530 MB of chain + symbol slabs in global memory against a 126 MB L2), three of
// four warps waiting while warp 0 parsed, and byte loads of the input from global memory.  Here
//   * phase 1 builds the chains of the whole stream: head[32768] (u16) and one back-link per
//     position, link[65536] (u16), both in shared memory; the head table is dead afterwards and
//   * phase 2 stages the input over it (64 KiB), so every probe of a chain walk — link, quick
//     reject byte, 4-byte compare, match extension — is a shared-memory access;
//   * find_match is parse-independent (every position is inserted exactly once, in ascending order,
//     whatever the parse does — SURVEY §7), so the stream is searched window by window (2040
//     positions) by all 32 warps, each lane running its own chain walk and taking the next
//     position of its warp's range as soon as it is done (no lane waits for the longest chain);
//   * the parse is parallel too: step(p) — where the greedy / lazy rule goes from p — is a function
//     of the search results at p, p+1, p+2.  Every 17-position segment is swept backwards by one
//     thread (exit of the segment from EVERY entry point), segments are composed in groups of
//     eight by warps, one thread walks the 15 groups, and the entries found that way are walked
//     down again in parallel; each segment then emits its own symbol records.  The only serial
//     walks are 15 + 8 + 17 steps long;
//   * BlockSplitStats only acts where 2048 observations are pending: that step is located from the
//     per-segment observation counts, symbols in front of it count for the current block, symbols
//     behind it for the next (two histogram sets), and should_end_block itself runs unchanged;
//   * a block is packed by the whole CTA: 1024 symbol records per round, bit offsets from a block
//     prefix sum, OR-ed into a shared staging buffer, bytes flushed coalesced.
// Byte-identical to deflate_hc.cuh / the oracle (tests/test_gpu_checksum_compress.py, test_gpu_fuzz.py).
#pragma once
#include "deflate_hc.cuh"

namespace bdf {

constexpr int HCS_THREADS = 1024;
constexpr int HCS_WARPS = HCS_THREADS / 32;
constexpr uint32_t HCS_SEG = 17;                         // odd: the per-segment threads hit different banks
constexpr uint32_t HCS_NSEG = 120;
constexpr uint32_t HCS_W = HCS_SEG * HCS_NSEG;           // 2040 positions parsed per window: < 2048 observations,
                                                         // so at most one block-split check falls into a window
constexpr uint32_t HCS_GSEG = 8;                         // segments per group
constexpr uint32_t HCS_NGRP = HCS_NSEG / HCS_GSEG;       // 15
constexpr uint32_t HCS_GLEN = HCS_SEG * HCS_GSEG;        // 136
constexpr uint32_t HCS_SEARCH = HCS_W + 8;               // positions searched per window (the lazy rule looks 2 ahead)
constexpr int HCS_BURST = 8;                             // chain candidates a lane walks between two refills of its warp
constexpr uint32_t HCS_IBLK = 2048;                      // positions hashed per insertion round
constexpr uint32_t HCS_INS_WARPS = 16;                   // warps that link positions (warp w: hash & 15 == w)
constexpr uint32_t HCS_LIST = 288;                       // pending positions per inserting warp
constexpr uint32_t HCS_STAGE_WORDS = 1024 * 2 + 64;      // bit staging of one emission round (<= 47 bits per record)
constexpr uint32_t HCS_REC_MATCH = 0x80000000u;          // record: literal byte, or MATCH | len << 16 | (offset - 1)
constexpr size_t HCS_SCRATCH_PER_CTA = (65536 + 64) * sizeof(uint32_t);
constexpr uint32_t HCS_NONE = 0xFFFFu;

struct HcsWindow {                 // live while a window is searched and parsed
    uint32_t res[HCS_SEARCH];      // find_match result per position: len | offset << 16
    uint16_t nxt[HCS_W + 8];       // step from a position: delta | lazy literals << 9 | match << 11
    uint32_t sw[HCS_W];            // segment sweep: exit - segment end | steps << 9 (to the exit, from this position)
    uint16_t ex2[HCS_W + 8];       // exit from the position's group (window-relative); afterwards the step list
};
struct HcsInsert {                 // live while the chains are built
    uint16_t hbuf[HCS_IBLK];       // hash of every position of the round
    uint32_t cls[HCS_INS_WARPS][HCS_IBLK / 32];      // bit k of row w: position k of the round belongs to warp w
    uint16_t list[HCS_INS_WARPS][HCS_LIST];
};
struct HcsEncode {                 // live while a block is encoded
    uint32_t litlen_code[288], offset_code[32];
    uint32_t scratch[288];
    uint8_t hdr_lens[320];
    uint16_t hdr_items[320];
    uint32_t pre_freq[19], pre_code[19];
    uint8_t pre_len[19];
    uint32_t stage[HCS_STAGE_WORDS];
    uint32_t warp_tot[HCS_WARPS];
};

constexpr uint32_t HCS_DP_CHUNK = 504;                   // positions whose match lists are staged at a time (multiple of 3)
struct HcsDp {                     // live while the near-optimal parser runs its cost pass over a block
    uint32_t ring[512];            // cost to the end of the block for the 512 positions behind the current one
    uint32_t lst[HCS_DP_CHUNK * 8];    // match lists of the chunk (8 entries per position; 16-byte aligned)
    uint32_t ch[HCS_DP_CHUNK];         // choices of the chunk
    uint8_t lit_cost[256], len_cost[260], slot_cost[32];
};
static_assert(offsetof(HcsDp, lst) % 16 == 0, "the lists are staged with 16-byte copies");

struct __align__(16) HcsSmem {
    union {
        uint8_t in[65536 + 32];    // phase 2: the stream (zero padded)
        uint16_t head[32768];      // phase 1: bucket -> most recent position, 0xFFFF = empty
    };
    uint16_t link[65536];          // position -> distance to the previous position with the same hash, 0 = none
    union {
        HcsWindow w;
        HcsInsert ins;
        HcsEncode enc;
        HcsDp dp;
    };
    // block state (names shared with HcSmem: hc_should_end / hc_prepare_header run unchanged)
    uint32_t litlen_freq[288], offset_freq[32];          // current block, symbols of finished windows
    uint32_t freq_a[320], freq_b[320];                   // this window: in front of / behind the split check
    uint8_t litlen_len[288], offset_len[32];
    uint32_t new_obs[14], obs[14], obs_a[14], obs_b[14];
    uint32_t num_new, num_obs, cnt_a, cnt_b;
    uint32_t seg_rec[HCS_NSEG + 8], seg_obs[HCS_NSEG + 8];   // exclusive scans over the entered segments
    uint16_t seg_entry[HCS_NSEG + 8], grp_entry[HCS_NGRP + 1];
    uint32_t scan_tmp[8];
    // control words (written by one thread, read by all after a barrier)
    uint32_t c_next_entry, c_pc, c_rec_at_pc, c_win_rec, c_win_obs, c_split, c_search_next;
    // bit sink of the CTA
    uint32_t sink_bits;            // bits pending in enc.stage (after a round: < 8)
    uint32_t sink_carry;           // those bits (the staging buffer itself is overlaid between blocks)
    uint32_t sink_overflow;
    unsigned long long sink_flushed;   // bytes stored so far
};
static_assert(sizeof(HcsSmem) <= 232448, "one CTA per SM: everything has to fit the 227 KB a CTA may use");

// ---- CTA bit sink.  All threads call put(): thread t contributes up to 64 bits (nb may be 0) in
// thread order.  LSB-first like Bitstream (src/compress/bitstream.rs); capacity checked per round
// the way the reference's checked writes fail (the stream fails once a completed byte does not fit).
template <bool COUNT>
struct CtaSink {
    uint8_t *out;
    unsigned long long cap;

    __device__ __forceinline__ void init(HcsSmem &sm, uint8_t *dst, unsigned long long capacity)
    {
        out = dst; cap = capacity;
        if (threadIdx.x == 0) { sm.sink_bits = 0; sm.sink_carry = 0; sm.sink_overflow = 0; sm.sink_flushed = 0; }
    }
    // zero the staging buffer (the encode region was used by something else before)
    __device__ __forceinline__ void begin_block(HcsSmem &sm)
    {
        for (uint32_t i = threadIdx.x; i < HCS_STAGE_WORDS; i += HCS_THREADS) sm.enc.stage[i] = i ? 0u : sm.sink_carry;
        __syncthreads();
    }
    __device__ __forceinline__ void put(HcsSmem &sm, unsigned long long bits, uint32_t nb)
    {
        const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
        uint32_t incl = nb;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(BDF_FULL_MASK, incl, d);
            if (lane >= (unsigned)d) incl += t;
        }
        if (lane == 31) sm.enc.warp_tot[warp] = incl;
        __syncthreads();
        uint32_t base = sm.sink_bits, total = 0;
        {
            // every thread sums the warp totals in front of its warp (32 broadcast reads)
            uint32_t v = lane < HCS_WARPS ? sm.enc.warp_tot[lane] : 0u;
            uint32_t pre = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(BDF_FULL_MASK, pre, d);
                if (lane >= (unsigned)d) pre += t;
            }
            total = __shfl_sync(BDF_FULL_MASK, pre, 31);
            base += __shfl_sync(BDF_FULL_MASK, pre - v, warp);
        }
        if (nb && !COUNT) {
            const uint32_t at = base + incl - nb;
            const uint32_t w = at >> 5, sh = at & 31u;
            BDF_ASSERT(nb <= 64 && w + 2 < HCS_STAGE_WORDS);
            const uint32_t lo = (uint32_t)bits, hi = (uint32_t)(bits >> 32);
            atomicOr(&sm.enc.stage[w], lo << sh);
            const unsigned long long up = sh ? ((unsigned long long)hi << 32 | lo) >> (32 - sh) : hi;
            if (sh + nb > 32) atomicOr(&sm.enc.stage[w + 1], (uint32_t)up);
            if (sh + nb > 64) atomicOr(&sm.enc.stage[w + 2], (uint32_t)(up >> 32));
        }
        __syncthreads();
        // flush the complete bytes, carry the rest
        const uint32_t nbits = sm.sink_bits + total;
        const uint32_t nbytes = nbits >> 3;
        const unsigned long long flushed = sm.sink_flushed;
        const bool over = flushed + nbytes > cap || sm.sink_overflow;
        if (!COUNT && !over)
            for (uint32_t b = threadIdx.x; b < nbytes; b += HCS_THREADS)
                out[flushed + b] = (uint8_t)(sm.enc.stage[b >> 2] >> (8 * (b & 3)));
        const uint32_t carry = COUNT ? 0u : (sm.enc.stage[nbytes >> 2] >> (8 * (nbytes & 3))) & 0xFFu;
        __syncthreads();
        if (!COUNT)
            for (uint32_t i = threadIdx.x; i <= (nbits >> 5) + 1 && i < HCS_STAGE_WORDS; i += HCS_THREADS) sm.enc.stage[i] = 0;
        __syncthreads();
        if (threadIdx.x == 0) {
            if (!COUNT) sm.enc.stage[0] = carry;
            sm.sink_carry = carry;
            sm.sink_bits = nbits & 7u;
            sm.sink_flushed = flushed + nbytes;
            if (over) sm.sink_overflow = 1;
        }
        __syncthreads();
    }
};

// 4 bytes of the staged stream at any position (two aligned words)
__device__ __forceinline__ uint32_t hcs_ld32(const uint8_t *in, uint32_t p)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(in) + (p >> 2);
    return __funnelshift_r(w[0], w[1], 8u * (p & 3u));
}
__device__ __forceinline__ unsigned long long hcs_ld64(const uint8_t *in, uint32_t p)
{
    const uint32_t *w = reinterpret_cast<const uint32_t *>(in) + (p >> 2);
    const uint32_t sh = 8u * (p & 3u), w0 = w[0], w1 = w[1], w2 = w[2];
    return (unsigned long long)__funnelshift_r(w1, w2, sh) << 32 | __funnelshift_r(w0, w1, sh);
}
// common prefix of in[a..] and in[b..], at most maxlen bytes (match_len_*, src/compress/matchfinder.rs:245-694)
__device__ __forceinline__ uint32_t hcs_prefix(const uint8_t *in, uint32_t a, uint32_t b, uint32_t maxlen)
{
    uint32_t n = 0;
    while (n + 8 <= maxlen) {
        const unsigned long long x = hcs_ld64(in, a + n) ^ hcs_ld64(in, b + n);
        if (x) return n + ((__ffsll((long long)x) - 1) >> 3);
        n += 8;
    }
    if (n < maxlen) {
        unsigned long long x = hcs_ld64(in, a + n) ^ hcs_ld64(in, b + n);
        x |= 1ull << (8 * (maxlen - n));                 // stop at maxlen (maxlen - n < 8)
        return n + ((__ffsll((long long)x) - 1) >> 3);
    }
    return n;
}

// what hc_prepare_header works on (member names of HcSmem)
struct HcsHeaderView {
    uint8_t *litlen_len, *offset_len, *hdr_lens;
    uint16_t *hdr_items;
    uint32_t *pre_freq, *pre_code;
    uint8_t *pre_len;
    uint32_t *scratch;
};

// ---- the pieces of a stream's life, shared by the greedy / lazy kernel and the near-optimal one.
// All of them are called by every thread of the CTA (they contain barriers).
struct HcsStream {                 // uniform per stream
    const uint8_t *gin;            // the stream in global memory
    uint32_t len;
    uint32_t gmis;                 // gin & 3
    const uint32_t *gw;            // aligned words that hold the stream
    __device__ __forceinline__ void set(const uint8_t *p, uint32_t n)
    {
        gin = p; len = n;
        gmis = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
        gw = reinterpret_cast<const uint32_t *>(p - gmis);
    }
    // 4 bytes of the stream at p (only words that hold a byte of the stream are touched)
    __device__ __forceinline__ uint32_t g32(uint32_t p) const
    {
        const uint32_t q = p + gmis, k = q >> 2, sh = 8u * (q & 3u);
        const uint32_t w0 = __ldg(gw + k);
        if (sh == 0) return w0;
        const uint32_t w1 = (k + 1) * 4u < len + gmis ? __ldg(gw + k + 1) : 0u;
        return __funnelshift_r(w0, w1, sh);
    }
};

// phase 1: chains of the whole stream (head + link).  H4: chains of 4-byte hashes (near-optimal tier
// only; the reference's hash chains — and therefore levels 2..9 — hash 3 bytes).
__device__ __forceinline__ uint32_t hash4(uint32_t v32) { return (v32 * 0x1E35A7BDu) >> 17; }
template <bool H4 = false>
__device__ __forceinline__ void hcs_build_chains(HcsSmem &sm, const HcsStream &st)
{
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t len = st.len;
    auto g32 = [&](uint32_t p) { return st.g32(p); };
    for (uint32_t i = tid; i < 32768 / 8; i += HCS_THREADS) reinterpret_cast<uint4 *>(sm.head)[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
    for (uint32_t i = tid; i < 65536 / 8; i += HCS_THREADS) reinterpret_cast<uint4 *>(sm.link)[i] = make_uint4(0, 0, 0, 0);
    uint32_t l_head = 0, l_tail = 0;       // inserting warps: ring indices into their list (uniform per warp)
    for (uint32_t b0 = 0; b0 < len; b0 += HCS_IBLK) {
        __syncthreads();
        for (uint32_t k = tid; k < HCS_INS_WARPS * (HCS_IBLK / 32); k += HCS_THREADS) (&sm.ins.cls[0][0])[k] = 0;
        __syncthreads();
        // (i) hash of every position of the round, and one bit in the row of the warp that will link
        // it; the last two positions of a stream are never inserted
        for (uint32_t k = tid; k < HCS_IBLK; k += HCS_THREADS) {
            const uint32_t p = b0 + k;
            if (p + (H4 ? 4u : 3u) <= len) {
                const uint32_t h = H4 ? hash4(g32(p)) : hash3(g32(p) & 0xFFFFFFu);
                sm.ins.hbuf[k] = (uint16_t)h;
                atomicOr(&sm.ins.cls[h & (HCS_INS_WARPS - 1)][k >> 5], 1u << (k & 31u));
            }
        }
        __syncthreads();
        // (ii) warp w links the positions whose hash is w modulo 16, in ascending order: it
        // collects them from its bit row (8 positions per lane and step) and links 32 at a time.
        // Chains of different hash values never touch, so the warps do not wait for one another.
        if (warp < HCS_INS_WARPS) {
            uint16_t *list = sm.ins.list[warp];
            const uint32_t nblk = len - b0 < HCS_IBLK ? len - b0 : HCS_IBLK;
            auto link_some = [&](uint32_t have) {
                const bool ok = lane < have;
                uint32_t k = 0, h = 0x10000u + lane;          // unique key for idle lanes
                if (ok) { k = list[(l_head + lane) % HCS_LIST]; h = sm.ins.hbuf[k]; }
                const uint32_t p = b0 + k;
                const unsigned peers = __match_any_sync(BDF_FULL_MASK, h);
                const unsigned lower = peers & lanemask_lt();
                const uint32_t peer_pos = __shfl_sync(BDF_FULL_MASK, p, lower ? 31 - __clz(lower) : 0);
                if (ok) {
                    const uint32_t prev = lower ? peer_pos : (uint32_t)sm.head[h];
                    sm.link[p] = (!lower && prev == HCS_NONE) ? (uint16_t)0 : (uint16_t)(p - prev);
                }
                __syncwarp();
                if (ok && (peers >> lane) == 1u) sm.head[h] = (uint16_t)p;
                __syncwarp();
                l_head += have;
            };
            for (uint32_t s0 = 0; s0 < nblk; s0 += 256) {
                uint32_t mine = (sm.ins.cls[warp][(s0 >> 5) + (lane >> 2)] >> (8u * (lane & 3u))) & 0xFFu;
                const uint32_t cnt = __popc(mine);
                uint32_t incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t t = __shfl_up_sync(BDF_FULL_MASK, incl, d);
                    if (lane >= (unsigned)d) incl += t;
                }
                uint32_t at = l_tail + incl - cnt;
                while (mine) {
                    const uint32_t k = __ffs(mine) - 1;
                    mine &= mine - 1;
                    list[at % HCS_LIST] = (uint16_t)(s0 + 8 * lane + k);    // position inside the round
                    at++;
                }
                l_tail += __shfl_sync(BDF_FULL_MASK, incl, 31);
                __syncwarp();
                while (l_tail - l_head >= 32u) link_some(32u);
            }
            while (l_tail != l_head) link_some(l_tail - l_head < 32u ? l_tail - l_head : 32u);   // the buffer is rewritten next round
        }
    }
    __syncthreads();
}

// phase 2: the stream moves into shared memory (over the head table), zero padded
__device__ __forceinline__ void hcs_stage_input(HcsSmem &sm, const HcsStream &st)
{
    const unsigned tid = threadIdx.x;
    const uint32_t len = st.len;
    auto g32 = [&](uint32_t p) { return st.g32(p); };
    {
        uint32_t *iw = reinterpret_cast<uint32_t *>(sm.in);
        const uint32_t nw = (len + 3) >> 2;
        for (uint32_t k = tid; k < (65536 + 32) / 4; k += HCS_THREADS) {
            uint32_t v = 0;
            if (k < nw) {
                v = g32(4 * k);
                if (4 * k + 4 > len) v &= 0xFFFFFFFFu >> (8 * (4 * k + 4 - len));     // zero padding behind the stream
            }
            iw[k] = v;
        }
    }
}

// find_match for the positions [entry, entry + nsearch): results to sm.w.res[position - entry].
// LISTS: every improvement along the chain is also kept — up to HCS_NLIST (len | offset << 16)
// entries of strictly increasing length per position in lists[position * HCS_NLIST ..] (global), the
// match list the near-optimal parser relaxes (find_matches, src/compress/matchfinder.rs:1283-1296).
constexpr uint32_t HCS_NLIST = 8;
// Quick reject of a candidate once a match of `best` bytes is known: a longer match has to agree on the
// bytes best-3 .. best, so the walk compares that WORD (for best == 3 it is the 4-byte head itself)
// instead of the single byte at best the reference looks at (matchfinder.rs:829-836).  Same result —
// only candidates that cannot improve are dropped — and about half as many reach the compare, whose
// divergent path is what the walk pays for (profiles/r2_hcs_mixed_summary.md).  BDF_HCS_TAIL4=0: the byte test.
#ifndef BDF_HCS_TAIL4
#define BDF_HCS_TAIL4 1
#endif
#if BDF_HCS_TAIL4
#define HCS_TAIL_AT(q) hcs_ld32(sm.in, (q) - 3u)
#else
#define HCS_TAIL_AT(q) ((uint32_t)sm.in[(q)])
#endif
// near3 (near-optimal tier): the chains in shared memory are chains of 4-byte hashes; near3[p] is the
// distance from p to the closest earlier position with the same 3-byte hash (0 = none), tried first,
// so that 3-byte matches are not lost (the shape of the reference's own near-optimal matchfinder: one
// bucket of 3-byte hashes in front of the structure that holds the 4-byte ones, matchfinder.rs:1344-1463).
template <bool LISTS>
__device__ __forceinline__ void hcs_search(HcsSmem &sm, uint32_t len, const HcParams &prm, uint32_t entry, uint32_t nsearch,
                                           uint32_t *lists, const uint16_t *near3 = nullptr)
{
    const unsigned lane = threadIdx.x & 31u;
    // ---- search: warps take chunks of 32 positions from a counter; every lane walks one chain at a
    // time and takes the next position of the chunk as soon as it is done
    {
        uint32_t next = 0, range_end = 0;
        bool active = false, exhausted = false;
        uint32_t p = 0, cur = 0, best = 0, boff = 0, depth = 0, first = 0, room = 0, src4 = 0, tb = 0, nl = 0;
        bool can4 = false;
        for (;;) {
            if (next >= range_end && !exhausted) {  // uniform: take the next chunk
                uint32_t c = 0;
                if (lane == 0) c = atomicAdd(&sm.c_search_next, 32u);
                c = __shfl_sync(BDF_FULL_MASK, c, 0);
                next = c < nsearch ? c : nsearch;
                range_end = c + 32u < nsearch ? c + 32u : nsearch;
                exhausted = c + 32u >= nsearch;
            }
            const unsigned idle = __ballot_sync(BDF_FULL_MASK, !active);
            if (idle && next < range_end) {
                if (!active) {
                    const uint32_t q = next + __popc(idle & lanemask_lt());
                    if (q < range_end) {
                        p = entry + q;
                        first = sm.link[p];
                        if (LISTS) {
#pragma unroll
                            for (uint32_t k = 0; k < HCS_NLIST; k += 4)
                                *reinterpret_cast<uint4 *>(lists + (size_t)p * HCS_NLIST + k) = make_uint4(0, 0, 0, 0);
                            nl = 0;
                        }
                        if (LISTS && near3) {
                            best = 0; boff = 0; depth = 0;
                            src4 = hcs_ld32(sm.in, p);
                            room = len - p < 258u ? len - p : 258u;
                            can4 = p + 4 <= len;
                            bool over = false;                  // nothing can follow the 3-byte candidate
                            const uint32_t n3 = p + 3 <= len ? (uint32_t)near3[p] : 0u;
                            if (n3 != 0 && n3 <= 32768u) {
                                const uint32_t m4 = hcs_ld32(sm.in, p - n3);
                                if (((m4 ^ src4) & 0xFFFFFFu) == 0) {
                                    const uint32_t l = (can4 && m4 == src4) ? 4 + hcs_prefix(sm.in, p - n3 + 4, p + 4, room - 4) : 3u;
                                    best = l; boff = n3;
                                    lists[(size_t)p * HCS_NLIST] = l | n3 << 16; nl = 1;
                                    if (l >= prm.nice_len || l == 258 || p + l >= len) over = true;
                                    else tb = HCS_TAIL_AT(p + l);
                                }
                            }
                            if (over || !can4 || first == 0) sm.w.res[q] = best | boff << 16;
                            else { cur = p - first; active = true; }
                        } else if (p + 3 > len || first == 0) sm.w.res[q] = 0;
                        else {
                            cur = p - first; best = 0; boff = 0; depth = 0;
                            src4 = hcs_ld32(sm.in, p);
                            room = len - p < 258u ? len - p : 258u;
                            can4 = p + 4 <= len;
                            active = true;
                        }
                    }
                }
                next += __popc(idle);
            }
            if (!__any_sync(BDF_FULL_MASK, active)) {
                if (next >= range_end && exhausted) break;
                continue;
            }
            if (active) {
                // a burst of candidates of find_match_impl (src/compress/matchfinder.rs:812-887); most of
                // them fall at the quick reject (the byte at best_len), which is all the loop carries
                bool done = false;
#pragma unroll 1
                for (int burst = 0; burst < HCS_BURST; burst++) {
                    const uint32_t off = p - cur;
                    if (off > 32768u) { done = true; break; }
                    if (!(best >= 3 && HCS_TAIL_AT(cur + best) != tb)) {
                        const uint32_t m4 = hcs_ld32(sm.in, cur);
                        const bool eq3 = ((m4 ^ src4) & 0xFFFFFFu) == 0;
                        if (can4) {
                            if (m4 == src4) {
                                const uint32_t l = 4 + hcs_prefix(sm.in, cur + 4, p + 4, room - 4);
                                if (l > best) {
                                    best = l; boff = off;
                                    if (LISTS) { lists[(size_t)p * HCS_NLIST + (nl < HCS_NLIST ? nl : HCS_NLIST - 1)] = l | off << 16; nl++; }
                                    // nice_len / 258 reached, or nothing longer can follow (the reference
                                    // leaves its loop at the next candidate: pos + best_len >= len)
                                    if (l >= prm.nice_len || l == 258 || p + l >= len) { done = true; break; }
                                    tb = HCS_TAIL_AT(p + l);
                                }
                            } else if (best < 3 && eq3) {
                                best = 3; boff = off;
                                if (LISTS) { lists[(size_t)p * HCS_NLIST] = 3u | off << 16; nl = 1; }
                                if (p + 3 >= len) { done = true; break; }
                                tb = HCS_TAIL_AT(p + 3);
                            }
                        } else if (eq3 && best < 3) {          // room == 3: p + 3 == len
                            best = 3; boff = off;
                            if (LISTS) { lists[(size_t)p * HCS_NLIST] = 3u | off << 16; nl = 1; }
                            done = true;
                            break;
                        }
                    }
                    // prev_tab is indexed modulo 32768 in the reference: a candidate exactly one
                    // window back reads the slot the current position has just overwritten
                    const uint32_t lk = off == 32768u ? first : (uint32_t)sm.link[cur];
                    if (!lk || lk > cur) { done = true; break; }      // end of the chain (the aliased link can point in front of the stream)
                    cur -= lk;
                    if (++depth >= prm.max_depth) { done = true; break; }
                }
                if (done) {
                    BDF_ASSERT(p - entry < HCS_SEARCH && best <= 258 && boff <= 32768 && (best < 3 || (boff >= 1 && boff <= p)));
                    sm.w.res[p - entry] = best | boff << 16;
                    active = false;
                }
            }
        }
    }
}

// the step from every position of the window (decide_greedy_sequences, src/compress/mod.rs:1290-1340)
__device__ __forceinline__ void hcs_steps_greedy(HcsSmem &sm, uint32_t len, const HcParams &prm, uint32_t entry, uint32_t wvalid)
{
    const unsigned tid = threadIdx.x;
        for (uint32_t i = tid; i < wvalid; i += HCS_THREADS) {
        const uint32_t p = entry + i;
        const uint32_t l = sm.w.res[i] & 0xFFFFu;
        uint32_t v;
        if (l < 3) v = 1u;
        else {
            uint32_t nl = 0, L = l;
            if (prm.lazy >= 1 && p + 1 < len && l < prm.nice_len) {
                const uint32_t l1 = sm.w.res[i + 1] & 0xFFFFu;
                if (l1 > l) {
                    nl = 1; L = l1;
                    if (prm.lazy >= 2 && p + 2 < len) {
                        const uint32_t l2 = sm.w.res[i + 2] & 0xFFFFu;
                        if (l2 > l1) { nl = 2; L = l2; }
                    }
                }
            }
            v = (nl + L) | nl << 9 | 1u << 11;
        }
        sm.w.nxt[i] = (uint16_t)v;
    }
}

// The parse of one window whose steps are in sm.w.nxt (and whose matches are in sm.w.res): path,
// records, symbol counts; with check_split the place where BlockSplitStats acts and its verdict.
// Results: sm.c_next_entry (window-relative), sm.c_win_rec, sm.c_pc, sm.c_rec_at_pc, sm.c_split;
// freq_a / obs_a / cnt_a have been added to the block, freq_b / obs_b / cnt_b are left to the caller.
__device__ __forceinline__ void hcs_parse_window(HcsSmem &sm, uint32_t len, uint32_t entry, uint32_t wvalid, uint32_t block_start,
                                                 uint32_t nrec, uint32_t *recs, bool check_split)
{
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (tid < HCS_NSEG + 8) sm.seg_entry[tid] = (uint16_t)HCS_NONE;
    if (tid < HCS_NGRP + 1) sm.grp_entry[tid] = (uint16_t)HCS_NONE;
    for (uint32_t i = tid; i < 320; i += HCS_THREADS) { sm.freq_a[i] = 0; sm.freq_b[i] = 0; }
    if (tid < 14) { sm.obs_a[tid] = 0; sm.obs_b[tid] = 0; }
    if (tid == 0) { sm.cnt_a = 0; sm.cnt_b = 0; sm.c_pc = 0xFFFFFFFFu; sm.c_rec_at_pc = 0; sm.c_split = 0; sm.c_search_next = 0; }
    __syncthreads();
    // ---- P2: one thread per segment, backwards: exit and number of steps from every entry
    if (tid < HCS_NSEG) {
        const uint32_t s0 = tid * HCS_SEG, s1 = s0 + HCS_SEG;
        const uint32_t top = s1 < wvalid ? s1 : wvalid;
        for (uint32_t i = top; i-- > s0; ) {
            const uint32_t j = i + (sm.w.nxt[i] & 511u);
            uint32_t ex = j >= s1 ? j - s1 : 0u, st = 1u;       // (0: the stream ends inside this segment)
            if (j < top) {
                const uint32_t t = sm.w.sw[j];
                ex = t & 511u; st += t >> 9;
            }
            sm.w.sw[i] = ex | st << 9;
        }
    }
    __syncthreads();
    // ---- P2b: one warp per group of eight segments, last segment first: exit from the group
    if (warp < HCS_NGRP) {
        const uint32_t g0 = warp * HCS_GLEN, g1 = g0 + HCS_GLEN;
        for (uint32_t s = HCS_GSEG; s-- > 0; ) {
            const uint32_t i = g0 + s * HCS_SEG + lane, s1 = g0 + (s + 1) * HCS_SEG;
            if (lane < HCS_SEG && i < wvalid) {
                const uint32_t x = s1 + (sm.w.sw[i] & 511u);
                sm.w.ex2[i] = (uint16_t)((x >= g1 || x >= wvalid) ? x : (uint32_t)sm.w.ex2[x]);
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- P3: one thread walks the groups
    if (tid == 0) {
        uint32_t cur = 0;
        while (cur < wvalid) {
            sm.grp_entry[cur / HCS_GLEN] = (uint16_t)cur;
            cur = sm.w.ex2[cur];
        }
        sm.c_next_entry = cur;
    }
    __syncthreads();
    // ---- one thread per group walks its segments
    if (tid < HCS_NGRP) {
        uint32_t cur = sm.grp_entry[tid];
        const uint32_t g1 = (tid + 1) * HCS_GLEN;
        if (cur != HCS_NONE) {
            while (cur < g1 && cur < wvalid) {
                const uint32_t s = cur / HCS_SEG;
                sm.seg_entry[s] = (uint16_t)cur;
                cur = (s + 1) * HCS_SEG + (sm.w.sw[cur] & 511u);
            }
        }
    }
    __syncthreads();
    // ---- exclusive scan of the step counts over the entered segments (128 threads)
    if (tid < 128) {
        const uint32_t e = tid < HCS_NSEG ? (uint32_t)sm.seg_entry[tid] : HCS_NONE;
        const uint32_t st = e != HCS_NONE ? sm.w.sw[e] >> 9 : 0u;
        uint32_t ist = st;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(BDF_FULL_MASK, ist, d);
            if (lane >= (unsigned)d) ist += x;
        }
        if (lane == 31) sm.scan_tmp[warp] = ist;
        // (warps 0..3 only; a named barrier keeps the other 28 warps out of it)
        asm volatile("bar.sync 1, 128;" ::: "memory");
        uint32_t base = 0;
        for (unsigned k = 0; k < warp; k++) base += sm.scan_tmp[k];
        if (tid < HCS_NSEG) sm.seg_rec[tid] = base + ist - st;
        if (tid == 127) sm.c_win_obs = base + ist;           // steps of this window
    }
    __syncthreads();
    // ---- every entered segment lists its steps (ex2 is free now); all that follows is per step
    uint16_t *steps = sm.w.ex2;
    if (tid < HCS_NSEG) {
        uint32_t cur = sm.seg_entry[tid];
        if (cur != HCS_NONE) {
            uint32_t k = sm.seg_rec[tid];
            const uint32_t s1 = (tid + 1) * HCS_SEG;
            while (cur < s1 && cur < wvalid) {
                steps[k++] = (uint16_t)cur;
                cur += sm.w.nxt[cur] & 511u;
            }
        }
    }
    __syncthreads();
    // ---- records and observations in front of every step: thread t has steps 2t and 2t + 1
    const uint32_t nsteps = sm.c_win_obs;
    uint32_t cur0 = 0, cur1 = 0, n0 = 0, n1 = 0, rc0 = 0, rc1 = 0, ob0 = 0, ob1 = 0;
    if (2 * tid < nsteps) {
        cur0 = steps[2 * tid]; n0 = sm.w.nxt[cur0];
        rc0 = ((n0 >> 9) & 3u) + 1u; ob0 = rc0 + ((n0 >> 11) & 1u);
    }
    if (2 * tid + 1 < nsteps) {
        cur1 = steps[2 * tid + 1]; n1 = sm.w.nxt[cur1];
        rc1 = ((n1 >> 9) & 3u) + 1u; ob1 = rc1 + ((n1 >> 11) & 1u);
    }
    uint32_t base_rc, base_ob;
    {
        const uint32_t mine = (rc0 + rc1) | (ob0 + ob1) << 16;       // < 65536 each
        uint32_t inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(BDF_FULL_MASK, inc, d);
            if (lane >= (unsigned)d) inc += x;
        }
        if (lane == 31) sm.seg_obs[warp] = inc;                      // (seg_obs doubles as the scratch of this scan)
        __syncthreads();
        uint32_t v = sm.seg_obs[lane], pre = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t x = __shfl_up_sync(BDF_FULL_MASK, pre, d);
            if (lane >= (unsigned)d) pre += x;
        }
        const uint32_t tot = __shfl_sync(BDF_FULL_MASK, pre, 31);
        const uint32_t wbase = __shfl_sync(BDF_FULL_MASK, pre - v, warp);
        const uint32_t ex = wbase + inc - mine;
        base_rc = ex & 0xFFFFu; base_ob = ex >> 16;
        if (tid == 0) sm.c_win_rec = tot & 0xFFFFu;
    }
    // ---- where BlockSplitStats acts (src/compress/mod.rs:387-415): the first step whose top sees
    // >= 2048 pending observations, a block of >= 5000 bytes and > 5000 bytes left
    {
        const uint32_t num_new0 = sm.num_new;
        if (check_split && 2 * tid < nsteps) {
            const uint32_t p = entry + cur0;
            if (num_new0 + base_ob >= 2048u && p - block_start >= 5000u && len - p > 5000u) atomicMin(&sm.c_pc, cur0);
        }
        if (check_split && 2 * tid + 1 < nsteps) {
            const uint32_t p = entry + cur1;
            if (num_new0 + base_ob + ob0 >= 2048u && p - block_start >= 5000u && len - p > 5000u) atomicMin(&sm.c_pc, cur1);
        }
    }
    __syncthreads();
    // ---- every step writes its records and counts its symbols
    const uint32_t pc = sm.c_pc;             // window-relative, 0xFFFFFFFF = no check in this window
#pragma unroll
    for (int h = 0; h < 2; h++) {
        if (2 * tid + h < nsteps) {
            const uint32_t cur = h ? cur1 : cur0, n = h ? n1 : n0;
            uint32_t r = nrec + base_rc + (h ? rc0 : 0u);
            BDF_ASSERT(r + 3 <= 65536 + 64 && entry + cur < len);
            const bool behind = cur >= pc;
            if (cur == pc) sm.c_rec_at_pc = r;
            uint32_t *fq = behind ? sm.freq_b : sm.freq_a;
            uint32_t *ob = behind ? sm.obs_b : sm.obs_a;
            const uint32_t nl = (n >> 9) & 3u;
            const uint32_t p = entry + cur;
            if (n & 0x800u) {
                for (uint32_t k = 0; k < nl; k++) {
                    const uint32_t b = sm.in[p + k];
                    recs[r++] = b;
                    atomicAdd(&fq[b], 1u);
                    atomicAdd(&ob[b >> 5], 1u);
                }
                const uint32_t m = sm.w.res[cur + nl], L = m & 0xFFFFu, O = m >> 16;
                recs[r] = HCS_REC_MATCH | L << 16 | (O - 1u);
                const unsigned slot = offset_slot_of(O);
                atomicAdd(&fq[257 + length_slot_of(L)], 1u);
                atomicAdd(&fq[288 + slot], 1u);
                atomicAdd(&ob[8 + (L >= 8)], 1u);
                atomicAdd(&ob[10 + (slot < 16 ? 0 : slot < 24 ? 1 : slot < 30 ? 2 : 0)], 1u);
                atomicAdd(behind ? &sm.cnt_b : &sm.cnt_a, nl + 2u);
            } else {
                const uint32_t b = sm.in[p];
                recs[r] = b;
                atomicAdd(&fq[b], 1u);
                atomicAdd(&ob[b >> 5], 1u);
                atomicAdd(behind ? &sm.cnt_b : &sm.cnt_a, 1u);
            }
        }
    }
    __syncthreads();
    // ---- bookkeeping: what is in front of the check joins the block; the check; the rest
    for (uint32_t i = tid; i < 320; i += HCS_THREADS) {
        if (i < 288) sm.litlen_freq[i] += sm.freq_a[i];
        else sm.offset_freq[i - 288] += sm.freq_a[i];
    }
    if (tid < 14) sm.new_obs[tid] += sm.obs_a[tid];
    __syncthreads();
    if (tid == 0) {
        sm.num_new += sm.cnt_a;
        if (check_split && pc != 0xFFFFFFFFu) sm.c_split = hc_should_end(sm, entry + pc - block_start, len - (entry + pc)) ? 1u : 0u;
    }
    __syncthreads();
}

// One block: the two Huffman codes from the block's histograms, the dynamic header
// (write_dynamic_huffman_header_impl, src/compress/mod.rs:1775-1883) and the symbols recs[rec_begin, rec_end).
template <bool SIZE>
__device__ __forceinline__ void hcs_encode_block(HcsSmem &sm, CtaSink<SIZE> &sink, const uint32_t *recs, uint32_t blk_rec_begin,
                                                 uint32_t blk_rec_end, bool is_final)
{
    const unsigned tid = threadIdx.x;
    __syncthreads();
    unsigned nlit_syms = 0, noff_syms = 0, npre = 0, nitems = 0;
    if (tid == 0) sm.litlen_freq[256]++;
    __syncthreads();
    make_huffman_code_cta(288, 14, sm.litlen_freq, sm.litlen_len, sm.enc.litlen_code, sm.enc.scratch);
    make_huffman_code_cta(32, 15, sm.offset_freq, sm.offset_len, sm.enc.offset_code, sm.enc.scratch);
    if (tid == 0) {
        HcsHeaderView hv{sm.litlen_len, sm.offset_len, sm.enc.hdr_lens, sm.enc.hdr_items, sm.enc.pre_freq,
                         sm.enc.pre_code, sm.enc.pre_len, sm.enc.scratch};
        hc_prepare_header(hv, nlit_syms, noff_syms, npre, nitems);
        sm.scan_tmp[0] = nlit_syms; sm.scan_tmp[1] = noff_syms; sm.scan_tmp[2] = npre; sm.scan_tmp[3] = nitems;
    }
    sink.begin_block(sm);          // zeroes the staging words (and is the barrier behind thread 0's work)
    nlit_syms = sm.scan_tmp[0]; noff_syms = sm.scan_tmp[1]; npre = sm.scan_tmp[2]; nitems = sm.scan_tmp[3];
    {
        // BFINAL, BTYPE = 2, HLIT, HDIST, HCLEN (thread 0), then the precode lengths (threads 1..19)
        unsigned long long bits = 0;
        uint32_t nb = 0;
        if (tid == 0) {
            bits = (is_final ? 1u : 0u) | (2u << 1) | ((nlit_syms - 257) << 3) | ((noff_syms - 1) << 8) | ((npre - 4) << 13);
            nb = 17;
        } else if (tid <= npre) {
            const uint8_t perm[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
            bits = sm.enc.pre_len[perm[tid - 1]];
            nb = 3;
        }
        sink.put(sm, bits, nb);
    }
    for (uint32_t base = 0; base < nitems; base += HCS_THREADS) {
        unsigned long long bits = 0;
        uint32_t nb = 0;
        if (base + tid < nitems) {
            const unsigned it = sm.enc.hdr_items[base + tid], sym = it >> 8, extra = it & 0xFF;
            const unsigned cl = sm.enc.pre_len[sym];
            bits = sm.enc.pre_code[sym] | (extra << cl);
            nb = cl + (sym == 16 ? 2 : sym == 17 ? 3 : sym == 18 ? 7 : 0);
        }
        sink.put(sm, bits, nb);
    }
    for (uint32_t base = blk_rec_begin; base < blk_rec_end; base += HCS_THREADS) {
        unsigned long long bits = 0;
        uint32_t nb = 0;
        if (base + tid < blk_rec_end) {
            const uint32_t rec = recs[base + tid];
            if (!(rec & HCS_REC_MATCH)) {
                bits = sm.enc.litlen_code[rec]; nb = sm.litlen_len[rec];
            } else {
                const uint32_t L = (rec >> 16) & 0x1FFu, O = (rec & 0x7FFFu) + 1u;
                unsigned lslot = length_slot_of(L), lb, le;
                length_slot_info(lslot, lb, le);
                const unsigned lcl = sm.litlen_len[257 + lslot];
                const uint32_t lbits = sm.enc.litlen_code[257 + lslot] | ((L - lb) << lcl);
                const uint32_t lnb = lcl + le;
                unsigned oslot = offset_slot_of(O), ob_, oe;
                offset_slot_info(oslot, ob_, oe);
                const unsigned ocl = sm.offset_len[oslot];
                const uint32_t obits = sm.enc.offset_code[oslot] | ((O - ob_) << ocl);
                bits = (unsigned long long)obits << lnb | lbits;
                nb = lnb + ocl + oe;
            }
        }
        sink.put(sm, bits, nb);
    }
    sink.put(sm, tid == 0 ? sm.enc.litlen_code[256] : 0u, tid == 0 ? sm.litlen_len[256] : 0u);
}

// The end of a stream: sync flush / padding, footer, results (warp 0).
template <bool SIZE>
__device__ __forceinline__ void hcs_finish_stream(HcsSmem &sm, const DeflateArgs &a, unsigned long long idx, const uint8_t *gin,
                                                  uint32_t len, uint8_t *out, unsigned hdr, unsigned uflags)
{
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    __syncthreads();
    if (warp == 0) {
        // FlushMode::Sync (:662-681): 3 zero bits, pad to a byte, 00 00 FF FF; otherwise pad to a byte
        unsigned long long sz = sm.sink_flushed;
        uint32_t pend = sm.sink_bits;
        bool over = sm.sink_overflow != 0;
        const unsigned long long cap = SIZE ? ~0ull : unit_cap(len, uflags);
        uint32_t tail_bytes = 0;
        uint8_t tail[6];
        uint32_t carry = SIZE ? 0u : (sm.sink_carry & 0xFFu);
        if (uflags & UNIT_SYNC) {
            pend += 3;
            if (pend > 8) { tail[tail_bytes++] = (uint8_t)carry; carry = 0; pend -= 8; }
            tail[tail_bytes++] = (uint8_t)carry;          // padded to the byte
            tail[tail_bytes++] = 0; tail[tail_bytes++] = 0; tail[tail_bytes++] = 0xFF; tail[tail_bytes++] = 0xFF;
        } else if (pend) {
            tail[tail_bytes++] = (uint8_t)carry;
        }
        if (sz + tail_bytes > cap) over = true;
        if (!SIZE && !over && lane < tail_bytes) out[hdr + sz + lane] = tail[lane];
        sz += tail_bytes;
        int st = BDF_OK;
        if (over) { st = BDF_INSUFFICIENT_SPACE; sz = 0; }
        else if (!SIZE) sz = frame_footer(a.format, gin, len, out, hdr + sz, g_crc_tables.slice, g_crc_tables.x2n, lane);
        if (lane == 0) { a.status[idx] = st; a.out_size[idx] = sz; }
    }
}

template <bool SIZE>
__global__ void __launch_bounds__(HCS_THREADS, 1) deflate_hcs_kernel(DeflateArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HcsSmem &sm = *reinterpret_cast<HcsSmem *>(smem_raw);
    __shared__ unsigned long long s_idx;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const HcParams prm = hc_params(a.level);
    uint32_t *recs = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(a.scratch) + a.scratch_stride * blockIdx.x);

    for (;;) {
        __syncthreads();
        if (tid == 0) s_idx = atomicAdd(a.work_counter, 1ull);
        __syncthreads();
        const unsigned long long idx = s_idx;
        if (idx >= a.n) break;
        if (a.klass && a.klass[idx] != a.want) continue;       // the other kernel's stream
        const uint8_t *gin = a.in + a.in_off[idx];
        const uint64_t len64 = a.in_off[idx + 1] - a.in_off[idx];
        uint8_t *out = SIZE ? nullptr : a.out + a.out_off[idx];
        const unsigned uflags = unit_flags_of(a, idx);
        if (SIZE && len64 == 0) {             // the estimator's block loop never runs (:808)
            if (tid == 0) { a.status[idx] = BDF_OK; a.out_size[idx] = 0; }
            continue;
        }
        if (len64 > 65536) {
            if (tid == 0) { a.status[idx] = BDF_STREAM_UNSUPPORTED; a.out_size[idx] = 0; }
            continue;
        }
        const uint32_t len = (uint32_t)len64;
        HcsStream st;
        st.set(gin, len);
        hcs_build_chains(sm, st);
        hcs_stage_input(sm, st);
        CtaSink<SIZE> sink;
        if (warp == 0 && !SIZE) frame_header(a.format, a.level, out, lane);
        const unsigned hdr = SIZE ? 0u : a.format == BDF_ZLIB ? 2u : a.format == BDF_GZIP ? 10u : 0u;   // what frame_header wrote
        sink.init(sm, SIZE ? nullptr : out + hdr, SIZE ? ~0ull : unit_cap(len, uflags));
        for (uint32_t i = tid; i < 288; i += HCS_THREADS) sm.litlen_freq[i] = 0;
        if (tid < 32) sm.offset_freq[tid] = 0;
        if (tid < 14) { sm.new_obs[tid] = 0; sm.obs[tid] = 0; }
        if (tid == 0) { sm.num_new = 0; sm.num_obs = 0; sm.c_search_next = 0; }
        __syncthreads();

        // ================================================= phase 3: windows
        uint32_t entry = 0;                 // parse position (uniform)
        uint32_t block_start = 0;           // uniform
        uint32_t blk_rec_begin = 0;         // first record of the current block (uniform)
        uint32_t nrec = 0;                  // records written so far (uniform)
        bool more = true;
        while (more) {
            const uint32_t wvalid = len - entry < HCS_W ? len - entry : HCS_W;      // positions parsed in this window
            const uint32_t nsearch = len - entry < HCS_SEARCH ? len - entry : HCS_SEARCH;
            hcs_search<false>(sm, len, prm, entry, nsearch, nullptr);
            __syncthreads();
            hcs_steps_greedy(sm, len, prm, entry, wvalid);
            hcs_parse_window(sm, len, entry, wvalid, block_start, nrec, recs, true);
            const uint32_t pc = sm.c_pc;
            const uint32_t win_rec = sm.c_win_rec;
            const uint32_t next_entry = entry + sm.c_next_entry;
            const bool split = sm.c_split != 0;
            const bool last = next_entry >= len;
            // blocks that end here: the one cut by the check, and the last one of the stream
            for (int pass = 0; pass < 2; pass++) {
                const bool do_split = pass == 0 && split;
                const bool do_last = pass == 1 && last;
                if (!do_split && !do_last) {
                    if (pass == 0) {
                        // no cut: what lies behind the check belongs to the same block
                        for (uint32_t i = tid; i < 320; i += HCS_THREADS) {
                            if (i < 288) sm.litlen_freq[i] += sm.freq_b[i];
                            else sm.offset_freq[i - 288] += sm.freq_b[i];
                        }
                        if (tid < 14) sm.new_obs[tid] += sm.obs_b[tid];
                        if (tid == 0) sm.num_new += sm.cnt_b;
                        __syncthreads();
                    }
                    continue;
                }
                const uint32_t blk_rec_end = do_split ? sm.c_rec_at_pc : nrec + win_rec;
                const bool is_final = do_last && (uflags & UNIT_FINISH);
                hcs_encode_block<SIZE>(sm, sink, recs, blk_rec_begin, blk_rec_end, is_final);
                // ---------------------------------------------- the next block starts empty
                if (do_split) {
                    block_start = entry + pc;
                    blk_rec_begin = blk_rec_end;
                    for (uint32_t i = tid; i < 320; i += HCS_THREADS) {
                        if (i < 288) sm.litlen_freq[i] = sm.freq_b[i];
                        else sm.offset_freq[i - 288] = sm.freq_b[i];
                    }
                    if (tid < 14) { sm.new_obs[tid] = sm.obs_b[tid]; sm.obs[tid] = 0; }
                    if (tid == 0) { sm.num_new = sm.cnt_b; sm.num_obs = 0; }
                    __syncthreads();
                }
            }
            nrec += win_rec;
            entry = next_entry;
            more = !last;
        }
        // an empty input is one block that holds only the end-of-block symbol (src/compress/mod.rs:648-660):
        // the window loop above ran once with nothing to parse and encoded it as the last block
        hcs_finish_stream<SIZE>(sm, a, idx, gin, len, out, hdr, uflags);
    }
}

// ---- levels 10..12 for streams of at most 64 KiB: a near-optimal parser on the same machinery.
// Replaces compress_near_optimal_block (src/compress/mod.rs:1586-1773) and the binary-tree
// matchfinder it drives (matchfinder.rs:1234-1776) for the batch path; this tier is held to the
// reference's SIZE (within 0.5 %), not to its bytes, so the parser is free:
//   * matches come from the hash chains: one pass over all positions keeps, per position, the best
//     match (for the greedy pass) and the list of improvements along the chain (len, offset of
//     strictly increasing length — what find_matches hands the reference's DP);
//   * pass 1, like the reference's: a greedy parse with BlockSplitStats decides where the block
//     ends and gives the histograms the symbol costs come from;
//   * pass 2 is a shortest path over the block like the reference's, but run BACKWARDS (cost to the
//     end of the block): one warp, three positions per step (a match is at least 3 long, so the
//     match edges of three neighbours only look at costs that are already final; their literal
//     edges are then chained), eight list entries per position in parallel lanes.  Running it
//     backwards makes the result a step function — exactly what the parallel parse of this file
//     consumes, so there is no serial backtrack;
//   * pass 3: that parse over the block (records, final histograms), then the block is encoded.
constexpr size_t NOS_SCRATCH_PER_CTA = HCS_SCRATCH_PER_CTA + 65536 * sizeof(uint32_t) * 2 + 65536 * HCS_NLIST * sizeof(uint32_t) +
                                      65536 * sizeof(uint16_t);        // + near3 (behind the lists)
constexpr uint32_t NOS_INF = 0x0FFFFFFFu;
constexpr uint32_t NOS_SHORTER = 7;

__device__ __forceinline__ HcParams nos_params(int level, int depth_override = 0)
{
    // depth / nice length of the chain walks.  The reference's binary trees search 35 / 100 / 300 nodes
    // deep (nice 75 / 150 / 258); a chain of 3-byte hashes has to walk further to see the same matches
    // (measured on text: depth 70 -> 3.1 % above the reference's size at level 10, 160 -> 1.2 %).
    HcParams p;
    if (level <= 10) { p.max_depth = 250; p.nice_len = 75; }
    else if (level == 11) { p.max_depth = 400; p.nice_len = 150; }
    else { p.max_depth = 600; p.nice_len = 258; }
#ifdef BDF_NOS_H4
    // chains of 4-byte hashes behind one 3-byte candidate: the same matches at a fraction of the depth
    if (level <= 10) p.max_depth = BDF_NOS_H4_D10;
    else if (level == 11) p.max_depth = BDF_NOS_H4_D11;
    else p.max_depth = BDF_NOS_H4_D12;
#endif
    if (depth_override > 0) p.max_depth = (uint32_t)depth_override;      // experiments (BDF_NOS_DEPTH)
    p.lazy = 0;
    return p;
}

// 3-byte chains first (only the closest candidate of every position is kept, in global memory), then
// the 4-byte chains the search walks
__device__ __forceinline__ void nos_build_chains_h4(HcsSmem &sm, const HcsStream &st, uint16_t *near3)
{
    hcs_build_chains<false>(sm, st);
    for (uint32_t i = threadIdx.x; i < 65536 / 8; i += HCS_THREADS)
        reinterpret_cast<uint4 *>(near3)[i] = reinterpret_cast<const uint4 *>(sm.link)[i];
    __syncthreads();
    hcs_build_chains<true>(sm, st);
}

__global__ void __launch_bounds__(HCS_THREADS, 1) deflate_nos_kernel(DeflateArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HcsSmem &sm = *reinterpret_cast<HcsSmem *>(smem_raw);
    __shared__ unsigned long long s_idx;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const HcParams prm = nos_params(a.level, a.nos_depth);
    uint8_t *slab = static_cast<uint8_t *>(a.scratch) + a.scratch_stride * blockIdx.x;
    uint32_t *recs = reinterpret_cast<uint32_t *>(slab);
    uint32_t *gbest = reinterpret_cast<uint32_t *>(slab + HCS_SCRATCH_PER_CTA);
    uint32_t *choice = gbest + 65536;
    uint32_t *lists = choice + 65536;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_idx = atomicAdd(a.work_counter, 1ull);
        __syncthreads();
        const unsigned long long idx = s_idx;
        if (idx >= a.n) break;
        const uint8_t *gin = a.in + a.in_off[idx];
        const uint64_t len64 = a.in_off[idx + 1] - a.in_off[idx];
        uint8_t *out = a.out + a.out_off[idx];
        const unsigned uflags = unit_flags_of(a, idx);
        if (len64 > 65536) {
            if (tid == 0) { a.status[idx] = BDF_STREAM_UNSUPPORTED; a.out_size[idx] = 0; }
            continue;
        }
        const uint32_t len = (uint32_t)len64;
        HcsStream st;
        st.set(gin, len);
        uint16_t *near3 = nullptr;
#ifdef BDF_NOS_H4
        near3 = reinterpret_cast<uint16_t *>(lists + (size_t)65536 * HCS_NLIST);
        nos_build_chains_h4(sm, st, near3);
#else
        hcs_build_chains(sm, st);
#endif
        hcs_stage_input(sm, st);
        CtaSink<false> sink;
        if (warp == 0) frame_header(a.format, a.level, out, lane);
        const unsigned hdr = a.format == BDF_ZLIB ? 2u : a.format == BDF_GZIP ? 10u : 0u;
        sink.init(sm, out + hdr, unit_cap(len, uflags));
        if (tid == 0) sm.c_search_next = 0;
        __syncthreads();
        // ---- every position: best match and match list
        for (uint32_t pos0 = 0; pos0 < len; pos0 += HCS_SEARCH) {
            const uint32_t ns = len - pos0 < HCS_SEARCH ? len - pos0 : HCS_SEARCH;
            hcs_search<true>(sm, len, prm, pos0, ns, lists, near3);
            __syncthreads();
            for (uint32_t i = tid; i < ns; i += HCS_THREADS) gbest[pos0 + i] = sm.w.res[i];
            if (tid == 0) sm.c_search_next = 0;
            __syncthreads();
        }
        // ---- blocks
        uint32_t block_start = 0;
        do {
            // pass 1: greedy parse (BlockSplitStats) -> end of the block, histograms for the costs
            for (uint32_t i = tid; i < 288; i += HCS_THREADS) sm.litlen_freq[i] = 0;
            if (tid < 32) sm.offset_freq[tid] = 0;
            if (tid < 14) { sm.new_obs[tid] = 0; sm.obs[tid] = 0; }
            if (tid == 0) { sm.num_new = 0; sm.num_obs = 0; }
            __syncthreads();
            uint32_t block_end = len;
            for (uint32_t entry = block_start; entry < len; ) {
                const uint32_t wvalid = len - entry < HCS_W ? len - entry : HCS_W;
                const uint32_t nload = len - entry < HCS_SEARCH ? len - entry : HCS_SEARCH;
                for (uint32_t i = tid; i < nload; i += HCS_THREADS) sm.w.res[i] = gbest[entry + i];
                __syncthreads();
                hcs_steps_greedy(sm, len, prm, entry, wvalid);
                hcs_parse_window(sm, len, entry, wvalid, block_start, 0, recs, true);
                if (sm.c_split) { block_end = entry + sm.c_pc; break; }
                for (uint32_t i = tid; i < 320; i += HCS_THREADS) {
                    if (i < 288) sm.litlen_freq[i] += sm.freq_b[i];
                    else sm.offset_freq[i - 288] += sm.freq_b[i];
                }
                if (tid < 14) sm.new_obs[tid] += sm.obs_b[tid];
                if (tid == 0) sm.num_new += sm.cnt_b;
                const uint32_t next_entry = entry + sm.c_next_entry;
                __syncthreads();
                entry = next_entry;
            }
            // symbol costs from the greedy histograms (update_costs, :2209-2224); a symbol the greedy
            // pass never produced costs what a rare symbol would
            if (tid == 0) sm.litlen_freq[256]++;
            __syncthreads();
            make_huffman_code_cta(288, 14, sm.litlen_freq, sm.litlen_len, sm.enc.litlen_code, sm.enc.scratch);
            make_huffman_code_cta(32, 15, sm.offset_freq, sm.offset_len, sm.enc.offset_code, sm.enc.scratch);
            for (uint32_t i = tid; i < 256 + 260 + 32; i += HCS_THREADS) {
                if (i < 256) sm.dp.lit_cost[i] = sm.litlen_len[i] ? sm.litlen_len[i] : 12;
                else if (i < 256 + 260) {
                    const uint32_t l = i - 256;
                    uint32_t c = 0;
                    if (l >= 3 && l <= 258) {
                        unsigned slot = length_slot_of(l), lb, le;
                        length_slot_info(slot, lb, le);
                        c = (sm.litlen_len[257 + slot] ? sm.litlen_len[257 + slot] : 10) + le;
                    }
                    sm.dp.len_cost[l] = (uint8_t)c;
                } else {
                    const uint32_t slot = i - 516;
                    unsigned ob_, oe;
                    offset_slot_info(slot < 30 ? slot : 29, ob_, oe);
                    sm.dp.slot_cost[slot] = (uint8_t)((sm.offset_len[slot] ? sm.offset_len[slot] : 8) + oe);
                }
            }
            __syncthreads();
            // pass 2: cost to the end of the block, backwards.  The match lists of 504 positions are staged
            // in shared memory by everybody, warp 0 walks them three positions per step, everybody writes
            // the chunk's choices out.
            if (tid == 0) sm.dp.ring[block_end & 511u] = 0;
            for (uint32_t chi = block_end; chi > block_start; ) {
                const uint32_t clo = chi - block_start > HCS_DP_CHUNK ? chi - HCS_DP_CHUNK : block_start;
                for (uint32_t i = tid; i < (chi - clo) * 2; i += HCS_THREADS)
                    reinterpret_cast<uint4 *>(sm.dp.lst)[i] = reinterpret_cast<const uint4 *>(lists + (size_t)clo * HCS_NLIST)[i];
                __syncthreads();
                if (warp == 0) {
                    const uint32_t g = lane >> 3, k = lane & 7u;
                    for (uint32_t hi = chi; hi > clo; hi = hi - clo > 3 ? hi - 3 : clo) {
                        const uint32_t p = hi - 1 - g;                  // this lane's position (groups 0..2)
                        const bool mine = g < 3 && hi >= clo + 1 + g;
                        const uint32_t e = mine ? sm.dp.lst[(p - clo) * 8 + k] : 0u;
                        // (the idle fourth group fetches the literal costs of the three positions)
                        uint32_t litc = 0;
                        if (g == 3 && k < 3 && hi >= clo + 1 + k) litc = sm.dp.lit_cost[sm.in[hi - 1 - k]];
                        // entry k stands for every length above entry k - 1's up to its own, at its offset: the
                        // lane tries its own length and up to NOS_SHORTER shorter ones (the reference's DP only
                        // relaxes the list lengths themselves, src/compress/mod.rs:1683-1700; trying the shorter
                        // ones as well makes up for what a chain walk finds less than the tree search)
                        const uint32_t e_prev = __shfl_up_sync(BDF_FULL_MASK, e, 1);
                        uint32_t v = NOS_INF << 4, vlen = 0;
                        if (e) {
                            const uint32_t l = e & 0xFFFFu, off = e >> 16;
                            const uint32_t lmin0 = k ? (e_prev & 0xFFFFu) + 1u : 3u;
                            const uint32_t lmin = l > lmin0 + NOS_SHORTER ? l - NOS_SHORTER : lmin0;
                            const uint32_t sc = sm.dp.slot_cost[offset_slot_of(off)];
#pragma unroll
                            for (uint32_t d = 0; d <= NOS_SHORTER; d++) {
                                const uint32_t t = l - d;
                                if (t >= lmin && t <= l && p + t <= block_end) {
                                    const uint32_t c = (sm.dp.len_cost[t] + sc + sm.dp.ring[(p + t) & 511u]) << 4 | k;
                                    if (c < v) { v = c; vlen = t; }
                                }
                            }
                        }
#pragma unroll
                        for (int d = 1; d < 8; d <<= 1) {
                            const uint32_t t = __shfl_xor_sync(BDF_FULL_MASK, v, d);
                            v = t < v ? t : v;
                        }
                        const uint32_t m0 = __shfl_sync(BDF_FULL_MASK, v, 0), m1 = __shfl_sync(BDF_FULL_MASK, v, 8), m2 = __shfl_sync(BDF_FULL_MASK, v, 16);
                        const uint32_t ew = (e & 0xFFFF0000u) | vlen;                   // this lane's best (length, offset)
                        const uint32_t w0 = __shfl_sync(BDF_FULL_MASK, ew, m0 & 7u), w1 = __shfl_sync(BDF_FULL_MASK, ew, 8 + (m1 & 7u)),
                                       w2 = __shfl_sync(BDF_FULL_MASK, ew, 16 + (m2 & 7u));
                        const uint32_t l0 = __shfl_sync(BDF_FULL_MASK, litc, 24), l1 = __shfl_sync(BDF_FULL_MASK, litc, 25),
                                       l2 = __shfl_sync(BDF_FULL_MASK, litc, 26);
                        if (lane == 0) {
                            uint32_t c = sm.dp.ring[hi & 511u];
                            const uint32_t mm[3] = {m0, m1, m2}, ww[3] = {w0, w1, w2}, ll[3] = {l0, l1, l2};
#pragma unroll
                            for (int j = 0; j < 3; j++) {
                                if (hi >= clo + 1u + j) {
                                    const uint32_t q = hi - 1 - j;
                                    const uint32_t lit = c + ll[j];
                                    const uint32_t mc = mm[j] >> 4;
                                    uint32_t ch = 1u;
                                    c = lit;
                                    if (mc < lit) { c = mc; ch = ww[j]; }
                                    BDF_ASSERT(q - clo < HCS_DP_CHUNK && q < 65536 && q + (ch & 0xFFFFu) <= block_end);
                                    sm.dp.ring[q & 511u] = c;
                                    sm.dp.ch[q - clo] = ch;
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
                __syncthreads();
                for (uint32_t i = tid; i < chi - clo; i += HCS_THREADS) choice[clo + i] = sm.dp.ch[i];
                __syncthreads();
                chi = clo;
            }
            // pass 3: the parse along the choices: records and the block's final histograms
            for (uint32_t i = tid; i < 288; i += HCS_THREADS) sm.litlen_freq[i] = 0;
            if (tid < 32) sm.offset_freq[tid] = 0;
            __syncthreads();
            uint32_t nrec = 0;
            for (uint32_t entry = block_start; entry < block_end; ) {
                const uint32_t wvalid = block_end - entry < HCS_W ? block_end - entry : HCS_W;
                for (uint32_t i = tid; i < wvalid; i += HCS_THREADS) {
                    const uint32_t ch = choice[entry + i], l = ch & 0xFFFFu;
                    sm.w.res[i] = ch;
                    sm.w.nxt[i] = (uint16_t)(l == 1 ? 1u : (l | 1u << 11));
                }
                __syncthreads();
                hcs_parse_window(sm, block_end, entry, wvalid, block_start, nrec, recs, false);
                nrec += sm.c_win_rec;
                const uint32_t next_entry = entry + sm.c_next_entry;
                __syncthreads();
                entry = next_entry;
            }
            const bool is_final = block_end >= len && (uflags & UNIT_FINISH);
            hcs_encode_block<false>(sm, sink, recs, 0, nrec, is_final);
            block_start = block_end;
        } while (block_start < len);
        hcs_finish_stream<false>(sm, a, idx, gin, len, out, hdr, uflags);
    }
}

}  // namespace bdf
