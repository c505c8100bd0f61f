// deflate_l1.cuh — level 1: single-probe hash table + static Huffman block.
//
// Replaces compress_greedy_block's level-1 branch with HtMatchFinder
// (reference src/compress/mod.rs:1499-1583, src/compress/matchfinder.rs:1139-1231)
// for inputs of at most 65536 bytes (one block, no split statistics, :1505-1529).
//
// The parse is inherently serial because only positions where find_match is
// CALLED enter the table (skip_positions is a no-op, :1231).  One warp owns a
// stream and speculates 32 consecutive positions at a time as "all literals":
// each lane probes as if every lower lane had been inserted (same-hash lower
// lanes are found with match.any), the first lane that really finds a match
// ends the speculation, lanes up to it commit their table writes, and the
// literals plus the match are bit-packed with a warp prefix sum.  The result
// is exactly the serial parse.
#pragma once
#include "deflate_common.cuh"

namespace bdf {

constexpr uint32_t L1_MAX_LEN = 65536;
constexpr uint32_t L1_EMPTY = 0xFFFFu;

constexpr int L1_WARPS = 4;                      // streams per CTA
constexpr size_t L1_TABLE_BYTES = 32768 * sizeof(uint16_t);

// The 64 KiB hash table of each stream (last position per bucket, 0xFFFF = empty) lives in
// a per-warp global slab that stays L2-resident while the stream is parsed: in shared
// memory it would cap the SM at three streams, and this parse is a chain of dependent
// probes that only many streams in flight can hide.
struct __align__(16) L1Smem {
    uint32_t sink[L1_WARPS][SINK_WORDS];
    uint32_t crc[4][256];
    uint32_t x2n[32];
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

__global__ void __launch_bounds__(L1_WARPS * 32) deflate_l1_kernel(DeflateArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    L1Smem &sm = *reinterpret_cast<L1Smem *>(smem_raw);
    const unsigned lane = lane_id(), warp = threadIdx.x >> 5;
    uint16_t *table = reinterpret_cast<uint16_t *>(static_cast<uint8_t *>(a.scratch) +
                                                   a.scratch_stride * (blockIdx.x * L1_WARPS + warp));
    if (a.format == BDF_GZIP) load_crc_tables_to_smem(sm.crc, sm.x2n);
    for (;;) {
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(a.work_counter, 1ull);
        idx = __shfl_sync(BDF_FULL_MASK, idx, 0);
        if (idx >= a.n) break;
        const uint8_t *in = a.in + a.in_off[idx];
        const uint64_t len64 = a.in_off[idx + 1] - a.in_off[idx];
        uint8_t *out = a.out + a.out_off[idx];
        if (len64 > L1_MAX_LEN) {
            if (lane == 0) { a.status[idx] = BDF_STREAM_UNSUPPORTED; a.out_size[idx] = 0; }
            continue;
        }
        const uint32_t len = (uint32_t)len64;
        for (unsigned i = lane; i < 32768 / 8; i += 32) reinterpret_cast<uint4 *>(table)[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
        const unsigned hdr = frame_header(a.format, 1, out, lane);
        BitSink bs;
        bs.init(sm.sink[warp], out + hdr, deflate_bound(len), lane);
        bs.put1(3, 3, lane);                     // BFINAL = 1, BTYPE = 01

        uint32_t pos = 0;
        while (pos < len) {
            const uint32_t p = pos + lane;
            const bool hashable = p + 3 <= len;          // the last two positions are never hashed (:1140)
            uint32_t v = 0, h = 0x10000u + lane;         // unique key for lanes that do not hash
            if (hashable) { v = ld24(in + p); h = hash3(v); }
            const unsigned peers = __match_any_sync(BDF_FULL_MASK, h);
            const unsigned lower = peers & lanemask_lt();
            uint32_t cand = L1_EMPTY;
            if (hashable) cand = lower ? pos + (31 - __clz(lower)) : table[h];
            const bool found = hashable && cand != L1_EMPTY && p - cand <= 32768u && ld24(in + cand) == v;
            const unsigned fb = __ballot_sync(BDF_FULL_MASK, found);
            const unsigned k = fb ? __ffs(fb) - 1 : 32;            // first lane whose probe hits
            // commit bucket writes of lanes 0..k (the match start is inserted too, :1162-1163)
            const unsigned committing = __ballot_sync(BDF_FULL_MASK, hashable && lane <= k);
            const unsigned mine = peers & committing;
            if (hashable && lane <= k && (mine >> lane) == 1u) table[h] = (uint16_t)p;
            // literals: lanes below k that are inside the input
            uint32_t bits = 0, nb = 0;
            if (lane < k && p < len) static_lit_code(in[p], bits, nb);
            bs.put(bits, nb, lane);
            if (k == 32) { pos += 32; continue; }
            const uint32_t mp = pos + k;
            const uint32_t mc = __shfl_sync(BDF_FULL_MASK, cand, k);
            const unsigned room = len - mp < 258 ? len - mp : 258;
            const unsigned mlen = warp_match_len(in + mc, in + mp, room, lane);
            uint32_t b0, n0, b1, n1;
            static_len_code(mlen, b0, n0);
            static_off_code(mp - mc, b1, n1);
            bs.put(lane == 0 ? b0 : lane == 1 ? b1 : 0, lane == 0 ? n0 : lane == 1 ? n1 : 0, lane);
            pos = mp + mlen;
            __syncwarp();
        }
        bs.put1(0, 7, lane);                     // end of block (symbol 256 = 0000000)
        uint64_t sz = bs.finish(lane);
        int st = BDF_OK;
        if (sz == ~0ull) { st = BDF_INSUFFICIENT_SPACE; sz = 0; }
        else sz = frame_footer(a.format, in, len, out, hdr + sz, sm.crc, sm.x2n, lane);
        if (lane == 0) { a.status[idx] = st; a.out_size[idx] = sz; }
        __syncwarp();
    }
}

}  // namespace bdf
