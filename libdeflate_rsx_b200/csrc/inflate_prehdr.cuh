// inflate_prehdr.cuh — first-block header pre-pass of a decompress batch, one LANE per stream (sm_100a).
//
// The code-length run of a dynamic block header (read_dynamic_huffman_header,
// src/decompress/mod.rs:403-507) is the one strictly serial piece of a stream: up to 320 precode
// symbols, each depending on the bit position the previous one left.  Inside inflate_kernel the 16
// lanes of a group all execute that chain for ONE stream (ncu, config 2: 31 % of the kernel's warp
// instructions, 24 % of its stall samples, profiles/r1_inflate_summary.md); inside
// inflate_lane_kernel the whole warp does, while 31 other streams wait.  Every stream of a batch
// starts with a block header, so the first one is decoded here for the whole batch with 32 streams
// per warp instruction: a lane parses the framing, the block bits and the precode of its stream,
// builds a 128-entry precode table of its own and decodes the litlen + offset code lengths into a
// row of 320 bytes.  The two inflate kernels pick the row up (load_code_lengths in inflate.cuh),
// seek their bit reader behind the header and go straight to the table build.
//
// The pre-pass never decides anything: a header it cannot vouch for (first block not dynamic,
// short input, precode that is not a complete code, repeat of "previous" at position 0, framing
// error) gets meta = 0 and the inflate kernels read that header themselves, with the reference's
// accept / reject rules.  A row is only marked valid if the in-kernel reader would have produced
// exactly the same lengths and ended on the same bit.
#pragma once
#include "inflate.cuh"

namespace bdf {

constexpr int PREHDR_THREADS = 64;

struct PrehdrArgs {
    const uint8_t *in;
    const uint64_t *in_off;
    uint32_t *rows;          // n x PREHDR_ROW_WORDS
    uint32_t *meta;          // n
    uint32_t n;
};

// ---- a dynamic block header read by ONE lane, in two pieces so that a lane can also spread the
// work over the rounds of its own decode loop (inflate_lane.cuh):
//   lane_hdr_begin    block bits, counts, precode lengths -> the lane's precode table
//   lane_hdr_lengths  up to `budget` code-length symbols of the run-length coded litlen + offset
//                     lengths (:441-497) into the lane's row
// ptab: 128 entries (sym << 3 | codeword bits), entry i at ptab[i * pstride]; row: bytes, any address
// space.  Every anomaly (not a dynamic block, short input, a precode that is not a complete code,
// "repeat previous" at position 0) is reported as failure WITHOUT a verdict: the caller hands the
// header to the group-wide reader, which applies the reference's accept / reject rules.
struct LaneHdr {
    uint32_t nlit, noff, final;
    uint32_t i, prev;            // progress of the run-length decode
};

__device__ __forceinline__ bool lane_hdr_begin(BitReader &br, uint32_t dlen, uint8_t *ptab, unsigned pstride, LaneHdr &h)
{
    br.refill();
    if (br.consumed_bits() + 3 > (int64_t)dlen * 8) return false;
    h.final = br.take(1);
    if (br.take(2) != 2) return false;
    br.refill();
    h.nlit = 257 + br.take(5);
    h.noff = 1 + br.take(5);
    const unsigned npre = 4 + br.take(4);
    br.refill();
    const unsigned n_lo = npre < 10 ? npre : 10;
    const uint32_t lo = br.take(3 * n_lo);
    br.refill();
    const uint32_t hi = npre > 10 ? br.take(3 * (npre - 10)) : 0;
    if (br.overrun()) return false;
    // precode lengths by symbol, 3 bits each (order 16,17,18,0,8,7,9,6,10,5,11,4,12,3,13,2,14,1,15)
    const uint64_t perm_lo = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 |
                             9ull << 30 | 6ull << 35 | 10ull << 40 | 5ull << 45;
    const uint64_t perm_hi = 11ull | 4ull << 5 | 12ull << 10 | 3ull << 15 | 13ull << 20 | 2ull << 25 |
                             14ull << 30 | 1ull << 35 | 15ull << 40;
    uint64_t pl = 0;         // length of precode symbol s at bits [3s, 3s+3)
    uint64_t cnt = 0;        // codewords of length l at bits [8l, 8l+8)
#pragma unroll
    for (unsigned k = 0; k < 19; k++) {
        const unsigned sym = k < 10 ? (unsigned)(perm_lo >> (5 * k)) & 31u : (unsigned)(perm_hi >> (5 * (k - 10))) & 31u;
        const unsigned v = k < 10 ? (lo >> (3 * k)) & 7u : (hi >> (3 * (k - 10))) & 7u;
        const unsigned l = k < npre ? v : 0;
        pl |= (uint64_t)l << (3 * sym);
        cnt += 1ull << (8 * l);
    }
    // complete code or nothing (Kraft sum in units of 2^-7), canonical first codewords
    uint32_t kraft = 0;
    uint64_t next = 0;       // next codeword of length l at bits [8l, 8l+8)
    {
        uint32_t code = 0;
#pragma unroll
        for (unsigned l = 1; l <= 7; l++) {
            const uint32_t c = (uint32_t)(cnt >> (8 * l)) & 255u;
            kraft += c << (7 - l);
            next |= (uint64_t)code << (8 * l);
            code = (code + c) << 1;
        }
    }
    if (kraft != 128u) return false;
#pragma unroll 1
    for (unsigned s = 0; s < 19; s++) {
        const unsigned l = (unsigned)(pl >> (3 * s)) & 7u;
        if (l == 0) continue;
        const uint32_t code = (uint32_t)(next >> (8 * l)) & 255u;
        next += 1ull << (8 * l);
        const unsigned rev = __brev(code) >> (32 - l);
        const uint8_t e = (uint8_t)(s << 3 | l);
        BDF_ASSERT(rev < (1u << l));
        for (unsigned i = rev; i < 128; i += 1u << l) ptab[i * pstride] = e;
    }
    h.i = 0; h.prev = 0;
    return true;
}

// -> 1 all lengths decoded (the caller still checks br.overrun()), 0 budget used up, -1 failure
struct PlainRefill {
    __device__ __forceinline__ void operator()(BitReader &b) const { b.refill(); }
};
template <class Refill = PlainRefill>
__device__ __forceinline__ int lane_hdr_lengths(BitReader &br, const uint8_t *ptab, unsigned pstride, uint8_t *rowb, LaneHdr &h,
                                                unsigned budget, Refill refill = Refill())
{
    const unsigned total = h.nlit + h.noff;
    unsigned i = h.i, prev = h.prev;
    int done = 0;
#pragma unroll 1
    while (i < total) {
        if (budget == 0) goto out;
        budget--;
        refill(br);                        // >= 33 valid bits: two plain lengths, or one and a repeat symbol
        uint32_t e = ptab[br.peek(7) * pstride];
        if ((e >> 3) < 16) {
            rowb[i++] = (uint8_t)(e >> 3);
            prev = e >> 3;
            br.drop(e & 7u);
            if (i >= total) break;
            e = ptab[br.peek(7) * pstride];
            if ((e >> 3) < 16) {
                rowb[i++] = (uint8_t)(e >> 3);
                prev = e >> 3;
                br.drop(e & 7u);
                continue;
            }
        }
        br.drop(e & 7u);
        const unsigned sym = e >> 3;
        unsigned rep, val;
        if (sym == 16) {
            if (i == 0) { done = -1; goto out; }
            rep = 3 + br.take(2);
            val = prev;
        } else if (sym == 17) {
            rep = 3 + br.take(3);
            val = 0;
        } else {
            rep = 11 + br.take(7);
            val = 0;
        }
        if (rep > total - i) rep = total - i;           // overruns of a repeat are clamped (:462-493)
        BDF_ASSERT(i + rep <= (unsigned)PREHDR_ROW_BYTES);
        for (unsigned q = 0; q < rep; q++) rowb[i + q] = (uint8_t)val;
        prev = val;
        i += rep;
    }
    done = 1;
out:
    h.i = i; h.prev = prev;
    return done;
}

// One stream's first header.  ptab: this warp's precode tables, entry i of lane l at ptab[i * 32 + l];
// row: this lane's 320-byte row in shared memory.  Returns the meta word, 0 if the header is left to
// the inflate kernels.
__device__ __forceinline__ uint32_t prehdr_decode(const uint8_t *p, uint32_t dlen, uint8_t *ptab, uint32_t *row, unsigned lane)
{
    BitReader br;
    br.init(p, dlen);
    LaneHdr h;
    if (!lane_hdr_begin(br, dlen, ptab + lane, 32, h)) return 0;
    uint8_t *rowb = reinterpret_cast<uint8_t *>(row);
    if (lane_hdr_lengths(br, ptab + lane, 32, rowb, h, ~0u) != 1) return 0;
    if (br.overrun()) return 0;
    for (unsigned q = h.nlit + h.noff; q < (unsigned)PREHDR_ROW_BYTES; q++) rowb[q] = 0;
    const int64_t cb = br.consumed_bits();
    if (cb < 0 || cb > 0xFFFF) return 0;
    return PREHDR_VALID | (uint32_t)cb | (h.nlit - 257) << 16 | (h.noff - 1) << 21 | h.final << 26;
}

template <int FORMAT>
__global__ void __launch_bounds__(PREHDR_THREADS) inflate_prehdr_kernel(PrehdrArgs a)
{
    constexpr int WARPS = PREHDR_THREADS / 32;
    constexpr int ROW_WORDS_PAD = PREHDR_ROW_WORDS + 1;          // odd stride: equal indices of different lanes in different banks
    __shared__ uint8_t s_ptab[WARPS][128 * 32];
    __shared__ uint32_t s_rows[WARPS][32 * ROW_WORDS_PAD];
    const unsigned w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint32_t *my_row = s_rows[w] + lane * ROW_WORDS_PAD;
    for (unsigned long long base = ((unsigned long long)blockIdx.x * WARPS + w) * 32ull; base < a.n;
         base += (unsigned long long)gridDim.x * PREHDR_THREADS) {
        const unsigned long long idx = base + lane;
        uint32_t meta = 0;
        if (idx < a.n) {
            const uint64_t o0 = a.in_off[idx], len64 = a.in_off[idx + 1] - o0;
            if (len64 <= 0xFFFFFFF0ull) {
                uint32_t at = 0, dlen = 0;
                const uint8_t *p = a.in + o0;
                if (inflate_frame_header<FORMAT>(p, (uint32_t)len64, at, dlen) == BDF_OK)
                    meta = prehdr_decode(p + at, dlen, s_ptab[w], my_row, lane);
            }
            a.meta[idx] = meta;
        }
        __syncwarp();
        // rows -> global memory, one coalesced row at a time
        unsigned valid = __ballot_sync(BDF_FULL_MASK, (meta & PREHDR_VALID) != 0);
        while (valid) {
            const unsigned r = __ffs(valid) - 1;
            valid &= valid - 1;
            uint32_t *dst = a.rows + (base + r) * PREHDR_ROW_WORDS;
            const uint32_t *src = s_rows[w] + r * ROW_WORDS_PAD;
            for (unsigned j = lane; j < (unsigned)PREHDR_ROW_WORDS; j += 32) dst[j] = src[j];
        }
        __syncwarp();
    }
}

}  // namespace bdf
