// deflate_hc.cuh — levels 2..9: hash chains + greedy / lazy / lazy2 parse,
// dynamic Huffman blocks.  One CTA per stream (inputs up to 65536 bytes).
//
// Replaces, for the batch path, the reference's
//   MatchFinder::{find_match_impl, skip_match, skip_positions}  src/compress/matchfinder.rs:754-1106
//   decide_greedy_sequences + BlockSplitStats                    src/compress/mod.rs:1261-1373, 271-416
//   make_huffman_code                                            src/compress/huffman_comp.rs
//   write_dynamic_huffman_header_impl / write_sequences          src/compress/mod.rs:1775-1883, 1952-2155
//
// GPU decomposition: every position is inserted into the chains
// exactly once and in ascending order whatever the parse does, so the result
// of find_match(p) depends only on the data.  The CTA therefore
//   1. inserts a window of positions (warp 0; same-hash positions inside a
//      32-wide group are linked with match.any, the rest through the bucket
//      heads), keeping one 16-bit back-link PER POSITION (no modulo-32768
//      aliasing; the reference's aliasing quirk — a candidate at distance
//      exactly 32768 reads the current position's link — is applied
//      explicitly when the chain is walked),
//   2. searches all positions of the window in parallel (one lane each),
//   3. lets warp 0 run the reference's serial parse over the stored results
//      (lazy probes become look-ups), with histograms and split statistics,
//   4. at a block boundary builds the two Huffman codes and bit-packs header
//      and symbols through the warp bit sink.
#pragma once
#include "deflate_common.cuh"

namespace bdf {

constexpr int HC_THREADS = 128;
constexpr int HC_WARPS = HC_THREADS / 32;
constexpr uint32_t HC_WINDOW_AFTER_JUMP = 4;   // >= lazy depth + 2
constexpr uint32_t HC_NSPEC = 32;              // small windows searched per round after a long match (see the kernel)
constexpr uint32_t HC_WINDOW = 1024;          // positions searched per round (max): fewer CTA barriers, better balance

// Symbol records written by the parse and consumed by the emitter, in stream order:
// a literal is its byte value; a match is a length record followed by an offset record.
constexpr uint32_t SYM_LEN = 0x10000u, SYM_OFF = 0x20000u;

struct __align__(16) HcSmem {
    uint16_t mlen[HC_WINDOW], moff[HC_WINDOW];
    uint32_t litlen_freq[288], offset_freq[32];
    uint32_t litlen_code[288], offset_code[32];     // bit-reversed codewords
    uint8_t litlen_len[288], offset_len[32];
    uint32_t scratch[288];             // make_huffman_code counters
    uint32_t sink[SINK_WORDS];
    uint32_t crc[4][256];
    uint32_t x2n[32];
    uint32_t new_obs[14], obs[14];
    uint32_t num_new, num_obs;
    // control words written by thread 0, read by everyone after a barrier
    uint32_t c_search_from, c_search_to, c_done, c_pos, c_ins_from, c_ins_to, c_nspec;
    uint32_t ins_pos[HC_WARPS][128];   // hc_insert_par: this warp's positions of the next 128, in order
    uint16_t ins_hash[HC_WARPS][128];
    uint8_t hdr_lens[320];
    uint16_t hdr_items[320];
    uint32_t pre_freq[19], pre_code[19];
    uint8_t pre_len[19];
};

// Hash chains of one stream.  They live in global memory (a per-CTA slab that stays
// L2-resident while the stream is being parsed) rather than in shared memory, so that
// many streams are in flight per SM and the serial parse of one hides behind the
// chain walks of the others.
//   head[32768]: bucket -> most recent position, all-ones = empty
//   link[MAX_LEN]: position -> distance to the previous same-hash position, 0 = none / further
//                  than 65535 (matchfinder.rs:799-805 stores 0 for those too)
// Two instances: streams up to 64 KiB (16-bit heads, the BASELINE shapes) and units up to
// 256 KiB (the chunk size of Compressor::compress, src/compress/mod.rs:699-772; 32-bit heads).
template <bool BIG>
struct HcChains {
    using head_t = typename std::conditional<BIG, uint32_t, uint16_t>::type;
    static constexpr uint32_t MAX_LEN = BIG ? 262144u : 65536u;
    static constexpr uint32_t NOPOS = BIG ? 0xFFFFFFFFu : 0xFFFFu;
    static constexpr size_t CHAIN_BYTES = 32768 * sizeof(head_t) + (size_t)MAX_LEN * sizeof(uint16_t);
    static constexpr size_t SCRATCH_PER_CTA = CHAIN_BYTES + ((size_t)MAX_LEN + 64) * sizeof(uint32_t);   // + symbol records
    head_t *head;
    uint16_t *link;
};

struct HcParams { unsigned max_depth, nice_len, lazy; };
__device__ __forceinline__ HcParams hc_params(int level)
{
    // init_params, src/compress/mod.rs:543-602; lazy depth :624-630
    HcParams p;
    switch (level) {
        case 2: p.max_depth = 6; p.nice_len = 10; break;
        case 3: p.max_depth = 12; p.nice_len = 14; break;
        case 4: case 5: p.max_depth = 16; p.nice_len = 30; break;
        case 6: p.max_depth = 35; p.nice_len = 65; break;
        case 7: p.max_depth = 100; p.nice_len = 130; break;
        case 8: p.max_depth = 300; p.nice_len = 258; break;
        default: p.max_depth = 600; p.nice_len = 258; break;
    }
    p.lazy = level >= 8 ? 2 : level >= 5 ? 1 : 0;
    return p;
}

// Insert positions [from, to) (warp 0, all lanes).  Positions with fewer than
// 3 bytes left are never inserted (matchfinder.rs:765,1021).
template <class CH>
__device__ __forceinline__ void hc_insert_range(const CH &ch, const uint8_t *in, uint32_t len, uint32_t from,
                                                uint32_t to, unsigned lane)
{
    for (uint32_t base = from; base < to; base += 32) {
        const uint32_t p = base + lane;
        const bool ok = p < to && p + 3 <= len;
        uint32_t h = 0x10000u + lane;
        if (ok) h = hash3(ld24(in + p));
        const unsigned peers = __match_any_sync(BDF_FULL_MASK, h);
        const unsigned lower = peers & lanemask_lt();
        if (ok) {
            uint32_t prev = lower ? base + (31 - __clz(lower)) : ch.head[h];
            ch.link[p] = (prev == CH::NOPOS || p - prev > 0xFFFFu) ? 0 : (uint16_t)(p - prev);
        }
        __syncwarp();
        if (ok && (peers >> lane) == 1u) ch.head[h] = (typename CH::head_t)p;
        __syncwarp();
    }
}

// The same insertion by all warps of the CTA at once.  Chains of different hash values never
// touch each other, so warp w takes the positions whose hash is congruent to w modulo HC_WARPS:
// it gathers them from the next 128 positions (in order) and links them 32 per step.  Every step
// is a dependent L2 round trip (bucket head read, then written), and one warp inserting 32
// positions per step was what bounded highly compressible data (config 4 on corpus A: nine
// steps per 258-byte match); four warps working on four times as many positions per step cut
// the number of steps by four.
template <class CH, class S>
__device__ __forceinline__ void hc_insert_par(const CH &ch, const uint8_t *in, uint32_t len, uint32_t from,
                                              uint32_t to, unsigned lane, unsigned warp, S &sm)
{
    uint32_t *lpos = sm.ins_pos[warp];
    uint16_t *lhash = sm.ins_hash[warp];
    for (uint32_t base = from; base < to; base += 128) {
        uint32_t m = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t p = base + 32 * k + lane;
            const bool ok = p < to && p + 3 <= len;
            uint32_t h = 0;
            if (ok) h = hash3(ld24(in + p));
            const bool own = ok && (h & (HC_WARPS - 1)) == warp;
            const unsigned b = __ballot_sync(BDF_FULL_MASK, own);
            if (own) {
                const unsigned at = m + __popc(b & lanemask_lt());
                lpos[at] = p;
                lhash[at] = (uint16_t)h;
            }
            m += __popc(b);
        }
        __syncwarp();
        for (uint32_t s0 = 0; s0 < m; s0 += 32) {
            const bool ok = s0 + lane < m;
            uint32_t p = 0, h = 0x10000u + lane;           // unique key for idle lanes
            if (ok) { p = lpos[s0 + lane]; h = lhash[s0 + lane]; }
            const unsigned peers = __match_any_sync(BDF_FULL_MASK, h);
            const unsigned lower = peers & lanemask_lt();
            const uint32_t peer_pos = __shfl_sync(BDF_FULL_MASK, p, lower ? 31 - __clz(lower) : 0);
            if (ok) {
                const uint32_t prev = lower ? peer_pos : (uint32_t)ch.head[h];
                ch.link[p] = (prev == CH::NOPOS || p - prev > 0xFFFFu) ? 0 : (uint16_t)(p - prev);
            }
            __syncwarp();
            if (ok && (peers >> lane) == 1u) ch.head[h] = (typename CH::head_t)p;
            __syncwarp();
        }
    }
}

// 4 bytes at any alignment from the aligned words that contain them (never touches a word
// without a requested byte, so it stays inside the stream)
__device__ __forceinline__ uint32_t ld32_any(const uint8_t *p)
{
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(p) & 3u);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(p - sh);
    const uint32_t lo = w[0];
    if (sh == 0) return lo;
    return __funnelshift_r(lo, w[1], 8 * sh);
}
__device__ __forceinline__ unsigned prefix_len_bytes(const uint8_t *a, const uint8_t *b, unsigned maxlen)
{
    unsigned n = 0;
    // most candidates differ within the first word: one 4-byte step, then 8 bytes per step (a
    // 258-byte match of periodic data is 33 steps instead of 65)
    if (n + 4 <= maxlen) {
        const uint32_t x = ld32_any(a) ^ ld32_any(b);
        if (x) return (__ffs(x) - 1) >> 3;
        n = 4;
    }
    while (n + 8 <= maxlen) {
        const uint64_t x = ld64_any(a + n) ^ ld64_any(b + n);
        if (x) return n + ((__ffsll((long long)x) - 1) >> 3);
        n += 8;
    }
    if (n + 4 <= maxlen) {
        const uint32_t x = ld32_any(a + n) ^ ld32_any(b + n);
        if (x) return n + ((__ffs(x) - 1) >> 3);
        n += 4;
    }
    while (n < maxlen && a[n] == b[n]) n++;
    return n;
}

// find_match_impl without the insertion (already done): walk the chain of p.
template <class CH>
__device__ __forceinline__ void hc_search(const CH &ch, const uint8_t *in, uint32_t len, uint32_t p,
                                          const HcParams &prm, unsigned &out_len, unsigned &out_off)
{
    out_len = 0; out_off = 0;
    if (p + 3 > len) return;
    const uint32_t first = ch.link[p];
    if (!first) return;
    const bool can4 = p + 4 <= len;
    const uint8_t *src = in + p;
    const uint32_t v3 = ld24(src);
    const uint32_t b3 = can4 ? src[3] : 0;
    const unsigned room = len - p < 258 ? len - p : 258;
    unsigned best = 0, best_off = 0, depth = 0;
    int32_t cur = (int32_t)p - (int32_t)first;
    while (cur >= 0 && depth < prm.max_depth) {
        const uint32_t off = p - (uint32_t)cur;
        if (off > 32768u) break;
        if (p + best >= len) break;
        const uint8_t *m = in + cur;
        if (!(best >= 3 && m[best] != src[best])) {
            const bool eq3 = ld24(m) == v3;
            if (can4) {
                if (eq3 && m[3] == b3) {
                    unsigned l = 4 + prefix_len_bytes(m + 4, src + 4, room - 4);
                    if (l > best) {
                        best = l; best_off = off;
                        if (l >= prm.nice_len || l == 258) break;
                    }
                } else if (best < 3 && eq3) {
                    best = 3; best_off = off;
                }
            } else if (eq3) {
                unsigned l = 3;          // room == 3 here
                if (l > best) {
                    best = l; best_off = off;
                    if (l >= prm.nice_len || l == 258) break;
                }
            }
        }
        // prev_tab is indexed modulo 32768 in the reference: a candidate exactly one
        // window back reads the slot the current position has just overwritten
        const uint32_t lk = off == 32768u ? first : ch.link[cur];
        if (!lk) break;
        cur -= (int32_t)lk;
        depth++;
    }
    out_len = best; out_off = best_off;
}

// BlockSplitStats::should_end_block, src/compress/mod.rs:387-415 (+ :359-384); thread-0 only
template <class S>
__device__ bool hc_should_end(S &sm, uint32_t block_len, uint32_t remaining)
{
    if (sm.num_new < 2048 && block_len < 300000u) return false;
    if (remaining <= 5000u) return false;
    if (block_len >= 300000u) return true;
    if (block_len >= 5000u) {
        if (sm.num_obs != 0) {
            uint32_t old_bits = 0, new_bits = 0;
            const uint32_t lg_all = 31 - __clz(sm.num_obs), lg_new = 31 - __clz(sm.num_new);
            for (int i = 0; i < 14; i++) {
                uint32_t k = sm.new_obs[i];
                if (k) {
                    uint32_t lo = 31 - __clz(sm.obs[i] + 1), ln = 31 - __clz(k + 1);
                    old_bits += k * (lg_all > lo ? lg_all - lo : 0);
                    new_bits += k * (lg_new > ln ? lg_new - ln : 0);
                }
            }
            if ((int32_t)old_bits - (int32_t)new_bits > (int32_t)block_len / 16) return true;
        }
        for (int i = 0; i < 14; i++) { sm.obs[i] += sm.new_obs[i]; sm.new_obs[i] = 0; }
        sm.num_obs += sm.num_new;
        sm.num_new = 0;
    }
    return false;
}

// write_dynamic_huffman_header_impl, src/compress/mod.rs:1775-1883.  Thread 0 computes
// the run-length items and the precode; the caller's warp then packs the bits.
template <class S>
__device__ void hc_prepare_header(S &sm, unsigned &nlit, unsigned &noff, unsigned &npre, unsigned &nitems)
{
    nlit = 288; noff = 32;
    while (nlit > 257 && sm.litlen_len[nlit - 1] == 0) nlit--;
    while (noff > 1 && sm.offset_len[noff - 1] == 0) noff--;
    const unsigned total = nlit + noff;
    for (unsigned i = 0; i < nlit; i++) sm.hdr_lens[i] = sm.litlen_len[i];
    for (unsigned i = 0; i < noff; i++) sm.hdr_lens[nlit + i] = sm.offset_len[i];
    for (int i = 0; i < 19; i++) sm.pre_freq[i] = 0;
    nitems = 0;
    for (unsigned i = 0; i < total;) {
        const unsigned l = sm.hdr_lens[i];
        unsigned run = 1;
        while (i + run < total && sm.hdr_lens[i + run] == l) run++;
        i += run;
        if (l == 0) {
            while (run >= 11) {
                unsigned k = run < 138 ? run : 138;
                sm.hdr_items[nitems++] = (uint16_t)(18u << 8 | (k - 11));
                sm.pre_freq[18]++;
                run -= k;
            }
            if (run >= 3) {
                unsigned k = run < 10 ? run : 10;
                sm.hdr_items[nitems++] = (uint16_t)(17u << 8 | (k - 3));
                sm.pre_freq[17]++;
                run -= k;
            }
        } else if (run >= 4) {
            sm.hdr_items[nitems++] = (uint16_t)(l << 8);
            sm.pre_freq[l]++;
            run--;
            while (run >= 3) {
                unsigned k = run < 6 ? run : 6;
                sm.hdr_items[nitems++] = (uint16_t)(16u << 8 | (k - 3));
                sm.pre_freq[16]++;
                run -= k;
            }
        }
        while (run--) {
            sm.hdr_items[nitems++] = (uint16_t)(l << 8);
            sm.pre_freq[l]++;
        }
    }
    make_huffman_code_serial(19, 7, sm.pre_freq, sm.pre_len, sm.pre_code, sm.scratch);
    const uint8_t perm[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    npre = 19;
    while (npre > 4 && sm.pre_len[perm[npre - 1]] == 0) npre--;
}

template <bool BIG, bool SIZE = false>
__global__ void __launch_bounds__(HC_THREADS) deflate_hc_kernel(DeflateArgs a)
{
    using CH = HcChains<BIG>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HcSmem &sm = *reinterpret_cast<HcSmem *>(smem_raw);
    __shared__ unsigned long long s_idx;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const HcParams prm = hc_params(a.level);
    uint8_t *slab = static_cast<uint8_t *>(a.scratch) + a.scratch_stride * blockIdx.x;
    CH ch;
    ch.head = reinterpret_cast<typename CH::head_t *>(slab);
    ch.link = reinterpret_cast<uint16_t *>(ch.head + 32768);
    uint32_t *syms = reinterpret_cast<uint32_t *>(slab + CH::CHAIN_BYTES);
    if (a.format == BDF_GZIP) load_crc_tables_to_smem(sm.crc, sm.x2n);

    for (;;) {
        __syncthreads();
        if (tid == 0) s_idx = atomicAdd(a.work_counter, 1ull);
        __syncthreads();
        const unsigned long long idx = s_idx;
        if (idx >= a.n) break;
        if (a.klass && a.klass[idx] != a.want) continue;       // the other kernel's stream
        const uint8_t *in = a.in + a.in_off[idx];
        const uint64_t len64 = a.in_off[idx + 1] - a.in_off[idx];
        uint8_t *out = SIZE ? nullptr : a.out + a.out_off[idx];
        const unsigned uflags = unit_flags_of(a, idx);
        if (SIZE && len64 == 0) {             // the estimator's block loop never runs (:808)
            if (tid == 0) { a.status[idx] = BDF_OK; a.out_size[idx] = 0; }
            continue;
        }
        if (len64 > CH::MAX_LEN) {
            if (tid == 0) { a.status[idx] = BDF_STREAM_UNSUPPORTED; a.out_size[idx] = 0; }
            continue;
        }
        const uint32_t len = (uint32_t)len64;
        for (unsigned i = tid; i < 32768 * sizeof(typename CH::head_t) / 16; i += HC_THREADS)
            reinterpret_cast<uint4 *>(ch.head)[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
        unsigned hdr = 0;
        BitSinkT<SIZE> bs;
        if (warp == 0) {
            hdr = SIZE ? 0 : frame_header(a.format, a.level, out, lane);
            bs.init(sm.sink, SIZE ? nullptr : out + hdr, SIZE ? ~0ull : unit_cap(len, uflags), lane);
        }
        __syncthreads();

        // p (parse position) is owned by warp 0 and republished through sm.c_pos after every
        // stage, so that all warps take the same trips through the loops below
        uint32_t p = 0;
        uint32_t ins_end = 0;         // warp 0: positions below this are inserted
        uint32_t window = 32;         // warp 0: positions to search next

        do {
            // ------------------------------------------------ one block
            const uint32_t block_start = p;
            uint32_t nsym = 0;          // warp 0: symbol records of this block
            for (unsigned i = tid; i < 288; i += HC_THREADS) sm.litlen_freq[i] = 0;
            if (tid < 32) sm.offset_freq[tid] = 0;
            if (tid < 14) { sm.new_obs[tid] = 0; sm.obs[tid] = 0; }
            if (tid == 0) { sm.num_new = 0; sm.num_obs = 0; }
            __syncthreads();
            bool block_done = len == 0;
            while (!block_done) {
                // 1. make sure [p, p + window) is inserted, then search it
                // After a long match the next round searches HC_NSPEC small windows instead of one: at p
                // and at p + 258, p + 516, ... — in run-length / periodic data every match has the
                // maximum length, so the parse lands exactly on the next window and one
                // insert / search / parse round (three CTA barriers, ~20 k cycles) yields up to
                // HC_NSPEC matches instead of one.  A window the parse does not land on is ignored;
                // find_match(q) does not depend on what was inserted behind q, so inserting further
                // ahead changes nothing.
                if (warp == 0) {
                    const uint32_t nspec = window == HC_WINDOW_AFTER_JUMP ? HC_NSPEC : 1u;
                    uint32_t to = p + window < len ? p + window : len;
                    uint32_t ito = to;
                    if (nspec > 1) {
                        const uint32_t far = p + 258u * (nspec - 1) + HC_WINDOW_AFTER_JUMP;
                        ito = far < len ? far : len;
                    }
                    if (lane == 0) {
                        sm.c_ins_from = ins_end; sm.c_ins_to = ito > ins_end ? ito : ins_end;
                        sm.c_search_from = p; sm.c_search_to = to; sm.c_nspec = nspec;
                    }
                    if (ito > ins_end) ins_end = ito;
                }
                __syncthreads();
                if (sm.c_ins_to > sm.c_ins_from) hc_insert_par(ch, in, len, sm.c_ins_from, sm.c_ins_to, lane, warp, sm);
                __syncthreads();
                const uint32_t sfrom = sm.c_search_from, sto = sm.c_search_to, nspec = sm.c_nspec;
                if (nspec == 1) {
                    for (uint32_t q = sfrom + tid; q < sto; q += HC_THREADS) {
                        unsigned l, o;
                        hc_search(ch, in, len, q, prm, l, o);
                        sm.mlen[q - sfrom] = (uint16_t)l;
                        sm.moff[q - sfrom] = (uint16_t)o;
                    }
                } else if (tid < nspec * HC_WINDOW_AFTER_JUMP) {
                    // results of window w at indices [4w, 4w + 4)
                    const uint32_t q = sfrom + 258u * (tid / HC_WINDOW_AFTER_JUMP) + (tid % HC_WINDOW_AFTER_JUMP);
                    unsigned l = 0, o = 0;
                    if (q < len) hc_search(ch, in, len, q, prm, l, o);
                    sm.mlen[tid] = (uint16_t)l;
                    sm.moff[tid] = (uint16_t)o;
                }
                __syncthreads();
                // 2. parse by warp 0 (decide_greedy_sequences).  Literal runs are handled 32 positions
                //    at a time; the split check, which only acts once 2048 observations are pending,
                //    is evaluated exactly where the serial loop would act on it (see n_safe below).
                if (warp == 0) {
                    bool jumped = false;
                    for (uint32_t w = 0; w < nspec && !block_done; w++) {
                    // window w: positions [wfrom, wto), results at index (position + woff)
                    const uint32_t wfrom = sfrom + 258u * w;
                    if (p != wfrom) break;                       // the parse did not land on this window
                    const uint32_t wto = nspec == 1 ? sto : (wfrom + HC_WINDOW_AFTER_JUMP < len ? wfrom + HC_WINDOW_AFTER_JUMP : len);
                    const uint32_t woff = HC_WINDOW_AFTER_JUMP * w - wfrom;      // modulo 2^32
                    while (p < len) {
                        // lazy look-ups must stay inside the searched window
                        if (p + prm.lazy >= wto && wto < len) break;
                        bool end_block = false;
                        if (lane == 0) end_block = hc_should_end(sm, p - block_start, len - p);
                        end_block = __shfl_sync(BDF_FULL_MASK, end_block, 0);
                        __syncwarp();
                        if (end_block) { block_done = true; break; }
                        unsigned l = sm.mlen[p + woff], o = sm.moff[p + woff];
                        if (l < 3) {
                            // literals up to the next match start, the window end, or the next position
                            // at which should_end_block could change state
                            const uint32_t pending = sm.num_new;
                            uint32_t n_safe;
                            if (pending < 2048) n_safe = 2048 - pending;
                            else if (len - p <= 5000) n_safe = 0xFFFFFFFFu;
                            else n_safe = 5000 - (p - block_start);     // block_len < 5000 here
                            uint32_t limit = wto - p < n_safe ? wto : p + n_safe;
                            uint32_t n = 0;
                            for (;;) {
                                const uint32_t q = p + n + lane;
                                const bool lit = q < limit && sm.mlen[q + woff] < 3;
                                const unsigned stop = __ballot_sync(BDF_FULL_MASK, !lit);
                                const unsigned take = stop ? __ffs(stop) - 1 : 32;
                                if (lane < take) {
                                    const unsigned b = in[q];
                                    atomicAdd(&sm.litlen_freq[b], 1u);
                                    atomicAdd(&sm.new_obs[b >> 5], 1u);
                                    syms[nsym + n + lane] = b;
                                }
                                n += take;
                                if (take < 32) break;
                            }
                            __syncwarp();
                            if (lane == 0) sm.num_new = pending + n;
                            __syncwarp();
                            nsym += n;
                            p += n;
                            continue;
                        }
                        unsigned nlit = 0;      // literals emitted by a lazy decision
                        if (prm.lazy >= 1 && p + 1 < len && l < prm.nice_len) {
                            const unsigned l1 = sm.mlen[p + 1 + woff];
                            if (l1 > l) {
                                nlit = 1; l = l1; o = sm.moff[p + 1 + woff];
                                if (prm.lazy >= 2 && p + 2 < len) {
                                    const unsigned l2 = sm.mlen[p + 2 + woff];
                                    if (l2 > l1) { nlit = 2; l = l2; o = sm.moff[p + 2 + woff]; }
                                }
                            }
                        }
                        if (lane == 0) {
                            for (unsigned k = 0; k < nlit; k++) {
                                const unsigned b = in[p + k];
                                sm.new_obs[b >> 5]++; sm.num_new++;
                                sm.litlen_freq[b]++;
                                syms[nsym + k] = b;
                            }
                            const unsigned slot = offset_slot_of(o);
                            syms[nsym + nlit] = SYM_LEN | l;
                            syms[nsym + nlit + 1] = SYM_OFF | o;
                            sm.new_obs[8 + (l >= 8)]++;
                            sm.new_obs[10 + (slot < 16 ? 0 : slot < 24 ? 1 : slot < 30 ? 2 : 0)]++;
                            sm.num_new += 2;
                            sm.litlen_freq[257 + length_slot_of(l)]++;
                            sm.offset_freq[slot]++;
                        }
                        __syncwarp();
                        nsym += nlit + 2;
                        p += nlit + l;
                        if (p >= wto) jumped = l >= 32;
                    }
                    }
                    if (p >= len) block_done = true;
                    // after a long match only a few positions are worth searching (the next one
                    // is likely another long match, and each of its neighbours costs a 258-byte
                    // compare that the jump throws away); the window grows back geometrically
                    window = jumped ? HC_WINDOW_AFTER_JUMP : (window * 4 < HC_WINDOW ? window * 4 : HC_WINDOW);
                    if (lane == 0) { sm.c_done = block_done ? 1u : 0u; sm.c_pos = p; }
                }
                __syncthreads();
                block_done = sm.c_done != 0;
                p = sm.c_pos;
            }
            // ------------------------------------------------ block end: codes + emission (warp 0)
            if (warp == 0) {
                if (lane == 0) sm.litlen_freq[256]++;
                __syncwarp();
                __threadfence_block();
                unsigned nlit_syms = 0, noff_syms = 0, npre = 0, nitems = 0;
                if (lane == 0) {
                    make_huffman_code_serial(288, 14, sm.litlen_freq, sm.litlen_len, sm.litlen_code, sm.scratch);
                    make_huffman_code_serial(32, 15, sm.offset_freq, sm.offset_len, sm.offset_code, sm.scratch);
                    hc_prepare_header(sm, nlit_syms, noff_syms, npre, nitems);
                }
                nlit_syms = __shfl_sync(BDF_FULL_MASK, nlit_syms, 0);
                noff_syms = __shfl_sync(BDF_FULL_MASK, noff_syms, 0);
                npre = __shfl_sync(BDF_FULL_MASK, npre, 0);
                nitems = __shfl_sync(BDF_FULL_MASK, nitems, 0);
                __syncwarp();
                const bool is_final = p >= len && (uflags & UNIT_FINISH);
                bs.put1((is_final ? 1u : 0u) | (2u << 1), 3, lane);
                bs.put1((nlit_syms - 257) | ((noff_syms - 1) << 5) | ((npre - 4) << 10), 14, lane);
                {
                    const uint64_t perm = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 |
                                          9ull << 30 | 6ull << 35 | 10ull << 40 | 5ull << 45 | 11ull << 50 | 4ull << 55;
                    const uint64_t perm2 = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 | 1ull << 25 | 15ull << 30;
                    unsigned sym = lane < 12 ? (unsigned)(perm >> (5 * lane)) & 31u
                                 : lane < 19 ? (unsigned)(perm2 >> (5 * (lane - 12))) & 31u : 0u;
                    bs.put(lane < npre ? sm.pre_len[sym] : 0, lane < npre ? 3 : 0, lane);
                }
                for (unsigned base = 0; base < nitems; base += 32) {
                    uint32_t bits = 0, nb = 0;
                    if (base + lane < nitems) {
                        const unsigned it = sm.hdr_items[base + lane], sym = it >> 8, extra = it & 0xFF;
                        const unsigned cl = sm.pre_len[sym];
                        bits = sm.pre_code[sym] | (extra << cl);
                        nb = cl + (sym == 16 ? 2 : sym == 17 ? 3 : sym == 18 ? 7 : 0);
                    }
                    bs.put(bits, nb, lane);
                }
                // symbols, 32 records per step: literal / length / offset records in stream order
                for (uint32_t base = 0; base < nsym; base += 32) {
                    uint32_t bits = 0, nb = 0;
                    if (base + lane < nsym) {
                        const uint32_t rec = syms[base + lane], v = rec & 0xFFFFu;
                        if (rec < SYM_LEN) {
                            bits = sm.litlen_code[v]; nb = sm.litlen_len[v];
                        } else if (rec < SYM_OFF) {
                            unsigned slot = length_slot_of(v), lb, le;
                            length_slot_info(slot, lb, le);
                            const unsigned cl = sm.litlen_len[257 + slot];
                            bits = sm.litlen_code[257 + slot] | ((v - lb) << cl);
                            nb = cl + le;
                        } else {
                            unsigned slot = offset_slot_of(v), ob, oe;
                            offset_slot_info(slot, ob, oe);
                            const unsigned cl = sm.offset_len[slot];
                            bits = sm.offset_code[slot] | ((v - ob) << cl);
                            nb = cl + oe;
                        }
                    }
                    bs.put(bits, nb, lane);
                }
                bs.put1(sm.litlen_code[256], sm.litlen_len[256], lane);
            }
            __syncthreads();
        } while (p < len);
        if (warp == 0) {
            if (uflags & UNIT_SYNC) bs.sync_marker(lane);
            uint64_t sz = bs.finish(lane);
            int st = BDF_OK;
            if (sz == ~0ull) { st = BDF_INSUFFICIENT_SPACE; sz = 0; }
            else if (!SIZE) sz = frame_footer(a.format, in, len, out, hdr + sz, sm.crc, sm.x2n, lane);
            if (lane == 0) { a.status[idx] = st; a.out_size[idx] = sz; }
        }
    }
}

}  // namespace bdf
