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

constexpr int HC_THREADS = 256;
constexpr int HC_WARPS = HC_THREADS / 32;
constexpr uint32_t HC_MAX_LEN = 65536;
constexpr uint32_t HC_WINDOW = 256;           // positions searched per round (max)
constexpr uint32_t HC_NOPOS = 0xFFFFu;

struct HcSeq {
    uint32_t litrun;
    uint16_t len;       // 0 terminates the block
    uint16_t off;
};

struct __align__(16) HcSmem {
    uint16_t head[32768];              // bucket -> most recent position, 0xFFFF = empty
    uint16_t link[65536];              // position -> distance to previous same-hash position, 0 = none
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
    uint32_t c_search_from, c_search_to, c_done, c_pos;
    uint8_t hdr_lens[320];
    uint16_t hdr_items[320];
    uint32_t pre_freq[19], pre_code[19];
    uint8_t pre_len[19];
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
__device__ __forceinline__ void hc_insert_range(HcSmem &sm, const uint8_t *in, uint32_t len, uint32_t from,
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
            uint32_t prev = lower ? base + (31 - __clz(lower)) : sm.head[h];
            sm.link[p] = prev == HC_NOPOS ? 0 : (uint16_t)(p - prev);   // distance <= 65535 always fits
        }
        __syncwarp();
        if (ok && (peers >> lane) == 1u) sm.head[h] = (uint16_t)p;
        __syncwarp();
    }
}

__device__ __forceinline__ unsigned prefix_len_bytes(const uint8_t *a, const uint8_t *b, unsigned maxlen)
{
    unsigned n = 0;
    while (n < maxlen && a[n] == b[n]) n++;
    return n;
}

// find_match_impl without the insertion (already done): walk the chain of p.
__device__ __forceinline__ void hc_search(const HcSmem &sm, const uint8_t *in, uint32_t len, uint32_t p,
                                          const HcParams &prm, unsigned &out_len, unsigned &out_off)
{
    out_len = 0; out_off = 0;
    if (p + 3 > len) return;
    const uint32_t first = sm.link[p];
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
        const uint32_t lk = off == 32768u ? first : sm.link[cur];
        if (!lk) break;
        cur -= (int32_t)lk;
        depth++;
    }
    out_len = best; out_off = best_off;
}

// BlockSplitStats::should_end_block, src/compress/mod.rs:387-415 (+ :359-384); thread-0 only
__device__ bool hc_should_end(HcSmem &sm, uint32_t block_len, uint32_t remaining)
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
__device__ void hc_prepare_header(HcSmem &sm, unsigned &nlit, unsigned &noff, unsigned &npre, unsigned &nitems)
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

__global__ void __launch_bounds__(HC_THREADS, 1) deflate_hc_kernel(DeflateArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HcSmem &sm = *reinterpret_cast<HcSmem *>(smem_raw);
    __shared__ unsigned long long s_idx;
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const HcParams prm = hc_params(a.level);
    HcSeq *seqs = reinterpret_cast<HcSeq *>(static_cast<uint8_t *>(a.scratch) + a.scratch_stride * blockIdx.x);
    if (a.format == BDF_GZIP) load_crc_tables_to_smem(sm.crc, sm.x2n);

    for (;;) {
        __syncthreads();
        if (tid == 0) s_idx = atomicAdd(a.work_counter, 1ull);
        __syncthreads();
        const unsigned long long idx = s_idx;
        if (idx >= a.n) break;
        const uint8_t *in = a.in + a.in_off[idx];
        const uint64_t len64 = a.in_off[idx + 1] - a.in_off[idx];
        uint8_t *out = a.out + a.out_off[idx];
        if (len64 > HC_MAX_LEN) {
            if (tid == 0) { a.status[idx] = BDF_STREAM_UNSUPPORTED; a.out_size[idx] = 0; }
            continue;
        }
        const uint32_t len = (uint32_t)len64;
        for (unsigned i = tid; i < 32768 / 2; i += HC_THREADS) reinterpret_cast<uint32_t *>(sm.head)[i] = 0xFFFFFFFFu;
        unsigned hdr = 0;
        BitSink bs;
        if (warp == 0) {
            hdr = frame_header(a.format, a.level, out, lane);
            bs.init(sm.sink, out + hdr, deflate_bound(len), lane);
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
            uint32_t nseq = 0, litrun = 0;
            for (unsigned i = tid; i < 288; i += HC_THREADS) sm.litlen_freq[i] = 0;
            if (tid < 32) sm.offset_freq[tid] = 0;
            if (tid < 14) { sm.new_obs[tid] = 0; sm.obs[tid] = 0; }
            if (tid == 0) { sm.num_new = 0; sm.num_obs = 0; }
            __syncthreads();
            bool block_done = len == 0;
            while (!block_done) {
                // 1. make sure [p, p + window) is inserted, then search it
                if (warp == 0) {
                    uint32_t to = p + window < len ? p + window : len;
                    if (to > ins_end) { hc_insert_range(sm, in, len, ins_end, to, lane); ins_end = to; }
                    if (lane == 0) { sm.c_search_from = p; sm.c_search_to = to; }
                }
                __syncthreads();
                const uint32_t sfrom = sm.c_search_from, sto = sm.c_search_to;
                for (uint32_t q = sfrom + tid; q < sto; q += HC_THREADS) {
                    unsigned l, o;
                    hc_search(sm, in, len, q, prm, l, o);
                    sm.mlen[q - sfrom] = (uint16_t)l;
                    sm.moff[q - sfrom] = (uint16_t)o;
                }
                __syncthreads();
                // 2. serial parse by warp 0 (decide_greedy_sequences)
                if (warp == 0) {
                    bool jumped = false;
                    while (p < len) {
                        // lazy look-ups must stay inside the searched window
                        if (p + prm.lazy >= sto && sto < len) break;
                        bool end_block = false;
                        if (lane == 0) end_block = hc_should_end(sm, p - block_start, len - p);
                        end_block = __shfl_sync(BDF_FULL_MASK, end_block, 0);
                        if (end_block) { block_done = true; break; }
                        unsigned l = sm.mlen[p - sfrom], o = sm.moff[p - sfrom];
                        if (l >= 3) {
                            unsigned nlit = 0;      // literals emitted by a lazy decision
                            if (prm.lazy >= 1 && p + 1 < len && l < prm.nice_len) {
                                const unsigned l1 = sm.mlen[p + 1 - sfrom];
                                if (l1 > l) {
                                    nlit = 1; l = l1; o = sm.moff[p + 1 - sfrom];
                                    if (prm.lazy >= 2 && p + 2 < len) {
                                        const unsigned l2 = sm.mlen[p + 2 - sfrom];
                                        if (l2 > l1) { nlit = 2; l = l2; o = sm.moff[p + 2 - sfrom]; }
                                    }
                                }
                            }
                            if (lane == 0) {
                                for (unsigned k = 0; k < nlit; k++) {
                                    const unsigned b = in[p + k];
                                    sm.new_obs[b >> 5]++; sm.num_new++;
                                    sm.litlen_freq[b]++;
                                }
                                const unsigned slot = offset_slot_of(o);
                                seqs[nseq].litrun = litrun + nlit;
                                seqs[nseq].len = (uint16_t)l;
                                seqs[nseq].off = (uint16_t)o;
                                sm.new_obs[8 + (l >= 8)]++;
                                sm.new_obs[10 + (slot < 16 ? 0 : slot < 24 ? 1 : slot < 30 ? 2 : 0)]++;
                                sm.num_new += 2;
                                sm.litlen_freq[257 + length_slot_of(l)]++;
                                sm.offset_freq[slot]++;
                            }
                            nseq++;
                            litrun = 0;
                            p += nlit + l;
                            if (p >= sto) jumped = l >= 32;
                        } else {
                            if (lane == 0) {
                                const unsigned b = in[p];
                                sm.new_obs[b >> 5]++; sm.num_new++;
                                sm.litlen_freq[b]++;
                            }
                            litrun++;
                            p++;
                        }
                    }
                    if (p >= len) block_done = true;
                    window = jumped ? 32 : HC_WINDOW;
                    if (lane == 0) { sm.c_done = block_done ? 1u : 0u; sm.c_pos = p; }
                }
                __syncthreads();
                block_done = sm.c_done != 0;
                p = sm.c_pos;
            }
            // ------------------------------------------------ block end: codes + emission (warp 0)
            if (warp == 0) {
                if (lane == 0) {
                    seqs[nseq].litrun = litrun; seqs[nseq].len = 0; seqs[nseq].off = 0;
                    sm.litlen_freq[256]++;
                }
                nseq++;
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
                const bool is_final = p >= len;
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
                // symbols: literals 32 at a time, then the match as two items
                uint32_t at = block_start;
                for (uint32_t s = 0; s < nseq; s++) {
                    const HcSeq sq = seqs[s];
                    for (uint32_t base = 0; base < sq.litrun; base += 32) {
                        uint32_t bits = 0, nb = 0;
                        if (base + lane < sq.litrun) {
                            const unsigned b = in[at + base + lane];
                            bits = sm.litlen_code[b]; nb = sm.litlen_len[b];
                        }
                        bs.put(bits, nb, lane);
                    }
                    at += sq.litrun;
                    if (sq.len >= 3) {
                        uint32_t bits = 0, nb = 0;
                        if (lane == 0) {
                            unsigned slot = length_slot_of(sq.len), base, extra;
                            length_slot_info(slot, base, extra);
                            const unsigned cl = sm.litlen_len[257 + slot];
                            bits = sm.litlen_code[257 + slot] | ((sq.len - base) << cl);
                            nb = cl + extra;
                        } else if (lane == 1) {
                            unsigned slot = offset_slot_of(sq.off), base, extra;
                            offset_slot_info(slot, base, extra);
                            const unsigned cl = sm.offset_len[slot];
                            bits = sm.offset_code[slot] | ((sq.off - base) << cl);
                            nb = cl + extra;
                        }
                        bs.put(bits, nb, lane);
                        at += sq.len;
                    }
                }
                bs.put1(sm.litlen_code[256], sm.litlen_len[256], lane);
            }
            __syncthreads();
        } while (p < len);
        if (warp == 0) {
            uint64_t sz = bs.finish(lane);
            int st = BDF_OK;
            if (sz == ~0ull) { st = BDF_INSUFFICIENT_SPACE; sz = 0; }
            else sz = frame_footer(a.format, in, len, out, hdr + sz, sm.crc, sm.x2n, lane);
            if (lane == 0) { a.status[idx] = st; a.out_size[idx] = sz; }
        }
    }
}

}  // namespace bdf
