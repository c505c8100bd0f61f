// deflate_nos_split.cuh — levels 10..12 for streams of at most 64 KiB as THREE kernels per wave of streams.
//
// deflate_nos_kernel (deflate_hcs.cuh) runs the whole near-optimal pipeline of a stream inside one
// CTA: match lists from the shared-memory chains (32 warps), greedy pass for block ends and symbol
// costs (32 warps), backward cost pass (ONE warp — the recurrence is serial: cost[p] needs
// cost[p + 1 .. p + 258]), parallel parse and encoding (32 warps).  With one CTA per SM the cost pass
// leaves 31 warps of every SM waiting: 31 M cycles per 64 KiB stream, ~95 % of the stream's time on
// run-length / periodic data (bench config `compress_l12`) and about half on text.
// The cost pass needs nothing that lives in shared memory — match lists and results are in global
// memory, the cost tables are 548 bytes — so it does not have to sit inside the CTA that owns the
// chains.  Here a wave of streams goes through
//   A  deflate_nos_search_kernel  one CTA per stream: chains, match lists, greedy pass for EVERY block
//                                 of the stream -> block ends + cost tables (global);
//   B  deflate_nos_cost_kernel    one WARP per stream, 8 warps per SM: the backward cost pass of all
//                                 blocks, match lists prefetched four steps ahead -> choices (global);
//   C  deflate_nos_emit_kernel    one CTA per stream: the parse along the choices, codes, bit packing.
// so the serial pass runs for 8 x 148 streams at once instead of 148.  Same functions, same parse
// and therefore the same bytes as deflate_nos_kernel (tests/test_gpu_determinism.py runs both);
// BDF_NOS_SPLIT=0 selects the single kernel.
#pragma once
#include "deflate_hcs.cuh"

namespace bdf {

constexpr uint32_t NOS_MAX_BLOCKS = 20;       // BlockSplitStats never cuts a block below 5000 bytes: at most 13 per 64 KiB
struct NosBlock {
    uint32_t end;                             // the block is [previous end, end)
    uint8_t lit_cost[256], len_cost[260], slot_cost[32];
};
struct NosPlan {                              // per stream, written by kernel A
    uint32_t nblocks;                         // 0: nothing to do (unsupported length)
    uint32_t len;
    NosBlock blk[NOS_MAX_BLOCKS];
};
constexpr size_t NOS_PLAN_BYTES = (sizeof(NosPlan) + 255) & ~(size_t)255;
constexpr size_t NOS_SPLIT_PER_STREAM = NOS_SCRATCH_PER_CTA + NOS_PLAN_BYTES;

struct NosWave {
    uint32_t first, count;                    // streams [first, first + count) of the batch
};
struct NosSlab {
    uint32_t *recs, *gbest, *choice, *lists;
    uint16_t *near3;
    NosPlan *plan;
    __device__ __forceinline__ NosSlab(const DeflateArgs &a, uint32_t slot)
    {
        uint8_t *slab = static_cast<uint8_t *>(a.scratch) + a.scratch_stride * slot;
        recs = reinterpret_cast<uint32_t *>(slab);
        gbest = reinterpret_cast<uint32_t *>(slab + HCS_SCRATCH_PER_CTA);
        choice = gbest + 65536;
        lists = choice + 65536;
        near3 = reinterpret_cast<uint16_t *>(lists + (size_t)65536 * HCS_NLIST);
        plan = reinterpret_cast<NosPlan *>(slab + NOS_SCRATCH_PER_CTA);
    }
};

// ---- A: match lists and the greedy pass (pass 1 of deflate_nos_kernel) for every block
__global__ void __launch_bounds__(HCS_THREADS, 1) deflate_nos_search_kernel(DeflateArgs a, NosWave wv)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HcsSmem &sm = *reinterpret_cast<HcsSmem *>(smem_raw);
    const unsigned tid = threadIdx.x;
    const HcParams prm = nos_params(a.level, a.nos_depth);
    for (uint32_t slot = blockIdx.x; slot < wv.count; slot += gridDim.x) {
        __syncthreads();
        const unsigned long long idx = (unsigned long long)wv.first + slot;
        NosSlab sl(a, slot);
        const uint8_t *gin = a.in + a.in_off[idx];
        const uint64_t len64 = a.in_off[idx + 1] - a.in_off[idx];
        if (len64 > 65536) {
            if (tid == 0) { sl.plan->nblocks = 0; sl.plan->len = 0; a.status[idx] = BDF_STREAM_UNSUPPORTED; a.out_size[idx] = 0; }
            continue;
        }
        const uint32_t len = (uint32_t)len64;
        HcsStream st;
        st.set(gin, len);
        const uint16_t *near3 = nullptr;
#ifdef BDF_NOS_H4
        nos_build_chains_h4(sm, st, sl.near3);
        near3 = sl.near3;
#else
        hcs_build_chains(sm, st);
#endif
        hcs_stage_input(sm, st);
        if (tid == 0) sm.c_search_next = 0;
        __syncthreads();
        for (uint32_t pos0 = 0; pos0 < len; pos0 += HCS_SEARCH) {
            const uint32_t ns = len - pos0 < HCS_SEARCH ? len - pos0 : HCS_SEARCH;
            hcs_search<true>(sm, len, prm, pos0, ns, sl.lists, near3);
            __syncthreads();
            for (uint32_t i = tid; i < ns; i += HCS_THREADS) sl.gbest[pos0 + i] = sm.w.res[i];
            if (tid == 0) sm.c_search_next = 0;
            __syncthreads();
        }
        uint32_t block_start = 0, nb = 0;
        do {
            for (uint32_t i = tid; i < 288; i += HCS_THREADS) sm.litlen_freq[i] = 0;
            if (tid < 32) sm.offset_freq[tid] = 0;
            if (tid < 14) { sm.new_obs[tid] = 0; sm.obs[tid] = 0; }
            if (tid == 0) { sm.num_new = 0; sm.num_obs = 0; }
            __syncthreads();
            uint32_t block_end = len;
            for (uint32_t entry = block_start; entry < len; ) {
                const uint32_t wvalid = len - entry < HCS_W ? len - entry : HCS_W;
                const uint32_t nload = len - entry < HCS_SEARCH ? len - entry : HCS_SEARCH;
                for (uint32_t i = tid; i < nload; i += HCS_THREADS) sm.w.res[i] = sl.gbest[entry + i];
                __syncthreads();
                hcs_steps_greedy(sm, len, prm, entry, wvalid);
                hcs_parse_window(sm, len, entry, wvalid, block_start, 0, sl.recs, true);
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
            // the last slot takes whatever is left (cannot happen with blocks of at least 5000 bytes)
            if (nb == NOS_MAX_BLOCKS - 1) block_end = len;
            // symbol costs from the greedy histograms (update_costs, src/compress/mod.rs:2209-2224)
            if (tid == 0) sm.litlen_freq[256]++;
            __syncthreads();
            make_huffman_code_cta(288, 14, sm.litlen_freq, sm.litlen_len, sm.enc.litlen_code, sm.enc.scratch);
            make_huffman_code_cta(32, 15, sm.offset_freq, sm.offset_len, sm.enc.offset_code, sm.enc.scratch);
            NosBlock &B = sl.plan->blk[nb];
            for (uint32_t i = tid; i < 256 + 260 + 32; i += HCS_THREADS) {
                if (i < 256) B.lit_cost[i] = sm.litlen_len[i] ? sm.litlen_len[i] : 12;
                else if (i < 256 + 260) {
                    const uint32_t l = i - 256;
                    uint32_t c = 0;
                    if (l >= 3 && l <= 258) {
                        unsigned slot_ = length_slot_of(l), lb, le;
                        length_slot_info(slot_, lb, le);
                        c = (sm.litlen_len[257 + slot_] ? sm.litlen_len[257 + slot_] : 10) + le;
                    }
                    B.len_cost[l] = (uint8_t)c;
                } else {
                    const uint32_t slot_ = i - 516;
                    unsigned ob_, oe;
                    offset_slot_info(slot_ < 30 ? slot_ : 29, ob_, oe);
                    B.slot_cost[slot_] = (uint8_t)((sm.offset_len[slot_] ? sm.offset_len[slot_] : 8) + oe);
                }
            }
            if (tid == 0) B.end = block_end;
            __syncthreads();
            nb++;
            block_start = block_end;
        } while (block_start < len);
        if (tid == 0) { sl.plan->nblocks = nb; sl.plan->len = len; }
    }
}

// ---- B: the backward cost pass (pass 2 of deflate_nos_kernel), one warp per stream
constexpr int NOS_COST_WARPS = 4;             // per CTA
struct NosCostSmem {
    uint32_t ring[512];
    uint8_t lit_cost[256], len_cost[260], slot_cost[32];
    uint32_t pad;
};
__global__ void __launch_bounds__(NOS_COST_WARPS * 32) deflate_nos_cost_kernel(DeflateArgs a, NosWave wv)
{
    __shared__ NosCostSmem s_all[NOS_COST_WARPS];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    NosCostSmem &S = s_all[warp];
    const uint32_t slot = blockIdx.x * NOS_COST_WARPS + warp;
    if (slot >= wv.count) return;
    const unsigned long long idx = (unsigned long long)wv.first + slot;
    NosSlab sl(a, slot);
    const uint32_t nblocks = sl.plan->nblocks;
    const uint8_t *gin = a.in + a.in_off[idx];
    const uint32_t g = lane >> 3, k = lane & 7u;
    uint32_t block_start = 0;
    for (uint32_t b = 0; b < nblocks; b++) {
        const NosBlock &B = sl.plan->blk[b];
        const uint32_t block_end = B.end;
        __syncwarp();
        for (uint32_t i = lane; i < 256 + 260 + 32; i += 32) {
            if (i < 256) S.lit_cost[i] = B.lit_cost[i];
            else if (i < 516) S.len_cost[i - 256] = B.len_cost[i - 256];
            else S.slot_cost[i - 516] = B.slot_cost[i - 516];
        }
        if (lane == 0) S.ring[block_end & 511u] = 0;
        __syncwarp();
        // what a step needs from global memory does not depend on the costs: this lane's list entry (groups
        // 0..2: positions hi-1, hi-2, hi-3, entry k) or literal byte (group 3, k < 3), fetched PF steps ahead
        constexpr int PF = 4;
        auto fetch = [&](uint32_t hi) -> uint32_t {
            if (hi <= block_start) return 0u;
            if (g < 3) return hi >= block_start + 1 + g ? __ldg(sl.lists + (size_t)(hi - 1 - g) * HCS_NLIST + k) : 0u;
            return (k < 3 && hi >= block_start + 1 + k) ? (uint32_t)__ldg(gin + (hi - 1 - k)) : 0u;
        };
        uint32_t q[PF];
#pragma unroll
        for (int t = 0; t < PF; t++) q[t] = fetch(block_end > 3u * t ? block_end - 3u * t : 0u);
        for (uint32_t hi = block_end; hi > block_start; hi = hi - block_start > 3 ? hi - 3 : block_start) {
            const uint32_t cur = q[0];
#pragma unroll
            for (int t = 0; t + 1 < PF; t++) q[t] = q[t + 1];
            q[PF - 1] = fetch(hi > 3u * PF ? hi - 3u * PF : 0u);
            const uint32_t p = hi - 1 - g;                  // this lane's position (groups 0..2)
            const bool mine = g < 3 && hi >= block_start + 1 + g;
            const uint32_t e = mine ? cur : 0u;
            uint32_t litc = 0;
            if (g == 3 && k < 3 && hi >= block_start + 1 + k) litc = S.lit_cost[cur & 0xFFu];
            if (!__any_sync(BDF_FULL_MASK, e != 0)) {
                // no match starts at any of the three positions: literals only
                const uint32_t l0 = __shfl_sync(BDF_FULL_MASK, litc, 24), l1 = __shfl_sync(BDF_FULL_MASK, litc, 25),
                               l2 = __shfl_sync(BDF_FULL_MASK, litc, 26);
                if (lane == 0) {
                    uint32_t c = S.ring[hi & 511u];
                    const uint32_t ll[3] = {l0, l1, l2};
#pragma unroll
                    for (int j = 0; j < 3; j++) {
                        if (hi >= block_start + 1u + j) {
                            const uint32_t qq = hi - 1 - j;
                            c += ll[j];
                            S.ring[qq & 511u] = c;
                            sl.choice[qq] = 1u;
                        }
                    }
                }
                __syncwarp();
                continue;
            }
            const uint32_t e_prev = __shfl_up_sync(BDF_FULL_MASK, e, 1);
            uint32_t v = NOS_INF << 4, vlen = 0;
            if (e) {
                const uint32_t l = e & 0xFFFFu, off = e >> 16;
                const uint32_t lmin0 = k ? (e_prev & 0xFFFFu) + 1u : 3u;
                const uint32_t lmin = l > lmin0 + NOS_SHORTER ? l - NOS_SHORTER : lmin0;
                const uint32_t sc = S.slot_cost[offset_slot_of(off)];
#pragma unroll
                for (uint32_t d = 0; d <= NOS_SHORTER; d++) {
                    const uint32_t t = l - d;
                    if (t >= lmin && t <= l && p + t <= block_end) {
                        const uint32_t c = (S.len_cost[t] + sc + S.ring[(p + t) & 511u]) << 4 | k;
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
            const uint32_t ew = (e & 0xFFFF0000u) | vlen;
            const uint32_t w0 = __shfl_sync(BDF_FULL_MASK, ew, m0 & 7u), w1 = __shfl_sync(BDF_FULL_MASK, ew, 8 + (m1 & 7u)),
                           w2 = __shfl_sync(BDF_FULL_MASK, ew, 16 + (m2 & 7u));
            const uint32_t l0 = __shfl_sync(BDF_FULL_MASK, litc, 24), l1 = __shfl_sync(BDF_FULL_MASK, litc, 25),
                           l2 = __shfl_sync(BDF_FULL_MASK, litc, 26);
            if (lane == 0) {
                uint32_t c = S.ring[hi & 511u];
                const uint32_t mm[3] = {m0, m1, m2}, ww[3] = {w0, w1, w2}, ll[3] = {l0, l1, l2};
#pragma unroll
                for (int j = 0; j < 3; j++) {
                    if (hi >= block_start + 1u + j) {
                        const uint32_t qq = hi - 1 - j;
                        const uint32_t lit = c + ll[j];
                        const uint32_t mc = mm[j] >> 4;
                        uint32_t ch = 1u;
                        c = lit;
                        if (mc < lit) { c = mc; ch = ww[j]; }
                        BDF_ASSERT(qq < 65536 && qq + (ch & 0xFFFFu) <= block_end);
                        S.ring[qq & 511u] = c;
                        sl.choice[qq] = ch;
                    }
                }
            }
            __syncwarp();
        }
        block_start = block_end;
    }
}

// ---- C: the parse along the choices (pass 3 of deflate_nos_kernel) and the encoding, block by block
__global__ void __launch_bounds__(HCS_THREADS, 1) deflate_nos_emit_kernel(DeflateArgs a, NosWave wv)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    HcsSmem &sm = *reinterpret_cast<HcsSmem *>(smem_raw);
    const unsigned tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    for (uint32_t slot = blockIdx.x; slot < wv.count; slot += gridDim.x) {
        __syncthreads();
        const unsigned long long idx = (unsigned long long)wv.first + slot;
        NosSlab sl(a, slot);
        const uint32_t nblocks = sl.plan->nblocks;
        if (nblocks == 0) continue;
        const uint32_t len = sl.plan->len;
        const uint8_t *gin = a.in + a.in_off[idx];
        uint8_t *out = a.out + a.out_off[idx];
        const unsigned uflags = unit_flags_of(a, idx);
        HcsStream st;
        st.set(gin, len);
        hcs_stage_input(sm, st);
        CtaSink<false> sink;
        if (warp == 0) frame_header(a.format, a.level, out, lane);
        const unsigned hdr = a.format == BDF_ZLIB ? 2u : a.format == BDF_GZIP ? 10u : 0u;
        sink.init(sm, out + hdr, unit_cap(len, uflags));
        __syncthreads();
        uint32_t block_start = 0;
        for (uint32_t b = 0; b < nblocks; b++) {
            const uint32_t block_end = sl.plan->blk[b].end;
            for (uint32_t i = tid; i < 288; i += HCS_THREADS) sm.litlen_freq[i] = 0;
            if (tid < 32) sm.offset_freq[tid] = 0;
            __syncthreads();
            uint32_t nrec = 0;
            for (uint32_t entry = block_start; entry < block_end; ) {
                const uint32_t wvalid = block_end - entry < HCS_W ? block_end - entry : HCS_W;
                for (uint32_t i = tid; i < wvalid; i += HCS_THREADS) {
                    const uint32_t ch = sl.choice[entry + i], l = ch & 0xFFFFu;
                    sm.w.res[i] = ch;
                    sm.w.nxt[i] = (uint16_t)(l == 1 ? 1u : (l | 1u << 11));
                }
                __syncthreads();
                hcs_parse_window(sm, block_end, entry, wvalid, block_start, nrec, sl.recs, false);
                nrec += sm.c_win_rec;
                const uint32_t next_entry = entry + sm.c_next_entry;
                __syncthreads();
                entry = next_entry;
            }
            const bool is_final = block_end >= len && (uflags & UNIT_FINISH);
            hcs_encode_block<false>(sm, sink, sl.recs, 0, nrec, is_final);
            block_start = block_end;
        }
        hcs_finish_stream<false>(sm, a, idx, gin, len, out, hdr, uflags);
    }
}

}  // namespace bdf
