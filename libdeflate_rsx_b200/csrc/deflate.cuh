// deflate.cuh — batch DEFLATE compression: level dispatch (sm_100a).
//
// Replaces Compressor::compress / compress_zlib / compress_gzip as called from
// BatchCompressor::compress_batch (reference src/batch.rs:20-58,
// src/compress/mod.rs:693-790,2248-2357).
//   level 0      deflate_stored_kernel   (this file)
//   level 1      deflate_l1_kernel       (deflate_l1.cuh)
//   levels 2..9  deflate_hc_kernel       (deflate_hc.cuh)
//   levels 10..12 deflate_bt_kernel    (deflate_bt.cuh)
#pragma once
#include "deflate_common.cuh"
#include "deflate_bt.cuh"
#include "deflate_hc.cuh"
#include "deflate_hcs.cuh"
#include "deflate_nos_split.cuh"
#include "deflate_l1.cuh"
#include "gather.cuh"

namespace bdf {

struct DeflateScratch {
    void *p = nullptr;
    size_t cap = 0;
    bool l1_ready = false, hc_ready = false, hcs_ready = false, nos_ready = false, nos_split_ready = false;
};
inline void deflate_scratch_free(DeflateScratch &s)
{
    if (s.p) cudaFree(s.p);
    s.p = nullptr;
    s.cap = 0;
}

// ---- level 0: stored blocks (compress_uncompressed, src/compress/mod.rs:1400-1464)
constexpr int L0_WARPS_PER_BLOCK = 8;
__global__ void __launch_bounds__(L0_WARPS_PER_BLOCK * 32) deflate_stored_kernel(DeflateArgs a)
{
    __shared__ uint32_t s_crc[4][256];
    __shared__ uint32_t s_x2n[32];
    const unsigned lane = lane_id();
    if (a.format == BDF_GZIP) load_crc_tables_to_smem(s_crc, s_x2n);
    const uint32_t warps = gridDim.x * L0_WARPS_PER_BLOCK;
    for (uint32_t idx = blockIdx.x * L0_WARPS_PER_BLOCK + (threadIdx.x >> 5); idx < a.n; idx += warps) {
        const uint8_t *in = a.in + a.in_off[idx];
        const uint64_t len = a.in_off[idx + 1] - a.in_off[idx];
        if (a.size_only) {
            // compress_to_size at level 0, src/compress/mod.rs:1073-1082
            const uint64_t blocks = len / 65535 + ((len % 65535 != 0 || (len == 0 && a.final_block)) ? 1 : 0);
            if (lane == 0) { a.out_size[idx] = len + blocks * 5; a.status[idx] = BDF_OK; }
            continue;
        }
        uint8_t *out = a.out + a.out_off[idx];
        const unsigned uflags = unit_flags_of(a, idx);
        uint64_t op = frame_header(a.format, 0, out, lane);
        // an empty input produces ZERO deflate bytes at level 0 (the block loop never runs, :1408)
        for (uint64_t ip = 0; ip < len;) {
            uint64_t blk = len - ip < 65535 ? len - ip : 65535;
            unsigned bfinal = ip + 65535 >= len && (uflags & UNIT_FINISH);
            if (lane < 5) {
                uint8_t b = lane == 0 ? (uint8_t)bfinal
                          : lane == 1 ? (uint8_t)blk
                          : lane == 2 ? (uint8_t)(blk >> 8)
                          : lane == 3 ? (uint8_t)~blk
                                      : (uint8_t)(~blk >> 8);
                out[op + lane] = b;
            }
            op += 5;
            for (uint64_t i = lane; i < blk; i += 32) out[op + i] = in[ip + i];
            op += blk;
            ip += blk;
        }
        if (uflags & UNIT_SYNC) {
            if (lane < 5) out[op + lane] = lane >= 3 ? 0xFF : 0;
            op += 5;
        }
        op = frame_footer(a.format, in, len, out, op, s_crc, s_x2n, lane);
        if (lane == 0) { a.out_size[idx] = op; a.status[idx] = BDF_OK; }
    }
}


// ---- which hash-chain kernel a stream goes to (levels 2..9, at most 64 KiB).  deflate_hcs keeps a
// whole stream in shared memory (one stream per SM) and searches every position; data that is one
// long run or a short period — where nearly every step of the parse is a maximal match — is far
// cheaper in deflate_hc, which searches only where the parse lands.  One warp per stream looks at
// four places: the 8 bytes there must occur again within the 512 bytes in front of them, and the
// 64 bytes there must repeat at that distance.  Three hits of four make the stream "runny".
constexpr int CLASSIFY_WARPS = 8;
__global__ void __launch_bounds__(CLASSIFY_WARPS * 32) deflate_classify_kernel(const uint8_t *in, const uint64_t *in_off,
                                                                                uint32_t n, uint8_t *klass)
{
    const unsigned lane = lane_id();
    for (uint32_t idx = blockIdx.x * CLASSIFY_WARPS + (threadIdx.x >> 5); idx < n; idx += gridDim.x * CLASSIFY_WARPS) {
        const uint8_t *d = in + in_off[idx];
        const uint64_t len = in_off[idx + 1] - in_off[idx];
        unsigned hits = 0;
        if (len >= 4096) {
            for (unsigned k = 1; k <= 4; k++) {
                const uint64_t p = len / 5 * k;                 // >= 819: 512 bytes in front, 64 behind
                unsigned long long pat = 0;
                for (int b = 0; b < 8; b++) pat |= (unsigned long long)d[p + b] << (8 * b);
                unsigned best = 0xFFFFu;
                for (unsigned o = 1 + lane; o <= 512; o += 32) {
                    unsigned long long v = 0;
                    for (int b = 0; b < 8; b++) v |= (unsigned long long)d[p - o + b] << (8 * b);
                    if (v == pat && o < best) best = o;
                }
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) {
                    const unsigned t = __shfl_xor_sync(BDF_FULL_MASK, best, s);
                    best = t < best ? t : best;
                }
                if (best != 0xFFFFu) {
                    const bool same = d[p + lane] == d[p + lane - best] && d[p + 32 + lane] == d[p + 32 + lane - best];
                    if (__all_sync(BDF_FULL_MASK, same)) hits++;
                }
            }
        }
        if (lane == 0) klass[idx] = hits >= 3 ? 1 : 0;
    }
}

// Host-side dispatcher.  *why != nullptr with cudaSuccess means "unsupported".
// ---- chunked streams: Compressor::compress, src/compress/mod.rs:699-772.  An input above 256 KiB
// is cut into 256 KiB chunks, each compressed by a fresh compressor (units of the deflate kernels,
// raw DEFLATE, all but the last followed by a sync flush) and the results are concatenated; the
// zlib / gzip wrapper goes around the whole (:2248-2357).  One warp per stream.
struct JoinArgs {
    const uint8_t *in;
    const uint64_t *in_off;        // n + 1, the streams
    const uint32_t *unit_begin;    // n + 1: units of stream i are [unit_begin[i], unit_begin[i+1])
    const uint8_t *tmp;            // unit outputs
    const uint64_t *tmp_off, *unit_size;
    const int32_t *unit_status;
    uint8_t *out;
    const uint64_t *out_off;
    uint64_t *out_size;
    int32_t *status;
    uint32_t n;
    int level, format;
};
constexpr int JOIN_WARPS = 4;
__global__ void __launch_bounds__(JOIN_WARPS * 32) deflate_join_kernel(JoinArgs a)
{
    __shared__ uint32_t s_crc[4][256];
    __shared__ uint32_t s_x2n[32];
    const unsigned lane = lane_id();
    if (a.format == BDF_GZIP) load_crc_tables_to_smem(s_crc, s_x2n);
    for (uint32_t idx = blockIdx.x * JOIN_WARPS + (threadIdx.x >> 5); idx < a.n; idx += gridDim.x * JOIN_WARPS) {
        const uint8_t *in = a.in + a.in_off[idx];
        const uint64_t len = a.in_off[idx + 1] - a.in_off[idx];
        uint8_t *out = a.out + a.out_off[idx];
        const uint64_t hdr = frame_header(a.format, a.level, out, lane);
        const uint64_t cap = deflate_bound(len);
        uint64_t op = 0;
        int st = BDF_OK;
        for (uint32_t u = a.unit_begin[idx]; u < a.unit_begin[idx + 1]; u++) {
            if (a.unit_status[u] != BDF_OK) { st = a.unit_status[u]; break; }
            const uint64_t sz = a.unit_size[u];
            if (op + sz > cap) { st = BDF_INSUFFICIENT_SPACE; break; }
            warp_copy_bytes(out + hdr + op, a.tmp + a.tmp_off[u], sz, lane);
            op += sz;
        }
        uint64_t total = 0;
        if (st == BDF_OK) total = frame_footer(a.format, in, len, out, hdr + op, s_crc, s_x2n, lane);
        if (lane == 0) { a.status[idx] = st; a.out_size[idx] = total; }
        __syncwarp();
    }
}

// max_len: an upper bound of the entry lengths when the caller knows one (host API), 0 otherwise
// (device API: the 64 KiB instances run and longer entries get BDF_STREAM_UNSUPPORTED).
inline cudaError_t launch_deflate(DeflateArgs a, DeflateScratch &scratch, int sm_count, cudaStream_t s,
                                  int *nlaunch, const char **why, uint64_t max_len = 0)
{
    // the level-1 estimator always runs the block-split path, which only the 256 KiB instance holds
    const bool big = max_len > 65536 || (a.size_only && a.level == 1);
    *nlaunch = 0;
    *why = nullptr;
    cudaError_t e;
    if (a.level == 0) {
        unsigned long long want = ((unsigned long long)a.n + L0_WARPS_PER_BLOCK - 1) / L0_WARPS_PER_BLOCK;
        unsigned long long full = (unsigned long long)sm_count * 8;
        deflate_stored_kernel<<<(unsigned)(want < full ? want : full), L0_WARPS_PER_BLOCK * 32, 0, s>>>(a);
        *nlaunch = 1;
        return cudaGetLastError();
    }
    if (a.level == 1) {
        const size_t smem = sizeof(L1Smem);
        // read per launch (two getenv calls against a kernel of milliseconds) so that one process can
        // compare settings
        const char *env = getenv("BDF_L1_CTAS_PER_SM");
        const int l1_ctas_per_sm = env && atoi(env) > 0 && atoi(env) <= 16 ? atoi(env) : BDF_L1_MIN_CTAS;
        // 0 = one match per round (the round-1 parse); bit 0 = whole-window rounds, bits 1 / 2 = bucket prefetch into
        // L1 / L2, bit 3 = whole-window rounds also where blocks are split (units above 64 KiB, size estimation)
        env = getenv("BDF_L1_WINDOW");
        a.l1_window = env ? atoi(env) : 11;
        const unsigned long long full = (unsigned long long)sm_count * l1_ctas_per_sm;
        const unsigned long long want = ((unsigned long long)a.n + L1_WARPS - 1) / L1_WARPS;
        const unsigned grid = (unsigned)(want < full ? want : full);
        const size_t per_warp = big ? L1Cfg<true>::TABLE_BYTES : L1Cfg<false>::TABLE_BYTES;
        const size_t need = per_warp * L1_WARPS * (size_t)(big ? grid : full);
        if (scratch.cap < need) {
            if (scratch.p) cudaFree(scratch.p);
            scratch.p = nullptr;
            scratch.cap = 0;
            *why = "cudaMalloc(deflate scratch)";
            e = cudaMalloc(&scratch.p, need);
            if (e != cudaSuccess) return e;
            scratch.cap = need;
            *why = nullptr;
        }
        a.scratch = scratch.p;
        a.scratch_stride = per_warp;
        if (a.size_only) deflate_l1_kernel<true, true><<<grid, L1_WARPS * 32, smem, s>>>(a);
        else if (big) deflate_l1_kernel<true><<<grid, L1_WARPS * 32, smem, s>>>(a);
        else deflate_l1_kernel<false><<<grid, L1_WARPS * 32, smem, s>>>(a);
        *nlaunch = 1;
        return cudaGetLastError();
    }
    static int hc_old = -1;
    if (hc_old < 0) {
        const char *env = getenv("BDF_HC_KERNEL");
        hc_old = env && !strcmp(env, "old") ? 1 : env && !strcmp(env, "new") ? 2 : 0;
    }
    if (a.level <= 9 && !big && hc_old != 1) {
        // streams of at most 64 KiB: runs / short periods -> deflate_hc_kernel (searches where the parse
        // lands), everything else -> deflate_hcs_kernel (everything in shared memory, one CTA per SM)
        const size_t smem_s = sizeof(HcsSmem), smem_o = sizeof(HcSmem);
        if (!scratch.hcs_ready) {
            *why = "cudaFuncSetAttribute(deflate_hcs_kernel)";
            e = cudaFuncSetAttribute(deflate_hcs_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(deflate_hcs_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(deflate_hc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_o);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(deflate_hc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_o);
            if (e != cudaSuccess) return e;
            scratch.hcs_ready = true;
            *why = nullptr;
        }
        static int hc_ctas = 0;
        if (hc_ctas == 0) {
            const char *env = getenv("BDF_HC_CTAS_PER_SM");
            hc_ctas = env && atoi(env) > 0 ? atoi(env) : 8;
        }
        const unsigned old_full = (unsigned)sm_count * (unsigned)hc_ctas;
        const size_t hcs_bytes = HCS_SCRATCH_PER_CTA * (size_t)sm_count;
        const size_t old_bytes = HcChains<false>::SCRATCH_PER_CTA * (size_t)old_full;
        const size_t klass_bytes = ((size_t)a.n + 255) & ~(size_t)255;
        const size_t need = hcs_bytes + old_bytes + klass_bytes + 2 * sizeof(unsigned long long) + 256;
        if (scratch.cap < need) {
            if (scratch.p) cudaFree(scratch.p);
            scratch.p = nullptr;
            scratch.cap = 0;
            *why = "cudaMalloc(deflate scratch)";
            e = cudaMalloc(&scratch.p, need);
            if (e != cudaSuccess) return e;
            scratch.cap = need;
            *why = nullptr;
        }
        uint8_t *base = static_cast<uint8_t *>(scratch.p);
        uint8_t *klass = base + hcs_bytes + old_bytes;
        unsigned long long *ctr2 = reinterpret_cast<unsigned long long *>(klass + klass_bytes);
        const bool split = hc_old == 0;                  // BDF_HC_KERNEL=new: everything through deflate_hcs_kernel
        if (split) {
            unsigned long long want = ((unsigned long long)a.n + CLASSIFY_WARPS - 1) / CLASSIFY_WARPS;
            unsigned long long full = (unsigned long long)sm_count * 8;
            deflate_classify_kernel<<<(unsigned)(want < full ? want : full), CLASSIFY_WARPS * 32, 0, s>>>(a.in, a.in_off, a.n, klass);
            e = cudaMemsetAsync(ctr2, 0, sizeof(unsigned long long), s);
            if (e != cudaSuccess) return e;
            DeflateArgs o = a;
            o.klass = klass; o.want = 1;
            o.work_counter = ctr2;
            o.scratch = base + hcs_bytes;
            o.scratch_stride = HcChains<false>::SCRATCH_PER_CTA;
            const unsigned grid_o = a.n < old_full ? a.n : old_full;
            if (a.size_only) deflate_hc_kernel<false, true><<<grid_o, HC_THREADS, smem_o, s>>>(o);
            else deflate_hc_kernel<false><<<grid_o, HC_THREADS, smem_o, s>>>(o);
            *nlaunch += 2;
            a.klass = klass; a.want = 0;
        }
        const unsigned grid = a.n < (unsigned)sm_count ? a.n : (unsigned)sm_count;
        a.scratch = base;
        a.scratch_stride = HCS_SCRATCH_PER_CTA;
        if (a.size_only) deflate_hcs_kernel<true><<<grid, HCS_THREADS, smem_s, s>>>(a);
        else deflate_hcs_kernel<false><<<grid, HCS_THREADS, smem_s, s>>>(a);
        *nlaunch += 1;
        return cudaGetLastError();
    }
    if (a.level <= 9) {
        const size_t smem = sizeof(HcSmem);
        static int hc_ctas_per_sm = 0;
        if (hc_ctas_per_sm == 0) {
            const char *env = getenv("BDF_HC_CTAS_PER_SM");
            hc_ctas_per_sm = env && atoi(env) > 0 ? atoi(env) : 8;
        }
        const unsigned full_grid = (unsigned)sm_count * (unsigned)hc_ctas_per_sm;
        unsigned grid = a.n < full_grid ? a.n : full_grid;
        if (!scratch.hc_ready) {
            *why = "cudaFuncSetAttribute(deflate_hc_kernel)";
            e = cudaFuncSetAttribute(deflate_hc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(deflate_hc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(deflate_hc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(deflate_hc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            scratch.hc_ready = true;
            *why = nullptr;
        }
        // small instance: one slab per resident CTA (stays L2-resident); big instance: only as many
        // slabs as CTAs are launched (1.7 MiB each)
        const size_t per_cta = big ? HcChains<true>::SCRATCH_PER_CTA : HcChains<false>::SCRATCH_PER_CTA;
        const size_t need = per_cta * (size_t)(big ? grid : full_grid);
        if (scratch.cap < need) {
            if (scratch.p) cudaFree(scratch.p);
            scratch.p = nullptr;
            scratch.cap = 0;
            *why = "cudaMalloc(deflate scratch)";
            e = cudaMalloc(&scratch.p, need);
            if (e != cudaSuccess) return e;
            scratch.cap = need;
            *why = nullptr;
        }
        a.scratch = scratch.p;
        a.scratch_stride = per_cta;
        if (a.size_only && big) deflate_hc_kernel<true, true><<<grid, HC_THREADS, smem, s>>>(a);
        else if (a.size_only) deflate_hc_kernel<false, true><<<grid, HC_THREADS, smem, s>>>(a);
        else if (big) deflate_hc_kernel<true><<<grid, HC_THREADS, smem, s>>>(a);
        else deflate_hc_kernel<false><<<grid, HC_THREADS, smem, s>>>(a);
        *nlaunch = 1;
        return cudaGetLastError();
    }
    static int no_old = -1;
    if (no_old < 0) {
        const char *env = getenv("BDF_NO_KERNEL");
        no_old = env && !strcmp(env, "old") ? 1 : 0;
    }
    static int nos_split = -1, nos_wave = 0, nos_depth = 0;
    if (nos_split < 0) {
        const char *denv = getenv("BDF_NOS_DEPTH");
        nos_depth = denv && atoi(denv) > 0 ? atoi(denv) : 0;
        const char *env = getenv("BDF_NOS_SPLIT");
        nos_split = env ? (atoi(env) != 0) : 1;
        const char *wenv = getenv("BDF_NOS_WAVE");
        // streams per SM and wave.  The cost kernel takes ~18.5 ms per wave whatever the wave holds up to a
        // few dozen warps per SM (a chain of ~300 dependent instructions per step), so a wave should be as
        // large as the scratch allows: 2.9 MB per stream, 6 GB for 14 x 148 (gpurun_out/nos_probe_r3m_waves.txt:
        // corpus A level 12 1.60 / 2.64 / 3.77 GB/s at 4 / 8 / 14)
        nos_wave = wenv && atoi(wenv) > 0 ? atoi(wenv) : 14;
    }
    a.nos_depth = nos_depth;
    if (!big && !a.size_only && !no_old && nos_split) {
        // levels 10..12, streams of at most 64 KiB, three kernels per wave of streams (deflate_nos_split.cuh):
        // the serial cost pass runs one warp per stream for the whole wave instead of one warp per SM
        const size_t smem = sizeof(HcsSmem);
        if (!scratch.nos_split_ready) {
            *why = "cudaFuncSetAttribute(deflate_nos_search_kernel / deflate_nos_emit_kernel)";
            e = cudaFuncSetAttribute(deflate_nos_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(deflate_nos_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            scratch.nos_split_ready = true;
            *why = nullptr;
        }
        const unsigned wave_max = (unsigned)sm_count * (unsigned)nos_wave;
        unsigned wave = a.n < wave_max ? a.n : wave_max;
        if (scratch.cap < NOS_SPLIT_PER_STREAM * (size_t)wave) {
            if (scratch.p) cudaFree(scratch.p);
            scratch.p = nullptr;
            scratch.cap = 0;
            *why = "cudaMalloc(deflate scratch)";
            // a smaller wave if the device cannot spare the memory
            for (;;) {
                e = cudaMalloc(&scratch.p, NOS_SPLIT_PER_STREAM * (size_t)wave);
                if (e == cudaSuccess || wave <= (unsigned)sm_count) break;
                (void)cudaGetLastError();
                wave = wave / 2 > (unsigned)sm_count ? wave / 2 : (unsigned)sm_count;
            }
            if (e != cudaSuccess) return e;
            scratch.cap = NOS_SPLIT_PER_STREAM * (size_t)wave;
            *why = nullptr;
        }
        a.scratch = scratch.p;
        a.scratch_stride = NOS_SPLIT_PER_STREAM;
        *nlaunch = 0;
        for (unsigned first = 0; first < a.n; first += wave) {
            NosWave wv;
            wv.first = first;
            wv.count = a.n - first < wave ? a.n - first : wave;
            const unsigned grid = wv.count < (unsigned)sm_count ? wv.count : (unsigned)sm_count;
            deflate_nos_search_kernel<<<grid, HCS_THREADS, smem, s>>>(a, wv);
            deflate_nos_cost_kernel<<<(wv.count + NOS_COST_WARPS - 1) / NOS_COST_WARPS, NOS_COST_WARPS * 32, 0, s>>>(a, wv);
            deflate_nos_emit_kernel<<<grid, HCS_THREADS, smem, s>>>(a, wv);
            *nlaunch += 3;
        }
        return cudaGetLastError();
    }
    if (!big && !a.size_only && !no_old) {
        // levels 10..12, streams of at most 64 KiB: the near-optimal parser on the shared-memory
        // hash-chain machinery (size within 0.5 % of the reference's, not its bytes)
        const size_t smem = sizeof(HcsSmem);
        if (!scratch.nos_ready) {
            *why = "cudaFuncSetAttribute(deflate_nos_kernel)";
            e = cudaFuncSetAttribute(deflate_nos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            scratch.nos_ready = true;
            *why = nullptr;
        }
        const unsigned grid = a.n < (unsigned)sm_count ? a.n : (unsigned)sm_count;
        const size_t need = NOS_SCRATCH_PER_CTA * (size_t)sm_count;
        if (scratch.cap < need) {
            if (scratch.p) cudaFree(scratch.p);
            scratch.p = nullptr;
            scratch.cap = 0;
            *why = "cudaMalloc(deflate scratch)";
            e = cudaMalloc(&scratch.p, need);
            if (e != cudaSuccess) return e;
            scratch.cap = need;
            *why = nullptr;
        }
        a.scratch = scratch.p;
        a.scratch_stride = NOS_SCRATCH_PER_CTA;
        deflate_nos_kernel<<<grid, HCS_THREADS, smem, s>>>(a);
        *nlaunch = 1;
        return cudaGetLastError();
    }
    {
        static int bt_threads_per_sm = 0;
        if (bt_threads_per_sm == 0) {
            const char *env = getenv("BDF_BT_THREADS_PER_SM");
            bt_threads_per_sm = env && atoi(env) > 0 ? atoi(env) : 128;
        }
        // the 256 KiB instance needs 4.3 MiB per thread: a quarter of the threads
        const unsigned long long full = (unsigned long long)sm_count * (big ? bt_threads_per_sm / 4 : bt_threads_per_sm) / BT_THREADS;
        const unsigned long long want = ((unsigned long long)a.n + BT_THREADS - 1) / BT_THREADS;
        const unsigned grid = (unsigned)(want < full ? want : full);
        const size_t per_thread = big ? BtTables<true>::SLAB_BYTES : BtTables<false>::SLAB_BYTES;
        const size_t need = per_thread * (size_t)grid * BT_THREADS;
        if (scratch.cap < need) {
            if (scratch.p) cudaFree(scratch.p);
            scratch.p = nullptr;
            scratch.cap = 0;
            *why = "cudaMalloc(deflate scratch)";
            e = cudaMalloc(&scratch.p, need);
            if (e != cudaSuccess) return e;
            scratch.cap = need;
            *why = nullptr;
        }
        a.scratch = scratch.p;
        a.scratch_stride = per_thread;
        if (big) deflate_bt_kernel<true><<<grid, BT_THREADS, 0, s>>>(a);
        else deflate_bt_kernel<false><<<grid, BT_THREADS, 0, s>>>(a);
        *nlaunch = 1;
        return cudaGetLastError();
    }
}

}  // namespace bdf
