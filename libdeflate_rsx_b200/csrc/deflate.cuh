// deflate.cuh — batch DEFLATE compression: level dispatch (sm_100a).
//
// Replaces Compressor::compress / compress_zlib / compress_gzip as called from
// BatchCompressor::compress_batch (reference src/batch.rs:20-58,
// src/compress/mod.rs:693-790,2248-2357).
//   level 0      deflate_stored_kernel   (this file)
//   level 1      deflate_l1_kernel       (deflate_l1.cuh)
//   levels 2..9  deflate_hc_kernel       (deflate_hc.cuh)
//   levels 10..12: not built yet (ratio-tolerance tier) -> BDF_E_UNSUPPORTED
#pragma once
#include "deflate_common.cuh"
#include "deflate_hc.cuh"
#include "deflate_l1.cuh"

namespace bdf {

struct DeflateScratch {
    void *p = nullptr;
    size_t cap = 0;
    bool l1_ready = false, hc_ready = false;
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
        uint8_t *out = a.out + a.out_off[idx];
        uint64_t op = frame_header(a.format, 0, out, lane);
        // an empty input produces ZERO deflate bytes at level 0 (the block loop never runs, :1408)
        for (uint64_t ip = 0; ip < len;) {
            uint64_t blk = len - ip < 65535 ? len - ip : 65535;
            unsigned bfinal = ip + 65535 >= len;
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
        op = frame_footer(a.format, in, len, out, op, s_crc, s_x2n, lane);
        if (lane == 0) { a.out_size[idx] = op; a.status[idx] = BDF_OK; }
    }
}

constexpr size_t HC_SCRATCH_PER_CTA = 32768 * sizeof(HcSeq);   // >= 65536/3 + 2 sequences

// Host-side dispatcher.  *why != nullptr with cudaSuccess means "unsupported".
inline cudaError_t launch_deflate(DeflateArgs a, DeflateScratch &scratch, int sm_count, cudaStream_t s,
                                  int *nlaunch, const char **why)
{
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
        if (!scratch.l1_ready) {
            *why = "cudaFuncSetAttribute(deflate_l1_kernel)";
            e = cudaFuncSetAttribute(deflate_l1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(deflate_l1_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
            if (e != cudaSuccess) return e;
            scratch.l1_ready = true;
            *why = nullptr;
        }
        unsigned long long full = (unsigned long long)sm_count * 3;
        unsigned grid = (unsigned)(a.n < full ? a.n : full);
        deflate_l1_kernel<<<grid, 32, smem, s>>>(a);
        *nlaunch = 1;
        return cudaGetLastError();
    }
    if (a.level <= 9) {
        const size_t smem = sizeof(HcSmem);
        unsigned grid = (unsigned)(a.n < (unsigned)sm_count ? a.n : (unsigned)sm_count);
        if (!scratch.hc_ready) {
            *why = "cudaFuncSetAttribute(deflate_hc_kernel)";
            e = cudaFuncSetAttribute(deflate_hc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            scratch.hc_ready = true;
            *why = nullptr;
        }
        const size_t need = HC_SCRATCH_PER_CTA * (size_t)sm_count;
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
        a.scratch_stride = HC_SCRATCH_PER_CTA;
        deflate_hc_kernel<<<grid, HC_THREADS, smem, s>>>(a);
        *nlaunch = 1;
        return cudaGetLastError();
    }
    *why = "compression level not implemented yet (levels 10-12)";
    return cudaSuccess;
}

}  // namespace bdf
