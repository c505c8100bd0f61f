// deflate.cuh — batch DEFLATE compression kernels (sm_100a).
//
// Replaces Compressor::compress / compress_zlib / compress_gzip as called from
// BatchCompressor::compress_batch (reference src/batch.rs:20-58,
// src/compress/mod.rs:693-790,2248-2357).
#pragma once
#include "checksum.cuh"
#include "common.cuh"

namespace bdf {

struct DeflateArgs {
    const uint8_t *in;
    const uint64_t *in_off;
    uint8_t *out;
    const uint64_t *out_off;
    uint64_t *out_size;
    int32_t *status;
    unsigned long long *work_counter;
    uint32_t n;
    int level;
    int format;
};

struct DeflateScratch {
    void *p = nullptr;
    size_t cap = 0;
};
inline void deflate_scratch_free(DeflateScratch &s)
{
    if (s.p) cudaFree(s.p);
    s.p = nullptr;
    s.cap = 0;
}

__host__ __device__ inline uint64_t deflate_bound(uint64_t len) { return len + (len / 65535 + 1) * 5 + 10; }

// Framing shared by every level (compress_zlib / compress_gzip,
// src/compress/mod.rs:2248-2357).  Warp-uniform; lane 0 stores.
__device__ __forceinline__ unsigned frame_header(int format, int level, uint8_t *out, unsigned lane)
{
    if (format == BDF_ZLIB) {
        unsigned hint = level < 2 ? 0 : level < 6 ? 1 : level < 8 ? 2 : 3;
        unsigned hdr = (8u << 8) | (7u << 12) | (hint << 6);
        hdr |= 31 - (hdr % 31);
        if (lane == 0) { out[0] = (uint8_t)(hdr >> 8); out[1] = (uint8_t)hdr; }
        return 2;
    }
    if (format == BDF_GZIP) {
        if (lane < 10) {
            uint8_t b = 0;
            if (lane == 0) b = 0x1F;
            else if (lane == 1) b = 0x8B;
            else if (lane == 2) b = 8;
            else if (lane == 8) b = level < 2 ? 4 : level >= 8 ? 2 : 0;
            else if (lane == 9) b = 255;
            out[lane] = b;
        }
        return 10;
    }
    return 0;
}
// Writes the footer after `at` deflate bytes; returns total framed size.
__device__ __forceinline__ uint64_t frame_footer(int format, const uint8_t *in, uint64_t len, uint8_t *out,
                                                 uint64_t at, const uint32_t (*s_crc)[256],
                                                 const uint32_t *s_x2n, unsigned lane)
{
    if (format == BDF_ZLIB) {
        uint32_t a = warp_adler32(in, len, lane);
        if (lane < 4) out[at + lane] = (uint8_t)(a >> (24 - 8 * lane));     // big-endian
        return at + 4;
    }
    if (format == BDF_GZIP) {
        uint32_t c = warp_crc32(in, len, s_crc, s_x2n, lane);
        uint32_t isz = (uint32_t)len;
        if (lane < 4) out[at + lane] = (uint8_t)(c >> (8 * lane));
        else if (lane < 8) out[at + lane] = (uint8_t)(isz >> (8 * (lane - 4)));
        return at + 8;
    }
    return at;
}

__device__ __forceinline__ void load_crc_tables_to_smem(uint32_t (*s_crc)[256], uint32_t *s_x2n)
{
    for (unsigned i = threadIdx.x; i < 1024; i += blockDim.x) s_crc[i >> 8][i & 255] = g_crc_tables.slice[i >> 8][i & 255];
    if (threadIdx.x < 32) s_x2n[threadIdx.x] = g_crc_tables.x2n[threadIdx.x];
    __syncthreads();
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

// Host-side dispatcher.  *why != nullptr with cudaSuccess means "unsupported".
inline cudaError_t launch_deflate(const DeflateArgs &a, DeflateScratch &scratch, int sm_count, cudaStream_t s,
                                  int *nlaunch, const char **why)
{
    (void)scratch;
    *nlaunch = 0;
    *why = nullptr;
    if (a.level == 0) {
        unsigned long long want = ((unsigned long long)a.n + L0_WARPS_PER_BLOCK - 1) / L0_WARPS_PER_BLOCK;
        unsigned long long full = (unsigned long long)sm_count * 8;
        deflate_stored_kernel<<<(unsigned)(want < full ? want : full), L0_WARPS_PER_BLOCK * 32, 0, s>>>(a);
        *nlaunch = 1;
        return cudaGetLastError();
    }
    *why = "compression level not implemented yet";
    return cudaSuccess;
}

}  // namespace bdf
