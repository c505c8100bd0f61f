// checksum.cuh — batch Adler-32 / CRC-32, one warp per stream (sm_100a).
//
// Replaces adler32(1, data) / crc32(0, data) (reference src/adler32/mod.rs:114-152,
// src/crc32/mod.rs:331-365) over many independent buffers.
//   Adler-32: with 0-based byte index i, s1 = 1 + Σb and s2 = n + n·Σb − Σ i·b
//   (mod 65521), so lanes sum (Σb, Σ i·b) over interleaved 16-byte vectors with
//   dp4a and one warp reduction finishes the stream — the "vectorised reduction".
//   CRC-32: see grp_crc32 (inflate.cuh) — per-lane slice-by-4 over contiguous
//   slices, recombined with x^(8k) mod P multipliers.
#pragma once
#include "inflate.cuh"

namespace bdf {

constexpr int CK_WARPS_PER_BLOCK = 8;

__device__ uint32_t warp_adler32(const uint8_t *d, uint64_t n, unsigned lane)
{
    uint32_t A = 0;
    uint64_t B = 0;
    // head bytes up to 16-byte alignment
    uint64_t head = (16 - ((uintptr_t)d & 15)) & 15;
    if (head > n) head = n;
    if (lane < head) { A += d[lane]; B += (uint64_t)lane * d[lane]; }
    const uint64_t nvec = (n - head) / 16;
    const uint4 *v = reinterpret_cast<const uint4 *>(d + head);
    uint32_t a32 = 0;
    uint64_t b64 = 0;
    for (uint64_t k = lane; k < nvec; k += 32) {
        uint4 q = __ldg(v + k);
        uint32_t s0 = __dp4a(q.x, 0x01010101u, 0u), s1 = __dp4a(q.y, 0x01010101u, 0u);
        uint32_t s2 = __dp4a(q.z, 0x01010101u, 0u), s3 = __dp4a(q.w, 0x01010101u, 0u);
        uint32_t w = __dp4a(q.x, 0x03020100u, 0u) + __dp4a(q.y, 0x07060504u, 0u) +
                     __dp4a(q.z, 0x0B0A0908u, 0u) + __dp4a(q.w, 0x0F0E0D0Cu, 0u);
        uint32_t s = s0 + s1 + s2 + s3;
        a32 += s;
        b64 += (head + 16 * k) * (uint64_t)s + w;
        if ((k >> 5 & 0xFFF) == 0xFFF) {      // fold long before 64-bit overflow
            b64 %= 65521u;
            a32 %= 65521u;
        }
    }
    A += a32 % 65521u;
    B += b64 % 65521u;
    uint64_t tail = head + nvec * 16;
    if (tail + lane < n) { uint32_t b = d[tail + lane]; A += b; B += (tail + lane) * (uint64_t)b; }
    return warp_adler_finish(A, B, n);
}

struct ChecksumArgs {
    const uint8_t *in;
    const uint64_t *in_off;
    uint32_t *out;
    uint32_t n;
};

template <int KIND>
__global__ void __launch_bounds__(CK_WARPS_PER_BLOCK * 32)
checksum_kernel(ChecksumArgs a)
{
    __shared__ uint32_t s_crc[KIND == BDF_CRC32 ? 4 : 1][256];
    __shared__ uint32_t s_x2n[32];
    const unsigned lane = lane_id();
    if (KIND == BDF_CRC32) {
        for (unsigned i = threadIdx.x; i < 1024; i += blockDim.x) s_crc[i >> 8][i & 255] = g_crc_tables.slice[i >> 8][i & 255];
        if (threadIdx.x < 32) s_x2n[threadIdx.x] = g_crc_tables.x2n[threadIdx.x];
        __syncthreads();
    }
    const uint32_t warps = gridDim.x * CK_WARPS_PER_BLOCK;
    for (uint32_t idx = blockIdx.x * CK_WARPS_PER_BLOCK + (threadIdx.x >> 5); idx < a.n; idx += warps) {
        const uint8_t *p = a.in + a.in_off[idx];
        uint64_t len = a.in_off[idx + 1] - a.in_off[idx];
        uint32_t r = KIND == BDF_CRC32 ? warp_crc32(p, len, s_crc, s_x2n, lane) : warp_adler32(p, len, lane);
        if (lane == 0) a.out[idx] = r;
    }
}

__global__ void crc_tables_init_kernel()
{
    // slice[0] = byte-at-a-time table of the reflected polynomial; slice[k] advances k more bytes
    unsigned b = threadIdx.x;
    uint32_t c = b;
    for (int k = 0; k < 8; k++) c = (c >> 1) ^ (BDF_CRC_POLY & (0u - (c & 1u)));
    g_crc_tables.slice[0][b] = c;
    __syncthreads();
    uint32_t t = c;
    for (int s = 1; s < 4; s++) {
        t = (t >> 8) ^ g_crc_tables.slice[0][t & 0xFF];
        g_crc_tables.slice[s][b] = t;
    }
    if (b == 0) {
        uint32_t p = 1u << 30;          // x^1
        g_crc_tables.x2n[0] = p;
        for (int n = 1; n < 32; n++) { p = gf2_mulmod(p, p); g_crc_tables.x2n[n] = p; }
    }
}

}  // namespace bdf
