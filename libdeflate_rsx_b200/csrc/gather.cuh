// gather.cuh — packs the bound-spaced output slab of a batch compression into a dense flat
// buffer (the layout the batch decompressor takes as input).  The reference does the same step on
// the host: every result is sliced out of the slab by (out_off[i], out_size[i]) and copied into
// its own Vec (src/batch_cuda.rs:123-137, src/batch.rs:51).  One warp per stream, 16-byte
// destination-aligned chunks.
#pragma once
#include "common.cuh"

namespace bdf {

struct GatherArgs {
    const uint8_t *src;
    const uint64_t *src_off;     // n
    const uint64_t *size;        // n
    uint8_t *dst;
    const uint64_t *dst_off;     // n + 1 (exclusive prefix sums of size)
    uint32_t n;
};

constexpr int GATHER_WARPS = 4;

// d[0..len) = s[0..len), one warp, 16-byte destination-aligned chunks, source realigned with PRMT
__device__ __forceinline__ void warp_copy_bytes(uint8_t *d, const uint8_t *s, uint64_t len, unsigned lane)
{
    uint64_t head = (16u - (reinterpret_cast<uintptr_t>(d) & 15u)) & 15u;
    if (head > len) head = len;
    if (lane < head) d[lane] = s[lane];
    const uint64_t nbody = (len - head) >> 4;
    const uint8_t *sb = s + head;
    const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(sb) & 3u);
    const uint32_t sel = 0x3210u + 0x1111u * sh;
    for (uint64_t c = lane; c < nbody; c += 32) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(sb + 16 * c - sh);
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = sh ? w[4] : 0u;
        *reinterpret_cast<uint4 *>(d + head + 16 * c) =
            make_uint4(__byte_perm(w0, w1, sel), __byte_perm(w1, w2, sel), __byte_perm(w2, w3, sel),
                       __byte_perm(w3, w4, sel));
    }
    const uint64_t done = head + 16 * nbody;
    if (done + lane < len) d[done + lane] = s[done + lane];
}

__global__ void __launch_bounds__(GATHER_WARPS * 32) gather_kernel(GatherArgs a)
{
    const unsigned lane = lane_id();
    for (uint64_t i = (uint64_t)blockIdx.x * GATHER_WARPS + threadIdx.x / 32; i < a.n;
         i += (uint64_t)gridDim.x * GATHER_WARPS)
        warp_copy_bytes(a.dst + a.dst_off[i], a.src + a.src_off[i], a.size[i], lane);
}

// dst_off[0..n] = exclusive prefix sums of size[0..n) (one CTA; the batch sizes of this engine,
// up to a few million streams, take microseconds)
constexpr int SCAN_THREADS = 1024;
__global__ void __launch_bounds__(SCAN_THREADS) size_scan_kernel(const uint64_t *size, uint64_t *dst_off, uint32_t n)
{
    __shared__ uint64_t part[SCAN_THREADS];
    const uint32_t per = (n + SCAN_THREADS - 1) / SCAN_THREADS;
    const uint32_t beg = threadIdx.x * per < n ? threadIdx.x * per : n;
    const uint32_t end = beg + per < n ? beg + per : n;
    uint64_t sum = 0;
    for (uint32_t i = beg; i < end; i++) sum += size[i];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (int d = 1; d < SCAN_THREADS; d <<= 1) {
        const uint64_t v = threadIdx.x >= (unsigned)d ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    uint64_t run = part[threadIdx.x] - sum;
    for (uint32_t i = beg; i < end; i++) { dst_off[i] = run; run += size[i]; }
    if (threadIdx.x == SCAN_THREADS - 1) dst_off[n] = part[SCAN_THREADS - 1];
}

}  // namespace bdf
