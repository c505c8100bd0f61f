// write_probe.cu — ceiling of the inflate kernel's WRITE PATTERN without any decoding: many
// concurrent 64 KiB output streams per SM (one per 16-lane group, 64-thread CTAs, 14 CTAs per SM),
// each written front to back.  mode 0: 16 lanes x STG.128; mode 1: one lane issues cp.async.bulk
// shared -> global copies of `piece` bytes (the replay of inflate.cuh); mode 2: plain grid-stride
// fill (every warp writes consecutive 512-byte rows).  Prints GB/s per mode.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_scripts/probe/write_probe gpurun_scripts/probe/write_probe.cu
// run:   gpurun -- 'gpurun_scripts/probe/write_probe 65536'
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void __launch_bounds__(64, 14) probe_streams(uint8_t *out, unsigned n, unsigned long long *counter,
                                                        unsigned piece, int mode)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const unsigned g = threadIdx.x >> 4, lane = threadIdx.x & 15u, wl = threadIdx.x & 31u;
    const unsigned mask = 0xFFFFu << (wl & 16u);
    uint8_t *buf = smem + g * 3184u;                       // the kernel's per-stream shared memory
    for (unsigned i = lane * 16; i < 2048; i += 256) *reinterpret_cast<uint4 *>(buf + (i % 3168u)) = make_uint4(i, g, 3, 4);
    __syncwarp(mask);
    for (;;) {
        unsigned long long idx = 0;
        if (lane == 0) idx = atomicAdd(counter, 1ull);
        idx = __shfl_sync(mask, idx, 0, 16);
        if (idx >= n) break;
        uint8_t *dst = out + idx * 65536ull;
        if (mode == 0) {
            const uint4 v = make_uint4((unsigned)idx, lane, 1, 2);
            for (unsigned i = lane; i < 4096; i += 16) reinterpret_cast<uint4 *>(dst)[i] = v;
        } else {
            asm volatile("fence.proxy.async;" ::: "memory");
            __syncwarp(mask);
            if (lane == 0) {
                const uint32_t src = (uint32_t)__cvta_generic_to_shared(buf);
                unsigned off = 0;
                for (; off + piece <= 65536u; off += piece)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(src), "r"(piece) : "memory");
                if (off < 65536u)
                    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(src), "r"(65536u - off) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            }
            __syncwarp(mask);
        }
    }
}

__global__ void __launch_bounds__(256) probe_fill(uint4 *out, size_t n16)
{
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) out[i] = v;
}

int main(int argc, char **argv)
{
    const unsigned n = argc > 1 ? atoi(argv[1]) : 65536;
    uint8_t *out;
    unsigned long long *counter;
    const size_t bytes = (size_t)n * 65536;
    if (cudaMalloc(&out, bytes) != cudaSuccess || cudaMalloc(&counter, 8 * 64) != cudaSuccess) { puts("alloc failed"); return 1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int smem = 4 * 3184;
    struct { int mode; unsigned piece; int ctas; const char *name; } runs[] = {
        {2, 0, 0, "grid-stride fill"},
        {0, 0, 14, "streams, 16-lane STG.128, 14 CTAs/SM"},
        {0, 0, 7, "streams, 16-lane STG.128, 7 CTAs/SM"},
        {0, 0, 4, "streams, 16-lane STG.128, 4 CTAs/SM"},
        {1, 1600, 14, "streams, bulk 1600 B, 14 CTAs/SM"},
        {1, 1600, 7, "streams, bulk 1600 B, 7 CTAs/SM"},
        {1, 1600, 4, "streams, bulk 1600 B, 4 CTAs/SM"},
        {1, 3072, 14, "streams, bulk 3072 B, 14 CTAs/SM"},
        {1, 1024, 14, "streams, bulk 1024 B, 14 CTAs/SM"},
        {1, 512, 14, "streams, bulk 512 B, 14 CTAs/SM"},
    };
    for (auto &r : runs) {
        float best = 1e30f, sum = 0;
        const int iters = 8;
        for (int it = 0; it < iters + 2; it++) {
            cudaMemsetAsync(counter, 0, 8 * 64);
            cudaEventRecord(e0);
            if (r.mode == 2) probe_fill<<<sms * 8, 256>>>(reinterpret_cast<uint4 *>(out), bytes / 16);
            else probe_streams<<<sms * r.ctas, 64, smem>>>(out, n, counter, r.piece, r.mode);
            cudaEventRecord(e1);
            if (cudaEventSynchronize(e1) != cudaSuccess) { printf("%s: %s\n", r.name, cudaGetErrorString(cudaGetLastError())); return 1; }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it >= 2) { sum += ms; if (ms < best) best = ms; }
        }
        printf("%-44s mean %.3f ms  %.0f GB/s   best %.0f GB/s\n", r.name, sum / iters, bytes / (sum / iters) / 1e6, bytes / best / 1e6);
    }
    return 0;
}
