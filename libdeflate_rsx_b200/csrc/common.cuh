// common.cuh — shared device helpers for the batch DEFLATE kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>
#include "../../include/bdeflate.h"

#define BDF_FULL_MASK 0xFFFFFFFFu

// Debug build (python -m libdeflate_rsx_b200.build --check -> libbdeflate_check.so): explicit bounds /
// invariant assertions at the kernels' shared- and global-memory indices.  The first failing source
// line is kept in a device word that bdf_debug_check_failures() reads back; compute-sanitizer is
// refused on the pool this was developed on, so this (plus tests/test_gpu_determinism.py) is the
// memcheck / racecheck stand-in.
#ifdef BDF_CHECK
__device__ unsigned int g_bdf_check_fail;
#define BDF_ASSERT(c) do { if (!(c)) atomicCAS(&g_bdf_check_fail, 0u, (unsigned)__LINE__ | 0x80000000u); } while (0)
#else
#define BDF_ASSERT(c) do { } while (0)
#endif

namespace bdf {

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// RFC 1951 length / offset code geometry, computed instead of tabulated so
// that divergent lanes never serialise on constant memory.
__device__ __forceinline__ void length_slot_info(unsigned slot, unsigned &base, unsigned &extra)
{
    if (slot < 8) { base = 3 + slot; extra = 0; }
    else if (slot >= 28) { base = 258; extra = 0; }   // 286/287 decode as 258 (tables.rs:342-343)
    else { extra = (slot >> 2) - 1; base = ((4 + (slot & 3)) << extra) + 3; }
}
__device__ __forceinline__ void offset_slot_info(unsigned slot, unsigned &base, unsigned &extra)
{
    if (slot > 29) slot = 29;                          // 30/31 alias 29 (tables.rs:377-378)
    if (slot < 4) { base = slot + 1; extra = 0; }
    else { extra = (slot >> 1) - 1; base = ((2 + (slot & 1)) << extra) + 1; }
}

// ---- CRC-32 (reflected 0xEDB88320) in GF(2): no CLMUL on sm_100, so streams
// are split across lanes and the partial CRCs are recombined with x^(8k) mod P
// multipliers — the GF(2) analogue of the fold constants the reference keeps
// in src/crc32_tables.rs:36-49.
#define BDF_CRC_POLY 0xEDB88320u
__device__ __forceinline__ uint32_t gf2_mulmod(uint32_t a, uint32_t b)
{
    uint32_t p = 0;
#pragma unroll 1
    for (int i = 0; i < 32; i++) {
        p ^= b & (0u - (a >> 31));
        a <<= 1;
        b = (b >> 1) ^ (BDF_CRC_POLY & (0u - (b & 1u)));
    }
    return p;
}
// x^(8*nbytes) mod P; x2n[k] = x^(2^k) mod P lives in shared or constant memory.
__device__ __forceinline__ uint32_t gf2_xpow8n(uint64_t nbytes, const uint32_t *x2n)
{
    uint32_t p = 0x80000000u;  // x^0
    unsigned k = 3;
    while (nbytes) {
        if (nbytes & 1) p = gf2_mulmod(x2n[k & 31], p);
        nbytes >>= 1;
        k++;
    }
    return p;
}

// ---- lane groups.  A warp is cut into 32/G groups of G consecutive lanes; each
// group owns one stream and runs in lock-step on its own (sub-)mask, so one
// warp instruction advances 32/G streams.  G = 32 is the classic warp-per-stream.
template <int G>
struct Grp {
    static_assert(G == 4 || G == 8 || G == 16 || G == 32, "group size");
    unsigned lane;    // lane index inside the group
    unsigned shift;   // first warp lane of the group
    unsigned mask;    // member mask of the group
    __device__ __forceinline__ Grp()
    {
        const unsigned wl = threadIdx.x & 31u;
        lane = wl & (unsigned)(G - 1);
        shift = wl & ~(unsigned)(G - 1);
        mask = G == 32 ? 0xFFFFFFFFu : (((1u << (G & 31)) - 1u) << shift);
    }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
    template <class T> __device__ __forceinline__ T shfl(T v, unsigned src) const { return __shfl_sync(mask, v, src, G); }
    template <class T> __device__ __forceinline__ T shfl_xor(T v, unsigned d) const { return __shfl_xor_sync(mask, v, d, G); }
    template <class T> __device__ __forceinline__ T shfl_up(T v, unsigned d) const { return __shfl_up_sync(mask, v, d, G); }
    __device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(mask, p) >> shift; }
    __device__ __forceinline__ unsigned match_any(unsigned v) const { return __match_any_sync(mask, v) >> shift; }
    __device__ __forceinline__ unsigned lt_mask() const { return (1u << lane) - 1u; }
};

struct CrcTables {
    uint32_t slice[4][256];  // slice-by-4
    uint32_t x2n[32];
};
__device__ CrcTables g_crc_tables;            // single translation unit (bdeflate.cu); filled by crc_tables_init_kernel

}  // namespace bdf
