// inflate_resume.cuh — resumable DEFLATE decoding for MANY decoder states, one LANE per state (sm_100a).
//
// Replaces Decompressor::decompress_streaming as DeflateDecoder::read calls it (reference
// src/decompress/mod.rs:204-372, src/stream.rs:263-376).  A streaming reader decodes a window's worth
// of output, hands it on, and continues later: per call it is a small, strictly serial piece of work
// whose state has to survive between calls.  What a GPU can add is width — a server holds thousands
// of such readers — so the unit here is a BATCH of decoder states advanced by one launch: every lane
// owns one state (bdf_inflate_state, 368 bytes of plain data in global memory) and runs the scalar
// step of inflate_resume_core.h on it with decode tables of its own in shared memory (3.4 KB per
// lane: 10-bit litlen and 8-bit offset direct tables, sorted symbol lists and first-code arrays for
// longer codewords; rebuilt from the saved code lengths whenever a call resumes inside a block).
// Lanes of a warp diverge freely — each is its own decoder — which is the price of keeping the step
// identical to the host-tested core; the batch engines (inflate.cuh, inflate_lane.cuh) remain the
// fast path for whole streams.
#pragma once
#include "common.cuh"
#include "inflate_resume_core.h"

namespace bdf {

struct ResumeArgs {
    bdf_inflate_state *states;
    const uint8_t *in;
    const uint64_t *in_off;
    const uint8_t *in_final;
    uint8_t *window;
    const uint64_t *win_off, *win_cap;
    uint64_t *win_pos, *in_consumed;
    int32_t *status;
    uint32_t n;
};

constexpr int RESUME_THREADS = 32;
constexpr size_t RESUME_SMEM = RESUME_THREADS * sizeof(bdf_rs::Tables);

__global__ void __launch_bounds__(RESUME_THREADS) inflate_resume_kernel(ResumeArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    bdf_rs::Tables &T = reinterpret_cast<bdf_rs::Tables *>(smem_raw)[threadIdx.x];
    const unsigned long long i = (unsigned long long)blockIdx.x * RESUME_THREADS + threadIdx.x;
    if (i >= a.n) return;
    const uint64_t o0 = a.in_off[i], len = a.in_off[i + 1] - o0;
    uint64_t pos = a.win_pos[i], used = 0;
    const int st = bdf_rs::resume_step(a.states[i], a.in + o0, len, a.in_final[i] != 0, a.window + a.win_off[i], a.win_cap[i],
                                       &pos, &used, T);
    a.win_pos[i] = pos;
    a.in_consumed[i] = used;
    a.status[i] = st;
}

}  // namespace bdf
