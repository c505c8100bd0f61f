// bdeflate.cu — the C ABI of include/bdeflate.h: context, launches, host staging.
//
// Single translation unit: the kernels live in the .cuh files included below.
// Built by libdeflate_rsx_b200/build.py (nvcc -gencode arch=compute_100a,code=sm_100a).
// There is no CPU path here: every entry point either runs the CUDA kernels or
// returns a negative bdf_error.
#include <cuda_runtime.h>
#include <ctype.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/bdeflate.h"
#include "checksum.cuh"
#include "inflate.cuh"
#include "inflate_lane.cuh"
#include "inflate_prehdr.cuh"
#include "inflate_resume.cuh"
#include "deflate.cuh"

namespace {

// staging slabs are sized from caller-supplied offsets: refuse anything above 2^46 bytes outright
// (far above any GPU's memory) so that sums cannot wrap
constexpr uint64_t BDF_MAX_SLAB_BYTES = 1ull << 46;
constexpr unsigned BDF_CHUNK_EVENTS = 32;      // input chunks in flight per host compress call

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct bdf_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t copy_streams[8] = {};          // large device-to-host results go out as 8 concurrent copies
    int d2h_streams = 8;                        // how many of them a call uses (BDF_D2H_STREAMS)
    unsigned copy_waited = 0;
    cudaStream_t aux_stream = nullptr;          // second engine of a decompress call (runs beside the first)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // Launches that use ctx-owned device scratch (compressor slabs, work-queue heads) are ordered
    // one after the other even when the callers pass different CUDA streams: every such launch
    // waits for ev_order of the previous one and records it again.
    cudaEvent_t ev_order = nullptr;
    bool order_pending = false;
    std::mutex mu;
    char err[320] = {0};
    uint64_t launches = 0;
    float last_ms = 0.f;
    unsigned long long *d_counters = nullptr;   // work-queue heads, one slot per launch in flight
    unsigned counter_slot = 0;
    int inflate_blocks_per_sm[3] = {0, 0, 0};
    int lane_blocks_per_sm[3][2] = {};          // [format][tables in global memory]
    int inflate_group = 16;                     // lanes per stream in inflate_kernel (BDF_INFLATE_GROUP)
    int inflate_mode = 0;                       // 0 = both engines, split by expansion ratio; 1 = lane groups only; 2 = lane per stream only (BDF_INFLATE_MODE)
    int inflate_split = 16;                     // expansion ratio from which a stream goes to the lane-group kernel (BDF_INFLATE_SPLIT)
    int lane_cfg = 5;                           // inflate_lane_kernel tables: 0 = (8, 7) bits in shared memory, 7 warps / SM; 1 = (9, 6), 5 warps; 2 = (8, 6), 8 warps; 3 / 4 = (8, 7) / (9, 7) in global memory behind L1, 16 warps; 5 = 4 with the litlen width chosen per block, 8 or 9 bits (BDF_LANE_CFG)
    int inflate_serial = 1;                     // the two engines of a call: 1 = lane groups, then lanes, on the caller's stream; 2 = the other order; 0 = side by side on two streams (BDF_INFLATE_SERIAL)
    int inflate_prehdr = 1;                     // first-block headers decoded by inflate_prehdr_kernel ahead of the engines (BDF_INFLATE_PREHDR=0: off)
    bool lane_cfg_auto = true;                  // BDF_LANE_CFG not set: small batches take the shared-memory tables (see launch_inflate_lane)
    int lane_warps_per_sm = 0;                  // cap on resident warps of inflate_lane_kernel, 0 = what fits (BDF_LANE_WARPS)
    bdf::DeflateScratch deflate_scratch;
    DevBuf in, out, in_off, out_off, max_out, out_size, status, checksum, lane_scratch, dense, dense_off, hdr_rows, hdr_meta;
    DevBuf rs_states, rs_in, rs_meta, rs_final, rs_win, rs_status;      // bdf_inflate_resume_batch_host staging
    bool resume_attr_set = false;
    void *h_stage[2] = {nullptr, nullptr};      // pinned: packing buffers of a scattered input
    size_t h_stage_cap[2] = {0, 0};
    void *h_result = nullptr;                   // pinned: dense result on its way to bound-spaced slots
    size_t h_result_cap = 0;
    cudaEvent_t ev_chunk[BDF_CHUNK_EVENTS] = {};
    cudaEvent_t ev_stage[2] = {};
    unsigned stage_no = 0;
    DevBuf u_in_off, u_tmp_off, u_size, u_status, u_flags, u_begin, u_tmp;     // chunked compression (units)
};

namespace {

// Work-queue heads: one slot per launch, handed out round robin.  A slot is reused after
// NUM_COUNTER_SLOTS further launches; launches through one ctx are ordered (ev_order), so the
// earlier user of a slot has finished long before.
constexpr unsigned NUM_COUNTER_SLOTS = 1024;

int fail(bdf_ctx *c, int code, const char *what, cudaError_t e = cudaSuccess)
{
    if (c) {
        if (e != cudaSuccess)
            snprintf(c->err, sizeof(c->err), "%s: %s", what, cudaGetErrorString(e));
        else
            snprintf(c->err, sizeof(c->err), "%s", what);
    }
    return code;
}

#define CK(call)                                                     \
    do {                                                             \
        cudaError_t e_ = (call);                                     \
        if (e_ != cudaSuccess) return fail(ctx, BDF_E_CUDA, #call, e_); \
    } while (0)

int ensure(bdf_ctx *ctx, DevBuf &b, size_t bytes)
{
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes == 0) bytes = 256;
    if (b.cap >= bytes) return 0;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) return fail(ctx, BDF_E_NOMEM, "cudaMalloc", e);
    b.cap = bytes;
    return 0;
}

unsigned long long *next_counter(bdf_ctx *ctx, cudaStream_t s)
{
    unsigned long long *c = ctx->d_counters + (ctx->counter_slot++ % NUM_COUNTER_SLOTS);
    if (cudaMemsetAsync(c, 0, sizeof(*c), s) != cudaSuccess) return nullptr;
    return c;
}

// see bdf_ctx::ev_order
int order_begin(bdf_ctx *ctx, cudaStream_t s)
{
    if (ctx->order_pending) CK(cudaStreamWaitEvent(s, ctx->ev_order, 0));
    return 0;
}
int order_end(bdf_ctx *ctx, cudaStream_t s)
{
    CK(cudaEventRecord(ctx->ev_order, s));
    ctx->order_pending = true;
    return 0;
}

template <int FORMAT, int G>
int launch_inflate_g(bdf_ctx *ctx, const bdf::InflateArgs &a, cudaStream_t s)
{
    constexpr int groups = bdf::INF_THREADS / G;
    const size_t smem = sizeof(bdf::InflateSmem<G>) * groups;
    int &bps = ctx->inflate_blocks_per_sm[FORMAT];
    if (bps == 0) {
        CK(cudaFuncSetAttribute(bdf::inflate_kernel<FORMAT, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(bdf::inflate_kernel<FORMAT, G>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, bdf::inflate_kernel<FORMAT, G>, bdf::INF_THREADS, smem));
        if (bps < 1) bps = 1;
    }
    unsigned long long want = ((unsigned long long)a.n + groups - 1) / groups;
    unsigned long long full = (unsigned long long)ctx->sm_count * bps;
    unsigned grid = (unsigned)(want < full ? want : full);
    bdf::inflate_kernel<FORMAT, G><<<grid, bdf::INF_THREADS, smem, s>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

template <int FORMAT, int LTB, int OTB, bool GT = false, bool ADAPT = false>
int launch_inflate_lane_c(bdf_ctx *ctx, bdf::InflateArgs a, cudaStream_t s)
{
    const size_t smem = sizeof(bdf::LaneSmem<LTB, OTB, GT>);
    int &bps = ctx->lane_blocks_per_sm[FORMAT][GT ? 1 : 0];
    if (bps == 0) {
        CK(cudaFuncSetAttribute(bdf::inflate_lane_kernel<FORMAT, LTB, OTB, GT, ADAPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // tables in global memory: leave the L1 as large as the shared memory in use allows
        if (!GT) CK(cudaFuncSetAttribute(bdf::inflate_lane_kernel<FORMAT, LTB, OTB, GT, ADAPT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, bdf::inflate_lane_kernel<FORMAT, LTB, OTB, GT, ADAPT>, 32, smem));
        if (bps < 1) bps = 1;
        if (ctx->lane_warps_per_sm > 0 && ctx->lane_warps_per_sm < bps) bps = ctx->lane_warps_per_sm;
    }
    unsigned long long want = ((unsigned long long)a.n + 31) / 32;
    unsigned long long full = (unsigned long long)ctx->sm_count * bps;
    unsigned grid = (unsigned)(want < full ? want : full);
    // the tables (GT) sit behind the symbol lists of the LAUNCHED grid: the kernel finds them from gridDim
    int rc = ensure(ctx, ctx->lane_scratch, (size_t)full * 32 * (bdf::LANE_SORTED_BYTES + (GT ? sizeof(bdf::LaneTab<LTB, OTB>) : 0)));
    if (rc) return rc;
    a.lane_scratch = (uint8_t *)ctx->lane_scratch.p;
    bdf::inflate_lane_kernel<FORMAT, LTB, OTB, GT, ADAPT><<<grid, 32, smem, s>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

template <int FORMAT>
int launch_inflate_lane(bdf_ctx *ctx, const bdf::InflateArgs &a, cudaStream_t s)
{
    // Tables in global memory pay off through the streams they let an SM hold (16 warps instead of 7); a
    // batch that fits the shared-memory variant in one go is faster there (mixed pipeline of bench.py,
    // 24576 light streams = 768 warps: 66 GB/s against 60).
    if (ctx->lane_cfg == 0 || (ctx->lane_cfg_auto && ((unsigned long long)a.n + 31) / 32 <= 7ull * (unsigned long long)ctx->sm_count))
        return launch_inflate_lane_c<FORMAT, 8, 7>(ctx, a, s);
    if (ctx->lane_cfg == 1) return launch_inflate_lane_c<FORMAT, 9, 6>(ctx, a, s);
    if (ctx->lane_cfg == 2) return launch_inflate_lane_c<FORMAT, 8, 6>(ctx, a, s);
    if (ctx->lane_cfg == 3) return launch_inflate_lane_c<FORMAT, 8, 7, true>(ctx, a, s);
    if (ctx->lane_cfg == 4) return launch_inflate_lane_c<FORMAT, 9, 7, true>(ctx, a, s);
    return launch_inflate_lane_c<FORMAT, 9, 7, true, true>(ctx, a, s);
}

template <int FORMAT>
int launch_inflate_group(bdf_ctx *ctx, const bdf::InflateArgs &a, cudaStream_t s)
{
    switch (ctx->inflate_group) {
        case 32: return launch_inflate_g<FORMAT, 32>(ctx, a, s);
        default: return launch_inflate_g<FORMAT, 16>(ctx, a, s);
    }
}

// Both engines of a decompress call, each taking the streams of its class: one after the other on
// the caller's stream (default), or side by side — the lane-per-stream kernel on the ctx's aux
// stream, fork / join with events.  classes: bit 0 = heavy streams present, bit 1 = light ones
// (3 when the caller cannot tell, i.e. the offsets live on the device).
template <int FORMAT>
int launch_inflate(bdf_ctx *ctx, bdf::InflateArgs a, cudaStream_t s, int classes)
{
    if (ctx->inflate_mode == 1) { a.split_ratio = 0; return launch_inflate_group<FORMAT>(ctx, a, s); }
    if (ctx->inflate_mode == 2) { a.split_ratio = 0; return launch_inflate_lane<FORMAT>(ctx, a, s); }
    a.split_ratio = (uint32_t)ctx->inflate_split;
    if (classes == 1) return launch_inflate_group<FORMAT>(ctx, a, s);
    if (classes == 2) return launch_inflate_lane<FORMAT>(ctx, a, s);
    if (ctx->inflate_serial) {
        // One after the other on the caller's stream.  Side by side (two streams, fork / join) was the
        // first design; measured with 65536 x 64 KiB streams per call it is never better and often much
        // worse — text 105 vs 140 GB/s, binary 64 vs 84, mixed 79 vs 86 (gpurun_out/inflate_modes_r3c.txt):
        // whichever kernel becomes resident first keeps shared memory the other one's CTAs need, and the
        // lane-per-stream kernel lives on its occupancy.  The heavy streams take a few percent of a
        // call's time at most (4.9 TB/s against ~0.1), so nothing is lost by running them first.
        int rc = ctx->inflate_serial == 1 ? launch_inflate_group<FORMAT>(ctx, a, s) : launch_inflate_lane<FORMAT>(ctx, a, s);
        if (rc) return rc;
        return ctx->inflate_serial == 1 ? launch_inflate_lane<FORMAT>(ctx, a, s) : launch_inflate_group<FORMAT>(ctx, a, s);
    }
    CK(cudaEventRecord(ctx->ev_fork, s));
    CK(cudaStreamWaitEvent(ctx->aux_stream, ctx->ev_fork, 0));
    int rc = launch_inflate_lane<FORMAT>(ctx, a, ctx->aux_stream);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev_join, ctx->aux_stream));
    rc = launch_inflate_group<FORMAT>(ctx, a, s);
    if (rc) return rc;
    CK(cudaStreamWaitEvent(s, ctx->ev_join, 0));
    return 0;
}

bool bad_format(int f) { return f != BDF_RAW && f != BDF_ZLIB && f != BDF_GZIP; }

// is_overlapping, src/api.rs:303-314: the safe API rejects calls whose input and output alias
bool overlaps(const void *a, size_t na, const void *b, size_t nb)
{
    const uintptr_t p1 = (uintptr_t)a, p2 = (uintptr_t)b;
    const uintptr_t e1 = p1 + na < p1 ? UINTPTR_MAX : p1 + na, e2 = p2 + nb < p2 ? UINTPTR_MAX : p2 + nb;
    return (p1 > p2 ? p1 : p2) < (e1 < e2 ? e1 : e2);
}

}  // namespace

extern "C" {

int bdf_version(void) { return BDF_VERSION; }

int bdf_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int bdf_ctx_create(int device, bdf_ctx **out)
{
    if (!out) return BDF_E_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return BDF_E_CUDA;
    bdf_ctx *ctx = new (std::nothrow) bdf_ctx();
    if (!ctx) return BDF_E_NOMEM;
    ctx->device = device;
    if (const char *e = getenv("BDF_INFLATE_GROUP")) {
        int v = atoi(e);
        if (v == 16 || v == 32) ctx->inflate_group = v;
    }
    if (const char *e = getenv("BDF_INFLATE_MODE")) {
        if (!strcmp(e, "group")) ctx->inflate_mode = 1;
        else if (!strcmp(e, "lane")) ctx->inflate_mode = 2;
    }
    if (const char *e = getenv("BDF_INFLATE_SPLIT")) {
        int v = atoi(e);
        if (v >= 1 && v <= 1032) ctx->inflate_split = v;
    }
    if (const char *e = getenv("BDF_LANE_CFG")) { ctx->lane_cfg = atoi(e) >= 0 && atoi(e) <= 5 ? atoi(e) : 5; ctx->lane_cfg_auto = false; }
    if (const char *e = getenv("BDF_INFLATE_PREHDR")) ctx->inflate_prehdr = atoi(e) != 0;
    if (const char *e = getenv("BDF_INFLATE_SERIAL")) ctx->inflate_serial = atoi(e);
    if (const char *e = getenv("BDF_LANE_WARPS")) ctx->lane_warps_per_sm = atoi(e) > 0 ? atoi(e) : 0;
    {
        // Device-to-host result copies per call: eight concurrent copies (56.6 vs 51 GB/s for one when a
        // single process owns the host).  With 2 / 4 / 8 ranks returning results to the same host the
        // aggregate is set by the host, not by this number: 70 / 73 / 104 GB/s in total whatever the
        // copies per rank (1..8) and the kind of host memory (profiles/r2_d2h_matrix.txt).
        ctx->d2h_streams = 8;
        if (const char *e = getenv("BDF_D2H_STREAMS")) {
            int v = atoi(e);
            if (v >= 1 && v <= 8) ctx->d2h_streams = v;
        }
    }
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    for (int i = 0; i < 8 && e == cudaSuccess; i++) e = cudaStreamCreateWithFlags(&ctx->copy_streams[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming);
    for (unsigned i = 0; i < BDF_CHUNK_EVENTS && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming);
    for (int i = 0; i < 2 && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&ctx->ev_stage[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_counters, NUM_COUNTER_SLOTS * sizeof(unsigned long long));
    if (e == cudaSuccess) {
        bdf::crc_tables_init_kernel<<<1, 256, 0, ctx->stream>>>();
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    if (e != cudaSuccess) {
        fprintf(stderr, "bdf_ctx_create: %s\n", cudaGetErrorString(e));
        bdf_ctx_destroy(ctx);
        return BDF_E_CUDA;
    }
    *out = ctx;
    return BDF_E_OK;
}

void bdf_ctx_destroy(bdf_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    DevBuf *bufs[] = {&ctx->in, &ctx->out, &ctx->in_off, &ctx->out_off, &ctx->max_out,
                      &ctx->out_size, &ctx->status, &ctx->checksum, &ctx->lane_scratch, &ctx->dense, &ctx->dense_off, &ctx->hdr_rows, &ctx->hdr_meta, &ctx->rs_states, &ctx->rs_in, &ctx->rs_meta, &ctx->rs_final, &ctx->rs_win, &ctx->rs_status, &ctx->u_in_off, &ctx->u_tmp_off,
                      &ctx->u_size, &ctx->u_status, &ctx->u_flags, &ctx->u_begin, &ctx->u_tmp};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    bdf::deflate_scratch_free(ctx->deflate_scratch);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    for (int i = 0; i < 8; i++)
        if (ctx->copy_streams[i]) cudaStreamDestroy(ctx->copy_streams[i]);
    if (ctx->aux_stream) { cudaStreamSynchronize(ctx->aux_stream); cudaStreamDestroy(ctx->aux_stream); }
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_order) cudaEventDestroy(ctx->ev_order);
    for (unsigned i = 0; i < BDF_CHUNK_EVENTS; i++) if (ctx->ev_chunk[i]) cudaEventDestroy(ctx->ev_chunk[i]);
    for (int i = 0; i < 2; i++) if (ctx->ev_stage[i]) cudaEventDestroy(ctx->ev_stage[i]);
    for (int i = 0; i < 2; i++) if (ctx->h_stage[i]) cudaFreeHost(ctx->h_stage[i]);
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *bdf_last_error(const bdf_ctx *ctx) { return ctx ? ctx->err : "null ctx"; }

// Debug builds (-DBDF_CHECK): source line of the first failed device assertion | 0x80000000, 0 if
// none; -1 in a build without the assertions.  Synchronises the device.
long long bdf_debug_check_failures(bdf_ctx *ctx)
{
#ifdef BDF_CHECK
    if (!ctx) return -1;
    std::lock_guard<std::mutex> g(ctx->mu);
    unsigned v = 0;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
        cudaMemcpyFromSymbol(&v, g_bdf_check_fail, sizeof(v)) != cudaSuccess)
        return -2;
    return (long long)v;
#else
    (void)ctx;
    return -1;
#endif
}
uint64_t bdf_kernel_launches(const bdf_ctx *ctx) { return ctx ? ctx->launches : 0; }
float bdf_last_kernel_ms(const bdf_ctx *ctx) { return ctx ? ctx->last_ms : 0.f; }

// NUMA node of the calling thread's current CUDA device (sysfs), -1 if unknown
static int current_device_numa_node()
{
    int dev = 0;
    char bus[64];
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), dev) != cudaSuccess)
        return -1;
    for (char *c = bus; *c; c++) *c = (char)tolower((unsigned char)*c);
    char path[160];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}
// CPUs of a NUMA node ("0-55,112-167"); false if the list cannot be read
static bool numa_node_cpus(int node, cpu_set_t *set)
{
    char path[96];
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen(path, "r");
    if (!f) return false;
    CPU_ZERO(set);
    int a = 0, b = 0, any = 0;
    for (;;) {
        if (fscanf(f, "%d", &a) != 1) break;
        b = a;
        int c = fgetc(f);
        if (c == '-') {
            if (fscanf(f, "%d", &b) != 1) break;
            c = fgetc(f);
        }
        for (int k = a; k <= b && k < CPU_SETSIZE; k++) { CPU_SET(k, set); any = 1; }
        if (c != ',') break;
    }
    fclose(f);
    return any != 0;
}

// Pinned memory is placed by first touch, i.e. on the NUMA node of the thread inside cudaHostAlloc.
// On a two-socket 8-GPU box a result slab on the far socket crosses the inter-socket link together
// with the slabs of the other GPUs, so the calling thread is moved next to its current CUDA device
// for the duration of the allocation.  BDF_HOST_ALLOC_NUMA=0 switches this off.  (The B200 pool this
// was developed on is virtualised: one NUMA node, numa_node = -1 for every GPU, so the call changes
// nothing there — 8 ranks returning 4 GiB each reach 91-93 GB/s in total with or without it.)
void *bdf_host_alloc(size_t bytes)
{
    void *p = nullptr;
    cpu_set_t old_set, node_set, want;
    bool moved = false;
    const char *env = getenv("BDF_HOST_ALLOC_NUMA");
    if (!(env && atoi(env) == 0) && sched_getaffinity(0, sizeof(old_set), &old_set) == 0) {
        const int node = current_device_numa_node();
        if (node >= 0 && numa_node_cpus(node, &node_set)) {
            CPU_AND(&want, &old_set, &node_set);
            if (CPU_COUNT(&want) > 0 && sched_setaffinity(0, sizeof(want), &want) == 0) moved = true;
        }
    }
    const cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault);
    if (moved) sched_setaffinity(0, sizeof(old_set), &old_set);
    return e == cudaSuccess ? p : nullptr;
}
void bdf_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

size_t bdf_compress_bound(int format, size_t len)
{
    size_t b = len + (len / 65535 + 1) * 5 + 10;
    return b + (format == BDF_ZLIB ? 6 : format == BDF_GZIP ? 18 : 0);
}

// One ordered launch of the compression kernels (they all work in ctx->deflate_scratch).
static int run_deflate(bdf_ctx *ctx, bdf::DeflateArgs &a, cudaStream_t s, uint64_t max_len)
{
    int rc = order_begin(ctx, s);
    if (rc) return rc;
    a.work_counter = next_counter(ctx, s);
    if (!a.work_counter) return fail(ctx, BDF_E_CUDA, "cudaMemsetAsync(work counter)");
    int nl = 0;
    const char *why = nullptr;
    cudaError_t e = bdf::launch_deflate(a, ctx->deflate_scratch, ctx->sm_count, s, &nl, &why, max_len);
    ctx->launches += nl;
    if (e != cudaSuccess) return fail(ctx, BDF_E_CUDA, why ? why : "launch_deflate", e);
    if (why) return fail(ctx, BDF_E_UNSUPPORTED, why);
    return order_end(ctx, s);
}

// ---------------------------------------------------------------- decompress
static int decompress_device_locked(bdf_ctx *ctx, int format, const uint8_t *in, const uint64_t *in_off,
                                    size_t n, uint8_t *out, const uint64_t *out_off, const uint64_t *max_out,
                                    uint64_t *out_size, uint32_t *checksum, int32_t *status, cudaStream_t s,
                                    int classes = 3)
{
    if (n == 0) return BDF_E_OK;
    if (n > 0xFFFFFFF0ull) return fail(ctx, BDF_E_ARG, "too many streams");
    int rc = order_begin(ctx, s);
    if (rc) return rc;
    bdf::InflateArgs a;
    a.in = in; a.in_off = in_off; a.out = out; a.out_off = out_off; a.max_out = max_out;
    a.out_size = out_size; a.checksum = checksum; a.status = status; a.n = (uint32_t)n;
    a.split_ratio = 0;
    a.lane_scratch = nullptr;
    a.work_counter = next_counter(ctx, s);
    a.work_counter2 = next_counter(ctx, s);
    if (!a.work_counter || !a.work_counter2) return fail(ctx, BDF_E_CUDA, "cudaMemsetAsync(work counter)");
    a.hdr_rows = nullptr;
    a.hdr_meta = nullptr;
    if (ctx->inflate_prehdr) {
        // every stream starts with a block header: decode the first one of the whole batch lane-parallel
        if ((rc = ensure(ctx, ctx->hdr_rows, n * (size_t)bdf::PREHDR_ROW_BYTES)) || (rc = ensure(ctx, ctx->hdr_meta, n * 4)))
            return rc;
        bdf::PrehdrArgs pa;
        pa.in = in; pa.in_off = in_off; pa.rows = (uint32_t *)ctx->hdr_rows.p; pa.meta = (uint32_t *)ctx->hdr_meta.p;
        pa.n = (uint32_t)n;
        const unsigned long long want = (n + bdf::PREHDR_THREADS - 1) / bdf::PREHDR_THREADS;
        const unsigned long long cap = (unsigned long long)ctx->sm_count * 32;
        const unsigned grid = (unsigned)(want < cap ? want : cap);
        switch (format) {
            case BDF_RAW: bdf::inflate_prehdr_kernel<BDF_RAW><<<grid, bdf::PREHDR_THREADS, 0, s>>>(pa); break;
            case BDF_ZLIB: bdf::inflate_prehdr_kernel<BDF_ZLIB><<<grid, bdf::PREHDR_THREADS, 0, s>>>(pa); break;
            default: bdf::inflate_prehdr_kernel<BDF_GZIP><<<grid, bdf::PREHDR_THREADS, 0, s>>>(pa); break;
        }
        ctx->launches++;
        CK(cudaGetLastError());
        a.hdr_rows = pa.rows;
        a.hdr_meta = pa.meta;
    }
    switch (format) {
        case BDF_RAW: rc = launch_inflate<BDF_RAW>(ctx, a, s, classes); break;
        case BDF_ZLIB: rc = launch_inflate<BDF_ZLIB>(ctx, a, s, classes); break;
        default: rc = launch_inflate<BDF_GZIP>(ctx, a, s, classes); break;
    }
    if (rc) return rc;
    return order_end(ctx, s);
}

int bdf_decompress_batch_device(bdf_ctx *ctx, int format, const uint8_t *in, const uint64_t *in_off, size_t n,
                                uint8_t *out, const uint64_t *out_off, const uint64_t *max_out,
                                uint64_t *out_size, uint32_t *checksum, int32_t *status, void *stream)
{
    if (!ctx) return BDF_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (bad_format(format)) return fail(ctx, BDF_E_ARG, "unknown format");
    if (n && (!in || !in_off || !out || !out_off || !max_out || !out_size || !status))
        return fail(ctx, BDF_E_ARG, "null pointer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    return decompress_device_locked(ctx, format, in, in_off, n, out, out_off, max_out, out_size, checksum,
                                    status, s);
}

// ------------------------------------------------------- resumable decoding
static int resume_device_locked(bdf_ctx *ctx, size_t n, bdf_inflate_state *states, const uint8_t *in,
                                const uint64_t *in_off, const uint8_t *in_final, uint8_t *window,
                                const uint64_t *win_off, const uint64_t *win_cap, uint64_t *win_pos,
                                uint64_t *in_consumed, int32_t *status, cudaStream_t s)
{
    if (n > 0xFFFFFFF0ull) return fail(ctx, BDF_E_ARG, "too many decoder states");
    if (!ctx->resume_attr_set) {
        CK(cudaFuncSetAttribute(bdf::inflate_resume_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bdf::RESUME_SMEM));
        ctx->resume_attr_set = true;
    }
    bdf::ResumeArgs a;
    a.states = states; a.in = in; a.in_off = in_off; a.in_final = in_final; a.window = window;
    a.win_off = win_off; a.win_cap = win_cap; a.win_pos = win_pos; a.in_consumed = in_consumed; a.status = status;
    a.n = (uint32_t)n;
    const unsigned grid = (unsigned)((n + bdf::RESUME_THREADS - 1) / bdf::RESUME_THREADS);
    bdf::inflate_resume_kernel<<<grid, bdf::RESUME_THREADS, bdf::RESUME_SMEM, s>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return BDF_E_OK;
}

int bdf_inflate_resume_batch_device(bdf_ctx *ctx, size_t n, bdf_inflate_state *states, const uint8_t *in,
                                    const uint64_t *in_off, const uint8_t *in_final, uint8_t *window,
                                    const uint64_t *win_off, const uint64_t *win_cap, uint64_t *win_pos,
                                    uint64_t *in_consumed, int32_t *status, void *stream)
{
    if (!ctx) return BDF_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (n == 0) return BDF_E_OK;
    if (!states || !in || !in_off || !in_final || !window || !win_off || !win_cap || !win_pos || !in_consumed || !status)
        return fail(ctx, BDF_E_ARG, "null pointer");
    CK(cudaSetDevice(ctx->device));
    return resume_device_locked(ctx, n, states, in, in_off, in_final, window, win_off, win_cap, win_pos, in_consumed,
                                status, stream ? (cudaStream_t)stream : ctx->stream);
}

int bdf_inflate_resume_batch_host(bdf_ctx *ctx, size_t n, bdf_inflate_state *states, const uint8_t *in,
                                  const uint64_t *in_off, const uint8_t *in_final, uint8_t *window,
                                  const uint64_t *win_off, const uint64_t *win_cap, uint64_t *win_pos,
                                  uint64_t *in_consumed, int32_t *status)
{
    if (!ctx) return BDF_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (n == 0) return BDF_E_OK;
    if (!states || !in || !in_off || !in_final || !window || !win_off || !win_cap || !win_pos || !in_consumed || !status)
        return fail(ctx, BDF_E_ARG, "null pointer");
    constexpr uint64_t HISTORY = 32768;
    size_t win_bytes = 0;
    for (size_t i = 0; i < n; i++) {
        if (in_off[i] > in_off[i + 1]) return fail(ctx, BDF_E_ARG, "in_off is not ascending");
        const uint64_t e = win_off[i] + win_cap[i];
        if (e < win_off[i] || e > BDF_MAX_SLAB_BYTES) return fail(ctx, BDF_E_ARG, "win_off + win_cap overflows");
        if (win_pos[i] > win_cap[i]) return fail(ctx, BDF_E_ARG, "win_pos beyond win_cap");
        if (e > win_bytes) win_bytes = (size_t)e;
    }
    if (in_off[n] > BDF_MAX_SLAB_BYTES) return fail(ctx, BDF_E_ARG, "input too large");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t in_bytes = (size_t)in_off[n];
    int rc;
    // rs_meta: in_off (n + 1) | win_off | win_cap | win_pos | in_consumed   (u64 each)
    if ((rc = ensure(ctx, ctx->rs_states, n * sizeof(bdf_inflate_state))) || (rc = ensure(ctx, ctx->rs_in, in_bytes + 8)) ||
        (rc = ensure(ctx, ctx->rs_meta, (5 * n + 1) * 8)) || (rc = ensure(ctx, ctx->rs_final, n)) ||
        (rc = ensure(ctx, ctx->rs_win, win_bytes + 8)) || (rc = ensure(ctx, ctx->rs_status, n * 4)))
        return rc;
    uint64_t *m = (uint64_t *)ctx->rs_meta.p;
    uint64_t *d_in_off = m, *d_win_off = m + (n + 1), *d_win_cap = d_win_off + n, *d_win_pos = d_win_cap + n, *d_used = d_win_pos + n;
    CK(cudaMemcpyAsync(ctx->rs_states.p, states, n * sizeof(bdf_inflate_state), cudaMemcpyHostToDevice, s));
    if (in_bytes) CK(cudaMemcpyAsync(ctx->rs_in.p, in, in_bytes, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_in_off, in_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_win_off, win_off, n * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_win_cap, win_cap, n * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(d_win_pos, win_pos, n * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->rs_final.p, in_final, n, cudaMemcpyHostToDevice, s));
    // the history a step can refer to: the 32 KiB below the write position
    for (size_t i = 0; i < n; i++) {
        const uint64_t hist = win_pos[i] < HISTORY ? win_pos[i] : HISTORY;
        const uint64_t at = win_off[i] + win_pos[i] - hist;
        if (hist) CK(cudaMemcpyAsync((uint8_t *)ctx->rs_win.p + at, window + at, (size_t)hist, cudaMemcpyHostToDevice, s));
    }
    std::vector<uint64_t> old_pos(win_pos, win_pos + n);
    rc = resume_device_locked(ctx, n, (bdf_inflate_state *)ctx->rs_states.p, (const uint8_t *)ctx->rs_in.p, d_in_off,
                              (const uint8_t *)ctx->rs_final.p, (uint8_t *)ctx->rs_win.p, d_win_off, d_win_cap, d_win_pos,
                              d_used, (int32_t *)ctx->rs_status.p, s);
    if (rc) return rc;
    CK(cudaMemcpyAsync(states, ctx->rs_states.p, n * sizeof(bdf_inflate_state), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(win_pos, d_win_pos, n * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(in_consumed, d_used, n * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(status, ctx->rs_status.p, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    bool any = false;
    for (size_t i = 0; i < n; i++) {
        if (win_pos[i] < old_pos[i] || win_pos[i] > win_cap[i]) return fail(ctx, BDF_E_CUDA, "resume step returned an invalid window position");
        const uint64_t at = win_off[i] + old_pos[i], cnt = win_pos[i] - old_pos[i];
        if (cnt) {
            CK(cudaMemcpyAsync(window + at, (const uint8_t *)ctx->rs_win.p + at, (size_t)cnt, cudaMemcpyDeviceToHost, s));
            any = true;
        }
    }
    if (any) CK(cudaStreamSynchronize(s));
    return BDF_E_OK;
}

// Copies [beg, end) of a device slab to the same offsets of a host buffer.  Large ranges go out as
// several concurrent copies (one copy engine stream does not saturate the link: 51 GB/s with one
// copy of 4 GiB on the development box, 56.6 GB/s with eight); `after` has been recorded on the
// stream that produced the data.
static int copy_range_d2h(bdf_ctx *ctx, uint8_t *host, const uint8_t *dev, size_t beg, size_t end, cudaEvent_t after,
                          int *next_stream)
{
    const size_t bytes = end - beg;
    const int nstreams = ctx->d2h_streams;
    int pieces = 1;
    if (bytes >= (64u << 20)) pieces = nstreams;
    const size_t piece = ((bytes + pieces - 1) / pieces + ((1u << 20) - 1)) & ~(size_t)((1u << 20) - 1);
    for (size_t o = 0; o < bytes; o += piece) {
        const size_t cnt = bytes - o < piece ? bytes - o : piece;
        cudaStream_t cs = ctx->copy_streams[*next_stream % nstreams];
        if (!(ctx->copy_waited >> (*next_stream % nstreams) & 1u)) {
            CK(cudaStreamWaitEvent(cs, after, 0));
            ctx->copy_waited |= 1u << (*next_stream % nstreams);
        }
        CK(cudaMemcpyAsync(host + beg + o, dev + beg + o, cnt, cudaMemcpyDeviceToHost, cs));
        (*next_stream)++;
    }
    return 0;
}

int bdf_decompress_batch_host(bdf_ctx *ctx, int format, const uint8_t *in, const uint64_t *in_off, size_t n,
                              uint8_t *out, const uint64_t *out_off, const uint64_t *max_out,
                              uint64_t *out_size, uint32_t *checksum, int32_t *status)
{
    if (!ctx) return BDF_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (bad_format(format)) return fail(ctx, BDF_E_ARG, "unknown format");
    if (n == 0) return BDF_E_OK;
    if (!in || !in_off || !out || !out_off || !max_out || !out_size || !status)
        return fail(ctx, BDF_E_ARG, "null pointer");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t in_bytes = (size_t)in_off[n];
    size_t out_bytes = 0;
    int classes = 0;
    const uint64_t split = (uint64_t)ctx->inflate_split;
    for (size_t i = 0; i < n; i++) {
        // the slab is sized from these sums: a wrapped sum would let the kernel write past it
        if (in_off[i] > in_off[i + 1] || in_off[i + 1] > in_off[n]) return fail(ctx, BDF_E_ARG, "in_off is not ascending");
        const uint64_t e = out_off[i] + max_out[i];
        if (e < out_off[i] || e > BDF_MAX_SLAB_BYTES) return fail(ctx, BDF_E_ARG, "out_off + max_out overflows");
        if (e > out_bytes) out_bytes = (size_t)e;
        const uint64_t len = in_off[i + 1] - in_off[i];
        classes |= max_out[i] >= split * (len ? len : 1) ? 1 : 2;
    }
    if (overlaps(in, in_bytes, out, out_bytes)) return fail(ctx, BDF_E_ARG, "Input and output buffers overlap");
    int rc;
    if ((rc = ensure(ctx, ctx->in, in_bytes + 8)) || (rc = ensure(ctx, ctx->out, out_bytes + 8)) ||
        (rc = ensure(ctx, ctx->in_off, (n + 1) * 8)) || (rc = ensure(ctx, ctx->out_off, n * 8)) ||
        (rc = ensure(ctx, ctx->max_out, n * 8)) || (rc = ensure(ctx, ctx->out_size, n * 8)) ||
        (rc = ensure(ctx, ctx->status, n * 4)) || (rc = ensure(ctx, ctx->checksum, n * 4)))
        return rc;
    CK(cudaMemcpyAsync(ctx->in.p, in, in_bytes, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_off.p, in_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->out_off.p, out_off, n * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->max_out.p, max_out, n * 8, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(ctx->ev0, s));
    rc = decompress_device_locked(ctx, format, (const uint8_t *)ctx->in.p, (const uint64_t *)ctx->in_off.p, n,
                                  (uint8_t *)ctx->out.p, (const uint64_t *)ctx->out_off.p,
                                  (const uint64_t *)ctx->max_out.p, (uint64_t *)ctx->out_size.p,
                                  (uint32_t *)ctx->checksum.p, (int32_t *)ctx->status.p, s, classes);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, s));
    CK(cudaMemcpyAsync(out_size, ctx->out_size.p, n * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(status, ctx->status.p, n * 4, cudaMemcpyDeviceToHost, s));
    if (checksum) CK(cudaMemcpyAsync(checksum, ctx->checksum.p, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    // Only what the streams produced comes back: [out_off[i], out_off[i] + out_size[i]) of every
    // stream, neighbours that touch merged into one range.  Nothing outside a slot is written, the
    // rest of a slot (and the whole slot of a failed stream) keeps the caller's bytes.
    ctx->copy_waited = 0;
    int next_stream = 0;
    size_t i = 0;
    while (i < n) {
        size_t beg = (size_t)out_off[i], end = beg + (size_t)out_size[i];
        size_t j = i + 1;
        while (j < n && (size_t)out_off[j] == end) {
            end += (size_t)out_size[j];
            j++;
        }
        if (end > beg && (rc = copy_range_d2h(ctx, out, (const uint8_t *)ctx->out.p, beg, end, ctx->ev1, &next_stream)))
            return rc;
        i = j;
    }
    for (int k = 0; k < ctx->d2h_streams; k++)
        if (ctx->copy_waited >> k & 1u) CK(cudaStreamSynchronize(ctx->copy_streams[k]));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    return BDF_E_OK;
}

// ------------------------------------------------------------------ compress
static int compress_device_locked(bdf_ctx *ctx, int level, int format, const uint8_t *in, const uint64_t *in_off,
                                  size_t n, uint8_t *out, const uint64_t *out_off, uint64_t *out_size,
                                  int32_t *status, cudaStream_t s)
{
    bdf::DeflateArgs a;
    a.in = in; a.in_off = in_off; a.out = out; a.out_off = out_off; a.out_size = out_size; a.status = status;
    a.n = (uint32_t)n; a.level = level > 12 ? 12 : level; a.format = format; a.unit_flags = nullptr;
    return run_deflate(ctx, a, s, 0);
}

int bdf_compress_batch_device(bdf_ctx *ctx, int level, int format, const uint8_t *in, const uint64_t *in_off,
                              size_t n, uint8_t *out, const uint64_t *out_off, uint64_t *out_size,
                              int32_t *status, void *stream)
{
    if (!ctx) return BDF_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (bad_format(format)) return fail(ctx, BDF_E_ARG, "unknown format");
    if (level < 0) return fail(ctx, BDF_E_ARG, "negative level");
    if (n == 0) return BDF_E_OK;
    if (!in || !in_off || !out || !out_off || !out_size || !status) return fail(ctx, BDF_E_ARG, "null pointer");
    if (n > 0xFFFFFFF0ull) return fail(ctx, BDF_E_ARG, "too many streams");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    return compress_device_locked(ctx, level, format, in, in_off, n, out, out_off, out_size, status, s);
}

// Streams above 64 KiB: units of at most 256 KiB (Compressor::compress, src/compress/mod.rs:699-772)
// through the 256 KiB kernel instances into a temporary slab, then deflate_join_kernel.  Called with
// the inputs already on the device and the ctx mutex held.
// Device buffers of a batch (flat input + offsets, bound-spaced output slab + per-stream results).
struct DevBatch {
    const uint8_t *in;
    const uint64_t *in_off;
    uint8_t *out;
    const uint64_t *out_off;
    uint64_t *out_size;
    int32_t *status;
};

// in_off: host copy of the offsets; d: where the batch lives on the device
static int compress_chunked_locked(bdf_ctx *ctx, int level, int format, const uint64_t *in_off, size_t n,
                                   cudaStream_t s, const DevBatch &d)
{
    constexpr uint64_t CHUNK = 256 * 1024;
    std::vector<uint64_t> uoff, toff;
    std::vector<uint8_t> uflags;
    std::vector<uint32_t> ubegin(n + 1);
    uint64_t tmp_bytes = 0, max_unit = 0;
    for (size_t i = 0; i < n; i++) {
        const uint64_t beg = in_off[i], len = in_off[i + 1] - in_off[i];
        ubegin[i] = (uint32_t)uoff.size();
        uint64_t ip = 0;
        do {
            const uint64_t ul = len - ip < CHUNK ? len - ip : CHUNK;
            const bool last = ip + ul >= len;
            uoff.push_back(beg + ip);
            uflags.push_back(last ? bdf::UNIT_FINISH : bdf::UNIT_SYNC);
            toff.push_back(tmp_bytes);
            tmp_bytes += (bdf::deflate_bound(ul) + 15) & ~15ull;
            if (ul > max_unit) max_unit = ul;
            ip += ul;
        } while (ip < len);
    }
    const size_t nu = uflags.size();
    if (nu > 0xFFFFFFF0ull) return fail(ctx, BDF_E_ARG, "too many chunks");
    ubegin[n] = (uint32_t)nu;
    uoff.push_back(in_off[n]);
    int rc;
    if ((rc = ensure(ctx, ctx->u_in_off, (nu + 1) * 8)) || (rc = ensure(ctx, ctx->u_tmp_off, nu * 8)) ||
        (rc = ensure(ctx, ctx->u_size, nu * 8)) || (rc = ensure(ctx, ctx->u_status, nu * 4)) ||
        (rc = ensure(ctx, ctx->u_flags, nu)) || (rc = ensure(ctx, ctx->u_begin, (n + 1) * 4)) ||
        (rc = ensure(ctx, ctx->u_tmp, tmp_bytes + 16)))
        return rc;
    CK(cudaMemcpyAsync(ctx->u_in_off.p, uoff.data(), (nu + 1) * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->u_tmp_off.p, toff.data(), nu * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->u_flags.p, uflags.data(), nu, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->u_begin.p, ubegin.data(), (n + 1) * 4, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));              // the vectors above die with this frame
    CK(cudaEventRecord(ctx->ev0, s));
    bdf::DeflateArgs a;
    a.in = d.in; a.in_off = (const uint64_t *)ctx->u_in_off.p;
    a.out = (uint8_t *)ctx->u_tmp.p; a.out_off = (const uint64_t *)ctx->u_tmp_off.p;
    a.out_size = (uint64_t *)ctx->u_size.p; a.status = (int32_t *)ctx->u_status.p;
    a.n = (uint32_t)nu; a.level = level > 12 ? 12 : level; a.format = BDF_RAW;
    a.unit_flags = (const uint8_t *)ctx->u_flags.p;
    if ((rc = run_deflate(ctx, a, s, max_unit))) return rc;
    bdf::JoinArgs j;
    j.in = d.in; j.in_off = d.in_off;
    j.unit_begin = (const uint32_t *)ctx->u_begin.p; j.tmp = (const uint8_t *)ctx->u_tmp.p;
    j.tmp_off = (const uint64_t *)ctx->u_tmp_off.p; j.unit_size = (const uint64_t *)ctx->u_size.p;
    j.unit_status = (const int32_t *)ctx->u_status.p; j.out = d.out;
    j.out_off = d.out_off; j.out_size = d.out_size;
    j.status = d.status; j.n = (uint32_t)n; j.level = a.level; j.format = format;
    unsigned long long want = (n + bdf::JOIN_WARPS - 1) / bdf::JOIN_WARPS;
    unsigned long long full = (unsigned long long)ctx->sm_count * 16;
    bdf::deflate_join_kernel<<<(unsigned)(want < full ? want : full), bdf::JOIN_WARPS * 32, 0, s>>>(j);
    ctx->launches++;
    CK(cudaGetLastError());
    return BDF_E_OK;
}

// Any stream length on device-resident data.  The kernels to run (64 KiB instances, or 256 KiB units +
// join) and the size of the unit slab depend on the lengths, so the offsets are read back once
// (8 (n + 1) bytes and one synchronisation of `stream`); everything else stays on the device and the
// call returns with the work enqueued, like bdf_compress_batch_device.
int bdf_compress_batch_device_any(bdf_ctx *ctx, int level, int format, const uint8_t *in, const uint64_t *in_off,
                                  size_t n, uint8_t *out, const uint64_t *out_off, uint64_t *out_size,
                                  int32_t *status, void *stream)
{
    if (!ctx) return BDF_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (bad_format(format)) return fail(ctx, BDF_E_ARG, "unknown format");
    if (level < 0) return fail(ctx, BDF_E_ARG, "negative level");
    if (n == 0) return BDF_E_OK;
    if (!in || !in_off || !out || !out_off || !out_size || !status) return fail(ctx, BDF_E_ARG, "null pointer");
    if (n > 0xFFFFFFF0ull) return fail(ctx, BDF_E_ARG, "too many streams");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    std::vector<uint64_t> h_off(n + 1);
    CK(cudaMemcpyAsync(h_off.data(), in_off, (n + 1) * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    uint64_t max_len = 0;
    for (size_t i = 0; i < n; i++) {
        if (h_off[i] > h_off[i + 1]) return fail(ctx, BDF_E_ARG, "in_off is not ascending");
        if (h_off[i + 1] - h_off[i] > max_len) max_len = h_off[i + 1] - h_off[i];
    }
    if (h_off[n] > BDF_MAX_SLAB_BYTES) return fail(ctx, BDF_E_ARG, "input too large");
    if (max_len <= (level == 0 ? 256u * 1024u : 65536u))
        return compress_device_locked(ctx, level, format, in, in_off, n, out, out_off, out_size, status, s);
    const DevBatch d{in, in_off, out, out_off, out_size, status};
    return compress_chunked_locked(ctx, level, format, h_off.data(), n, s, d);
}

// ---- host compress calls.  One implementation behind three entry points (flat input + bound-spaced
// output, flat input + dense output, scattered input + dense output):
//   * the input goes to the device in chunks on a copy stream while the kernels of earlier chunks
//     run (a scattered input is first packed into two pinned staging buffers, chunk by chunk);
//   * the result is packed on the device (size scan + gather_kernel) and comes back as ONE copy of
//     what was produced — not one copy per stream, and never the bound-sized slab.
struct HostInput {
    const uint8_t *flat;               // flat input, or
    const uint8_t *const *ptrs;        // n scattered buffers
    const uint64_t *off;               // n + 1 offsets into the (virtual) flat input
};

static int ensure_host(bdf_ctx *ctx, void *&p, size_t &cap, size_t bytes)
{
    if (cap >= bytes) return 0;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return fail(ctx, BDF_E_NOMEM, "cudaHostAlloc", e);
    cap = bytes;
    return 0;
}

// Packs the bound-spaced slab into ctx->dense; leaves offsets (n + 1) in ctx->dense_off on the
// device and in `dense_off` on the host (synchronises the stream).
static int pack_result_locked(bdf_ctx *ctx, size_t n, uint64_t *dense_off, cudaStream_t s)
{
    int rc;
    if ((rc = ensure(ctx, ctx->dense_off, (n + 1) * 8))) return rc;
    bdf::size_scan_kernel<<<1, bdf::SCAN_THREADS, 0, s>>>((const uint64_t *)ctx->out_size.p, (uint64_t *)ctx->dense_off.p, (uint32_t)n);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(dense_off, ctx->dense_off.p, (n + 1) * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    const size_t total = (size_t)dense_off[n];
    if ((rc = ensure(ctx, ctx->dense, total + 16))) return rc;
    bdf::GatherArgs ga{(const uint8_t *)ctx->out.p, (const uint64_t *)ctx->out_off.p, (const uint64_t *)ctx->out_size.p,
                       (uint8_t *)ctx->dense.p, (const uint64_t *)ctx->dense_off.p, (uint32_t)n};
    unsigned long long want = (n + bdf::GATHER_WARPS - 1) / bdf::GATHER_WARPS;
    unsigned long long full = (unsigned long long)ctx->sm_count * 16;
    bdf::gather_kernel<<<(unsigned)(want < full ? want : full), bdf::GATHER_WARPS * 32, 0, s>>>(ga);
    ctx->launches++;
    CK(cudaGetLastError());
    return 0;
}

// dense != nullptr: dense result (dense_cap bytes, offsets to dense_off[n + 1]); else the result is
// scattered into the caller's bound-spaced slots out + out_off[i] and sizes go to out_size.
static int compress_host_impl(bdf_ctx *ctx, int level, int format, const HostInput &hin, size_t n,
                              uint8_t *dense, size_t dense_cap, uint64_t *dense_off,
                              uint8_t *out, const uint64_t *out_off, uint64_t *out_size, int32_t *status)
{
    const uint64_t *in_off = hin.off;
    uint64_t max_len = 0;
    for (size_t i = 0; i < n; i++) {
        if (in_off[i] > in_off[i + 1]) return fail(ctx, BDF_E_ARG, "in_off is not ascending");
        if (in_off[i + 1] - in_off[i] > max_len) max_len = in_off[i + 1] - in_off[i];
    }
    if (in_off[n] > BDF_MAX_SLAB_BYTES) return fail(ctx, BDF_E_ARG, "input too large");
    // level 0 needs the unit path only above one chunk (the stored kernel handles any length)
    const bool chunked = max_len > (level == 0 ? 256u * 1024u : 65536u);
    // one lock for the whole call: the staging buffers of the ctx are shared by all callers
    std::lock_guard<std::mutex> g(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t in_bytes = (size_t)in_off[n];
    // device slab at the reference's bounds (src/batch.rs:39), 16-byte aligned slots
    std::vector<uint64_t> slab_off(n);
    size_t slab_bytes = 0;
    for (size_t i = 0; i < n; i++) {
        slab_off[i] = slab_bytes;
        slab_bytes += (bdf_compress_bound(format, (size_t)(in_off[i + 1] - in_off[i])) + 15) & ~(size_t)15;
    }
    if (dense) {
        if (overlaps(hin.flat, hin.flat ? in_bytes : 0, dense, dense_cap)) return fail(ctx, BDF_E_ARG, "Input and output buffers overlap");
    } else {
        size_t out_bytes = 0;
        for (size_t i = 0; i < n; i++) {
            const uint64_t e = out_off[i] + bdf_compress_bound(format, (size_t)(in_off[i + 1] - in_off[i]));
            if (e < out_off[i] || e > BDF_MAX_SLAB_BYTES) return fail(ctx, BDF_E_ARG, "out_off overflows");
            if (e > out_bytes) out_bytes = (size_t)e;
        }
        if (overlaps(hin.flat, hin.flat ? in_bytes : 0, out, out_bytes)) return fail(ctx, BDF_E_ARG, "Input and output buffers overlap");
    }
    int rc;
    if ((rc = ensure(ctx, ctx->in, in_bytes + 8)) || (rc = ensure(ctx, ctx->out, slab_bytes + 8)) ||
        (rc = ensure(ctx, ctx->in_off, (n + 1) * 8)) || (rc = ensure(ctx, ctx->out_off, n * 8)) ||
        (rc = ensure(ctx, ctx->out_size, n * 8)) || (rc = ensure(ctx, ctx->status, n * 4)))
        return rc;
    CK(cudaMemcpyAsync(ctx->in_off.p, in_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->out_off.p, slab_off.data(), n * 8, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));                  // slab_off is read by the copy above; also orders the staging reuse below
    // ---- input chunks: copy stream ahead of the kernels
    const size_t CHUNK = 64u << 20;
    if (hin.ptrs && (rc = ensure_host(ctx, ctx->h_stage[0], ctx->h_stage_cap[0], CHUNK)) ) return rc;
    if (hin.ptrs && (rc = ensure_host(ctx, ctx->h_stage[1], ctx->h_stage_cap[1], CHUNK)) ) return rc;
    cudaStream_t cs = ctx->copy_streams[0];
    CK(cudaEventRecord(ctx->ev0, s));
    unsigned chunk_no = 0;
    // kernels for streams [a, b) once their bytes are on their way (event on the copy stream)
    auto launch_range = [&](size_t a, size_t b) -> int {
        cudaEvent_t ev = ctx->ev_chunk[chunk_no++ % BDF_CHUNK_EVENTS];
        CK(cudaEventRecord(ev, cs));
        CK(cudaStreamWaitEvent(s, ev, 0));
        if (chunked || b <= a) return 0;
        return compress_device_locked(ctx, level, format, (const uint8_t *)ctx->in.p,
                                      (const uint64_t *)ctx->in_off.p + a, b - a, (uint8_t *)ctx->out.p,
                                      (const uint64_t *)ctx->out_off.p + a, (uint64_t *)ctx->out_size.p + a,
                                      (int32_t *)ctx->status.p + a, s);
    };
    if (hin.flat) {
        size_t a = 0;
        while (a < n) {
            size_t b = a;
            while (b < n && (b == a || in_off[b + 1] - in_off[a] <= CHUNK)) b++;
            const size_t beg = (size_t)in_off[a], bytes = (size_t)in_off[b] - beg;
            if (bytes) CK(cudaMemcpyAsync((uint8_t *)ctx->in.p + beg, hin.flat + beg, bytes, cudaMemcpyHostToDevice, cs));
            if ((rc = launch_range(a, b))) return rc;
            a = b;
        }
    } else {
        // pack the buffers into the two pinned staging buffers in turn; a kernel starts for the
        // streams that are complete after each buffer
        size_t k = 0, koff = 0, launched = 0, dev_pos = 0;
        while (k < n) {
            const unsigned b = ctx->stage_no++ & 1u;
            uint8_t *st = (uint8_t *)ctx->h_stage[b];
            CK(cudaEventSynchronize(ctx->ev_stage[b]));           // the copy that read this buffer last
            size_t fill = 0;
            while (k < n) {
                const size_t len = (size_t)(in_off[k + 1] - in_off[k]);
                size_t cnt = len - koff;
                if (cnt > CHUNK - fill) cnt = CHUNK - fill;
                if (cnt) memcpy(st + fill, hin.ptrs[k] + koff, cnt);
                fill += cnt;
                koff += cnt;
                if (koff < len) break;                            // buffer full in the middle of a stream
                k++;
                koff = 0;
                if (fill == CHUNK) break;
            }
            if (fill) CK(cudaMemcpyAsync((uint8_t *)ctx->in.p + dev_pos, st, fill, cudaMemcpyHostToDevice, cs));
            CK(cudaEventRecord(ctx->ev_stage[b], cs));
            dev_pos += fill;
            if (k > launched) {
                if ((rc = launch_range(launched, k))) return rc;
                launched = k;
            }
        }
    }
    if (chunked) {
        const DevBatch d{(const uint8_t *)ctx->in.p, (const uint64_t *)ctx->in_off.p, (uint8_t *)ctx->out.p,
                         (const uint64_t *)ctx->out_off.p, (uint64_t *)ctx->out_size.p, (int32_t *)ctx->status.p};
        rc = compress_chunked_locked(ctx, level, format, in_off, n, s, d);
        if (rc) return rc;
    }
    CK(cudaEventRecord(ctx->ev1, s));
    CK(cudaMemcpyAsync(status, ctx->status.p, n * 4, cudaMemcpyDeviceToHost, s));
    // ---- result: pack on the device, one copy back
    std::vector<uint64_t> tmp_off;
    uint64_t *doff = dense_off;
    if (!dense) { tmp_off.resize(n + 1); doff = tmp_off.data(); }
    if ((rc = pack_result_locked(ctx, n, doff, s))) return rc;
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    const size_t total = (size_t)doff[n];
    if (dense) {
        if (total > dense_cap) return fail(ctx, BDF_E_ARG, "dense output buffer too small (needed size is in out_off[n])");
        if (total) CK(cudaMemcpyAsync(dense, ctx->dense.p, total, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        return BDF_E_OK;
    }
    if ((rc = ensure_host(ctx, ctx->h_result, ctx->h_result_cap, total + 16))) return rc;
    if (total) CK(cudaMemcpyAsync(ctx->h_result, ctx->dense.p, total, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (size_t i = 0; i < n; i++) {
        const size_t sz = (size_t)(doff[i + 1] - doff[i]);
        out_size[i] = sz;
        if (sz) memcpy(out + out_off[i], (const uint8_t *)ctx->h_result + doff[i], sz);
    }
    return BDF_E_OK;
}

static int compress_host_args(bdf_ctx *ctx, int level, int format, size_t n)
{
    if (bad_format(format)) return fail(ctx, BDF_E_ARG, "unknown format");
    if (level < 0) return fail(ctx, BDF_E_ARG, "negative level");
    if (n > 0xFFFFFFF0ull) return fail(ctx, BDF_E_ARG, "too many streams");
    return 0;
}

int bdf_compress_batch_host(bdf_ctx *ctx, int level, int format, const uint8_t *in, const uint64_t *in_off,
                            size_t n, uint8_t *out, const uint64_t *out_off, uint64_t *out_size, int32_t *status)
{
    if (!ctx) return BDF_E_ARG;
    int rc = compress_host_args(ctx, level, format, n);
    if (rc) return rc;
    if (n == 0) return BDF_E_OK;
    if (!in || !in_off || !out || !out_off || !out_size || !status) return fail(ctx, BDF_E_ARG, "null pointer");
    HostInput hin{in, nullptr, in_off};
    return compress_host_impl(ctx, level, format, hin, n, nullptr, 0, nullptr, out, out_off, out_size, status);
}

int bdf_compress_batch_host_dense(bdf_ctx *ctx, int level, int format, const uint8_t *in, const uint64_t *in_off,
                                  size_t n, uint8_t *out, size_t out_cap, uint64_t *out_off, int32_t *status)
{
    if (!ctx) return BDF_E_ARG;
    int rc = compress_host_args(ctx, level, format, n);
    if (rc) return rc;
    if (!out_off) return fail(ctx, BDF_E_ARG, "null pointer");
    if (n == 0) { out_off[0] = 0; return BDF_E_OK; }
    if (!in || !in_off || !out || !status) return fail(ctx, BDF_E_ARG, "null pointer");
    HostInput hin{in, nullptr, in_off};
    return compress_host_impl(ctx, level, format, hin, n, out, out_cap, out_off, nullptr, nullptr, nullptr, status);
}

int bdf_compress_batch_host_sg(bdf_ctx *ctx, int level, int format, const uint8_t *const *in_ptrs,
                               const size_t *in_lens, size_t n, uint8_t *out, size_t out_cap, uint64_t *out_off,
                               int32_t *status)
{
    if (!ctx) return BDF_E_ARG;
    int rc = compress_host_args(ctx, level, format, n);
    if (rc) return rc;
    if (!out_off) return fail(ctx, BDF_E_ARG, "null pointer");
    if (n == 0) { out_off[0] = 0; return BDF_E_OK; }
    if (!in_ptrs || !in_lens || !out || !status) return fail(ctx, BDF_E_ARG, "null pointer");
    std::vector<uint64_t> off(n + 1);
    off[0] = 0;
    for (size_t i = 0; i < n; i++) {
        if (in_lens[i] && !in_ptrs[i]) return fail(ctx, BDF_E_ARG, "null pointer");
        if (overlaps(in_ptrs[i], in_lens[i], out, out_cap)) return fail(ctx, BDF_E_ARG, "Input and output buffers overlap");
        off[i + 1] = off[i] + in_lens[i];
        if (off[i + 1] < off[i]) return fail(ctx, BDF_E_ARG, "input too large");
    }
    HostInput hin{nullptr, in_ptrs, off.data()};
    return compress_host_impl(ctx, level, format, hin, n, out, out_cap, out_off, nullptr, nullptr, nullptr, status);
}

// --------------------------------------------------------------------- units
int bdf_compress_units_host(bdf_ctx *ctx, int level, const uint8_t *in, const uint64_t *unit_off,
                            const uint8_t *flush, size_t n, uint8_t *out, const uint64_t *out_off,
                            uint64_t *out_size, int32_t *status)
{
    if (!ctx) return BDF_E_ARG;
    if (level < 0) return fail(ctx, BDF_E_ARG, "negative level");
    if (n == 0) return BDF_E_OK;
    if (!in || !unit_off || !flush || !out || !out_off || !out_size || !status)
        return fail(ctx, BDF_E_ARG, "null pointer");
    if (n > 0xFFFFFFF0ull) return fail(ctx, BDF_E_ARG, "too many units");
    std::lock_guard<std::mutex> g(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    std::vector<uint8_t> uflags(n);
    uint64_t max_unit = 0;
    size_t out_bytes = 0;
    for (size_t i = 0; i < n; i++) {
        const uint64_t ul = unit_off[i + 1] - unit_off[i];
        if (ul > 256u * 1024u) return fail(ctx, BDF_E_ARG, "a unit is at most 262144 bytes");
        if (flush[i] != BDF_FLUSH_SYNC && flush[i] != BDF_FLUSH_FINISH) return fail(ctx, BDF_E_ARG, "unknown flush mode");
        // DeflateEncoder gives a chunk that ends in a sync flush 5 more bytes of room (src/stream.rs:66-69,113-116)
        uflags[i] = flush[i] == BDF_FLUSH_FINISH ? (uint8_t)bdf::UNIT_FINISH : (uint8_t)(bdf::UNIT_SYNC | bdf::UNIT_CAP5);
        if (ul > max_unit) max_unit = ul;
        const size_t e = (size_t)out_off[i] + (size_t)bdf::unit_cap(ul, uflags[i]);
        if (e > out_bytes) out_bytes = e;
    }
    const size_t in_bytes = (size_t)unit_off[n];
    if (overlaps(in, in_bytes, out, out_bytes)) return fail(ctx, BDF_E_ARG, "Input and output buffers overlap");
    int rc;
    if ((rc = ensure(ctx, ctx->in, in_bytes + 8)) || (rc = ensure(ctx, ctx->out, out_bytes + 8)) ||
        (rc = ensure(ctx, ctx->in_off, (n + 1) * 8)) || (rc = ensure(ctx, ctx->out_off, n * 8)) ||
        (rc = ensure(ctx, ctx->out_size, n * 8)) || (rc = ensure(ctx, ctx->status, n * 4)) ||
        (rc = ensure(ctx, ctx->u_flags, n)))
        return rc;
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->in.p, in, in_bytes, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_off.p, unit_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->out_off.p, out_off, n * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->u_flags.p, uflags.data(), n, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(ctx->ev0, s));
    bdf::DeflateArgs a;
    a.in = (const uint8_t *)ctx->in.p; a.in_off = (const uint64_t *)ctx->in_off.p;
    a.out = (uint8_t *)ctx->out.p; a.out_off = (const uint64_t *)ctx->out_off.p;
    a.out_size = (uint64_t *)ctx->out_size.p; a.status = (int32_t *)ctx->status.p;
    a.n = (uint32_t)n; a.level = level > 12 ? 12 : level; a.format = BDF_RAW;
    a.unit_flags = (const uint8_t *)ctx->u_flags.p;
    if ((rc = run_deflate(ctx, a, s, max_unit))) return rc;
    CK(cudaEventRecord(ctx->ev1, s));
    CK(cudaMemcpyAsync(out_size, ctx->out_size.p, n * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(status, ctx->status.p, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    for (size_t i = 0; i < n; i++)
        if (out_size[i])
            CK(cudaMemcpyAsync(out + out_off[i], (const uint8_t *)ctx->out.p + out_off[i], (size_t)out_size[i],
                               cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return BDF_E_OK;
}

// ---------------------------------------------------------- size estimation
static int compress_size_locked(bdf_ctx *ctx, int level, const uint8_t *in, const uint64_t *in_off, size_t n,
                                int final_block, uint64_t *out_size, int32_t *status, uint64_t max_len,
                                cudaStream_t s)
{
    bdf::DeflateArgs a;
    a.in = in; a.in_off = in_off; a.out = nullptr; a.out_off = nullptr; a.out_size = out_size; a.status = status;
    a.n = (uint32_t)n; a.level = level > 12 ? 12 : level; a.format = BDF_RAW; a.unit_flags = nullptr;
    a.size_only = 1; a.final_block = final_block ? 1 : 0;
    return run_deflate(ctx, a, s, max_len);
}

int bdf_compress_size_batch_device(bdf_ctx *ctx, int level, const uint8_t *in, const uint64_t *in_off,
                                   size_t n, int final_block, uint64_t *out_size, int32_t *status,
                                   void *stream)
{
    if (!ctx) return BDF_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (level < 0) return fail(ctx, BDF_E_ARG, "negative level");
    if (n == 0) return BDF_E_OK;
    if (!in || !in_off || !out_size || !status) return fail(ctx, BDF_E_ARG, "null pointer");
    if (n > 0xFFFFFFF0ull) return fail(ctx, BDF_E_ARG, "too many streams");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    return compress_size_locked(ctx, level, in, in_off, n, final_block, out_size, status, 0, s);
}

int bdf_compress_size_batch_host(bdf_ctx *ctx, int level, const uint8_t *in, const uint64_t *in_off,
                                 size_t n, int final_block, uint64_t *out_size, int32_t *status)
{
    if (!ctx) return BDF_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (level < 0) return fail(ctx, BDF_E_ARG, "negative level");
    if (n == 0) return BDF_E_OK;
    if (!in || !in_off || !out_size || !status) return fail(ctx, BDF_E_ARG, "null pointer");
    if (n > 0xFFFFFFF0ull) return fail(ctx, BDF_E_ARG, "too many streams");
    uint64_t max_len = 0;
    for (size_t i = 0; i < n; i++) {
        const uint64_t l = in_off[i + 1] - in_off[i];
        if (l > max_len) max_len = l;
    }
    if (level > 0 && max_len > 256u * 1024u) return fail(ctx, BDF_E_ARG, "a buffer is at most 262144 bytes");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t in_bytes = (size_t)in_off[n];
    int rc;
    if ((rc = ensure(ctx, ctx->in, in_bytes + 8)) || (rc = ensure(ctx, ctx->in_off, (n + 1) * 8)) ||
        (rc = ensure(ctx, ctx->out_size, n * 8)) || (rc = ensure(ctx, ctx->status, n * 4)))
        return rc;
    CK(cudaMemcpyAsync(ctx->in.p, in, in_bytes, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_off.p, in_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(ctx->ev0, s));
    rc = compress_size_locked(ctx, level, (const uint8_t *)ctx->in.p, (const uint64_t *)ctx->in_off.p, n,
                              final_block, (uint64_t *)ctx->out_size.p, (int32_t *)ctx->status.p, max_len, s);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, s));
    CK(cudaMemcpyAsync(out_size, ctx->out_size.p, n * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(status, ctx->status.p, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    return BDF_E_OK;
}

// ------------------------------------------------------------------ checksum
static int checksum_device_locked(bdf_ctx *ctx, int kind, const uint8_t *in, const uint64_t *in_off, size_t n,
                                  uint32_t *out, cudaStream_t s)
{
    bdf::ChecksumArgs a{in, in_off, out, (uint32_t)n};
    unsigned long long want = (n + bdf::CK_WARPS_PER_BLOCK - 1) / bdf::CK_WARPS_PER_BLOCK;
    unsigned long long full = (unsigned long long)ctx->sm_count * 8;
    unsigned grid = (unsigned)(want < full ? want : full);
    if (kind == BDF_CRC32)
        bdf::checksum_kernel<BDF_CRC32><<<grid, bdf::CK_WARPS_PER_BLOCK * 32, 0, s>>>(a);
    else
        bdf::checksum_kernel<BDF_ADLER32><<<grid, bdf::CK_WARPS_PER_BLOCK * 32, 0, s>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return BDF_E_OK;
}

int bdf_checksum_batch_device(bdf_ctx *ctx, int kind, const uint8_t *in, const uint64_t *in_off, size_t n,
                              uint32_t *out, void *stream)
{
    if (!ctx) return BDF_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (kind != BDF_ADLER32 && kind != BDF_CRC32) return fail(ctx, BDF_E_ARG, "unknown checksum kind");
    if (n == 0) return BDF_E_OK;
    if (!in || !in_off || !out) return fail(ctx, BDF_E_ARG, "null pointer");
    if (n > 0xFFFFFFF0ull) return fail(ctx, BDF_E_ARG, "too many streams");
    CK(cudaSetDevice(ctx->device));
    return checksum_device_locked(ctx, kind, in, in_off, n, out, stream ? (cudaStream_t)stream : ctx->stream);
}

int bdf_checksum_batch_host(bdf_ctx *ctx, int kind, const uint8_t *in, const uint64_t *in_off, size_t n,
                            uint32_t *out)
{
    if (!ctx) return BDF_E_ARG;
    if (kind != BDF_ADLER32 && kind != BDF_CRC32) return fail(ctx, BDF_E_ARG, "unknown checksum kind");
    if (n == 0) return BDF_E_OK;
    if (!in || !in_off || !out) return fail(ctx, BDF_E_ARG, "null pointer");
    if (n > 0xFFFFFFF0ull) return fail(ctx, BDF_E_ARG, "too many streams");
    std::lock_guard<std::mutex> g(ctx->mu);      // whole call: the staging buffers are shared
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ensure(ctx, ctx->in, (size_t)in_off[n] + 8)) || (rc = ensure(ctx, ctx->in_off, (n + 1) * 8)) ||
        (rc = ensure(ctx, ctx->checksum, n * 4)))
        return rc;
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->in.p, in, (size_t)in_off[n], cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_off.p, in_off, (n + 1) * 8, cudaMemcpyHostToDevice, s));
    CK(cudaEventRecord(ctx->ev0, s));
    rc = checksum_device_locked(ctx, kind, (const uint8_t *)ctx->in.p, (const uint64_t *)ctx->in_off.p, n,
                                (uint32_t *)ctx->checksum.p, s);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev1, s));
    CK(cudaMemcpyAsync(out, ctx->checksum.p, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1);
    return BDF_E_OK;
}

// ------------------------------------------------------------------ gather
int bdf_gather_streams_device(bdf_ctx *ctx, const uint8_t *src, const uint64_t *src_off, const uint64_t *size,
                              size_t n, uint8_t *dst, const uint64_t *dst_off, void *stream)
{
    if (!ctx) return BDF_E_ARG;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (n == 0) return BDF_E_OK;
    if (!src || !src_off || !size || !dst || !dst_off) return fail(ctx, BDF_E_ARG, "null pointer");
    if (n > 0xFFFFFFF0ull) return fail(ctx, BDF_E_ARG, "too many streams");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = stream ? (cudaStream_t)stream : ctx->stream;
    bdf::GatherArgs a{src, src_off, size, dst, dst_off, (uint32_t)n};
    unsigned long long want = (n + bdf::GATHER_WARPS - 1) / bdf::GATHER_WARPS;
    unsigned long long full = (unsigned long long)ctx->sm_count * 16;
    bdf::gather_kernel<<<(unsigned)(want < full ? want : full), bdf::GATHER_WARPS * 32, 0, s>>>(a);
    ctx->launches++;
    CK(cudaGetLastError());
    return BDF_E_OK;
}

}  // extern "C"
