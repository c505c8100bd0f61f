/*
 * bdeflate.h — C ABI of the B200-native batch DEFLATE engine (libbdeflate.so).
 *
 * This is the drop-in boundary for the batch path of 404Setup/libdeflate-rsx:
 * it is what a Rust `extern "C"` block behind src/batch.rs would bind (the
 * binding is shown in INTEGRATION.md).  Plain pointers and sizes only.
 *
 * Data layout (the crate's existing GPU convention, src/batch_cuda.rs:57-87 and
 * src/cuda/compress.cu:1-8): one flat byte buffer holding every stream back to
 * back, `in_off[n+1]` byte offsets into it, one flat output slab with
 * `out_off[n]` start offsets, and per-stream result arrays written by the
 * device.  Nothing allocated by the library crosses the boundary except
 * bdf_host_alloc() memory, which the caller frees with bdf_host_free().
 *
 * Error convention: the int return value is the CALL-level result (0 = the
 * batch ran; negative = bad argument / CUDA failure — there is NO CPU
 * fallback, unlike src/batch.rs:23-31).  Per-stream results are in-band, like
 * the reference's empty Vec / None (src/batch.rs:52-53,95-96): status[i] uses
 * the declaration order of DecompressResult (src/decompress/mod.rs:79-85) and
 * out_size[i] is 0 for a failed stream.
 */
#ifndef BDEFLATE_H
#define BDEFLATE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BDF_VERSION 0x000100 /* 0.1.0 */

/* Framing.  src/batch.rs only speaks raw DEFLATE; zlib and gzip follow
 * Compressor::compress_zlib/_gzip (src/compress/mod.rs:2248-2357) and
 * Decompressor::decompress_zlib/_gzip (src/decompress/mod.rs:1074-1240). */
enum bdf_format { BDF_RAW = 0, BDF_ZLIB = 1, BDF_GZIP = 2 };

/* Per-stream status == DecompressResult order; compression uses OK and
 * INSUFFICIENT_SPACE (CompressResult, src/compress/mod.rs:238-241). */
enum bdf_status {
    BDF_OK = 0,
    BDF_BAD_DATA = 1,
    BDF_SHORT_OUTPUT = 2, /* declared by the reference, never produced */
    BDF_INSUFFICIENT_SPACE = 3,
    BDF_SHORT_INPUT = 4,
    BDF_STREAM_UNSUPPORTED = 100 /* not a reference status: outside this build's range */
};

/* Call-level errors. */
enum bdf_error {
    BDF_E_OK = 0,
    BDF_E_ARG = -1,         /* null pointer, unknown format / level / kind */
    BDF_E_CUDA = -2,        /* a CUDA runtime call failed; see bdf_last_error */
    BDF_E_NOMEM = -3,       /* device or pinned allocation failed */
    BDF_E_UNSUPPORTED = -4  /* valid request this build does not implement */
};

enum bdf_checksum_kind { BDF_ADLER32 = 0, BDF_CRC32 = 1 };

typedef struct bdf_ctx bdf_ctx;

/* Library / device lifetime.  A ctx owns one CUDA stream, its scratch memory
 * and staging buffers on `device`; calls on one ctx are serialised by an
 * internal mutex, so a ctx may be shared between threads the way
 * BatchCompressor is `Sync` (src/batch.rs:5-18), and several ctxs (one per
 * GPU) may be driven concurrently. */
int bdf_version(void);
int bdf_device_count(void);
int bdf_ctx_create(int device, bdf_ctx **ctx);
void bdf_ctx_destroy(bdf_ctx *ctx);
const char *bdf_last_error(const bdf_ctx *ctx);
/* Debug builds of the library (-DBDF_CHECK, libbdeflate_check.so: bounds / invariant assertions in
 * the kernels): source line | 0x80000000 of the first assertion that failed on this process's
 * device, 0 if none; -1 from a build without the assertions. */
long long bdf_debug_check_failures(bdf_ctx *ctx);
/* Number of engine kernels launched through this ctx so far. */
uint64_t bdf_kernel_launches(const bdf_ctx *ctx);
/* Device time (ms, CUDA events on the ctx stream) of the kernels of the last
 * *_host call; 0 if none. */
float bdf_last_kernel_ms(const bdf_ctx *ctx);

/* Pinned host memory for the *_host entry points (pageable memory also works,
 * through the driver's staging copies).  The pages are placed on the NUMA node
 * of the calling thread's current CUDA device (the thread's CPU affinity is
 * narrowed to that node for the duration of the call; BDF_HOST_ALLOC_NUMA=0
 * disables it): call it after cudaSetDevice / bdf_ctx_create for that GPU. */
void *bdf_host_alloc(size_t bytes);
void bdf_host_free(void *p);

/* Compressor::{deflate,zlib,gzip}_compress_bound, src/compress/mod.rs:2236-2246:
 * len + (len / 65535 + 1) * 5 + 10, + 6 (zlib) or + 18 (gzip). */
size_t bdf_compress_bound(int format, size_t len);

/*
 * BatchDecompressor::decompress_batch (src/batch.rs:74-101).
 *   stream i : in[in_off[i] .. in_off[i+1])  ->  out[out_off[i] ..), at most max_out[i] bytes
 *   out_size[i] : bytes produced (0 on failure); status[i] : bdf_status
 *   checksum[i] (may be NULL): Adler-32 (zlib) / CRC-32 (gzip) of the output, 0 for raw
 * Like the reference, trailing input after the final block is ignored and
 * producing fewer than max_out[i] bytes is success.  The *_host calls reject input and output
 * ranges that overlap (is_overlapping, src/api.rs:303-314) with BDF_E_ARG.
 * *_device: every pointer is device memory valid on the ctx's device; the
 * work is enqueued on `stream` (a cudaStream_t; NULL means the ctx's own
 * stream — pass cudaStreamLegacy / cudaStreamPerThread explicitly to target
 * the default streams) and the call returns without synchronising.  *_host: every pointer is host
 * memory; the call copies in, runs, copies out and returns when the results
 * are in place.
 */
int bdf_decompress_batch_device(bdf_ctx *ctx, int format, const uint8_t *in,
                                const uint64_t *in_off, size_t n, uint8_t *out,
                                const uint64_t *out_off, const uint64_t *max_out,
                                uint64_t *out_size, uint32_t *checksum,
                                int32_t *status, void *stream);
int bdf_decompress_batch_host(bdf_ctx *ctx, int format, const uint8_t *in,
                              const uint64_t *in_off, size_t n, uint8_t *out,
                              const uint64_t *out_off, const uint64_t *max_out,
                              uint64_t *out_size, uint32_t *checksum,
                              int32_t *status);

/*
 * BatchCompressor::compress_batch (src/batch.rs:20-58) for `level` 0..12
 * (Compressor::new, src/compress/mod.rs:459-507; values above 12 behave as 12).
 * Stream i may use up to bdf_compress_bound(format, len_i) bytes at
 * out + out_off[i]; a stream whose encoding does not fit fails with
 * BDF_INSUFFICIENT_SPACE and out_size[i] = 0 — the reference has no
 * stored-block fallback (src/compress/mod.rs:641-644).
 * Stream length: the *_host call takes any length.  Streams above 64 KiB run through the
 * 256 KiB kernel instances; streams above 256 KiB are cut into 256 KiB chunks, each compressed
 * by a fresh compressor, all but the last followed by a sync flush, like Compressor::compress
 * (src/compress/mod.rs:699-772).  Levels 0..9 are byte-identical to this repository's restatement of
 * the reference compressor (oracle/; outputs of the real Rust binary cannot be captured here: no
 * toolchain); levels 10..12 on streams of at most 64 KiB are held to its size (within 0.5 %), not
 * its bytes.
 * The *_device call cannot see the lengths without synchronising: it runs the 64 KiB instances
 * and sets status[i] = BDF_STREAM_UNSUPPORTED for a longer stream (level 0 takes any length up
 * to 256 KiB there); bdf_compress_batch_device_any is the device-pointer call for any length.
 */
int bdf_compress_batch_device(bdf_ctx *ctx, int level, int format,
                              const uint8_t *in, const uint64_t *in_off, size_t n,
                              uint8_t *out, const uint64_t *out_off,
                              uint64_t *out_size, int32_t *status, void *stream);
int bdf_compress_batch_host(bdf_ctx *ctx, int level, int format, const uint8_t *in,
                            const uint64_t *in_off, size_t n, uint8_t *out,
                            const uint64_t *out_off, uint64_t *out_size,
                            int32_t *status);
/* Device pointers, any stream length (Compressor::compress takes any length,
 * src/compress/mod.rs:699-772): like bdf_compress_batch_device, but the call first reads the
 * offsets back (8 (n + 1) bytes, one synchronisation of `stream`) to choose the kernel instances and
 * size the slab of the 256 KiB units; the data never leaves the device and the call returns with
 * the work enqueued.  Same bytes as the *_host call. */
int bdf_compress_batch_device_any(bdf_ctx *ctx, int level, int format,
                                  const uint8_t *in, const uint64_t *in_off, size_t n,
                                  uint8_t *out, const uint64_t *out_off,
                                  uint64_t *out_size, int32_t *status, void *stream);

/*
 * The same call with the result packed: stream i is out[out_off[i] .. out_off[i+1]) (out_off has
 * n + 1 entries, written by the call; a failed stream is empty, like the reference's empty Vec,
 * src/batch.rs:52-53).  This is what a `Vec<Vec<u8>>` result is sliced from (src/batch.rs:34-57,
 * src/batch_cuda.rs:116-141) without a bound-sized host slab: out only needs room for what the batch
 * compresses to (out_cap bytes; BDF_E_ARG if it is too small, the needed size is then in out_off[n]).
 * The result is packed on the device and comes back in one copy.
 * bdf_compress_batch_host_sg takes the input the way `&[&[u8]]` holds it — n pointers and n lengths —
 * so the caller does not have to flatten it: the library packs the buffers into pinned staging
 * memory chunk by chunk while earlier chunks are already on their way to the device.
 */
int bdf_compress_batch_host_dense(bdf_ctx *ctx, int level, int format, const uint8_t *in,
                                  const uint64_t *in_off, size_t n, uint8_t *out, size_t out_cap,
                                  uint64_t *out_off, int32_t *status);
int bdf_compress_batch_host_sg(bdf_ctx *ctx, int level, int format, const uint8_t *const *in_ptrs,
                               const size_t *in_lens, size_t n, uint8_t *out, size_t out_cap,
                               uint64_t *out_off, int32_t *status);

/*
 * Compressor::compress(chunk, out, FlushMode) for many chunks at once (src/compress/mod.rs:693-790)
 * — the call DeflateEncoder::flush_buffer makes for every 256 KiB chunk of its buffer
 * (src/stream.rs:42-196).  Unit i = in[unit_off[i] .. unit_off[i+1]) (at most 262144 bytes),
 * raw DEFLATE, each through a fresh compressor; flush[i] = BDF_FLUSH_FINISH ends the stream
 * (last block has BFINAL), BDF_FLUSH_SYNC ends the unit with a sync flush (empty stored block
 * 00 00 FF FF, :662-681).  Room per unit: bdf_compress_bound(BDF_RAW, len), plus 5 for a sync
 * unit, like the encoder's own buffers (src/stream.rs:66-69).
 */
typedef enum bdf_flush { BDF_FLUSH_SYNC = 1, BDF_FLUSH_FINISH = 2 } bdf_flush;
int bdf_compress_units_host(bdf_ctx *ctx, int level, const uint8_t *in, const uint64_t *unit_off,
                            const uint8_t *flush, size_t n, uint8_t *out, const uint64_t *out_off,
                            uint64_t *out_size, int32_t *status);

/*
 * Compressor::compress_to_size(input, final_block) for many buffers at once
 * (src/compress/mod.rs:792-1094): out_size[i] = the reference's estimate, in bytes, of buffer i as
 * raw DEFLATE at `level` — what a caller sizes an output slab with instead of
 * bdf_compress_bound.  No output slab is needed: the compression kernels run with their bit
 * sinks in counting mode.  The estimate equals the compressed size at levels 2..9; at level 0 it
 * is the closed form of :1073-1082 (final_block adds the empty final block of an empty input);
 * at level 1 it counts the blocks the split statistics would cut (the compressor itself keeps
 * one block up to 64 KiB); at levels 10..12 it follows the estimator's own two-parse cost
 * seeding; an empty input costs 0 bytes at levels 1..12.
 * Buffer length: at most 262144 bytes through the *_host call; the *_device call runs the 64 KiB
 * instances (262144 at levels 0 and 1) and sets BDF_STREAM_UNSUPPORTED beyond.
 */
int bdf_compress_size_batch_device(bdf_ctx *ctx, int level, const uint8_t *in, const uint64_t *in_off,
                                   size_t n, int final_block, uint64_t *out_size, int32_t *status,
                                   void *stream);
int bdf_compress_size_batch_host(bdf_ctx *ctx, int level, const uint8_t *in, const uint64_t *in_off,
                                 size_t n, int final_block, uint64_t *out_size, int32_t *status);

/*
 * Decompressor::decompress_streaming(input, window, &mut write_pos) -> (result, in_consumed,
 * out_produced) for many decoders at once (src/decompress/mod.rs:204-372) — the call
 * DeflateDecoder::read makes whenever its window runs dry (src/stream.rs:263-376).  Raw DEFLATE.
 *
 * Decoder i continues where its state stands: it reads in[in_off[i] .. in_off[i+1]) — the bytes the
 * caller holds that the decoder has not consumed yet — and appends to its window
 * window[win_off[i] .. win_off[i] + win_cap[i]) at win_pos[i].  The bytes below win_pos[i] are the
 * output so far (match history: the caller keeps at least the last 32768 bytes there);
 * win_cap[i] must be at least 32768 + 258.  On return win_pos[i] has advanced by what was produced,
 * in_consumed[i] is the number of input bytes used up (drop them before the next call) and
 * status[i] says why the step stopped:
 *   BDF_OK                  the final block has ended (the state stays "done");
 *   BDF_SHORT_INPUT         more input is needed.  A decoder asks for up to 570 bytes in front of a
 *                           dynamic block header; set in_final[i] != 0 once no more input exists —
 *                           then BDF_SHORT_INPUT means the stream is truncated;
 *   BDF_INSUFFICIENT_SPACE  fewer than 258 bytes of room left: drain / shift the window;
 *   BDF_BAD_DATA            invalid stream (sticky).
 * A zeroed bdf_inflate_state is the start of a stream.  The state is plain data (368 bytes): it can
 * be stored, copied or moved between calls and ctxs.  Output bytes do not depend on how the input
 * and the window were cut.
 * Engine: one lane per decoder state (csrc/inflate_resume.cuh) — built for MANY concurrent
 * decoders, not for the speed of a single one.
 */
typedef struct bdf_inflate_state {
    uint32_t phase;       /* 0 block header next, 1 stored block body, 2 Huffman block body, 3 done, 4 failed */
    uint32_t final_block; /* BFINAL of the current block */
    uint32_t bit_off;     /* bits of the next input byte already consumed (0..7) */
    uint32_t stored_rem;  /* bytes left in a stored block */
    uint32_t nlit, noff;  /* code counts of the current Huffman block */
    uint32_t reserved[2];
    uint64_t total_in, total_out;
    uint8_t lens[320];    /* code lengths of the current Huffman block */
} bdf_inflate_state;
int bdf_inflate_resume_batch_device(bdf_ctx *ctx, size_t n, bdf_inflate_state *states, const uint8_t *in,
                                    const uint64_t *in_off, const uint8_t *in_final, uint8_t *window,
                                    const uint64_t *win_off, const uint64_t *win_cap, uint64_t *win_pos,
                                    uint64_t *in_consumed, int32_t *status, void *stream);
/* host pointers; only window bytes [old win_pos, new win_pos) are written back, and the history the
 * step can refer to (the 32768 bytes below win_pos) is what goes to the device */
int bdf_inflate_resume_batch_host(bdf_ctx *ctx, size_t n, bdf_inflate_state *states, const uint8_t *in,
                                  const uint64_t *in_off, const uint8_t *in_final, uint8_t *window,
                                  const uint64_t *win_off, const uint64_t *win_cap, uint64_t *win_pos,
                                  uint64_t *in_consumed, int32_t *status);

/* adler32(1, data) / crc32(0, data) per stream (src/adler32/mod.rs:114-152,
 * src/crc32/mod.rs:331-365). */
int bdf_checksum_batch_device(bdf_ctx *ctx, int kind, const uint8_t *in,
                              const uint64_t *in_off, size_t n, uint32_t *out,
                              void *stream);
int bdf_checksum_batch_host(bdf_ctx *ctx, int kind, const uint8_t *in,
                            const uint64_t *in_off, size_t n, uint32_t *out);

/*
 * Packs a bound-spaced result slab (stream i at src + src_off[i], size[i] bytes) into a dense
 * flat buffer: stream i goes to dst + dst_off[i], dst_off[n+1] being the exclusive prefix sums
 * of size[] (the caller computes them).  This is the device-side form of the slicing the
 * reference does on the host after its GPU call (src/batch_cuda.rs:123-137) and of the per-result
 * copy in src/batch.rs:51; it turns the output of bdf_compress_batch_device into the input layout
 * of bdf_decompress_batch_device without leaving the GPU.
 */
int bdf_gather_streams_device(bdf_ctx *ctx, const uint8_t *src, const uint64_t *src_off,
                              const uint64_t *size, size_t n, uint8_t *dst,
                              const uint64_t *dst_off, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BDEFLATE_H */
