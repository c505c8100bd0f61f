/*
 * oracle.h — CPU restatement of the reference's batch DEFLATE path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped
 * product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it.  The product path
 * (libdeflate_rsx_b200/csrc, libbdeflate.so) never links or falls back to it.
 *
 * What it restates (all paths relative to /root/reference):
 *   src/batch.rs                       batch fan-out, per-stream failure
 *   src/compress/{mod,matchfinder,huffman_comp,bitstream}.rs   levels 0..12
 *   src/decompress/{mod,tables}.rs (+ x86.rs:2193-2424 error tuples)
 *   src/adler32/mod.rs, src/crc32/mod.rs
 *
 * PARITY PINNING.  Decompression and checksums are pinned by the reference's
 * own known-answer tests (tests/unit_tests.rs, tests/adler32_test.rs; see
 * tests/golden/) and cross-checked against system zlib.  Compressed BYTES are
 * "parity unpinned": no test or fixture in the reference records compressed
 * output, and the Rust toolchain is absent, so byte-identity at levels 0..9 is
 * defined by this restatement (each function cites the lines it follows).
 */
#ifndef ORACLE_H
#define ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Framing selector shared with include/bdeflate.h. */
enum { ORC_FMT_RAW = 0, ORC_FMT_ZLIB = 1, ORC_FMT_GZIP = 2 };

/* Mirrors DecompressResult declaration order (src/decompress/mod.rs:79-85);
 * CompressResult::InsufficientSpace (src/compress/mod.rs:238-241) reuses 3. */
enum {
    ORC_OK = 0,
    ORC_BAD_DATA = 1,
    ORC_SHORT_OUTPUT = 2,
    ORC_INSUFFICIENT_SPACE = 3,
    ORC_SHORT_INPUT = 4
};

/* src/adler32/mod.rs:89-104 (seed 1) and src/crc32/mod.rs:364 (seed 0). */
uint32_t orc_adler32(uint32_t adler, const uint8_t *p, size_t n);
uint32_t orc_crc32(uint32_t crc, const uint8_t *p, size_t n);

/* src/compress/mod.rs:2236-2246 */
size_t orc_compress_bound(int format, size_t len);

/* One stream.  Returns ORC_OK or ORC_INSUFFICIENT_SPACE. */
int orc_compress(int level, int format, const uint8_t *in, size_t in_len,
                 uint8_t *out, size_t out_cap, size_t *out_size);

/* Compressor::compress(chunk, out, FlushMode), src/compress/mod.rs:693-790, for one chunk of at
 * most 256 KiB: raw DEFLATE; finish = FlushMode::Finish, sync = FlushMode::Sync (:662-681).
 * This is what DeflateEncoder::flush_buffer calls per chunk (src/stream.rs:42-196). */
int orc_compress_unit(int level, const uint8_t *in, size_t in_len, uint8_t *out, size_t out_cap,
                      int finish, int sync, size_t *out_size);

/* Compressor::compress_to_size(input, final_block), src/compress/mod.rs:792-1094,1073-1094: the
 * size estimate in bytes of raw DEFLATE (no framing); see size.inc.c. */
size_t orc_compress_to_size(int level, const uint8_t *in, size_t in_len, int final_block);

/* One stream.  in_consumed follows the reference (reader position, see
 * inflate.c header).  Returns an ORC_* status. */
int orc_decompress(int format, const uint8_t *in, size_t in_len, uint8_t *out,
                   size_t out_cap, size_t *in_consumed, size_t *out_size);

/* Batch calls over the flat layout of include/bdeflate.h: stream i is
 * in[in_off[i] .. in_off[i+1]), written at out + out_off[i].  One codec state
 * per worker thread like rayon's map_init (src/batch.rs:34-37,79-83).
 * nthreads <= 0 means "all online cores".  Per-stream failure is in-band:
 * status[i] != 0 and out_size[i] = 0 (src/batch.rs:52-53,95-96). */
int orc_compress_batch(int level, int format, const uint8_t *in,
                       const uint64_t *in_off, size_t n, uint8_t *out,
                       const uint64_t *out_off, uint64_t *out_size,
                       int32_t *status, int nthreads);
int orc_decompress_batch(int format, const uint8_t *in, const uint64_t *in_off,
                         size_t n, uint8_t *out, const uint64_t *out_off,
                         const uint64_t *max_out, uint64_t *out_size,
                         int32_t *status, int nthreads);
int orc_checksum_batch(int kind /*0 adler32, 1 crc32*/, const uint8_t *in,
                       const uint64_t *in_off, size_t n, uint32_t *out,
                       int nthreads);

int orc_num_cores(void);

#ifdef __cplusplus
}
#endif
#endif
