"""Host build of the pre-pass header decoder (csrc/inflate_prehdr.cuh: prehdr_decode) for CPU tests.

The device function is plain per-lane C++ apart from three intrinsics, so its TEXT is cut out of the
.cuh files (together with the BitReader it uses), the intrinsics are given host definitions, and the
result is compiled with g++ into tests/host_harness/_build/.  What runs in the CPU tests is therefore
the code the kernel runs, not a restatement of it.
"""
import ctypes
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(os.path.dirname(HERE)), "libdeflate_rsx_b200", "csrc")
BUILD = os.path.join(HERE, "_build")

PRELUDE = r"""
#include <cstdint>
#include <cstddef>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define BDF_OK 0
#include <cassert>
#define BDF_ASSERT(c) assert(c)
static inline uint32_t __ldg(const uint32_t *p) { return *p; }
static inline uint8_t __ldg(const uint8_t *p) { return *p; }
static inline uint32_t __brev(uint32_t x) { uint32_t r = 0; for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i); return r; }
constexpr int PREHDR_ROW_BYTES = 320, PREHDR_ROW_WORDS = 80;
constexpr uint32_t PREHDR_VALID = 1u << 31;
"""

WRAPPER = r"""
// the same header in pieces of `budget` symbols, the way the lane kernel spreads it over its rounds
extern "C" uint32_t prehdr_host_budget(const uint8_t *p, uint32_t dlen, uint8_t *row320, unsigned budget, uint64_t *end_bit)
{
    static uint8_t ptab[128];
    uint8_t row[324];
    memset(row, 0xEE, sizeof(row));
    BitReader br;
    br.init(p, dlen);
    LaneHdr h;
    if (!lane_hdr_begin(br, dlen, ptab, 1, h)) return 0;
    int r;
    while ((r = lane_hdr_lengths(br, ptab, 1, row, h, budget)) == 0) {}
    if (r != 1 || br.overrun()) return 0;
    memcpy(row320, row, 320);
    *end_bit = (uint64_t)br.consumed_bits();
    return PREHDR_VALID | (h.nlit - 257) << 16 | (h.noff - 1) << 21 | h.final << 26;
}
extern "C" uint32_t prehdr_host(const uint8_t *p, uint32_t dlen, uint8_t *row320, unsigned lane)
{
    static uint8_t ptab[128 * 32];
    uint32_t row[81];
    memset(row, 0xEE, sizeof(row));
    const uint32_t m = prehdr_decode(p, dlen, ptab, row, lane & 31u);
    memcpy(row320, row, 320);
    return m;
}
"""


def _cut(text, start_pat, end_pat):
    a = re.search(start_pat, text, re.M)
    assert a, start_pat
    b = re.search(end_pat, text[a.start():], re.M)
    assert b, end_pat
    return text[a.start():a.start() + b.end()]


def build():
    os.makedirs(BUILD, exist_ok=True)
    inflate = open(os.path.join(CSRC, "inflate.cuh")).read()
    prehdr = open(os.path.join(CSRC, "inflate_prehdr.cuh")).read()
    reader = _cut(inflate, r"^struct BitReader \{", r"^\};")
    decode = "\n".join([
        _cut(prehdr, r"^struct LaneHdr \{", r"^\};"),
        _cut(prehdr, r"^__device__ __forceinline__ bool lane_hdr_begin", r"^\}"),
        _cut(prehdr, r"^struct PlainRefill \{", r"^\};"),
        _cut(prehdr, r"^template <class Refill = PlainRefill>", r"^\}"),
        _cut(prehdr, r"^__device__ __forceinline__ uint32_t prehdr_decode", r"^\}"),
    ])
    src = os.path.join(BUILD, "prehdr_host.cpp")
    so = os.path.join(BUILD, "libprehdr_host.so")
    text = PRELUDE + reader + "\n" + decode + "\n" + WRAPPER
    if not os.path.exists(src) or open(src).read() != text or not os.path.exists(so):
        with open(src, "w") as f:
            f.write(text)
        subprocess.check_call(["g++", "-O1", "-g", "-shared", "-fPIC", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.prehdr_host.restype = ctypes.c_uint32
    lib.prehdr_host.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint]
    lib.prehdr_host_budget.restype = ctypes.c_uint32
    lib.prehdr_host_budget.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_char_p, ctypes.c_uint,
                                       ctypes.POINTER(ctypes.c_uint64)]
    return lib


def decode(lib, data: bytes, lane=0, misalign=0):
    """-> (meta, row bytes).  The stream is placed at `misalign` bytes past a 4-byte boundary, with
    guard bytes around it (the reader may only touch the words that cover the stream)."""
    pad = b"\xAA" * 8
    buf = ctypes.create_string_buffer(pad + b"\x55" * misalign + data + pad, 16 + misalign + len(data))
    base = ctypes.addressof(buf)
    # string buffers are at least 8-byte aligned
    p = ctypes.cast(base + 8 + misalign, ctypes.c_char_p)
    row = ctypes.create_string_buffer(320)
    m = lib.prehdr_host(p, len(data), row, lane)
    return m, row.raw


def decode_budget(lib, data: bytes, budget, misalign=0):
    """lane_hdr_begin + lane_hdr_lengths in pieces of `budget` symbols -> (meta without the bit count, row, end bit)"""
    pad = b"\xAA" * 8
    buf = ctypes.create_string_buffer(pad + b"\x55" * misalign + data + pad, 16 + misalign + len(data))
    p = ctypes.cast(ctypes.addressof(buf) + 8 + misalign, ctypes.c_char_p)
    row = ctypes.create_string_buffer(320)
    end = ctypes.c_uint64(0)
    m = lib.prehdr_host_budget(p, len(data), row, budget, ctypes.byref(end))
    return m, row.raw, int(end.value)
