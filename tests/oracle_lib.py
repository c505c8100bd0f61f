"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by the
product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "_build", "liboracle.so")

RAW, ZLIB, GZIP = 0, 1, 2
OK, BAD_DATA, SHORT_OUTPUT, INSUFFICIENT_SPACE, SHORT_INPUT = range(5)


def build(force=False):
    if force or not os.path.exists(_SO):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle")],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        u8p, u64p, i32p, u32p = (C.c_void_p,) * 4
        L.orc_adler32.restype = C.c_uint32
        L.orc_adler32.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
        L.orc_crc32.restype = C.c_uint32
        L.orc_crc32.argtypes = [C.c_uint32, C.c_char_p, C.c_size_t]
        L.orc_compress_bound.restype = C.c_size_t
        L.orc_compress_bound.argtypes = [C.c_int, C.c_size_t]
        L.orc_compress.restype = C.c_int
        L.orc_compress.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_size_t,
                                   C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.orc_compress_unit.restype = C.c_int
        L.orc_compress_unit.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                        C.c_int, C.c_int, C.POINTER(C.c_size_t)]
        L.orc_compress_to_size.restype = C.c_size_t
        L.orc_compress_to_size.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_int]
        L.orc_decompress.restype = C.c_int
        L.orc_decompress.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                     C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
        L.orc_inflate_last_ref_defect.restype = C.c_int
        L.orc_compress_batch.restype = C.c_int
        L.orc_compress_batch.argtypes = [C.c_int, C.c_int, u8p, u64p, C.c_size_t, u8p, u64p,
                                         u64p, i32p, C.c_int]
        L.orc_decompress_batch.restype = C.c_int
        L.orc_decompress_batch.argtypes = [C.c_int, u8p, u64p, C.c_size_t, u8p, u64p, u64p,
                                           u64p, i32p, C.c_int]
        L.orc_checksum_batch.restype = C.c_int
        L.orc_checksum_batch.argtypes = [C.c_int, u8p, u64p, C.c_size_t, u32p, C.c_int]
        L.orc_num_cores.restype = C.c_int
        _lib = L
    return _lib


def adler32(data, seed=1):
    return lib().orc_adler32(seed, bytes(data), len(data))


def crc32(data, seed=0):
    return lib().orc_crc32(seed, bytes(data), len(data))


def compress_bound(fmt, n):
    return lib().orc_compress_bound(fmt, n)


def compress(data, level, fmt=RAW, cap=None):
    """Returns compressed bytes, or None on InsufficientSpace (src/batch.rs:52-53)."""
    data = bytes(data)
    cap = compress_bound(fmt, len(data)) if cap is None else cap
    out = C.create_string_buffer(max(cap, 1))
    sz = C.c_size_t(0)
    st = lib().orc_compress(level, fmt, data, len(data), out, cap, C.byref(sz))
    return out.raw[:sz.value] if st == OK else None


def compress_unit(data, level, finish, sync, cap=None):
    """Compressor::compress(chunk, out, mode) for one chunk (<= 256 KiB); None on InsufficientSpace."""
    data = bytes(data)
    cap = compress_bound(RAW, len(data)) + (5 if sync else 0) if cap is None else cap
    out = C.create_string_buffer(max(cap, 1))
    sz = C.c_size_t(0)
    st = lib().orc_compress_unit(level, data, len(data), out, cap, int(finish), int(sync), C.byref(sz))
    return out.raw[:sz.value] if st == OK else None


def compress_to_size(data, level, final_block=True):
    """Compressor::compress_to_size (src/compress/mod.rs:1073-1094): estimated raw DEFLATE bytes."""
    data = bytes(data)
    return lib().orc_compress_to_size(level, data, len(data), int(final_block))


def decompress(data, max_out, fmt=RAW, full=False):
    """Returns bytes or None (src/batch.rs:93-97); full=True -> (status, bytes, in_consumed, ref_defect)."""
    data = bytes(data)
    out = C.create_string_buffer(max(max_out, 1))
    used, sz = C.c_size_t(0), C.c_size_t(0)
    st = lib().orc_decompress(fmt, data, len(data), out, max_out, C.byref(used), C.byref(sz))
    if full:
        return st, out.raw[:sz.value], used.value, lib().orc_inflate_last_ref_defect()
    return out.raw[:sz.value] if st == OK else None


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def flatten(bufs):
    """list of bytes -> (flat uint8 array, uint64 offsets[n+1])"""
    off = np.zeros(len(bufs) + 1, dtype=np.uint64)
    if bufs:
        off[1:] = np.cumsum([len(b) for b in bufs], dtype=np.uint64)
    flat = np.frombuffer(b"".join(bytes(b) for b in bufs), dtype=np.uint8).copy() \
        if int(off[-1]) else np.zeros(1, dtype=np.uint8)
    return flat, off


def compress_batch(flat, in_off, level, fmt=RAW, nthreads=0):
    n = len(in_off) - 1
    lens = np.diff(in_off).astype(np.uint64)
    bounds = np.array([compress_bound(fmt, int(x)) for x in lens], dtype=np.uint64) \
        if n else np.zeros(0, dtype=np.uint64)
    out_off = np.zeros(n, dtype=np.uint64)
    if n:
        out_off[1:] = np.cumsum(bounds)[:-1]
    out = np.zeros(int(bounds.sum()) + 1, dtype=np.uint8)
    out_size = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    lib().orc_compress_batch(level, fmt, _ptr(flat), _ptr(in_off), n, _ptr(out), _ptr(out_off),
                             _ptr(out_size), _ptr(status), nthreads)
    return out, out_off, out_size, status


def decompress_batch(flat, in_off, max_out, fmt=RAW, nthreads=0, out=None):
    n = len(in_off) - 1
    max_out = np.asarray(max_out, dtype=np.uint64)
    out_off = np.zeros(n, dtype=np.uint64)
    if n:
        out_off[1:] = np.cumsum(max_out)[:-1]
    if out is None:
        out = np.zeros(int(max_out.sum()) + 1, dtype=np.uint8)
    out_size = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    lib().orc_decompress_batch(fmt, _ptr(flat), _ptr(in_off), n, _ptr(out), _ptr(out_off),
                               _ptr(max_out), _ptr(out_size), _ptr(status), nthreads)
    return out, out_off, out_size, status


def checksum_batch(flat, in_off, kind, nthreads=0):
    n = len(in_off) - 1
    out = np.zeros(n, dtype=np.uint32)
    lib().orc_checksum_batch(kind, _ptr(flat), _ptr(in_off), n, _ptr(out), nthreads)
    return out


def num_cores():
    return lib().orc_num_cores()
