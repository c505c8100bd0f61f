"""Host-side mirror of the reference's batch API over the C ABI.

Same names, argument meaning and failure behaviour as reference src/batch.rs:
  BatchCompressor::new(level)            -> BatchCompressor(level)
  compress_batch(&[&[u8]]) -> Vec<Vec<u8>>         (failed stream = empty bytes, :52-53)
  BatchDecompressor::new()               -> BatchDecompressor()
  decompress_batch(&[&[u8]], &[usize]) -> Vec<Option<Vec<u8>>>   (failed stream = None, :95-96;
                                          result length = min(len(inputs), len(max_out_sizes)), :79-81)
plus the `format` selector (raw / zlib / gzip) the north star adds.  Everything
here is flatten / offsets / slicing; all codec work happens in libbdeflate.so
on the GPU, and a missing library or device is an exception, never a fallback.
"""
import ctypes as C
import threading

import numpy as np

from . import _native as N


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """One bdf_ctx (one GPU).  Thread-safe like the reference's `&self` API."""

    def __init__(self, device=0):
        self._lib = N.lib()
        h = C.c_void_p()
        rc = self._lib.bdf_ctx_create(int(device), C.byref(h))
        if rc != N.E_OK:
            raise N.BdfError(f"bdf_ctx_create(device={device}) failed with {rc}: "
                             "no usable CUDA device — this engine has no CPU fallback")
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self._lib.bdf_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != N.E_OK:
            msg = self._lib.bdf_last_error(self.handle)
            raise N.BdfError(f"libbdeflate call failed ({rc}): {msg.decode() if msg else ''}")

    @property
    def kernel_launches(self):
        return int(self._lib.bdf_kernel_launches(self.handle))

    @property
    def last_kernel_ms(self):
        return float(self._lib.bdf_last_kernel_ms(self.handle))


_default_ctx = {}
_default_lock = threading.Lock()


def default_context(device=0):
    with _default_lock:
        if device not in _default_ctx:
            _default_ctx[device] = Context(device)
        return _default_ctx[device]


def compress_bound(fmt, n):
    return int(N.lib().bdf_compress_bound(int(fmt), int(n)))


def flatten(bufs):
    """list of bytes-like -> (flat uint8 array, uint64 offsets[n+1])."""
    n = len(bufs)
    off = np.zeros(n + 1, dtype=np.uint64)
    if n:
        off[1:] = np.cumsum([len(b) for b in bufs], dtype=np.uint64)
    total = int(off[-1])
    flat = np.empty(max(total, 1), dtype=np.uint8)
    if total:
        flat[:total] = np.frombuffer(b"".join(bytes(b) for b in bufs), dtype=np.uint8)
    return flat, off


def exclusive_offsets(sizes):
    sizes = np.asarray(sizes, dtype=np.uint64)
    off = np.zeros(len(sizes), dtype=np.uint64)
    if len(sizes) > 1:
        off[1:] = np.cumsum(sizes[:-1], dtype=np.uint64)
    return off


class BatchCompressor:
    def __init__(self, level, format=N.RAW, context=None):
        self.level = int(level)
        self.format = int(format)
        self.ctx = context or default_context()

    def compress_flat(self, flat, in_off):
        """Flat-layout call: returns (out slab, out_off, out_size, status)."""
        n = len(in_off) - 1
        lens = np.diff(in_off)
        bounds = lens + (lens // np.uint64(65535) + np.uint64(1)) * np.uint64(5) + np.uint64(10)
        bounds = bounds + np.uint64({N.RAW: 0, N.ZLIB: 6, N.GZIP: 18}[self.format])
        out_off = exclusive_offsets(bounds)
        out = np.empty(max(int(bounds.sum()), 1), dtype=np.uint8)
        out_size = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        self.ctx.check(self.ctx._lib.bdf_compress_batch_host(
            self.ctx.handle, self.level, self.format, _ptr(flat), _ptr(in_off), n, _ptr(out),
            _ptr(out_off), _ptr(out_size), _ptr(status)))
        return out, out_off, out_size, status

    def compress_dense(self, flat, in_off):
        """Flat input, packed result: returns (out, out_off[n+1], status); stream i is
        out[out_off[i]:out_off[i+1]] (empty when it failed)."""
        n = len(in_off) - 1
        cap = int(in_off[-1]) + (int(in_off[-1]) // 65535 + n) * 5 + n * 28 + 64      # >= the sum of the bounds
        out = np.empty(cap, dtype=np.uint8)
        out_off = np.zeros(n + 1, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        self.ctx.check(self.ctx._lib.bdf_compress_batch_host_dense(
            self.ctx.handle, self.level, self.format, _ptr(flat), _ptr(in_off), n, _ptr(out), cap,
            _ptr(out_off), _ptr(status)))
        return out, out_off, status

    def compress_batch(self, inputs):
        """compress_batch(&[&[u8]]) -> Vec<Vec<u8>> (src/batch.rs:20-58): the buffers are handed over
        as they are (pointer + length each, no flattening on this side) and the packed result is
        sliced into one bytes object per stream."""
        n = len(inputs)
        if n == 0:
            return []
        keep = [b if isinstance(b, bytes) else bytes(b) for b in inputs]
        lens = (C.c_size_t * n)(*[len(b) for b in keep])
        ptrs = (C.c_char_p * n)(*keep)
        total = sum(len(b) for b in keep)
        cap = total + (total // 65535 + n) * 5 + n * 28 + 64
        out = np.empty(cap, dtype=np.uint8)
        out_off = np.zeros(n + 1, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        self.ctx.check(self.ctx._lib.bdf_compress_batch_host_sg(
            self.ctx.handle, self.level, self.format, C.cast(ptrs, C.c_void_p), C.cast(lens, C.c_void_p), n,
            _ptr(out), cap, _ptr(out_off), _ptr(status)))
        return [out[int(out_off[i]):int(out_off[i + 1])].tobytes() if status[i] == N.OK else b"" for i in range(n)]

    def compress_batch_slots(self, inputs):
        """The same through the bound-spaced slab call (bdf_compress_batch_host)."""
        if len(inputs) == 0:
            return []
        flat, in_off = flatten(inputs)
        out, out_off, out_size, status = self.compress_flat(flat, in_off)
        res = []
        for i in range(len(inputs)):
            if status[i] == N.OK:
                o = int(out_off[i])
                res.append(out[o:o + int(out_size[i])].tobytes())
            else:
                res.append(b"")
        return res


    def compress_to_size_batch(self, inputs, final_block=True):
        """Compressor::compress_to_size (src/compress/mod.rs:1073-1094) per buffer: the estimated
        raw-DEFLATE size in bytes, no output slab needed.  Buffers of at most 256 KiB."""
        if len(inputs) == 0:
            return []
        flat, in_off = flatten(inputs)
        n = len(inputs)
        out_size = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        self.ctx.check(self.ctx._lib.bdf_compress_size_batch_host(
            self.ctx.handle, self.level, _ptr(flat), _ptr(in_off), n, int(bool(final_block)),
            _ptr(out_size), _ptr(status)))
        if (status != N.OK).any():
            raise N.BdfError("compress_to_size: stream %d unsupported" % int(np.nonzero(status)[0][0]))
        return [int(v) for v in out_size]


class BatchDecompressor:
    def __init__(self, format=N.RAW, context=None):
        self.format = int(format)
        self.ctx = context or default_context()

    def decompress_flat(self, flat, in_off, max_out, want_checksum=False):
        n = len(in_off) - 1
        caps = [int(v) for v in max_out]
        total = sum(caps)                   # Python ints: a uint64 sum would wrap silently
        if any(v < 0 for v in caps) or total >= 1 << 46:
            raise N.BdfError("decompress_batch: output capacities add up to %d bytes" % total)
        max_out = np.array(caps, dtype=np.uint64)
        out_off = exclusive_offsets(max_out)
        out = np.empty(max(total, 1), dtype=np.uint8)
        out_size = np.zeros(n, dtype=np.uint64)
        status = np.zeros(n, dtype=np.int32)
        checksum = np.zeros(n, dtype=np.uint32)
        self.ctx.check(self.ctx._lib.bdf_decompress_batch_host(
            self.ctx.handle, self.format, _ptr(flat), _ptr(in_off), n, _ptr(out), _ptr(out_off),
            _ptr(max_out), _ptr(out_size), _ptr(checksum), _ptr(status)))
        if want_checksum:
            return out, out_off, out_size, status, checksum
        return out, out_off, out_size, status

    def decompress_batch(self, inputs, max_out_sizes):
        n = min(len(inputs), len(max_out_sizes))          # zip semantics, src/batch.rs:79-81
        if n == 0:
            return []
        flat, in_off = flatten(inputs[:n])
        out, out_off, out_size, status = self.decompress_flat(flat, in_off, list(max_out_sizes[:n]))
        res = []
        for i in range(n):
            if status[i] == N.OK:
                o = int(out_off[i])
                res.append(out[o:o + int(out_size[i])].tobytes())
            else:
                res.append(None)
        return res


def checksum_batch(inputs, kind, context=None):
    """adler32(1, x) / crc32(0, x) for each buffer."""
    ctx = context or default_context()
    if len(inputs) == 0:
        return []
    flat, in_off = flatten(inputs)
    out = np.zeros(len(inputs), dtype=np.uint32)
    ctx.check(ctx._lib.bdf_checksum_batch_host(ctx.handle, int(kind), _ptr(flat), _ptr(in_off),
                                               len(inputs), _ptr(out)))
    return [int(x) for x in out]
