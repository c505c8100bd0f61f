"""Host mirror of the reference's stream adapters (src/stream.rs) on the batch engine.

`DeflateEncoder` is the reference's encoder step for step: bytes are buffered
(1 MiB by default), a full buffer is cut into 256 KiB chunks, every chunk goes
through a fresh compressor — here: ONE `bdf_compress_units_host` call for all
chunks of the buffer — and ends in a sync flush, except the last chunk of
`finish()`, which ends the DEFLATE stream (stream.rs:42-196).  The output is
therefore byte-identical to the reference encoder's for the same sequence of
writes / flushes.

`DeflateDecoder` gives the same `read` interface (stream.rs:263-376) but is not
incremental: the reference keeps a 64 KiB window and resumes the decoder state
between reads; the batch engine inflates whole streams, so the first read
drains the inner reader and inflates everything (growing the output room on
BDF_INSUFFICIENT_SPACE up to DEFLATE's maximum expansion).
"""
import numpy as np

from . import _native as N
from .batch import _ptr, default_context

CHUNK = 256 * 1024


def compress_units(chunks, flush, level, context=None):
    """Compressor::compress(chunk, out, mode) for every chunk; returns a list of bytes (None = failed)."""
    ctx = context or default_context()
    n = len(chunks)
    if n == 0:
        return []
    lens = np.array([len(c) for c in chunks], dtype=np.uint64)
    off = np.zeros(n + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    flat = np.frombuffer(b"".join(bytes(c) for c in chunks) or b"\0", dtype=np.uint8)
    fl = np.array(flush, dtype=np.uint8)
    caps = lens + (lens // np.uint64(65535) + np.uint64(1)) * np.uint64(5) + np.uint64(10) + np.uint64(5)
    out_off = np.zeros(n, dtype=np.uint64)
    out_off[1:] = np.cumsum(caps)[:-1]
    out = np.empty(int(caps.sum()), dtype=np.uint8)
    out_size = np.zeros(n, dtype=np.uint64)
    status = np.zeros(n, dtype=np.int32)
    ctx.check(ctx._lib.bdf_compress_units_host(ctx.handle, int(level), _ptr(flat), _ptr(off), _ptr(fl), n,
                                               _ptr(out), _ptr(out_off), _ptr(out_size), _ptr(status)))
    return [out[int(out_off[i]):int(out_off[i]) + int(out_size[i])].tobytes() if status[i] == N.OK else None
            for i in range(n)]


class DeflateEncoder:
    """io.RawIOBase-like writer: write() / flush() / finish(); also a context manager."""

    def __init__(self, writer, level, buffer_size=1024 * 1024, context=None):
        self.writer = writer
        self.level = int(level)
        self.buffer_size = int(buffer_size)
        self.buffer = bytearray()
        self.ctx = context or default_context()

    def with_buffer_size(self, size):
        self.buffer_size = int(size)
        return self

    def _flush_buffer(self, final_block):
        # flush_buffer, stream.rs:42-196
        if not self.buffer and not final_block:
            return
        data = bytes(self.buffer)
        chunks = [data[i:i + CHUNK] for i in range(0, len(data), CHUNK)] or [b""]
        flush = [N.FLUSH_SYNC] * len(chunks)
        if final_block:
            flush[-1] = N.FLUSH_FINISH
        outs = compress_units(chunks, flush, self.level, self.ctx)
        for o in outs:
            if o is None:
                raise OSError("Compression failed")
            if self.writer is not None:
                self.writer.write(o)
        self.buffer.clear()

    def write(self, buf):
        self.buffer += buf
        if len(self.buffer) >= self.buffer_size:
            self._flush_buffer(False)
        return len(buf)

    def write_all(self, buf):
        self.write(buf)

    def flush(self):
        self._flush_buffer(False)
        if self.writer is not None and hasattr(self.writer, "flush"):
            self.writer.flush()

    def finish(self):
        self._flush_buffer(True)
        w, self.writer = self.writer, None
        return w

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        if self.writer is not None:      # Drop, stream.rs:234-240: errors are ignored there
            try:
                self._flush_buffer(True)
            except Exception:
                pass
            self.writer = None
        return False


class DeflateDecoder:
    def __init__(self, inner, context=None):
        self.inner = inner
        self.ctx = context or default_context()
        self._out = None
        self._pos = 0

    def _inflate_all(self):
        from .batch import BatchDecompressor
        data = self.inner.read()
        d = BatchDecompressor(format=N.RAW, context=self.ctx)
        flat = np.frombuffer(data or b"\0", dtype=np.uint8)
        off = np.array([0, len(data)], dtype=np.uint64)
        cap = max(4 * len(data), 1 << 16)
        limit = 1032 * len(data) + (1 << 16)           # DEFLATE cannot expand further
        while True:
            out, out_off, out_size, status = d.decompress_flat(flat, off, np.array([cap], dtype=np.uint64))
            if status[0] == N.OK:
                return out[:int(out_size[0])].tobytes()
            if status[0] != N.INSUFFICIENT_SPACE or cap >= limit:
                raise OSError("Decompression failed")        # io::ErrorKind::InvalidData, stream.rs:330-340
            cap = min(cap * 4, limit)

    def read(self, n=-1):
        if self._out is None:
            self._out = self._inflate_all()
        if n is None or n < 0:
            n = len(self._out) - self._pos
        chunk = self._out[self._pos:self._pos + n]
        self._pos += len(chunk)
        return chunk

    def read_to_end(self):
        return self.read(-1)
