"""Host mirror of the reference's safe API (src/api.rs) on top of the batch engine.

`Compressor` / `Decompressor` keep the reference's names, argument meaning and
error behaviour: level validation 0..=12 (api.rs:10-16), the zip-bomb guards of
`decompress_helper` (expected size <= len * limit_ratio + 4096 and <=
max_memory_limit, api.rs:213-239), overlap rejection (api.rs:303-314, enforced
inside the C ABI's host calls).  InvalidInput maps to ValueError, InvalidData /
"Compression failed" to BdfDataError.  Every method also has a `*_batch` form:
the guards are applied stream by stream, the work is one GPU call.
"""
import numpy as np

from . import _native as N
from .batch import BatchCompressor, BatchDecompressor, default_context

USIZE_MAX = (1 << 64) - 1


class BdfDataError(RuntimeError):
    """io::ErrorKind::InvalidData / io::Error::other of the reference."""


class Compressor:
    def __init__(self, level, context=None):
        if isinstance(level, bool) or not isinstance(level, (int, np.integer)) or not 0 <= int(level) <= 12:
            raise ValueError("Compression level must be between 0 and 12")        # api.rs:11-16
        self.level = int(level)
        self._ctx = context            # resolved on first use: the guards above need no GPU

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = default_context()
        return self._ctx

    # bounds (api.rs:59-69)
    def deflate_compress_bound(self, size):
        return int(N.lib().bdf_compress_bound(N.RAW, size))

    def zlib_compress_bound(self, size):
        return int(N.lib().bdf_compress_bound(N.ZLIB, size))

    def gzip_compress_bound(self, size):
        return int(N.lib().bdf_compress_bound(N.GZIP, size))

    def _batch(self, fmt, bufs, what):
        out = BatchCompressor(self.level, format=fmt, context=self.ctx).compress_batch(list(bufs))
        for i, (o, b) in enumerate(zip(out, bufs)):
            # the reference's only failure is an encoding that does not fit its bound; level 0 of an
            # empty input legitimately produces zero deflate bytes (src/compress/mod.rs:1408)
            if len(o) == 0 and not (self.level == 0 and len(b) == 0 and fmt == N.RAW):
                raise BdfDataError(f"{what} (stream {i})")
        return out

    def compress_deflate(self, data):
        return self._batch(N.RAW, [data], "Compression failed")[0]

    def compress_zlib(self, data):
        return self._batch(N.ZLIB, [data], "Compression failed")[0]

    def compress_gzip(self, data):
        return self._batch(N.GZIP, [data], "Compression failed")[0]

    def compress_deflate_batch(self, bufs):
        return self._batch(N.RAW, bufs, "Compression failed")

    def compress_zlib_batch(self, bufs):
        return self._batch(N.ZLIB, bufs, "Compression failed")

    def compress_gzip_batch(self, bufs):
        return self._batch(N.GZIP, bufs, "Compression failed")


class Decompressor:
    def __init__(self, context=None):
        self.max_memory_limit = USIZE_MAX      # api.rs:151
        self.limit_ratio = 2000                # api.rs:152
        self._ctx = context            # resolved on first use: the guards below need no GPU

    @property
    def ctx(self):
        if self._ctx is None:
            self._ctx = default_context()
        return self._ctx

    def set_max_memory_limit(self, limit):
        self.max_memory_limit = int(limit)

    def set_limit_ratio(self, ratio):
        self.limit_ratio = int(ratio)

    def _check(self, data_len, expected_size):
        # decompress_helper, api.rs:213-239 (saturating arithmetic)
        limit = min(min(data_len * self.limit_ratio, USIZE_MAX) + 4096, USIZE_MAX)
        if expected_size > limit:
            raise ValueError(f"Expected size {expected_size} exceeds safety limit for input size {data_len}")
        if expected_size > self.max_memory_limit:
            raise ValueError(f"Expected size {expected_size} exceeds maximum memory limit {self.max_memory_limit}")

    def _batch(self, fmt, bufs, sizes):
        bufs, sizes = list(bufs), [int(s) for s in sizes]
        if len(bufs) != len(sizes):
            raise ValueError("one expected size per stream")
        for b, s in zip(bufs, sizes):
            self._check(len(b), s)
        out = BatchDecompressor(format=fmt, context=self.ctx).decompress_batch(bufs, sizes)
        for i, o in enumerate(out):
            if o is None:
                raise BdfDataError(f"Decompression failed (stream {i})")
        return out

    def decompress_deflate(self, data, expected_size):
        return self._batch(N.RAW, [data], [expected_size])[0]

    def decompress_zlib(self, data, expected_size):
        return self._batch(N.ZLIB, [data], [expected_size])[0]

    def decompress_gzip(self, data, expected_size):
        return self._batch(N.GZIP, [data], [expected_size])[0]

    def decompress_deflate_batch(self, bufs, expected_sizes):
        return self._batch(N.RAW, bufs, expected_sizes)

    def decompress_zlib_batch(self, bufs, expected_sizes):
        return self._batch(N.ZLIB, bufs, expected_sizes)

    def decompress_gzip_batch(self, bufs, expected_sizes):
        return self._batch(N.GZIP, bufs, expected_sizes)
