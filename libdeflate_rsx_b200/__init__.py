"""libdeflate_rsx_b200 — B200-native batch DEFLATE engine behind the batch API of
404Setup/libdeflate-rsx (src/batch.rs).  The product is csrc/ (CUDA, sm_100a) +
include/bdeflate.h (C ABI); this package is the thin host mirror used by the
tests and the benchmark."""
from ._native import (ADLER32, BAD_DATA, CRC32, E_ARG, E_CUDA, E_NOMEM, E_OK, E_UNSUPPORTED, GZIP,
                      INSUFFICIENT_SPACE, OK, RAW, SHORT_INPUT, ZLIB, BdfError)
from .api import BdfDataError, Compressor, Decompressor
from .batch import (BatchCompressor, BatchDecompressor, Context, checksum_batch, compress_bound,
                    default_context, exclusive_offsets, flatten)
from .stream import DeflateDecoder, DeflateEncoder, compress_units

__all__ = [
    "BatchCompressor", "BatchDecompressor", "Context", "checksum_batch", "compress_bound",
    "default_context", "flatten", "exclusive_offsets", "RAW", "ZLIB", "GZIP", "OK", "BAD_DATA",
    "INSUFFICIENT_SPACE", "SHORT_INPUT", "ADLER32", "CRC32", "BdfError", "BdfDataError", "Compressor",
    "Decompressor", "DeflateEncoder", "DeflateDecoder", "compress_units", "E_OK", "E_ARG", "E_CUDA", "E_NOMEM",
    "E_UNSUPPORTED",
]
