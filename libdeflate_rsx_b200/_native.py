"""ctypes binding of include/bdeflate.h (libbdeflate.so).

The library is the product; there is no Python or CPU implementation behind
these calls.  Import fails loudly if the shared object has not been built.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# BDF_LIBRARY selects another build of the same ABI (the -DBDF_CHECK debug build, tests/test_gpu_check_build.py)
SO_PATH = os.environ.get("BDF_LIBRARY") or os.path.join(HERE, "libbdeflate.so")

RAW, ZLIB, GZIP = 0, 1, 2
OK, BAD_DATA, SHORT_OUTPUT, INSUFFICIENT_SPACE, SHORT_INPUT = range(5)
ADLER32, CRC32 = 0, 1
FLUSH_SYNC, FLUSH_FINISH = 1, 2
E_OK, E_ARG, E_CUDA, E_NOMEM, E_UNSUPPORTED = 0, -1, -2, -3, -4

EXPORTS = [
    "bdf_version", "bdf_device_count", "bdf_ctx_create", "bdf_ctx_destroy", "bdf_last_error",
    "bdf_kernel_launches", "bdf_last_kernel_ms", "bdf_host_alloc", "bdf_host_free",
    "bdf_compress_bound", "bdf_decompress_batch_device", "bdf_decompress_batch_host",
    "bdf_compress_batch_device", "bdf_compress_batch_host", "bdf_checksum_batch_device",
    "bdf_checksum_batch_host", "bdf_gather_streams_device", "bdf_compress_units_host",
    "bdf_compress_size_batch_device", "bdf_compress_size_batch_host",
    "bdf_compress_batch_host_dense", "bdf_compress_batch_host_sg", "bdf_debug_check_failures",
    "bdf_inflate_resume_batch_device", "bdf_inflate_resume_batch_host", "bdf_compress_batch_device_any",
]


class BdfError(RuntimeError):
    pass


def load():
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python -m libdeflate_rsx_b200.build` "
            "(or __graft_entry__.build()); there is no fallback implementation")
    L = C.CDLL(SO_PATH)
    vp, sz = C.c_void_p, C.c_size_t
    L.bdf_version.restype = C.c_int
    L.bdf_device_count.restype = C.c_int
    L.bdf_ctx_create.restype = C.c_int
    L.bdf_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.bdf_ctx_destroy.restype = None
    L.bdf_ctx_destroy.argtypes = [vp]
    L.bdf_last_error.restype = C.c_char_p
    L.bdf_last_error.argtypes = [vp]
    L.bdf_debug_check_failures.restype = C.c_longlong
    L.bdf_debug_check_failures.argtypes = [vp]
    L.bdf_kernel_launches.restype = C.c_uint64
    L.bdf_kernel_launches.argtypes = [vp]
    L.bdf_last_kernel_ms.restype = C.c_float
    L.bdf_last_kernel_ms.argtypes = [vp]
    L.bdf_host_alloc.restype = vp
    L.bdf_host_alloc.argtypes = [sz]
    L.bdf_host_free.restype = None
    L.bdf_host_free.argtypes = [vp]
    L.bdf_compress_bound.restype = sz
    L.bdf_compress_bound.argtypes = [C.c_int, sz]
    L.bdf_decompress_batch_device.restype = C.c_int
    L.bdf_decompress_batch_device.argtypes = [vp, C.c_int, vp, vp, sz, vp, vp, vp, vp, vp, vp, vp]
    L.bdf_decompress_batch_host.restype = C.c_int
    L.bdf_decompress_batch_host.argtypes = [vp, C.c_int, vp, vp, sz, vp, vp, vp, vp, vp, vp]
    L.bdf_compress_batch_device.restype = C.c_int
    L.bdf_compress_batch_device.argtypes = [vp, C.c_int, C.c_int, vp, vp, sz, vp, vp, vp, vp, vp]
    L.bdf_compress_batch_device_any.restype = C.c_int
    L.bdf_compress_batch_device_any.argtypes = [vp, C.c_int, C.c_int, vp, vp, sz, vp, vp, vp, vp, vp]
    L.bdf_compress_batch_host.restype = C.c_int
    L.bdf_compress_batch_host.argtypes = [vp, C.c_int, C.c_int, vp, vp, sz, vp, vp, vp, vp]
    L.bdf_compress_batch_host_dense.restype = C.c_int
    L.bdf_compress_batch_host_dense.argtypes = [vp, C.c_int, C.c_int, vp, vp, sz, vp, sz, vp, vp]
    L.bdf_compress_batch_host_sg.restype = C.c_int
    L.bdf_compress_batch_host_sg.argtypes = [vp, C.c_int, C.c_int, vp, vp, sz, vp, sz, vp, vp]
    L.bdf_checksum_batch_device.restype = C.c_int
    L.bdf_checksum_batch_device.argtypes = [vp, C.c_int, vp, vp, sz, vp, vp]
    L.bdf_checksum_batch_host.restype = C.c_int
    L.bdf_checksum_batch_host.argtypes = [vp, C.c_int, vp, vp, sz, vp]
    L.bdf_compress_units_host.restype = C.c_int
    L.bdf_compress_units_host.argtypes = [vp, C.c_int, vp, vp, vp, sz, vp, vp, vp, vp]
    L.bdf_compress_size_batch_device.restype = C.c_int
    L.bdf_compress_size_batch_device.argtypes = [vp, C.c_int, vp, vp, sz, C.c_int, vp, vp, vp]
    L.bdf_compress_size_batch_host.restype = C.c_int
    L.bdf_compress_size_batch_host.argtypes = [vp, C.c_int, vp, vp, sz, C.c_int, vp, vp]
    L.bdf_inflate_resume_batch_device.restype = C.c_int
    L.bdf_inflate_resume_batch_device.argtypes = [vp, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.bdf_inflate_resume_batch_host.restype = C.c_int
    L.bdf_inflate_resume_batch_host.argtypes = [vp, sz, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.bdf_gather_streams_device.restype = C.c_int
    L.bdf_gather_streams_device.argtypes = [vp, vp, vp, vp, sz, vp, vp, vp]
    return L


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = load()
    return _LIB
