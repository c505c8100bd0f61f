"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every
symbol include/bdeflate.h declares, and refuses to run without a GPU (no
fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "bdeflate.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bdf_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_batch_surface():
    names = declared_functions()
    for need in ("bdf_compress_batch_host", "bdf_compress_batch_device", "bdf_decompress_batch_host",
                 "bdf_decompress_batch_device", "bdf_checksum_batch_host", "bdf_compress_bound",
                 "bdf_ctx_create", "bdf_ctx_destroy"):
        assert need in names


def test_library_exports_every_declared_symbol():
    from libdeflate_rsx_b200 import _native as N
    lib = ctypes.CDLL(N.SO_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} declared in include/bdeflate.h but not exported"
    assert sorted(N.EXPORTS) == declared_functions()


def test_compress_bound_matches_reference_formula():
    # src/compress/mod.rs:2236-2246
    import libdeflate_rsx_b200 as bdf
    import oracle_lib as o
    for n in (0, 1, 65534, 65535, 65536, 131070, 1 << 20, (1 << 32) + 5):
        base = n + (n // 65535 + 1) * 5 + 10
        assert bdf.compress_bound(bdf.RAW, n) == base == o.compress_bound(o.RAW, n)
        assert bdf.compress_bound(bdf.ZLIB, n) == base + 6 == o.compress_bound(o.ZLIB, n)
        assert bdf.compress_bound(bdf.GZIP, n) == base + 18 == o.compress_bound(o.GZIP, n)


def test_no_cpu_fallback_without_gpu():
    from libdeflate_rsx_b200 import _native as N
    import libdeflate_rsx_b200 as bdf
    if N.lib().bdf_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(bdf.BdfError):
        bdf.Context(0)


def test_product_package_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "libdeflate_rsx_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_lib" not in text and "liboracle" not in text and "orc_" not in text, f


def test_safe_api_guards_need_no_gpu():
    """src/api.rs host logic: level validation (:10-16), zip-bomb guards (:213-239), bounds (:59-69)."""
    import pytest
    import libdeflate_rsx_b200 as bdf
    for bad in (-1, 13):
        with pytest.raises(ValueError, match="between 0 and 12"):
            bdf.Compressor(bad)
    c = bdf.Compressor(6)
    assert c.deflate_compress_bound(65536) == 65536 + 2 * 5 + 10
    assert c.zlib_compress_bound(0) == 5 + 10 + 6 and c.gzip_compress_bound(0) == 5 + 10 + 18
    d = bdf.Decompressor()
    with pytest.raises(ValueError, match="safety limit"):
        d.decompress_deflate(bytes(10), 30000)              # 30000 > 10 * 2000 + 4096
    d.set_max_memory_limit(50 << 20)
    with pytest.raises(ValueError, match="maximum memory limit"):
        d.decompress_zlib(bytes(1 << 20), 100 << 20)
    d.set_limit_ratio(10)
    with pytest.raises(ValueError, match="safety limit"):
        d.decompress_gzip_batch([bytes(10), bytes(10)], [4000, 5000])   # second stream: 5000 > 4196
    # tests/security_oom.rs: an 8 GiB request backed by 5 MiB of input passes the ratio guard and must
    # end in an ordinary error, never a crash (here: no GPU, so the engine refuses to start)
    if not __import__("torch").cuda.is_available():
        with pytest.raises(bdf.BdfError):
            bdf.Decompressor().decompress_deflate(bytes(5 << 20), 8 << 30)
