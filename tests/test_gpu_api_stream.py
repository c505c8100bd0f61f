"""GPU tests of the host mirrors of src/api.rs (safe API guards) and src/stream.rs
(stream adapters), ported from the reference's tests/security_limit.rs,
tests/security_overlap_test.rs, tests/stream_test.rs and tests/buffer_size_test.rs,
plus byte-identity of the encoder against the oracle's per-chunk compressor."""
import ctypes as C
import io
import zlib

import numpy as np
import pytest

import corpus
import oracle_lib as o
import tier

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- src/api.rs
def test_level_validation(engine):
    for bad in (-1, 13, 100):
        with pytest.raises(ValueError, match="between 0 and 12"):            # api.rs:11-16
            engine.Compressor(bad)
    for ok in (0, 1, 12):
        engine.Compressor(ok)


def test_memory_limit(engine):
    # tests/security_limit.rs:5-17
    d = engine.Decompressor()
    with pytest.raises(ValueError, match="safety limit"):
        d.decompress_deflate(bytes(10), 1_000_000)


def test_memory_limit_bypass_fixed(engine):
    # tests/security_limit.rs:20-41
    d = engine.Decompressor()
    d.set_max_memory_limit(50 * 1024 * 1024)
    with pytest.raises(ValueError, match="maximum memory limit"):
        d.decompress_deflate(bytes(1024 * 1024), 100 * 1024 * 1024)


def test_valid_decompression_within_limit(engine):
    # tests/security_limit.rs:44-57
    d = engine.Decompressor()
    d.set_max_memory_limit(1024 * 1024)
    original = b"Hello world" * 10
    comp = engine.Compressor(1).compress_deflate(original)
    assert comp == o.compress(original, 1)
    assert d.decompress_deflate(comp, len(original)) == original


def test_decompression_ratio_limit(engine):
    # tests/security_limit.rs:60-107
    d = engine.Decompressor()
    inp = bytes(10)
    with pytest.raises(engine.BdfDataError):            # within the limit: garbage input, not InvalidInput
        d.decompress_deflate(inp, 20000)
    with pytest.raises(ValueError):
        d.decompress_deflate(inp, 30000)
    d.set_limit_ratio(10)
    with pytest.raises(ValueError):
        d.decompress_deflate(inp, 5000)
    with pytest.raises(engine.BdfDataError):
        d.decompress_deflate(inp, 4000)


def test_memory_limit_with_real_data(engine):
    # tests/security_limit.rs:110-: 1 MB of zeroes compresses below 1/2000: the default ratio rejects it
    original = bytes(1_000_000)
    comp = engine.Compressor(1).compress_deflate(original)
    assert comp == o.compress(original, 1)
    d = engine.Decompressor()
    if len(comp) * 2000 + 4096 < len(original):
        with pytest.raises(ValueError, match="safety limit"):
            d.decompress_deflate(comp, len(original))
        d.set_limit_ratio(1_000_000)
    assert d.decompress_deflate(comp, len(original)) == original


def test_all_formats_and_batches(engine):
    bufs = [corpus.text_stream(1, 5000), corpus.binary_stream(2, 70000), b"", b"x"]
    c, d = engine.Compressor(6), engine.Decompressor()
    for name, fmt in (("deflate", 0), ("zlib", 1), ("gzip", 2)):
        comp = getattr(c, f"compress_{name}_batch")(bufs)
        assert comp == [o.compress(b, 6, fmt) for b in bufs]
        assert getattr(d, f"decompress_{name}_batch")(comp, [len(b) for b in bufs]) == bufs
        assert getattr(d, f"decompress_{name}")(getattr(c, f"compress_{name}")(bufs[0]), len(bufs[0])) == bufs[0]
    with pytest.raises(engine.BdfDataError):            # one bad stream fails the checked batch call
        d.decompress_deflate_batch([o.compress(bufs[0], 6), b"\x07garbage"], [5000, 100])
    rnd = np.random.default_rng(0).integers(0, 256, 65536, dtype=np.uint8).tobytes()
    with pytest.raises(engine.BdfDataError, match="Compression failed"):   # incompressible: no stored fallback
        c.compress_deflate(rnd)


def test_overlap_rejected_by_the_c_abi(engine):
    # tests/security_overlap_test.rs: input [0,100) / output [50,150) and the other three scenarios
    ctx = engine.default_context()
    lib = ctx._lib
    buf = np.zeros(4096, dtype=np.uint8)
    base = buf.ctypes.data
    size = np.zeros(1, dtype=np.uint64)
    status = np.zeros(1, dtype=np.int32)
    out_off = np.zeros(1, dtype=np.uint64)
    for (i0, i1), (o0, o1) in (((0, 100), (50, 150)), ((50, 150), (0, 100)), ((50, 100), (0, 150)), ((0, 150), (50, 100))):
        in_off = np.array([0, i1 - i0], dtype=np.uint64)
        rc = lib.bdf_compress_batch_host(ctx.handle, 1, 0, C.c_void_p(base + i0), in_off.ctypes.data, 1,
                                         C.c_void_p(base + o0), out_off.ctypes.data, size.ctypes.data,
                                         status.ctypes.data)
        assert rc == -1 and b"overlap" in lib.bdf_last_error(ctx.handle)
        max_out = np.array([o1 - o0], dtype=np.uint64)
        rc = lib.bdf_decompress_batch_host(ctx.handle, 0, C.c_void_p(base + i0), in_off.ctypes.data, 1,
                                           C.c_void_p(base + o0), out_off.ctypes.data, max_out.ctypes.data,
                                           size.ctypes.data, None, status.ctypes.data)
        assert rc == -1 and b"overlap" in lib.bdf_last_error(ctx.handle)


# ------------------------------------------------------------- src/stream.rs
class TrackingWriter:
    def __init__(self, fail_flush=False):
        self.data = bytearray()
        self.flush_count = 0
        self.fail_flush = fail_flush

    def write(self, b):
        self.data += b
        return len(b)

    def flush(self):
        if self.fail_flush:
            raise OSError("flush error")
        self.flush_count += 1


def reference_encoder_output(writes, level, buffer_size=1024 * 1024, flush_after=()):
    """DeflateEncoder (src/stream.rs:42-240) restated on the oracle's per-chunk compressor."""
    out, buf = bytearray(), bytearray()

    def flush_buffer(final):
        if not buf and not final:
            return
        data = bytes(buf)
        chunks = [data[i:i + 262144] for i in range(0, len(data), 262144)] or [b""]
        for k, c in enumerate(chunks):
            fin = final and k == len(chunks) - 1
            out.extend(o.compress_unit(c, level, fin, not fin))
        buf.clear()

    for i, w in enumerate(writes):
        buf.extend(w)
        if len(buf) >= buffer_size:
            flush_buffer(False)
        if i in flush_after:
            flush_buffer(False)
    flush_buffer(True)
    return bytes(out)


def test_stream_round_trip_and_small_reads(engine):
    # tests/stream_test.rs:41-80
    data = bytes(i % 256 for i in range(10000))
    enc = engine.DeflateEncoder(io.BytesIO(), 6)
    enc.write_all(data)
    comp = enc.finish().getvalue()
    assert comp == reference_encoder_output([data], 6)
    assert zlib.decompress(comp, -15) == data
    assert engine.DeflateDecoder(io.BytesIO(comp)).read_to_end() == data
    dec, got = engine.DeflateDecoder(io.BytesIO(comp)), bytearray()
    while True:
        piece = dec.read(10)
        if not piece:
            break
        got += piece
    assert bytes(got) == data


def test_encoder_flush_and_flush_error(engine):
    # tests/stream_test.rs:83-113
    w = TrackingWriter()
    enc = engine.DeflateEncoder(w, 6)
    enc.write_all(b"Hello World")
    enc.flush()
    assert len(w.data) > 0 and w.flush_count == 1
    assert bytes(w.data) == o.compress_unit(b"Hello World", 6, False, True)
    enc.finish()
    assert zlib.decompress(bytes(w.data), -15) == b"Hello World"
    bad = engine.DeflateEncoder(TrackingWriter(fail_flush=True), 6)
    bad.write_all(b"Hello World")
    with pytest.raises(OSError):
        bad.flush()


def test_with_buffer_size(engine):
    # tests/buffer_size_test.rs:24-58
    w = TrackingWriter()
    enc = engine.DeflateEncoder(w, 1).with_buffer_size(100)
    enc.write_all(bytes(150))
    n1 = len(w.data)
    assert n1 > 0
    enc.finish()
    assert len(w.data) > n1
    assert bytes(w.data) == reference_encoder_output([bytes(150)], 1, buffer_size=100)
    assert zlib.decompress(bytes(w.data), -15) == bytes(150)


@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_encoder_byte_identical_over_many_chunks(engine, level):
    """3.3 MiB in uneven writes: three full 1 MiB buffers (four 256 KiB sync chunks each) and a
    final partial one; an explicit flush in the middle; the whole output byte-identical to the
    restated reference encoder and valid raw DEFLATE."""
    parts = [corpus.text_stream(3), corpus.binary_stream(5), corpus.corpus_a_stream(2), corpus.lowentropy_stream(1)]
    blob = b"".join(parts) * 14
    blob = blob[:3_450_000]
    cuts = [0, 70_000, 1_048_576, 1_100_001, 2_500_000, len(blob)]
    writes = [blob[a:b] for a, b in zip(cuts, cuts[1:])]
    w = TrackingWriter()
    enc = engine.DeflateEncoder(w, level)
    for i, piece in enumerate(writes):
        enc.write(piece)
        if i == 2:
            enc.flush()
    enc.finish()
    assert bytes(w.data) == reference_encoder_output(writes, level, flush_after=(2,))
    assert zlib.decompress(bytes(w.data), -15) == blob
    assert engine.DeflateDecoder(io.BytesIO(bytes(w.data))).read_to_end() == blob


def test_empty_stream_and_context_manager(engine):
    enc = engine.DeflateEncoder(io.BytesIO(), 6)
    out = enc.finish().getvalue()
    assert out == o.compress_unit(b"", 6, True, False) and zlib.decompress(out, -15) == b""
    sink = io.BytesIO()
    with engine.DeflateEncoder(sink, 6) as e:       # Drop finishes the stream (stream.rs:234-240)
        e.write(b"dropped without finish")
    assert zlib.decompress(sink.getvalue(), -15) == b"dropped without finish"
    with pytest.raises(OSError):
        engine.DeflateDecoder(io.BytesIO(b"\x07not deflate")).read()


# ------------------------------------------ reference round-trip inputs through the safe API
def test_reference_offset_patterns_roundtrip(engine):
    """The 50 periodic-pattern round trips of the reference's tests/offset_tests.rs (inputs from
    tests/golden/reference_offset_cases.json), through the safe-API mirror; also byte-identity
    against the oracle, which the reference tests cannot check."""
    import json
    import os
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_offset_cases.json")))
    assert len(cases) == 50 and not any(c.get("unparsed") for c in cases)
    d = engine.Decompressor()
    by_level = {}
    for c in cases:
        pat = bytes.fromhex(c["pattern_hex"])
        data = (pat * (c["take"] // len(pat) + 1))[:c["take"]]
        by_level.setdefault(c["level"], []).append((c, data))
    for level, items in by_level.items():
        comp = engine.Compressor(level).compress_deflate_batch([x for _, x in items])
        back = d.decompress_deflate_batch(comp, [len(x) for _, x in items])
        for (c, data), z, b in zip(items, comp, back):
            assert b == data, c["cite"]
            tier.check_stream(z, data, o.compress(data, level), level, 0, c["cite"])


def test_reference_parallel_and_unit_roundtrips(engine):
    """tests/parallel_test.rs (1 MiB, 256 KiB + 1, 10 MiB and 5 MiB ramps: the chunked path) and the
    round trips / level ordering of tests/unit_tests.rs:70-149."""
    c6, d = engine.Compressor(6), engine.Decompressor()
    ramp = bytes(range(256))
    for n in (1024 * 1024, 256 * 1024 + 1, 10 * 1024 * 1024):          # parallel_test.rs:4-58
        data = (ramp * (n // 256 + 1))[:n]
        comp = c6.compress_deflate(data)
        assert len(comp) > 0 and d.decompress_deflate(comp, n) == data
    n = 5 * 1024 * 1024
    z = bytes((np.arange(n, dtype=np.uint64) * 3 % 251).astype(np.uint8))   # :61-76
    assert d.decompress_zlib(c6.compress_zlib(z), n) == z
    g = bytes((np.arange(n, dtype=np.uint64) * 7 % 251).astype(np.uint8))   # :79-94
    assert d.decompress_gzip(c6.compress_gzip(g), n) == g
    # unit_tests.rs:112-125: level ordering on 10000 x 'a'
    data = b"a" * 10000
    c0 = engine.Compressor(0).compress_deflate(data)
    c1 = engine.Compressor(1).compress_deflate(data)
    c12 = engine.Compressor(12).compress_deflate(data)
    assert len(c0) > 10000 and len(c1) < len(c0) and len(c12) <= len(c1)
    # unit_tests.rs:128-134 and :137-149
    for f in (d.decompress_deflate, d.decompress_zlib, d.decompress_gzip):
        with pytest.raises(engine.BdfDataError):
            f(bytes([0, 1, 2, 3]), 100)
    for data in (b"Data set 1", b"Data set 2 - different content"):
        assert d.decompress_deflate(c6.compress_deflate(data), len(data)) == data
