"""State isolation between calls and between the streams one lane group decodes in a row — the batch
analogue of the reference's decompressor-reuse tests (tests/reuse_decompressor.rs,
tests/security_state_reset.rs: a decode that stops inside a dynamic block must leave nothing
behind for the next one) and of tests/security_oom_panic.rs (1 MiB of zeros through the encoder)."""
import zlib

import numpy as np
import pytest

import corpus
import oracle_lib as o

pytestmark = pytest.mark.gpu


def reference_reuse_data():
    # tests/reuse_decompressor.rs:15-19
    data = bytearray()
    for i in range(1000):
        data += b"This is a repeating string to force dynamic huffman encoding. "
        data.append(i % 256)
    return bytes(data)


def test_reuse_after_a_decode_that_stops_inside_a_dynamic_block(engine):
    data = bytes(i % 251 for i in range(10000))                     # tests/security_state_reset.rs:10-13
    comp = engine.BatchCompressor(6).compress_batch([data])[0]
    assert comp == o.compress(data, 6) and len(comp) >= 200
    part1 = comp[:len(comp) // 2]                                    # :31-32
    other = bytes([66]) * 500                                        # :55
    other_comp = engine.BatchCompressor(6).compress_batch([other])[0]
    d = engine.BatchDecompressor()
    # one call after the other on the same context
    assert d.decompress_batch([part1], [20000]) == [None]
    assert d.decompress_batch([other_comp], [1000]) == [other]
    assert d.decompress_batch([comp], [len(data)]) == [data]
    # and inside one batch, where a lane group takes the next stream right after the broken one:
    # many copies so that every group of the grid meets both kinds back to back
    streams = [part1, other_comp, comp, part1[:40], other_comp] * 400
    caps = [20000, 1000, len(data), 20000, 500] * 400
    exp = [None, other, data, None, other] * 400
    assert d.decompress_batch(streams, caps) == exp


def test_reuse_mixed_one_shot_then_again(engine):
    data = reference_reuse_data()
    comp = engine.BatchCompressor(6).compress_batch([data])[0]
    assert zlib.decompress(comp, -15) == data
    d = engine.BatchDecompressor()
    for _ in range(3):                                               # tests/reuse_decompressor.rs:27-47
        assert d.decompress_batch([comp], [len(data)]) == [data]
    assert d.decompress_batch([comp], [len(data) - 1]) == [None]
    assert d.decompress_batch([comp], [len(data)]) == [data]


def test_encoder_one_mebibyte_of_zeros(engine):
    import io
    data = bytes(1024 * 1024)                                        # tests/security_oom_panic.rs:5-10
    sink = io.BytesIO()
    enc = engine.DeflateEncoder(sink, 6)
    enc.write(data)
    enc.finish()
    out = sink.getvalue()
    assert len(out) > 0
    assert zlib.decompress(out, -15) == data


def test_huge_expected_size_small_stream(engine):
    """tests/security_oom.rs scaled to 1 GiB: a tiny valid stream at the head of a large input slice,
    a huge expected size that passes the ratio guard — the call returns the 100 bytes (or an
    ordinary error), it does not crash."""
    comp = engine.Compressor(1).compress_deflate(b"A" * 100)
    big = comp + bytes(600 * 1024 - len(comp))
    got = engine.Decompressor().decompress_deflate(big, 1 << 30)
    assert got == b"A" * 100
