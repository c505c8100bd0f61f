"""CPU tests: pin the oracle against the reference's own known-answer vectors,
against system zlib, and against the committed golden digests."""
import hashlib
import json
import os
import zlib

import numpy as np
import pytest

import corpus
import kats
import oracle_lib as o

K = kats.load()
FMT_WBITS = {o.RAW: -15, o.ZLIB: 15, o.GZIP: 31}
FMT_NAME = {"raw": o.RAW, "zlib": o.ZLIB, "gzip": o.GZIP}


def zcompress(data, level, fmt):
    c = zlib.compressobj(level, zlib.DEFLATED, FMT_WBITS[fmt])
    return c.compress(data) + c.flush()


@pytest.mark.parametrize("data,expect,cite", K["adler32"])
def test_adler32_reference_kats(data, expect, cite):
    assert o.adler32(data) == expect, cite
    assert zlib.adler32(data) == expect


@pytest.mark.parametrize("data,expect,cite", K["crc32"])
def test_crc32_reference_kats(data, expect, cite):
    assert o.crc32(data) == expect, cite
    assert zlib.crc32(data) == expect


def test_checksum_tails_and_overflow():
    # tests/unit_tests.rs:352-368, tests/adler32_overflow.rs (there vs C libdeflate; here vs zlib)
    for n in K["crc_tail_sizes"]:
        d = bytes(i % 255 for i in range(n))
        assert o.crc32(d) == zlib.crc32(d)
        assert o.adler32(d) == zlib.adler32(d)
    for n in (100000, 1000000):
        d = b"\xff" * n
        assert o.adler32(d) == zlib.adler32(d)
        assert o.crc32(d) == zlib.crc32(d)


@pytest.mark.parametrize("stream,expect,cite", K["inflate"])
def test_inflate_reference_kats(stream, expect, cite):
    assert zlib.decompress(stream, -15) == expect          # the vector itself is valid
    assert o.decompress(stream, 1024) == expect, cite
    assert o.decompress(stream, len(expect)) == expect


@pytest.mark.parametrize("stream,formats,cite", K["inflate_must_fail"])
def test_inflate_reference_failures(stream, formats, cite):
    for f in formats:
        assert o.decompress(stream, 100, FMT_NAME[f]) is None, cite


def test_empty_level1_stream():
    # empty level-1 stream is 03 00 and inflates to nothing (tests/batch_test.rs:53-70)
    assert o.compress(b"", 1) == bytes([0x03, 0x00])
    assert o.decompress(bytes([0x03, 0x00]), 0) == b""
    assert o.compress(b"", 0) == b""          # level 0 + empty input: zero bytes (mod.rs:1408)


@pytest.mark.parametrize("level", list(range(0, 13)))
def test_oracle_compress_roundtrips_through_zlib(level):
    cases = corpus.small_cases() + [corpus.corpus_a_stream(0), corpus.text_stream(0, 30000),
                                    corpus.binary_stream(1, 20000), corpus.lowentropy_stream(3, 20000),
                                    corpus.offset_stream(3, 5000), corpus.offset_stream(32, 9000)]
    for fmt in (o.RAW, o.ZLIB, o.GZIP):
        for d in cases:
            c = o.compress(d, level, fmt)
            assert c is not None
            assert len(c) <= o.compress_bound(fmt, len(d))
            if level == 0 and len(d) == 0 and fmt == o.RAW:
                assert c == b""
                continue
            if level == 0 and len(d) == 0:
                continue  # header + footer around zero deflate bytes: not a valid stream, mirrored as-is
            assert zlib.decompress(c, FMT_WBITS[fmt]) == d
            assert o.decompress(c, len(d), fmt) == d


def test_oracle_inflates_zlib_streams():
    cases = corpus.small_cases() + [corpus.corpus_a_stream(5), corpus.text_stream(2),
                                    corpus.binary_stream(3), os.urandom(3000)]
    for fmt in (o.RAW, o.ZLIB, o.GZIP):
        for d in cases:
            for level in (0, 1, 6, 9):
                c = zcompress(d, level, fmt)
                assert o.decompress(c, len(d), fmt) == d
                if len(d):
                    assert o.decompress(c, len(d) - 1, fmt) is None   # tests/batch_test.rs:86-100


def test_level_ordering():
    # tests/unit_tests.rs:112-125
    d = b"a" * 10000
    c0, c1, c12 = (len(o.compress(d, l)) for l in (0, 1, 12))
    assert c0 > 10000 and c1 < c0 and c12 <= c1


def test_no_stored_fallback_for_incompressible_input():
    # SURVEY §7 hard part 6: the reference returns an empty result instead of storing
    d = np.random.default_rng(0).integers(0, 256, 65536, dtype=np.uint8).tobytes()
    assert o.compress(d, 1) is None
    assert o.compress(d, 6) is None
    assert o.compress(d, 0) is not None


def test_large_input_chunking():
    # > 256 KiB: independent 256 KiB chunks joined by sync-flush markers (mod.rs:699-772)
    d = (corpus.text_stream(7) * 5)[:300000]
    for level in (0, 1, 6):
        c = o.compress(d, level)
        assert zlib.decompress(c, -15) == d
        assert o.decompress(c, len(d)) == d
    assert b"\x00\x00\xff\xff" in o.compress(d, 6)


def test_truncated_and_corrupt_streams_fail():
    d = corpus.text_stream(3, 20000)
    c = o.compress(d, 6)
    for cut in (1, 2, 10, len(c) // 2):
        assert o.decompress(c[:-cut], len(d)) is None
    z = bytearray(o.compress(d, 6, o.ZLIB))
    z[-1] ^= 1
    assert o.decompress(bytes(z), len(d), o.ZLIB) is None      # Adler-32 mismatch
    g = bytearray(o.compress(d, 6, o.GZIP))
    g[-5] ^= 1
    assert o.decompress(bytes(g), len(d), o.GZIP) is None      # CRC-32 mismatch
    assert o.decompress(bytes([0x07]), 10) is None             # reserved block type 3


def test_batch_api_matches_single_stream_calls():
    bufs = corpus.small_cases() + [corpus.corpus_a_stream(1), corpus.text_stream(1, 10000)]
    flat, off = o.flatten(bufs)
    for level in (0, 1, 6):
        out, out_off, out_size, status = o.compress_batch(flat, off, level, o.RAW, nthreads=3)
        for i, b in enumerate(bufs):
            exp = o.compress(b, level)
            got = out[int(out_off[i]):int(out_off[i]) + int(out_size[i])].tobytes()
            assert status[i] == o.OK and got == exp
        comp = [out[int(out_off[i]):int(out_off[i]) + int(out_size[i])].tobytes()
                for i in range(len(bufs))]
        cflat, coff = o.flatten(comp)
        dout, doff, dsize, dst = o.decompress_batch(cflat, coff, [len(b) for b in bufs], o.RAW, 2)
        for i, b in enumerate(bufs):
            if level == 0 and len(b) == 0:
                continue
            assert dst[i] == o.OK
            assert dout[int(doff[i]):int(doff[i]) + int(dsize[i])].tobytes() == b
    sums = o.checksum_batch(flat, off, 1)
    assert [int(x) for x in sums] == [zlib.crc32(b) for b in bufs]


def test_reference_defect_flags():
    """The oracle reports when the REFERENCE decoder would mis-handle a valid
    stream (end-of-block code longer than its table, see oracle/inflate.c)."""
    d = corpus.binary_stream(1)
    c = o.compress(d, 6)
    st, out, used, defect = o.decompress(c, len(d), full=True)
    assert st == o.OK and out == d
    assert defect == 1          # 271 used symbols: the end-of-block code is 14 bits > 11
    c = o.compress(corpus.corpus_a_stream(0), 6)
    st, out, used, defect = o.decompress(c, 65536, full=True)
    assert st == o.OK and defect == 0


def test_golden_digests():
    """Committed digests of the oracle's compressed output (tests/golden/oracle_digests.json,
    made by tests/golden/gen_oracle_digests.py): guards the byte-identity definition
    against accidental edits of the restatement."""
    path = os.path.join(os.path.dirname(__file__), "golden", "oracle_digests.json")
    with open(path) as f:
        g = json.load(f)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import gen_oracle_digests as gen
    for name, data in gen.inputs():
        for level in gen.LEVELS:
            c = o.compress(data, level)
            rec = g[name][str(level)]
            if c is None:
                assert rec is None
            else:
                assert rec == [len(c), hashlib.sha256(c).hexdigest()], (name, level)


def test_golden_digests_large_inputs_and_units():
    """Digests of the oracle on inputs above 64 KiB / 256 KiB and of per-chunk flush-mode
    compression (tests/golden/oracle_digests_large.json)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import gen_oracle_digests_large as gen
    with open(os.path.join(os.path.dirname(__file__), "golden", "oracle_digests_large.json")) as f:
        assert gen.generate() == json.load(f)


def test_chunked_and_unit_streams_are_valid_deflate():
    """System zlib inflates what the oracle makes on the large-input paths: the level-1 split path,
    256 KiB chunks joined by sync flushes, and a sequence of sync-flushed units ended by a finish."""
    import zlib
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import gen_oracle_digests_large as gen
    wb = {0: -15, 1: 15, 2: 31}
    for name, data in gen.inputs():
        for level in gen.LEVELS:
            c = o.compress(data, level, level % 3)
            if c is not None:           # level 0 above 256 KiB overflows its bound: in-band failure
                assert zlib.decompress(c, wb[level % 3]) == data, (name, level)
    parts = [corpus.text_stream(1, 90000), corpus.binary_stream(2, 262144), b"tail"]
    for level in (0, 1, 6):
        stream = b"".join(o.compress_unit(p, level, i == len(parts) - 1, i != len(parts) - 1)
                          for i, p in enumerate(parts))
        assert zlib.decompress(stream, -15) == b"".join(parts)


def test_compress_to_size_restatement():
    """orc_compress_to_size (Compressor::compress_to_size, src/compress/mod.rs:792-1094): at levels
    2..9 the estimator parses like the compressor, so it must equal the emitted raw length; level 0
    is the closed form of :1073-1082; level 1 can only add whole static-block headers (10 bits
    each) over the single-block stream; an empty input costs nothing above level 0."""
    import corpus
    bufs = [corpus.text_stream(1, 65536), corpus.binary_stream(2, 65536)[:30011], corpus.lowentropy_stream(3, 65536),
            corpus.periodic_stream(4, 65536)[:777], corpus.corpus_a_stream(1), (corpus.text_stream(5, 65536) * 3)[:150000],
            b"a", b"abc" * 5]
    for level in range(2, 10):
        for b in bufs:
            c = o.compress(b, level)
            if c is not None:
                assert o.compress_to_size(b, level) == len(c), (level, len(b))
    for b in bufs:
        n = len(b)
        assert o.compress_to_size(b, 0) == n + 5 * (n // 65535 + (1 if n % 65535 else 0))
        c = o.compress(b, 1)
        est = o.compress_to_size(b, 1)
        assert len(c) <= est <= len(c) + 2 + (10 * (n // 5000 + 1) + 7) // 8, (n, est, len(c))
        for level in (10, 12):
            if n <= 70000:
                c = o.compress(b, level)
                assert abs(o.compress_to_size(b, level) - len(c)) <= max(8, len(c) // 100), (level, n)
    assert o.compress_to_size(b"", 0, True) == 5 and o.compress_to_size(b"", 0, False) == 0
    assert [o.compress_to_size(b"", lv) for lv in (1, 6, 12)] == [0, 0, 0]


def test_compress_to_size_random_buffers():
    """Seeded ragged buffers: at levels 2..9 the estimator's byte count is the length of the raw
    stream the compressor writes whenever that stream fits the bound; the stream itself inflates
    under system zlib."""
    import zlib
    import corpus
    rng = np.random.default_rng(23)
    gens = (corpus.text_stream, corpus.binary_stream, corpus.lowentropy_stream, corpus.periodic_stream)
    for trial in range(40):
        n = int(rng.integers(0, 70000)) if trial % 5 else int(rng.integers(0, 40))
        base = gens[trial % 4](int(rng.integers(0, 1000)), 65536)
        start = int(rng.integers(0, 2000))
        b = (base * 2)[start:start + n]
        level = int(rng.integers(2, 10))
        c = o.compress(b, level)
        est = o.compress_to_size(b, level)
        if len(b) == 0:
            assert est == 0
        elif c is not None:
            assert est == len(c), (trial, level, len(b))
            assert zlib.decompress(c, -15) == b


def test_bit_writer_capacity_boundary():
    """tests/bitstream_boundary.rs: a write that ends exactly at the end of the buffer succeeds.
    Through the compressor: a chunk fits a buffer of exactly its compressed size and fails in one
    byte less (Bitstream's checked writes, src/compress/bitstream.rs:143-222), at every level and
    for both flush modes."""
    import corpus
    bufs = [corpus.text_stream(3, 65536)[:5000], corpus.binary_stream(4, 65536)[:777], b"ab" * 40, b"x"]
    for level in (0, 1, 4, 6, 9, 10):
        for b in bufs:
            for finish, sync in ((True, False), (False, True)):
                full = o.compress_unit(b, level, finish, sync)
                assert full is not None
                assert o.compress_unit(b, level, finish, sync, cap=len(full)) == full
                assert o.compress_unit(b, level, finish, sync, cap=len(full) - 1) is None
