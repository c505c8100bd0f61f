"""GPU parity tests for batch decompression: the CUDA path (through the C ABI)
against the oracle, the reference's known-answer vectors and system zlib."""
import zlib

import numpy as np
import pytest

import corpus
import kats
import oracle_lib as o

pytestmark = pytest.mark.gpu
K = kats.load()
WBITS = {0: -15, 1: 15, 2: 31}
FMT_NAME = {"raw": 0, "zlib": 1, "gzip": 2}


def zcompress(data, level, fmt):
    c = zlib.compressobj(level, zlib.DEFLATED, WBITS[fmt])
    return c.compress(data) + c.flush()


def inputs():
    return corpus.small_cases() + [
        corpus.corpus_a_stream(0), corpus.corpus_a_stream(9), corpus.text_stream(1),
        corpus.binary_stream(2), corpus.lowentropy_stream(3), corpus.offset_stream(1),
        corpus.offset_stream(3), corpus.offset_stream(7), corpus.offset_stream(32),
        np.random.default_rng(5).integers(0, 256, 70000, dtype=np.uint8).tobytes(),
        (corpus.text_stream(5) * 6)[:300000],
        # long same-offset match runs (coalesced copies): periods 40, 300 and 20000, > 64 KiB each
        (corpus.text_stream(7, 40) * 8000)[:250001],
        (corpus.binary_stream(8, 300) * 900)[:200003],
        (corpus.text_stream(9, 20000) * 8)[:150000],
    ]


def test_reference_known_answer_streams(engine):
    d = engine.BatchDecompressor()
    streams = [s for s, _, _ in K["inflate"]] + [bytes([3, 0])]
    expect = [e for _, e, _ in K["inflate"]] + [b""]
    assert d.decompress_batch(streams, [1024] * len(streams)) == expect
    assert d.decompress_batch(streams, [len(e) for e in expect]) == expect


def test_reference_failure_cases(engine):
    for stream, formats, cite in K["inflate_must_fail"]:
        for f in formats:
            d = engine.BatchDecompressor(format=FMT_NAME[f])
            assert d.decompress_batch([stream], [100]) == [None], cite


def test_batch_semantics(engine):
    # tests/batch_test.rs: empty batch, empty input, garbage, max_out = len - 1, zip truncation
    d = engine.BatchDecompressor()
    assert d.decompress_batch([], []) == []
    good = o.compress(b"hello hello hello hello", 6)
    empty = o.compress(b"", 6)
    res = d.decompress_batch([good, empty, bytes([0, 1, 2, 3, 4, 5]), good, good[:-2], b""],
                             [23, 0, 100, 22, 100, 10])
    assert res == [b"hello hello hello hello", b"", None, None, None, None]
    assert d.decompress_batch([good, good, good], [100, 100]) == [b"hello hello hello hello"] * 2
    assert d.decompress_batch([good], [1000]) == [b"hello hello hello hello"]   # shorter than max_out is fine


@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_parity_with_oracle_and_zlib(engine, fmt):
    comp, plain = [], []
    for s in inputs():
        for level in (0, 1, 6, 9):
            comp.append(zcompress(s, level, fmt)); plain.append(s)
        for level in (0, 1, 3, 6, 9, 12):
            c = o.compress(s, level, fmt)
            if c is not None and not (level == 0 and len(s) == 0):
                comp.append(c); plain.append(s)
    d = engine.BatchDecompressor(format=fmt)
    flat, off = engine.flatten(comp)
    out, out_off, out_size, status, sums = d.decompress_flat(flat, off, [len(p) for p in plain],
                                                             want_checksum=True)
    oflat, ooff = o.flatten(comp)
    eout, eoff, esize, est = o.decompress_batch(oflat, ooff, [len(p) for p in plain], fmt)
    assert (status == est).all() and (status == 0).all()
    assert (out_size == esize).all()
    for i, p in enumerate(plain):
        got = out[int(out_off[i]):int(out_off[i]) + int(out_size[i])].tobytes()
        assert got == p, i
        if fmt == 1:
            assert int(sums[i]) == zlib.adler32(p)
        elif fmt == 2:
            assert int(sums[i]) == zlib.crc32(p)
    # every stream again with one byte too little room -> failure, like the oracle
    res = d.decompress_batch(comp, [max(len(p) - 1, 0) for p in plain])
    for r, p in zip(res, plain):
        assert r is None or len(p) == 0


def test_error_status_parity(engine):
    """Truncations and corruptions: success/failure must agree with the oracle
    stream by stream (the batch API only exposes success vs failure)."""
    rng = np.random.default_rng(11)
    base = [o.compress(corpus.text_stream(4, 20000), 6), o.compress(corpus.corpus_a_stream(2), 6),
            o.compress(corpus.binary_stream(6, 30000), 1), zcompress(corpus.text_stream(8, 9000), 9, 0),
            zcompress(np.random.default_rng(2).integers(0, 256, 5000, dtype=np.uint8).tobytes(), 6, 0)]
    sizes = [20000, 65536, 30000, 9000, 5000]
    streams, caps = [], []
    for c, n in zip(base, sizes):
        for cut in (1, 2, 3, 7, len(c) // 3, len(c) // 2, len(c) - 1):
            streams.append(c[:len(c) - cut]); caps.append(n)
        for _ in range(12):
            b = bytearray(c)
            b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
            streams.append(bytes(b)); caps.append(n)
        streams.append(c + b"trailing garbage"); caps.append(n)
    streams += [bytes([0x07]), bytes([0x05]), b"\x00", b"\x01\x00\x00\xff\xff", b"\x01\x05\x00\xfa\xffabc"]
    caps += [10, 10, 10, 10, 10]
    d = engine.BatchDecompressor()
    got = d.decompress_batch(streams, caps)
    flat, off = o.flatten(streams)
    eout, eoff, esize, est = o.decompress_batch(flat, off, caps, 0)
    for i, g in enumerate(got):
        exp = None if est[i] != 0 else eout[int(eoff[i]):int(eoff[i]) + int(esize[i])].tobytes()
        assert g == exp, (i, None if g is None else len(g), None if exp is None else len(exp))


def test_corrupt_footers(engine):
    s = corpus.text_stream(9, 12000)
    z = bytearray(o.compress(s, 6, 1)); z[-1] ^= 0x40
    g = bytearray(o.compress(s, 6, 2)); g[-6] ^= 0x01
    g2 = bytearray(o.compress(s, 6, 2)); g2[-1] ^= 0x01          # ISIZE
    assert engine.BatchDecompressor(format=1).decompress_batch([bytes(z)], [12000]) == [None]
    assert engine.BatchDecompressor(format=2).decompress_batch([bytes(g), bytes(g2)], [12000] * 2) == [None, None]
    zh = bytearray(o.compress(s, 6, 1)); zh[0] = 0x79
    assert engine.BatchDecompressor(format=1).decompress_batch([bytes(zh)], [12000]) == [None]
    # gzip with FNAME + FEXTRA + FHCRC header fields
    import gzip, io
    bio = io.BytesIO()
    with gzip.GzipFile(filename="name.txt", mode="wb", fileobj=bio, mtime=5) as f:
        f.write(s)
    assert engine.BatchDecompressor(format=2).decompress_batch([bio.getvalue()], [12000]) == [s]


def test_config2_shape_full_batch(engine):
    """BASELINE config 2 at full size: 65536 x 64 KiB zlib streams of corpus A.
    Size-independent checks: every status OK, every Adler-32 equals the
    checksum of the matching plain stream, sampled streams byte-exact."""
    n = 65536
    plain = [corpus.corpus_a_stream(k) for k in range(16)]
    base = [o.compress(p, 6, 1) for p in plain]
    adl = np.array([zlib.adler32(p) for p in plain], dtype=np.uint32)
    flat, off = engine.flatten([base[k % 16] for k in range(n)])
    d = engine.BatchDecompressor(format=1)
    out, out_off, out_size, status, sums = d.decompress_flat(flat, off, np.full(n, 65536, dtype=np.uint64),
                                                             want_checksum=True)
    assert (status == 0).all() and (out_size == 65536).all()
    assert (sums == np.tile(adl, n // 16)).all()
    for i in range(0, n, 4099):
        assert out[i * 65536:(i + 1) * 65536].tobytes() == plain[i % 16]
