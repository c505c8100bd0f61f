"""GPU parity tests: batch checksums and batch compression through the C ABI."""
import zlib

import numpy as np
import pytest

import corpus
import kats
import oracle_lib as o
import tier

pytestmark = pytest.mark.gpu
K = kats.load()
WBITS = {0: -15, 1: 15, 2: 31}


def inputs():
    return corpus.small_cases() + [
        corpus.corpus_a_stream(0), corpus.corpus_a_stream(11), corpus.text_stream(1),
        corpus.binary_stream(2), corpus.lowentropy_stream(3), corpus.offset_stream(3),
        corpus.offset_stream(32), corpus.text_stream(6, 65535), corpus.binary_stream(9, 65536),
        np.random.default_rng(5).integers(0, 256, 4000, dtype=np.uint8).tobytes(),
    ]


def test_checksum_reference_kats(engine):
    data = [d for d, _, _ in K["adler32"]]
    assert engine.checksum_batch(data, engine.ADLER32) == [e for _, e, _ in K["adler32"]]
    data = [d for d, _, _ in K["crc32"]]
    assert engine.checksum_batch(data, engine.CRC32) == [e for _, e, _ in K["crc32"]]


def test_checksum_tails_alignment_and_overflow(engine):
    bufs = [bytes(i % 255 for i in range(n)) for n in K["crc_tail_sizes"]]
    bufs += [b"\xff" * 100000, b"\xff" * 1000000, corpus.text_stream(3, 65537)]
    # misaligned starts inside the flat buffer come for free: sizes are odd
    for kind, ref in ((engine.ADLER32, zlib.adler32), (engine.CRC32, zlib.crc32)):
        assert engine.checksum_batch(bufs, kind) == [ref(b) for b in bufs]
        flat, off = o.flatten(bufs)
        assert engine.checksum_batch(bufs, kind) == [int(x) for x in o.checksum_batch(flat, off, kind)]
    assert engine.checksum_batch([], engine.CRC32) == []


def levels_implemented(engine):
    out = []
    for level in range(0, 13):
        try:
            engine.BatchCompressor(level).compress_batch([b"probe"])
            out.append(level)
        except engine.BdfError as e:
            assert "not implemented" in str(e)
    return out


def test_compress_byte_identical_to_oracle(engine):
    """Levels 0..9: byte-identical to the oracle (the repository's definition of
    the reference's output); every stream also inflates under system zlib.
    Levels 10..12 (streams up to 64 KiB: the parallel near-optimal parser) owe validity here and
    the 0.5 % size tolerance in test_compress_ratio_tier."""
    lv = levels_implemented(engine)
    assert lv == list(range(0, 13))
    for fmt in (0, 1, 2):
        for level in lv:
            if level >= 10 and fmt != (level - 10):      # one framing per near-optimal level keeps the test short
                continue
            got = engine.BatchCompressor(level, format=fmt).compress_batch(inputs())
            for g, s in zip(got, inputs()):
                exp = o.compress(s, level, fmt)
                tier.check_stream(g, s, exp, level, fmt)
                if exp is not None and not (level == 0 and len(s) == 0) and g:
                    assert zlib.decompress(g, WBITS[fmt]) == s


def large_inputs():
    t = corpus.text_stream(6, 65536)
    return [
        corpus.text_stream(6, 70000),                                   # one unit above 64 KiB
        (t * 5)[:262144], (t * 5)[:262145],                             # exactly one chunk / one byte more
        (corpus.binary_stream(4) * 10)[:600001],                        # three chunks
        b"".join(corpus.corpus_a_stream(k) for k in range(16)),         # 1 MiB of corpus A: four chunks
        corpus.lowentropy_stream(2) + corpus.text_stream(3) * 3,        # 256 KiB: stays one unit
        b"short one in the same batch",
    ]


LARGE_LEVELS = [0, 1, 2, 6, 9]


def test_large_streams_chunked_byte_identical(engine):
    """Streams above 64 KiB (one 256 KiB unit) and above 256 KiB (chunks joined by sync flushes,
    src/compress/mod.rs:699-772): byte-identical to the oracle, and valid under system zlib."""
    ins = large_inputs()
    for fmt in (0, 1, 2):
        for level in LARGE_LEVELS:
            got = engine.BatchCompressor(level, format=fmt).compress_batch(ins)
            for g, s in zip(got, ins):
                exp = o.compress(s, level, fmt)
                assert g == (exp if exp is not None else b""), (fmt, level, len(s), len(g), len(exp or b""))
                if exp is not None:       # level 0 above 256 KiB overflows its bound (sync markers): in-band failure
                    assert zlib.decompress(g, WBITS[fmt]) == s
    # and back through the batch decompressor
    comp = engine.BatchCompressor(6, format=2).compress_batch(ins)
    assert engine.BatchDecompressor(format=2).decompress_batch(comp, [len(s) for s in ins]) == ins


def test_large_streams_near_optimal_levels(engine):
    """Levels 10..12 above 64 KiB (one serial thread per 256 KiB unit: slow, so few and small)."""
    ins = [corpus.text_stream(6, 70000), (corpus.binary_stream(4) * 5)[:300000], b"tiny"]
    for level, fmt in ((10, 0), (12, 1)):
        batch = ins if level == 10 else ins[:1] + ins[2:]
        got = engine.BatchCompressor(level, format=fmt).compress_batch(batch)
        for g, s in zip(got, batch):
            assert g == o.compress(s, level, fmt), (level, len(s))
            assert zlib.decompress(g, WBITS[fmt]) == s


def test_compress_failure_is_in_band(engine):
    # incompressible input: empty result, not an exception and not a stored block (src/batch.rs:52-53)
    rnd = np.random.default_rng(0).integers(0, 256, 65536, dtype=np.uint8).tobytes()
    for level in [l for l in levels_implemented(engine) if l >= 1]:
        got = engine.BatchCompressor(level).compress_batch([rnd, b"abcabcabcabc" * 10])
        assert got[0] == b""
        tier.check_stream(got[1], b"abcabcabcabc" * 10, o.compress(b"abcabcabcabc" * 10, level), level, 0)
    assert engine.BatchCompressor(0).compress_batch([]) == []


def test_compress_ratio_tier(engine):
    """Levels 10..12: total compressed size within 0.5 % of the oracle's, and
    every stream round-trips through zlib."""
    lv = [l for l in levels_implemented(engine) if l >= 10]
    assert lv == [10, 11, 12]
    ins = [s for s in inputs() if len(s) >= 1000]
    for level in lv:
        got = engine.BatchCompressor(level).compress_batch(ins)
        tot_g = tot_o = 0
        for g, s in zip(got, ins):
            exp = o.compress(s, level)
            if exp is None:
                continue
            assert zlib.decompress(g, -15) == s
            tot_g += len(g); tot_o += len(exp)
        assert tot_g <= tot_o * 1.005, (level, tot_g, tot_o)
