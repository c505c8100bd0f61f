"""compress_to_size on the GPU (bdf_compress_size_batch_host) against the oracle's restatement of
Compressor::compress_to_size (src/compress/mod.rs:792-1094): equal byte counts, every level."""
import numpy as np
import pytest

import corpus
import oracle_lib as o

pytestmark = pytest.mark.gpu


def buffers(max_len):
    rng = np.random.default_rng(11)
    bufs = [b"", b"a", b"abc" * 5, bytes(300), corpus.corpus_a_stream(0)[:max_len], corpus.corpus_a_stream(3)[:max_len]]
    for k in range(3):
        for gen in (corpus.text_stream, corpus.binary_stream, corpus.lowentropy_stream, corpus.periodic_stream):
            n = int(rng.integers(1, max_len))
            bufs.append(gen(k + 20, 65536)[:n] if n <= 65536 else (gen(k + 20, 65536) * (n // 65536 + 1))[:n])
    bufs.append(rng.integers(0, 256, min(max_len, 40000), dtype=np.uint8).tobytes())    # incompressible
    return bufs


@pytest.mark.parametrize("level", [0, 1, 2, 4, 6, 9])
def test_size_matches_oracle(engine, level):
    for max_len in (65536, 262144):
        bufs = buffers(max_len)
        for final_block in (True, False):
            got = engine.BatchCompressor(level).compress_to_size_batch(bufs, final_block)
            exp = [o.compress_to_size(b, level, final_block) for b in bufs]
            assert got == exp, (level, max_len, [(i, g, e) for i, (g, e) in enumerate(zip(got, exp)) if g != e][:5])


@pytest.mark.parametrize("level", [10, 12])
def test_near_optimal_size_matches_oracle(engine, level):
    bufs = buffers(30000) + [corpus.text_stream(5, 65536), (corpus.binary_stream(6, 65536) * 2)[:100000]]
    got = engine.BatchCompressor(level).compress_to_size_batch(bufs)
    exp = [o.compress_to_size(b, level) for b in bufs]
    assert got == exp, [(i, g, e) for i, (g, e) in enumerate(zip(got, exp)) if g != e][:5]


def test_size_equals_compressed_length_at_greedy_levels(engine):
    """At levels 2..9 the estimator parses exactly like the compressor: its answer is the length of
    the raw stream the batch compressor writes (when that fits the bound)."""
    bufs = [b for b in buffers(65536) if len(b)]
    for level in (3, 6, 8):
        sizes = engine.BatchCompressor(level).compress_to_size_batch(bufs)
        outs = engine.BatchCompressor(level).compress_batch(bufs)
        for s, c in zip(sizes, outs):
            if c:
                assert s == len(c)


def test_size_large_batch_and_limits(engine):
    bufs = [corpus.corpus_a_stream(k % 16) for k in range(2048)]
    got = engine.BatchCompressor(6).compress_to_size_batch(bufs)
    exp = {k: o.compress_to_size(corpus.corpus_a_stream(k), 6) for k in range(16)}
    assert got == [exp[k % 16] for k in range(2048)]
    with pytest.raises(Exception):
        engine.BatchCompressor(6).compress_to_size_batch([bytes(262145)])
    assert engine.BatchCompressor(0).compress_to_size_batch([bytes(300000)]) == [300000 + 5 * 5]
