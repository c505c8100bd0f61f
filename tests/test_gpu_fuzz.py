"""Randomised differential tests: the CUDA path against the oracle (and system
zlib) on seeded, ragged batches — lengths from 0 to several chunks, every
level and framing, mixed content kinds inside one batch, corrupted streams."""
import zlib

import numpy as np
import pytest

import corpus
import oracle_lib as o
import tier

pytestmark = pytest.mark.gpu
WBITS = {0: -15, 1: 15, 2: 31}


def random_buffer(rng, max_len):
    """One buffer of a random kind and a length skewed towards the small and the boundary cases."""
    pick = rng.integers(0, 10)
    if pick == 0:
        n = int(rng.integers(0, 40))
    elif pick == 1:
        n = int(rng.choice([65535, 65536, 65537, 262143, 262144, 262145]))
    else:
        n = int(rng.integers(0, max_len))
    n = min(n, max_len)
    kind = int(rng.integers(0, 7))
    k = int(rng.integers(0, 1000))
    if kind == 0:
        base = corpus.text_stream(k, 65536)
    elif kind == 1:
        base = corpus.binary_stream(k, 65536)
    elif kind == 2:
        base = corpus.lowentropy_stream(k, 65536)
    elif kind == 3:
        base = corpus.periodic_stream(k, 65536)
    elif kind == 4:
        base = corpus.corpus_a_stream(k % 16)
    elif kind == 5:                                   # runs of one byte with sparse noise
        a = np.full(65536, int(rng.integers(0, 256)), dtype=np.uint8)
        idx = rng.integers(0, 65536, 40)
        a[idx] = rng.integers(0, 256, 40, dtype=np.uint8)
        base = a.tobytes()
    else:                                             # incompressible
        base = rng.integers(0, 256, 65536, dtype=np.uint8).tobytes()
    reps = n // len(base) + 1
    start = int(rng.integers(0, 4096))
    return (base * (reps + 1))[start:start + n]


@pytest.mark.parametrize("seed", [1, 2])
def test_compress_random_batches_match_oracle(engine, seed):
    rng = np.random.default_rng(seed)
    for level in (0, 1, 3, 5, 6, 8, 9):
        fmt = int(rng.integers(0, 3))
        bufs = [random_buffer(rng, 300000 if level in (1, 6) else 90000) for _ in range(24)]
        got = engine.BatchCompressor(level, format=fmt).compress_batch(bufs)
        for g, s in zip(got, bufs):
            exp = o.compress(s, level, fmt)
            assert g == (exp if exp is not None else b""), (seed, level, fmt, len(s))
            if exp is not None and not (level == 0 and len(s) == 0):
                assert zlib.decompress(g, WBITS[fmt]) == s


def test_near_optimal_random_batch_within_tolerance(engine):
    rng = np.random.default_rng(7)
    for level in (10, 11, 12):
        bufs = [random_buffer(rng, 30000) for _ in range(12)]
        got = engine.BatchCompressor(level).compress_batch(bufs)
        pairs = []
        for g, s in zip(got, bufs):
            exp = o.compress(s, level)
            tier.check_stream(g, s, exp, level, 0)
            pairs.append((g, exp))
        tier.check_total(pairs, level)


@pytest.mark.parametrize("seed", [3, 4])
def test_decompress_random_batches_match_oracle(engine, seed):
    """Streams from three producers (oracle, system zlib, the GPU compressor), ragged, with random
    output capacities (exact, generous, one byte short) and random corruptions: status and bytes
    must agree with the oracle stream by stream."""
    rng = np.random.default_rng(seed)
    for fmt in (0, 1, 2):
        bufs = [random_buffer(rng, 200000) for _ in range(40)]
        streams, caps = [], []
        for i, s in enumerate(bufs):
            level = int(rng.choice([0, 1, 4, 6, 9]))
            who = i % 3
            if who == 0:
                c = o.compress(s, level, fmt)
            elif who == 1:
                z = zlib.compressobj(max(level, 1), zlib.DEFLATED, WBITS[fmt])
                c = z.compress(s) + z.flush()
            else:
                c = engine.BatchCompressor(level, format=fmt).compress_batch([s])[0] or None
            if c is None or (level == 0 and len(s) == 0 and who != 1):
                continue
            mode = int(rng.integers(0, 6))
            if mode == 0 and len(s) > 0:
                streams.append(c); caps.append(len(s) - 1)                   # one byte short
            elif mode == 1:
                streams.append(c); caps.append(len(s) + int(rng.integers(0, 5000)))
            elif mode == 2 and len(c) > 8:
                b = bytearray(c)
                b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
                streams.append(bytes(b)); caps.append(len(s))                # one flipped bit
            elif mode == 3 and len(c) > 4:
                streams.append(c[:int(rng.integers(1, len(c)))]); caps.append(len(s))   # truncated
            else:
                streams.append(c); caps.append(len(s))
        got = engine.BatchDecompressor(format=fmt).decompress_batch(streams, caps)
        flat, off = o.flatten(streams)
        eout, eoff, esize, est = o.decompress_batch(flat, off, caps, fmt)
        for i, g in enumerate(got):
            exp = None if est[i] != 0 else eout[int(eoff[i]):int(eoff[i]) + int(esize[i])].tobytes()
            assert g == exp, (seed, fmt, i, None if g is None else len(g), None if exp is None else len(exp))
