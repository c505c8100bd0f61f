"""The two inflate engines (lane groups, inflate.cuh; lane per stream, inflate_lane.cuh) and the
split between them: every mode must give the oracle's bytes, sizes, checksums and statuses on the
same batches — good streams from three producers, ragged lengths, corrupted and truncated streams.
Running one batch through all modes (different kernels, grids and streams-per-warp) and demanding
identical results is also this repo's stand-in for a race check."""
import os
import zlib

import numpy as np
import pytest

import corpus
import oracle_lib as o
from test_gpu_fuzz import random_buffer
from test_gpu_inflate import inputs, zcompress

pytestmark = pytest.mark.gpu

MODES = [
    {"BDF_INFLATE_MODE": "group"},
    {"BDF_INFLATE_MODE": "group", "BDF_INFLATE_GROUP": "32"},
    {"BDF_INFLATE_MODE": "lane", "BDF_LANE_CFG": "0"},
    {"BDF_INFLATE_MODE": "lane", "BDF_LANE_CFG": "1"},
    {"BDF_INFLATE_MODE": "lane", "BDF_LANE_CFG": "2"},
    {"BDF_INFLATE_MODE": "auto", "BDF_INFLATE_SPLIT": "16"},
    {"BDF_INFLATE_MODE": "auto", "BDF_INFLATE_SPLIT": "3"},
]


def context_with(engine, env):
    keys = ["BDF_INFLATE_MODE", "BDF_INFLATE_GROUP", "BDF_LANE_CFG", "BDF_INFLATE_SPLIT"]
    old = {k: os.environ.get(k) for k in keys}
    try:
        for k in keys:
            os.environ.pop(k, None)
        os.environ.update(env)
        return engine.Context(0)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.fixture(scope="module", params=range(len(MODES)), ids=lambda i: "-".join(MODES[i].values()))
def ctx(engine, request):
    c = context_with(engine, MODES[request.param])
    yield c
    c.close()


def check_against_oracle(engine, ctx, fmt, streams, caps):
    d = engine.BatchDecompressor(format=fmt, context=ctx)
    flat, off = engine.flatten(streams)
    out, out_off, out_size, status, sums = d.decompress_flat(flat, off, caps, want_checksum=True)
    oflat, ooff = o.flatten(streams)
    eout, eoff, esize, est = o.decompress_batch(oflat, ooff, caps, fmt)
    for i in range(len(streams)):
        assert (status[i] == 0) == (est[i] == 0), (i, int(status[i]), int(est[i]))
        if est[i] == 0:
            assert int(out_size[i]) == int(esize[i]), i
            got = out[int(out_off[i]):int(out_off[i]) + int(out_size[i])].tobytes()
            exp = eout[int(eoff[i]):int(eoff[i]) + int(esize[i])].tobytes()
            assert got == exp, i
            if fmt == 1:
                assert int(sums[i]) == zlib.adler32(exp), i
            elif fmt == 2:
                assert int(sums[i]) == zlib.crc32(exp), i
        else:
            assert int(out_size[i]) == 0


@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_good_streams_every_engine(engine, ctx, fmt):
    comp, caps = [], []
    for k, s in enumerate(inputs()):
        for level in (0, 1, 6, 9):
            comp.append(zcompress(s, level, fmt)); caps.append(len(s) + (k % 3) * 100)
        for level in (1, 4, 6, 12):
            c = o.compress(s, level, fmt)
            if c is not None:
                comp.append(c); caps.append(len(s))
    check_against_oracle(engine, ctx, fmt, comp, caps)


@pytest.mark.parametrize("seed", [21, 22])
def test_ragged_and_corrupt_every_engine(engine, ctx, seed):
    rng = np.random.default_rng(seed)
    for fmt in (0, 1, 2):
        streams, caps = [], []
        for i in range(96):
            s = random_buffer(rng, 150000)
            level = int(rng.choice([0, 1, 4, 6, 9]))
            c = (o.compress(s, level, fmt) if i % 2 else None) or zcompress(s, max(level, 1), fmt)
            if c is None or (level == 0 and len(s) == 0 and i % 2):
                continue
            mode = int(rng.integers(0, 7))
            if mode == 0 and len(s) > 0:
                streams.append(c); caps.append(len(s) - 1)
            elif mode == 1:
                streams.append(c); caps.append(len(s) + int(rng.integers(0, 5000)))
            elif mode == 2 and len(c) > 8:
                b = bytearray(c)
                b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
                streams.append(bytes(b)); caps.append(len(s))
            elif mode == 3 and len(c) > 4:
                streams.append(c[:int(rng.integers(1, len(c)))]); caps.append(len(s))
            elif mode == 4:
                streams.append(c); caps.append(len(s) * 40 + 7)        # generous capacity: changes the class of the stream
            else:
                streams.append(c); caps.append(len(s))
        check_against_oracle(engine, ctx, fmt, streams, caps)


def test_many_small_and_unaligned_streams(engine, ctx):
    """Hundreds of short streams (slot turnover inside a warp, ragged output alignment)."""
    rng = np.random.default_rng(5)
    plain = []
    for i in range(700):
        n = int(rng.integers(0, 3000))
        kind = i % 4
        base = (corpus.text_stream, corpus.binary_stream, corpus.lowentropy_stream, corpus.periodic_stream)[kind](i % 50, 4096)
        plain.append(base[:n])
    for fmt in (0, 1, 2):
        comp = [o.compress(p, 1 + i % 9, fmt) or zcompress(p, 6, fmt) for i, p in enumerate(plain)]
        check_against_oracle(engine, ctx, fmt, comp, [len(p) + i % 5 for i, p in enumerate(plain)])


def test_mixed_corpus_shard(engine, ctx):
    """2048 x 64 KiB of corpus B (the BASELINE configs[4] mix) as gzip: CRC-32 of every stream."""
    base = [corpus.corpus_b_stream(k) for k in range(64)]
    comp = [o.compress(b, 6, 2) for b in base]
    n = 2048
    d = engine.BatchDecompressor(format=2, context=ctx)
    flat, off = engine.flatten([comp[k % 64] for k in range(n)])
    out, out_off, out_size, status, sums = d.decompress_flat(flat, off, np.full(n, 65536, dtype=np.uint64),
                                                             want_checksum=True)
    assert (status == 0).all() and (out_size == 65536).all()
    crc = np.array([zlib.crc32(b) for b in base], dtype=np.uint32)
    assert (sums == np.tile(crc, n // 64)).all()
    for i in range(0, n, 37):
        assert out[i * 65536:(i + 1) * 65536].tobytes() == base[i % 64]
