"""GPU tests at the shapes BASELINE.json names (configs[2..4]); config[1] is
test_gpu_inflate.py::test_config2_shape_full_batch.

At these sizes the oracle cannot compress every stream in seconds, so the
checks are the size-independent ones: corpus A has 16 distinct 64 KiB streams
(the file is periodic in 1 MiB), so all 65536 outputs are compared with the
16 oracle outputs; corpus B is checked through the compress -> decompress ->
CRC-32 round trip plus sampled byte-identity against the oracle."""
import zlib

import numpy as np
import pytest

import corpus
import oracle_lib as o

pytestmark = pytest.mark.gpu
STREAM = 65536


def _tiled(base, n):
    flat, off = o.flatten([base[k % len(base)] for k in range(n)])
    return flat, off


def _check_against_classes(out, out_off, out_size, status, expect):
    """Every stream i must equal expect[i % len(expect)]."""
    n, m = len(out_size), len(expect)
    assert (status == 0).all()
    for c in range(m):
        e = np.frombuffer(expect[c], dtype=np.uint8)
        idx = np.arange(c, n, m)
        assert (out_size[idx] == len(e)).all(), c
        starts = out_off[idx].astype(np.int64)
        rows = out[starts[:, None] + np.arange(len(e))[None, :]]
        assert (rows == e[None, :]).all(), c


def test_config3_level1_full_batch_byte_identical(engine):
    """configs[2]: 65536 x 64 KiB buffers at level 1 (greedy), raw DEFLATE, byte-identical."""
    n = 65536
    base = [corpus.corpus_a_stream(k) for k in range(16)]
    expect = [o.compress(b, 1) for b in base]
    flat, off = _tiled(base, n)
    out, out_off, out_size, status = engine.BatchCompressor(1).compress_flat(flat, off)
    _check_against_classes(out, out_off, out_size, status, expect)


def test_config4_level6_full_batch_and_level12_ratio(engine):
    """configs[3]: level 6 (lazy) byte-identical at full size; level 12 (near-optimal): total size
    within 0.5 % of the oracle's (tolerance of the north-star), every stream inflating under zlib."""
    n = 65536
    base = [corpus.corpus_a_stream(k) for k in range(16)]
    flat, off = _tiled(base, n)
    expect6 = [o.compress(b, 6) for b in base]
    out, out_off, out_size, status = engine.BatchCompressor(6).compress_flat(flat, off)
    _check_against_classes(out, out_off, out_size, status, expect6)
    ratio6 = n * STREAM / int(out_size.sum())
    # level 12 on a mixed batch (corpus A + corpus B), sized to finish in seconds
    mixed = [corpus.corpus_a_stream(k) for k in range(4)] + [corpus.corpus_b_stream(k) for k in range(12)]
    n12 = 2048
    flat, off = _tiled(mixed, n12)
    out, out_off, out_size, status = engine.BatchCompressor(12).compress_flat(flat, off)
    assert (status == 0).all()
    exp_sizes = np.array([len(o.compress(b, 12)) for b in mixed], dtype=np.int64)
    tot_ref = int(exp_sizes[np.arange(n12) % 16].sum())
    assert int(out_size.sum()) <= tot_ref * 1.005
    for i in range(0, n12, 97):
        s = int(out_off[i])
        assert zlib.decompress(out[s:s + int(out_size[i])].tobytes(), -15) == mixed[i % 16]
    assert ratio6 > 100          # corpus A is periodic: ~160:1 at level 6


def test_config5_mixed_pipeline_roundtrip_crc(engine):
    """configs[4] per-GPU shard shape (scaled to 512 MiB here): corpus B streams, compress (level 6,
    gzip framing) -> decompress -> CRC-32 of the output equals CRC-32 of the input, which the
    decompressor has also checked against the gzip footer."""
    n = 8192
    base = [corpus.corpus_b_stream(k) for k in range(64)]
    crc = np.array([zlib.crc32(b) for b in base], dtype=np.uint32)
    flat, off = _tiled(base, n)
    c = engine.BatchCompressor(6, format=engine.GZIP)
    out, out_off, out_size, status = c.compress_flat(flat, off)
    assert (status == 0).all()
    # compact the slab (bound-spaced) into a dense flat buffer for the decompressor
    sizes = out_size.astype(np.int64)
    comp = [out[int(out_off[i]):int(out_off[i]) + int(sizes[i])] for i in range(n)]
    cflat = np.concatenate(comp)
    coff = np.zeros(n + 1, dtype=np.uint64)
    coff[1:] = np.cumsum(sizes)
    d = engine.BatchDecompressor(format=engine.GZIP)
    dout, dout_off, dsize, dstatus, sums = d.decompress_flat(cflat, coff, np.full(n, STREAM, dtype=np.uint64),
                                                            want_checksum=True)
    assert (dstatus == 0).all() and (dsize == STREAM).all()
    assert (sums == crc[np.arange(n) % 64]).all()
    assert engine.checksum_batch([dout[i * STREAM:(i + 1) * STREAM].tobytes() for i in range(0, n, 1024)],
                                 engine.CRC32) == [int(crc[i % 64]) for i in range(0, n, 1024)]
    # sampled byte identity of the compressed streams against the oracle
    for i in range(0, 64, 7):
        assert comp[i].tobytes() == o.compress(base[i], 6, o.GZIP)
    for i in range(0, n, 509):
        assert dout[i * STREAM:(i + 1) * STREAM].tobytes() == base[i % 64]
