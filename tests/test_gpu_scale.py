"""GPU tests at the edges of the batch shape: one very large stream, very many tiny
streams, ragged mixes, and concurrent callers on one context."""
import threading
import zlib

import numpy as np
import pytest

import corpus
import oracle_lib as o

pytestmark = pytest.mark.gpu


def test_one_large_stream_every_framing(engine):
    """A 48 MiB stream (text, a long run of zeroes, periodic data, text again): positions above
    2^24, Adler-32 folding, coalesced matches at the 64 KiB cap, CRC-32 over tens of MiB."""
    t = corpus.text_stream(2)
    blob = (t * 180) + bytes(20 << 20) + corpus.corpus_a_stream(3) * 200 + (t * 100)
    for fmt, wb in ((0, -15), (1, 15), (2, 31)):
        z = zlib.compressobj(6, zlib.DEFLATED, wb)
        comp = z.compress(blob) + z.flush()
        d = engine.BatchDecompressor(format=fmt)
        flat, off = engine.flatten([comp])
        out, out_off, out_size, status, sums = d.decompress_flat(flat, off, [len(blob)], want_checksum=True)
        assert status[0] == 0 and int(out_size[0]) == len(blob)
        assert out[:len(blob)].tobytes() == blob
        if fmt == 1:
            assert int(sums[0]) == zlib.adler32(blob)
        if fmt == 2:
            assert int(sums[0]) == zlib.crc32(blob)
    # and the engine's own compressor on it (chunked path, 190 chunks), through its own decompressor
    comp = engine.BatchCompressor(1, format=2).compress_batch([blob])[0]
    assert zlib.decompress(comp, 31) == blob
    assert engine.BatchDecompressor(format=2).decompress_batch([comp], [len(blob)]) == [blob]


def test_many_tiny_streams(engine):
    n = 300_000
    rng = np.random.default_rng(3)
    words = [bytes(rng.integers(97, 123, int(k), dtype=np.uint8)) for k in rng.integers(0, 24, 64)]
    bufs = [words[i % 64] for i in range(n)]
    comp = engine.BatchCompressor(6).compress_batch(bufs)
    exp = [o.compress(w, 6) for w in words]
    assert all(comp[i] == exp[i % 64] for i in range(n))
    back = engine.BatchDecompressor().decompress_batch(comp, [len(b) for b in bufs])
    assert all(back[i] == bufs[i] for i in range(0, n, 997)) and all(b is not None for b in back)
    assert engine.checksum_batch(bufs[:5000], engine.CRC32) == [zlib.crc32(b) for b in bufs[:5000]]


def test_ragged_batch_sizes_spanning_five_orders(engine):
    sizes = [0, 1, 2, 3, 15, 16, 17, 255, 256, 257, 4095, 4096, 65535, 65536, 65537, 262144, 262145, 1_000_003]
    src = (corpus.text_stream(4) + corpus.binary_stream(4) + corpus.lowentropy_stream(4)) * 6
    bufs = [src[:s] for s in sizes] * 3
    for level, fmt in ((1, 1), (6, 0), (9, 2)):
        comp = engine.BatchCompressor(level, format=fmt).compress_batch(bufs)
        for c, b in zip(comp, bufs):
            assert c == o.compress(b, level, fmt)
        assert engine.BatchDecompressor(format=fmt).decompress_batch(comp, [len(b) for b in bufs]) == bufs


def test_concurrent_callers_share_one_context(engine):
    """The reference types are Sync (src/batch.rs: &self + OnceLock); the ctx serialises whole host
    calls (its staging buffers are shared).  Every thread has its own data, so a mix-up shows."""
    errors = []

    def worker(seed):
        try:
            bufs = [corpus.text_stream(seed * 16 + k, 20000 + 1000 * seed) for k in range(6)]
            comp = [o.compress(b, 6, 1) for b in bufs]
            sums = [zlib.crc32(b) for b in bufs]
            for it in range(8):
                kind = (seed + it) % 3
                if kind == 0:
                    assert engine.BatchDecompressor(format=1).decompress_batch(comp, [len(b) for b in bufs]) == bufs
                elif kind == 1:
                    assert engine.BatchCompressor(6, format=1).compress_batch(bufs) == comp
                else:
                    assert engine.checksum_batch(bufs, engine.CRC32) == sums
        except Exception as e:          # noqa: BLE001
            errors.append(repr(e)[:300])

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_sharded_batch_over_all_devices(engine):
    """Single-process sharding over every GPU of the box (1 on the standard test box, N under
    `gpurun --gpus N`): one ctx + one host thread per device, results in stream order."""
    from libdeflate_rsx_b200 import shard
    sb = shard.ShardedBatch()
    assert len(sb.contexts) >= 1
    bufs = [corpus.corpus_b_stream(k, 2000 + 997 * (k % 23)) for k in range(200)] + [b"", b"z"]
    comp = sb.compress_batch(bufs, 6, 2)
    exp = {}
    for i in range(0, len(bufs), 13):
        assert comp[i] == exp.setdefault(i, o.compress(bufs[i], 6, 2))
    assert sb.decompress_batch(comp, [len(b) for b in bufs], 2) == bufs
    assert sb.checksum_batch(bufs, engine.CRC32) == [zlib.crc32(b) for b in bufs]
