# small mixed workload for compute-sanitizer (memcheck): every kernel family, odd sizes and alignments
import sys, zlib
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import numpy as np
import corpus, oracle_lib as o
import libdeflate_rsx_b200 as b
bufs = corpus.small_cases() + [corpus.corpus_a_stream(1)[:20001], corpus.text_stream(2, 9999), corpus.binary_stream(3, 12345),
                               corpus.lowentropy_stream(4, 7777), bytes(5000), corpus.text_stream(5, 70001), b"ab" * 40000]
for fmt in (0, 1, 2):
    for lvl in (0, 1, 6, 10):
        comp = b.BatchCompressor(lvl, format=fmt).compress_batch(bufs)
        exp = [o.compress(s, lvl, fmt) or b"" for s in bufs]
        assert comp == exp, (fmt, lvl)
    comp = [o.compress(s, 6, fmt) for s in bufs]
    assert b.BatchDecompressor(format=fmt).decompress_batch(comp, [len(s) for s in bufs]) == bufs
    bad = [c[:len(c) // 2] for c in comp] + [bytes([7, 1, 2, 3])]
    b.BatchDecompressor(format=fmt).decompress_batch(bad, [len(s) for s in bufs] + [10])
assert b.checksum_batch(bufs, b.CRC32) == [zlib.crc32(s) for s in bufs]
assert b.checksum_batch(bufs, b.ADLER32) == [zlib.adler32(s) for s in bufs]
enc = b.DeflateEncoder(__import__("io").BytesIO(), 6)
enc.write(corpus.text_stream(6) * 5); out = enc.finish().getvalue()
assert zlib.decompress(out, -15) == corpus.text_stream(6) * 5
print("sanitize workload ok")
