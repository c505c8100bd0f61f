"""Helper of test_gpu_check_build.py: runs inflate / deflate batches of every kind through the
-DBDF_CHECK build (BDF_LIBRARY points at it) and prints the device assertion word."""
import sys

import numpy as np

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
import corpus
import oracle_lib as o
import libdeflate_rsx_b200 as bdf
from test_gpu_fuzz import random_buffer

ctx = bdf.default_context()
assert ctx._lib.bdf_debug_check_failures(ctx.handle) == 0, "not a -DBDF_CHECK build"
rng = np.random.default_rng(123)
bufs = [random_buffer(rng, 65536) for _ in range(120)] + [corpus.corpus_b_stream(k) for k in range(24)] + corpus.small_cases()
for level in (1, 2, 6, 9, 10, 12):
    got = bdf.BatchCompressor(level, format=level % 3).compress_batch(bufs)
    assert len(got) == len(bufs)
for fmt in (0, 1, 2):
    comp = [o.compress(b, 1 + i % 9, fmt) for i, b in enumerate(bufs)]
    keep = [(c, b) for c, b in zip(comp, bufs) if c is not None]
    # good streams, streams with a flipped bit, truncated streams, short output room
    streams, caps = [], []
    for i, (c, b) in enumerate(keep):
        if i % 4 == 1 and len(c) > 8:
            x = bytearray(c); x[len(x) // 2] ^= 0x10; c = bytes(x)
        elif i % 4 == 2 and len(c) > 4:
            c = c[:len(c) * 2 // 3]
        streams.append(c)
        caps.append(len(b) - 1 if i % 4 == 3 and len(b) else len(b) + i % 5)
    for mode in ("auto",):
        bdf.BatchDecompressor(format=fmt).decompress_batch(streams, caps)
print("check word", ctx._lib.bdf_debug_check_failures(ctx.handle))
