"""Helper of test_gpu_determinism.py: one process = one set of engine switches (the environment
variables are read once per process / ctx); prints one digest line per (operation, level, format)."""
import hashlib
import sys

import numpy as np

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
import corpus
import oracle_lib as o
import libdeflate_rsx_b200 as bdf
from test_gpu_fuzz import random_buffer

rng = np.random.default_rng(99)
bufs = [random_buffer(rng, 65536) for _ in range(150)] + [corpus.corpus_b_stream(k) for k in range(40)]
for level in (1, 2, 6, 9):            # level 1: whole-window rounds or one match per round (BDF_L1_WINDOW)
    for fmt in (0, 2):
        got = bdf.BatchCompressor(level, format=fmt).compress_batch(bufs)
        print("compress", level, fmt, hashlib.sha256(b"\0".join(got)).hexdigest())
# levels 10..12 (one CTA per stream or three kernels per wave, BDF_NOS_SPLIT): the same parse either way
for level in (10, 12):
    got = bdf.BatchCompressor(level, format=1).compress_batch(bufs[:24] + bufs[150:166])
    print("compress", level, 1, hashlib.sha256(b"\0".join(got)).hexdigest())
for fmt in (0, 1, 2):
    comp = [o.compress(b, 1 + i % 9, fmt) or o.compress(b, 0, fmt) for i, b in enumerate(bufs)]
    comp = [c if c is not None else b"" for c in comp]
    got = bdf.BatchDecompressor(format=fmt).decompress_batch(comp, [len(b) + i % 7 for i, b in enumerate(bufs)])
    print("decompress", fmt, hashlib.sha256(b"\0".join(g if g is not None else b"<none>" for g in got)).hexdigest())
