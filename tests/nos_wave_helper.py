"""Helper of test_gpu_nos_waves.py: level 10 / 12 over more streams than one wave holds; prints a digest."""
import hashlib
import sys
import zlib

import numpy as np

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
import libdeflate_rsx_b200 as bdf
from test_gpu_fuzz import random_buffer

rng = np.random.default_rng(5)
bufs = [random_buffer(rng, 6000) for _ in range(500)] + [b"", b"x"]
for level in (10, 12):
    got = bdf.BatchCompressor(level, format=0).compress_batch(bufs)
    for g, s in zip(got, bufs):
        assert (g == b"" and len(s) > 0) or zlib.decompress(g, -15) == s      # incompressible buffers fail in-band
    print("digest", level, hashlib.sha256(b"\0".join(got)).hexdigest())
