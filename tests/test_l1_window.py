"""Host check of the level-1 whole-window round (csrc/deflate_l1.cuh): the lane-level model in
tools/l1_window_sim.py must give the serial parse of the reference's HtMatchFinder
(src/compress/matchfinder.rs:1139-1231) token for token, and that parse, bit-packed as one static
block, must be the oracle's level-1 output.  The CUDA code follows the model step for step; the GPU
tests compare the kernel itself with the oracle."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import corpus
import oracle_lib as o
from l1_window_sim import encode_static, serial_tokens, window_tokens


def cases():
    rng = np.random.default_rng(7)
    out = [("text", corpus.text_stream(1, 12000)), ("binary", corpus.binary_stream(2, 12000)),
           ("lowent", corpus.lowentropy_stream(3, 12000)), ("corpusA", corpus.corpus_a_stream(0, 6000)),
           ("mixed", corpus.corpus_b_stream(5, 12000)),
           ("alphabet4", rng.integers(0, 4, 6000, dtype=np.uint8).tobytes()),
           ("alphabet2", rng.integers(0, 2, 3000, dtype=np.uint8).tobytes()),
           ("runs", b"".join(bytes([int(rng.integers(0, 256))]) * int(rng.integers(1, 700)) for _ in range(30)))]
    out += [("offset%d" % n, corpus.offset_stream(n, 3000)) for n in sorted(corpus.OFFSET_PATTERNS)]
    out += [("period%d" % n, (bytes(rng.integers(0, 256, n, dtype=np.uint8)) * (3000 // n + 1))[:3000]) for n in (5, 31, 33, 257, 259, 300)]
    out += [("small%d" % i, s) for i, s in enumerate(corpus.small_cases()) if len(s) <= 12000]
    return [(name, bytes(s)) for name, s in out]


@pytest.mark.parametrize("name,data", cases(), ids=[c[0] for c in cases()])
def test_window_round_is_the_serial_parse(name, data):
    serial = serial_tokens(data)
    for cap in (None, 16, 3):
        got, rounds = window_tokens(data, spec_cap=cap)
        assert got == serial, (name, cap)
    if data:
        assert rounds <= len(data) // 32 + 1 + sum(isinstance(t, tuple) for t in serial)
        exp = o.compress(data, 1)
        if exp is not None:                       # None: the static block does not fit the bound (in-band failure)
            assert encode_static(serial) == exp
