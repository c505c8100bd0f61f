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


def big_cases():
    rng = np.random.default_rng(3)
    het = bytes(corpus.text_stream(1))[:30000] + bytes(corpus.binary_stream(2))[:30000] + \
        bytes(corpus.lowentropy_stream(3))[:30000] + bytes(corpus.corpus_a_stream(0))[:20000] + \
        bytes(rng.integers(0, 64, 20000, dtype=np.uint8)) + bytes(corpus.text_stream(4))[:30000]
    return [("heterogeneous", het),
            ("text", (bytes(corpus.text_stream(1)) + bytes(corpus.text_stream(7)))[:100000]),
            ("just_above", bytes(corpus.corpus_b_stream(2)) + b"xyz")]


@pytest.mark.parametrize("name,data", big_cases(), ids=[c[0] for c in big_cases()])
def test_window_rounds_with_block_splitting(name, data):
    """Inputs above 64 KiB: the window's symbols go to the split statistics in lane order (bulk while
    no cut is possible, one by one otherwise); blocks and bytes must be the reference's."""
    from l1_window_sim import serial_blocks, window_blocks
    blocks = window_blocks(data)
    assert blocks == serial_blocks(data)
    bulk, slow = window_blocks.last_rounds
    assert bulk > slow and (slow > 0 or name == "just_above")
    if name == "heterogeneous":
        assert len(blocks) >= 3
    assert encode_static(None, blocks) == o.compress(data, 1)
