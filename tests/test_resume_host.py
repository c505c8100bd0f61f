"""CPU tests of the resumable decoder step (csrc/inflate_resume_core.h) and of the incremental
DeflateDecoder mirror (libdeflate_rsx_b200/stream.py, reference src/stream.rs:243-376).

The decoder core is plain C++ shared with the kernel; tests/host_harness compiles the same header
with g++ and the mirror runs on it instead of on bdf_inflate_resume_batch_host.  What is checked is
what the reference's stream tests check (tests/stream_test.rs) plus the property the resumable design
rests on: the output does not depend on how the input and the reads are cut."""
import io
import os
import random
import sys
import zlib

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import corpus  # noqa: E402
from host_harness import resume_host  # noqa: E402

stream = pytest.importorskip("libdeflate_rsx_b200.stream")

OK, BAD_DATA, INSUFFICIENT_SPACE, SHORT_INPUT = 0, 1, 3, 4


class Dribble(io.RawIOBase):
    """an inner reader that hands out at most `piece` bytes per read"""

    def __init__(self, data, pieces):
        self.data, self.pos, self.pieces, self.k = data, 0, pieces, 0

    def read(self, n=-1):
        piece = self.pieces[self.k % len(self.pieces)]
        self.k += 1
        if n is not None and n >= 0:
            piece = min(piece, n)
        chunk = self.data[self.pos:self.pos + piece]
        self.pos += len(chunk)
        return chunk


def raw(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()


def multi_block(parts):
    """several blocks of different kinds in one stream (full flushes in between)"""
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    out = b""
    for p in parts:
        out += c.compress(p) + c.flush(zlib.Z_FULL_FLUSH)
    return out + c.flush()


def plains():
    rnd = random.Random(9)
    return [
        b"", b"a", b"hello world " * 40,
        corpus.corpus_a_stream(2), corpus.text_stream(1, 200000), corpus.binary_stream(3),
        bytes(rnd.getrandbits(8) for _ in range(70000)),          # stored blocks
        b"\0" * 300000,                                             # long runs: offset 1, length 258
        bytes(rnd.choice(b"ab") for _ in range(50000)),
        corpus.text_stream(5, 40000) + bytes(rnd.getrandbits(8) for _ in range(40000)) + corpus.corpus_a_stream(0),
    ]


def decode_all(comp, pieces=(1 << 20,), reads=(1 << 20,)):
    dec = stream.DeflateDecoder(Dribble(comp, pieces), _step=resume_host.step)
    out = bytearray()
    k = 0
    while True:
        chunk = dec.read(reads[k % len(reads)])
        k += 1
        if not chunk:
            return bytes(out), dec
        out += chunk


def test_state_size():
    assert resume_host.build().resume_state_size() == stream.STATE_BYTES


@pytest.mark.parametrize("level,strategy", [(1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                                            (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (0, zlib.Z_DEFAULT_STRATEGY)])
def test_whole_streams(level, strategy):
    for p in plains():
        got, dec = decode_all(raw(p, level, strategy))
        assert got == p
        assert dec.done


def test_output_independent_of_cuts():
    rnd = random.Random(4)
    p = plains()[-1] + corpus.text_stream(7, 150000)
    comp = multi_block([p[i:i + 37000] for i in range(0, len(p), 37000)])
    assert zlib.decompress(comp, -15) == p
    for trial in range(12):
        pieces = [rnd.choice((1, 2, 3, 7, 64, 500, 571, 4096, 70000)) for _ in range(5)]
        reads = [rnd.choice((1, 10, 257, 258, 259, 4000, 32768, 65536, 100000)) for _ in range(5)]
        if 1 in reads and trial % 3:
            reads = [r for r in reads if r != 1] or [10]
        got, dec = decode_all(comp, pieces, reads)
        assert got == p, (pieces, reads)
        assert dec.done and not dec.input[:0]


def test_oracle_streams():
    import oracle_lib as o
    for k, p in enumerate(plains()[2:6]):
        for level in (1, 6, 12):
            got, _ = decode_all(o.compress(p[:65536], level, 0), pieces=(100 + 37 * k,), reads=(5000,))
            assert got == p[:65536]


def test_stream_rs_round_trip_shape():
    # tests/stream_test.rs: read in small pieces until Ok(0)
    data = corpus.text_stream(2, 300000)
    got, dec = decode_all(raw(data), pieces=(8192,), reads=(10,))
    assert got == data
    assert dec.steps < 60            # ~32 KiB of output per step, not one step per read


def test_truncated_stream_is_unexpected_eof():
    comp = raw(corpus.text_stream(4, 100000))
    for cut in (1, 2, 50, 569, 570, 571, len(comp) // 2, len(comp) - 1):
        dec = stream.DeflateDecoder(io.BytesIO(comp[:cut]), _step=resume_host.step)
        with pytest.raises(stream.UnexpectedEof):
            dec.read_to_end()
    # a stored block cut inside its body, and inside LEN / NLEN
    comp = raw(os.urandom(1000), 0)
    for cut in (2, 4, 600):
        with pytest.raises(EOFError):
            stream.DeflateDecoder(io.BytesIO(comp[:cut]), _step=resume_host.step).read_to_end()


def test_input_that_ends_between_blocks_is_a_clean_end():
    # stream.rs:367-369: EOF with the decoder at a block boundary is Ok(0)
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    part = c.compress(b"first part " * 100) + c.flush(zlib.Z_FULL_FLUSH)
    got, dec = decode_all(part)
    assert got == b"first part " * 100 and dec.done
    assert stream.DeflateDecoder(io.BytesIO(b""), _step=resume_host.step).read(10) == b""


def test_bad_data():
    good = raw(corpus.text_stream(6, 50000))
    for bad in (b"\x07not deflate", b"\x00\x05\x00\x00\x00", good[:200] + bytes(200) + good[400:]):
        dec = stream.DeflateDecoder(io.BytesIO(bad), _step=resume_host.step)
        with pytest.raises(OSError):
            dec.read_to_end()
    # failure is sticky
    st = bytearray(stream.STATE_BYTES)
    win = np.zeros(70000, dtype=np.uint8)
    assert resume_host.step(st, b"\x07", True, win, 0)[0] == BAD_DATA
    assert resume_host.step(st, raw(b"abc"), True, win, 0)[0] == BAD_DATA


def test_offset_beyond_history_is_rejected():
    # static block, match with distance 1 at output position 0
    bits = "1" + "10" + "0000001" + "00000" + "0000000"     # BFINAL, static; length code 257 (len 3); distance code 0; EOB
    v = int(bits[::-1], 2)
    data = v.to_bytes((len(bits) + 7) // 8, "little")
    st = bytearray(stream.STATE_BYTES)
    assert resume_host.step(st, data, True, np.zeros(70000, dtype=np.uint8), 0)[0] == BAD_DATA
    with pytest.raises(zlib.error):
        zlib.decompress(data, -15)


def test_step_contract_directly():
    """status / consumed / state across hand-made cuts: the step never consumes what it cannot decode
    and never needs a byte twice"""
    rnd = random.Random(21)
    p = corpus.text_stream(8, 120000) + bytes(rnd.getrandbits(8) for _ in range(3000)) + b"z" * 5000
    comp = multi_block([p[:50000], p[50000:121000], p[121000:]])
    st = bytearray(stream.STATE_BYTES)
    win = np.zeros(32768 + 258 + 700, dtype=np.uint8)       # a tight window: many shifts
    pos, ipos, out, pending = 0, 0, bytearray(), b""
    final = False
    for _ in range(100000):
        if len(win) - pos < 258:
            keep = 32768
            out += win[:pos - keep].tobytes()
            win[:keep] = win[pos - keep:pos].copy()
            pos = keep
        status, used, new_pos = resume_host.step(st, pending, final, win, pos)
        assert used <= len(pending) and new_pos >= pos
        pending = pending[used:]
        pos = new_pos
        if status == OK:
            break
        assert status in (SHORT_INPUT, INSUFFICIENT_SPACE)
        if status == SHORT_INPUT:
            assert not final
            take = rnd.choice((1, 5, 100, 600, 3000))
            more = comp[ipos:ipos + take]
            ipos += len(more)
            pending += more
            final = ipos >= len(comp)
    else:
        raise AssertionError("no progress")
    out += win[:pos].tobytes()
    assert bytes(out) == p
    assert int.from_bytes(st[40:48], "little") == len(p)             # total_out
    assert int.from_bytes(st[32:40], "little") + len(pending) == len(comp)      # total_in


def test_tampered_state_is_rejected():
    """the state is plain caller-owned memory: a resumed state with impossible fields fails cleanly"""
    comp = raw(corpus.text_stream(2, 60000))
    for field, value in (("lens", 200), ("nlit", 400), ("noff", 0), ("bit_off", 9), ("phase", 7)):
        st = bytearray(stream.STATE_BYTES)
        win = np.zeros(70000, dtype=np.uint8)
        status, used, pos = resume_host.step(st, comp[:3000], False, win, 0)
        assert status == SHORT_INPUT and int.from_bytes(st[0:4], "little") == 2        # inside a Huffman block
        if field == "lens":
            st[48 + 5] = value
        else:
            off = {"phase": 0, "bit_off": 8, "nlit": 16, "noff": 20}[field]
            st[off:off + 4] = value.to_bytes(4, "little")
        assert resume_host.step(st, comp[used:], True, win, pos)[0] == BAD_DATA
