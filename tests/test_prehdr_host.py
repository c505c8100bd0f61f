"""CPU tests of the first-block header pre-pass (csrc/inflate_prehdr.cuh).

prehdr_decode is compiled for the host from the kernel's own text (tests/host_harness) and compared
with an independent restatement of read_dynamic_huffman_header (reference
src/decompress/mod.rs:403-507) on streams from zlib, from the oracle's compressor and on damaged
streams.  The contract: a row marked valid holds exactly the code lengths the in-kernel reader
(read_code_lengths, csrc/inflate.cuh) produces and ends on the same bit; everything else is 0 and
is left to the inflate kernels.
"""
import os
import random
import sys
import zlib

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import corpus  # noqa: E402
from host_harness import prehdr_host  # noqa: E402

VALID = 1 << 31
ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]


class Bits:
    def __init__(self, data):
        self.d = data
        self.pos = 0

    def take(self, n):
        v = 0
        for k in range(n):
            byte = self.d[self.pos >> 3] if (self.pos >> 3) < len(self.d) else 0     # zero fill past the end
            v |= ((byte >> (self.pos & 7)) & 1) << k
            self.pos += 1
        return v


def ref_header(data):
    """-> None where the pre-pass must not claim the header, else (lens, bits, nlit, noff, final)."""
    b = Bits(data)
    if 3 > len(data) * 8:
        return None
    final = b.take(1)
    if b.take(2) != 2:
        return None
    nlit, noff, npre = 257 + b.take(5), 1 + b.take(5), 4 + b.take(4)
    pl = [0] * 19
    for k in range(npre):
        pl[ORDER[k]] = b.take(3)
    if b.pos > len(data) * 8:
        return None
    if sum(1 << (7 - l) for l in pl if l) != 128:
        return None          # not a complete precode: the kernels decide (incomplete codes have their own rules)
    # canonical codes, LSB-first lookup
    code, nxt = 0, [0] * 9
    cnt = [pl.count(l) for l in range(8)]
    cnt[0] = 0
    for l in range(1, 8):
        code = (code + cnt[l - 1]) << 1
        nxt[l] = code
    table = {}
    for s in range(19):
        l = pl[s]
        if l:
            c = nxt[l]
            nxt[l] += 1
            table[(l, int(format(c, "0%db" % l)[::-1], 2))] = s
    lens, prev, total = [], 0, nlit + noff
    while len(lens) < total:
        sym = None
        acc = 0
        for l in range(1, 8):
            acc |= b.take(1) << (l - 1)
            if (l, acc) in table:
                sym = table[(l, acc)]
                break
        assert sym is not None
        if sym < 16:
            lens.append(sym)
            prev = sym
            continue
        if sym == 16:
            if not lens:
                return None
            rep, val = 3 + b.take(2), prev
        elif sym == 17:
            rep, val = 3 + b.take(3), 0
        else:
            rep, val = 11 + b.take(7), 0
        rep = min(rep, total - len(lens))
        lens += [val] * rep
        prev = val
    if b.pos > len(data) * 8:
        return None
    return lens, b.pos, nlit, noff, final


@pytest.fixture(scope="module")
def lib():
    return prehdr_host.build()


def check(lib, data, lane=0, misalign=0):
    want = ref_header(data)
    meta, row = prehdr_host.decode(lib, data, lane, misalign)
    if want is None:
        assert meta == 0, (meta, len(data))
        return False
    lens, bits, nlit, noff, final = want
    assert meta & VALID
    assert meta & 0xFFFF == bits
    assert 257 + (meta >> 16 & 31) == nlit and 1 + (meta >> 21 & 31) == noff and (meta >> 26 & 1) == final
    assert meta & ~(VALID | 0x7FFFFFF) == 0
    assert list(row[:nlit + noff]) == lens
    assert not any(row[nlit + noff:])
    return True


def raw_deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()


def sample_streams():
    rnd = random.Random(5)
    plains = [corpus.corpus_a_stream(k) for k in range(6)]
    plains += [corpus.text_stream(k, 3000 + 9000 * k) for k in range(5)]
    plains += [bytes(rnd.getrandbits(8) for _ in range(n)) for n in (40, 700, 5000)]
    plains += [bytes(rnd.choice(b"abcdefgh") for _ in range(n)) for n in (10, 300, 20000)]
    plains += [b"", b"a", b"abc" * 50, bytes(range(256)) * 8, b"\0" * 70000]
    out = []
    for p in plains:
        for level in (1, 6, 9):
            out.append(raw_deflate(p, level))
        out.append(raw_deflate(p, 6, zlib.Z_HUFFMAN_ONLY))
        out.append(raw_deflate(p, 6, zlib.Z_FIXED))
    return out


def test_valid_headers_match_restatement(lib):
    n_valid = 0
    for k, s in enumerate(sample_streams()):
        n_valid += check(lib, s, lane=k % 32, misalign=k % 4)
    assert n_valid >= 40


def test_oracle_streams(lib):
    import oracle_lib as o
    n_valid = 0
    for k in range(8):
        p = corpus.text_stream(k, 12000) if k % 2 else corpus.corpus_a_stream(k)
        for level in (1, 2, 6, 9, 12):
            n_valid += check(lib, o.compress(p, level, 0), lane=k, misalign=(k + level) % 4)
    assert n_valid >= 20


def test_truncated_and_damaged(lib):
    rnd = random.Random(11)
    base = [raw_deflate(corpus.text_stream(3, 8000)), raw_deflate(corpus.corpus_a_stream(1)),
            raw_deflate(bytes(rnd.choice(b"abcdefghijklmnopq\0\1\2") for _ in range(3000)), 6, zlib.Z_HUFFMAN_ONLY)]
    seen = {True: 0, False: 0}
    for s in base:
        hdr_bytes = (ref_header(s)[1] + 7) // 8
        for cut in list(range(0, min(len(s), hdr_bytes + 3))):
            seen[check(lib, s[:cut], misalign=cut % 4)] += 1
        for _ in range(400):
            d = bytearray(s[:hdr_bytes + 8])
            for _ in range(rnd.randint(1, 3)):
                d[rnd.randrange(min(len(d), hdr_bytes))] ^= 1 << rnd.randrange(8)
            seen[check(lib, bytes(d), lane=rnd.randrange(32), misalign=rnd.randrange(4))] += 1
    assert seen[True] > 100 and seen[False] > 100


def test_random_bits(lib):
    rnd = random.Random(3)
    for _ in range(3000):
        n = rnd.randint(0, 120)
        d = bytearray(rnd.getrandbits(8) for _ in range(n))
        if d:
            d[0] = (d[0] & ~6) | 4          # block type 2
        check(lib, bytes(d), lane=rnd.randrange(32), misalign=rnd.randrange(4))


def test_header_in_pieces(lib):
    """the lane kernel decodes a later block's header a few symbols per round: same lengths, same end bit"""
    n = 0
    for k, s in enumerate(sample_streams()):
        want = ref_header(s)
        for budget in (1, 2, 7):
            meta, row, end = prehdr_host.decode_budget(lib, s, budget, misalign=k % 4)
            if want is None:
                assert meta == 0
                continue
            lens, bits, nlit, noff, final = want
            assert meta & VALID and end == bits
            assert 257 + (meta >> 16 & 31) == nlit and 1 + (meta >> 21 & 31) == noff and (meta >> 26 & 1) == final
            assert list(row[:nlit + noff]) == lens
            n += 1
    assert n >= 100
