"""Deterministic test / bench corpora.

corpus A ("gen_bench"): exactly reference scripts/gen_bench_files.py:4-27 —
100-byte pattern ((i*1234567) ^ (i*987654)) & 0xFF tiled into a 1 MiB chunk
that is truncated and repeated; stream k is bytes [k*65536, (k+1)*65536).
corpus B ("mixed"): SURVEY.md §8(d) — text / structured binary / periodic /
low-entropy kinds from a seeded xorshift64* generator (no uniform-random
streams: the reference returns an empty result for incompressible input).
"""
import numpy as np

CHUNK = 1 << 20
STREAM = 1 << 16


def base_pattern():
    return bytes(((i * 1234567) ^ (i * 987654)) & 0xFF for i in range(100))


_tile_cache = {}


def tile_1mib(pattern=None):
    pattern = pattern or base_pattern()
    if pattern not in _tile_cache:
        _tile_cache[pattern] = (pattern * (CHUNK // len(pattern) + 1))[:CHUNK]
    return _tile_cache[pattern]


def corpus_a_stream(k, size=STREAM):
    """Stream k of the gen_bench file (the file repeats with period 1 MiB)."""
    t = tile_1mib()
    start = (k * size) % CHUNK
    out = bytearray()
    while len(out) < size:
        take = min(size - len(out), CHUNK - start)
        out += t[start:start + take]
        start = (start + take) % CHUNK
    return bytes(out)


OFFSET_PATTERNS = {
    1: b"1", 2: b"12", 3: b"123", 4: b"1234", 5: b"12345", 7: b"1234567", 8: b"12345678",
    9: b"123456789", 10: b"1234567890", 11: b"12345678901", 12: b"123456789012",
    13: b"1234567890123", 14: b"12345678901234", 15: b"123456789012345",
    16: b"1234567890123456", 17: b"12345678901234567", 18: b"123456789012345678",
    19: b"1234567890123456789", 20: b"ABCDEFGHIJKLMNOPQRST", 21: b"ABCDEFGHIJKLMNOPQRSTU",
    22: b"ABCDEFGHIJKLMNOPQRSTUV", 23: b"ABCDEFGHIJKLMNOPQRSTUVW", 24: b"ABCDEFGHIJKLMNOPQRSTUVWX",
    25: b"ABCDEFGHIJKLMNOPQRSTUVWXY", 26: b"ABCDEFGHIJKLMNOPQRSTUVWXYZ",
    27: b"ABCDEFGHIJKLMNOPQRSTUVWXYZ0", 28: b"ABCDEFGHIJKLMNOPQRSTUVWXYZ01",
    29: b"ABCDEFGHIJKLMNOPQRSTUVWXYZ012", 30: b"ABCDEFGHIJKLMNOPQRSTUVWXYZ0123",
    31: b"ABCDEFGHIJKLMNOPQRSTUVWXYZ01234", 32: b"ABCDEFGHIJKLMNOPQRSTUVWXYZ012345",
}


def offset_stream(n, size=STREAM):
    """data_offsetN.bin (gen_bench_files.py:45-80), first `size` bytes."""
    return tile_1mib(OFFSET_PATTERNS[n])[:size]


class XorShift:
    def __init__(self, seed=0x5EED):
        self.s = np.uint64(seed or 1)

    def next(self):
        s = int(self.s)
        s ^= s >> 12
        s ^= (s << 25) & 0xFFFFFFFFFFFFFFFF
        s ^= s >> 27
        self.s = np.uint64(s)
        return (s * 0x2545F4914F6CDD1D) & 0xFFFFFFFFFFFFFFFF


_vocab = None


def _vocabulary():
    global _vocab
    if _vocab is None:
        r = XorShift(0xC0FFEE)
        v = []
        for _ in range(4096):
            n = 2 + r.next() % 9
            v.append(bytes(97 + r.next() % 26 for _ in range(n)))
        _vocab = v
    return _vocab


def text_stream(k, size=STREAM):
    rng = np.random.default_rng(0x5EED + 4 * k)
    vocab = _vocabulary()
    ranks = np.minimum(rng.zipf(1.1, size // 3 + 16) - 1, len(vocab) - 1)
    out = bytearray()
    col = 0
    for r in ranks:
        w = vocab[int(r)]
        out += w
        col += len(w) + 1
        if col >= 80:
            out += b"\n"
            col = 0
        else:
            out += b" "
        if len(out) >= size:
            break
    return bytes(out[:size])


def binary_stream(k, size=STREAM):
    rng = np.random.default_rng(0x5EED + 4 * k + 1)
    nrec = size // 32 + 1
    rec = np.zeros((nrec, 32), dtype=np.uint8)
    ctr = np.arange(nrec, dtype=np.uint32) + np.uint32(k * 1000)
    rec[:, 0:4] = ctr.view(np.uint8).reshape(-1, 4)
    rec[:, 4:8] = (ctr * np.uint32(7 + k % 13)).view(np.uint8).reshape(-1, 4)
    rec[:, 8:10] = rng.integers(0, 16, nrec, dtype=np.uint16).view(np.uint8).reshape(-1, 2)
    slow = np.repeat(rng.integers(0, 256, (nrec // 64 + 1, 6), dtype=np.uint8), 64, axis=0)[:nrec]
    rec[:, 10:16] = slow
    dic = rng.integers(0, 256, (256, 16), dtype=np.uint8)
    rec[:, 16:32] = dic[rng.integers(0, 256, nrec)]
    return rec.tobytes()[:size]


def periodic_stream(k, size=STREAM):
    keys = sorted(OFFSET_PATTERNS)
    return offset_stream(keys[k % len(keys)], size)


def lowentropy_stream(k, size=STREAM):
    rng = np.random.default_rng(0x5EED + 4 * k + 3)
    return np.minimum(rng.geometric(0.35, size) - 1, 15).astype(np.uint8).tobytes()


def corpus_b_stream(k, size=STREAM):
    return (text_stream, binary_stream, periodic_stream, lowentropy_stream)[k % 4](k, size)


def small_cases():
    """Edge cases from the reference's tests (tests/batch_test.rs:5-11,54;
    tests/unit_tests.rs:112-125; tests/parallel_test.rs ramps)."""
    cases = [
        b"",
        b"a",
        b"ab",
        b"abc",
        b"Short",
        b"Not empty",
        b"Hello world! This is a test string for deflate compression.",
        b"Another test string.",
        b"Repeating pattern repeating pattern repeating pattern repeating pattern.",
        bytes(1000),
        b"a" * 10000,
        bytes(i % 256 for i in range(5000)),
        bytes((i * 3) % 251 for i in range(7000)),
        bytes((i * 7) % 251 for i in range(9000)),
        b"ABC" * 333 + b"A",
        b"ABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789abcdefghijklmnopqr" * 185,
    ]
    return cases
