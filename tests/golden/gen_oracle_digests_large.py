"""Generates tests/golden/oracle_digests_large.json: size + SHA-256 of the oracle's output on
inputs above 64 KiB (level-1 block-split path, src/compress/mod.rs:1531-1564) and above 256 KiB
(chunks joined by sync flushes, :699-772), and of single chunks compressed with an explicit flush
mode (`orc_compress_unit`, the call DeflateEncoder::flush_buffer makes, src/stream.rs:42-196).
Like oracle_digests.json these pin THIS REPOSITORY'S restatement, not the reference binary."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

LEVELS = [0, 1, 2, 6, 9, 10]


def inputs():
    import corpus
    t = corpus.text_stream(6, 65536)
    yield "text_70000", corpus.text_stream(6, 70000)
    yield "text_262145", (t * 5)[:262145]
    yield "binary_600001", (corpus.binary_stream(4) * 10)[:600001]
    yield "corpus_a_1mib", b"".join(corpus.corpus_a_stream(k) for k in range(16))


def units():
    import corpus
    yield "unit_text_sync", corpus.text_stream(2, 100000), 0, 1
    yield "unit_text_finish", corpus.text_stream(2, 100000), 1, 0
    yield "unit_empty_finish", b"", 1, 0
    yield "unit_corpus_a_sync", corpus.corpus_a_stream(3) * 4, 0, 1


def digest(c):
    return None if c is None else [len(c), hashlib.sha256(c).hexdigest()]


def generate():
    import oracle_lib as o
    out = {}
    for name, data in inputs():
        out[name] = {str(level): digest(o.compress(data, level, level % 3)) for level in LEVELS}
    for name, data, fin, sync in units():
        out[name] = {str(level): digest(o.compress_unit(data, level, fin, sync)) for level in LEVELS}
    return out


if __name__ == "__main__":
    out = generate()
    with open(os.path.join(HERE, "oracle_digests_large.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", len(out), "entries")
