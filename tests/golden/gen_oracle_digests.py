"""Generates tests/golden/oracle_digests.json: size + SHA-256 of the oracle's
compressed output for a fixed set of inputs and every level.

These are digests of THIS REPOSITORY'S oracle (the reference's Rust code cannot
be run here — no cargo/rustc), so they pin the restatement against accidental
change; they are not outputs of the reference binary.  Re-run only when the
restatement is deliberately corrected:  python tests/golden/gen_oracle_digests.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

LEVELS = list(range(0, 13))


def inputs():
    import corpus
    yield "corpus_a_k0", corpus.corpus_a_stream(0)
    yield "corpus_a_k7", corpus.corpus_a_stream(7)
    yield "text_k0", corpus.text_stream(0)
    yield "binary_k1", corpus.binary_stream(1)
    yield "lowentropy_k3", corpus.lowentropy_stream(3)
    yield "offset9", corpus.offset_stream(9)
    for i, c in enumerate(corpus.small_cases()):
        yield f"small_{i}", c


if __name__ == "__main__":
    import oracle_lib as o
    out = {}
    for name, data in inputs():
        out[name] = {}
        for level in LEVELS:
            c = o.compress(data, level)
            out[name][str(level)] = None if c is None else [len(c), hashlib.sha256(c).hexdigest()]
    with open(os.path.join(HERE, "oracle_digests.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", len(out), "inputs")
