"""Loader for tests/golden/reference_kats.json (the reference's own known-answer vectors)."""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def _data(spec):
    if "ascii" in spec:
        return spec["ascii"].encode()
    if "repeat" in spec:
        return bytes.fromhex(spec["repeat"]) * spec["count"]
    if "mod255" in spec:
        return bytes(i % 255 for i in range(spec["mod255"]))
    if "ramp" in spec:
        return bytes(range(spec["ramp"]))
    raise KeyError(spec)


def load():
    with open(os.path.join(HERE, "golden", "reference_kats.json")) as f:
        k = json.load(f)
    out = {
        "adler32": [(_data(e["data"]), e["expect"], e["cite"]) for e in k["adler32"]],
        "crc32": [(_data(e["data"]), e["expect"], e["cite"]) for e in k["crc32"]],
        "inflate": [(bytes.fromhex(e["stream_hex"]), e["expect_ascii"].encode(), e["cite"])
                    for e in k["inflate"]],
        "inflate_must_fail": [(bytes.fromhex(e["stream_hex"]), e["formats"], e["cite"])
                              for e in k["inflate_must_fail"]],
    }
    # tests/unit_tests.rs:352-368 — tail sizes of (i % 255), checked there against C libdeflate
    out["crc_tail_sizes"] = [0, 1, 7, 8, 15, 16, 20, 28, 31, 32, 100, 108, 128, 1024, 1036]
    return out
