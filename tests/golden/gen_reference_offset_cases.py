"""Extracts the inputs of the reference's tests/offset_tests.rs (50 periodic-pattern round trips)
into tests/golden/reference_offset_cases.json: test name, compression level, pattern, length.
Only INPUTS are taken (the reference tests are round trips: no expected bytes exist).  Needs
/root/reference; the JSON is committed because the GPU box does not have the reference tree."""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/tests/offset_tests.rs"


def main():
    src = open(REF).read()
    cases = []
    for m in re.finditer(r"fn (test_offset\w+)\(\) \{(.*?)\n\}\n", src, re.S):
        name, body = m.group(1), m.group(2)
        lvl = re.search(r"Compressor::new\((\d+)\)", body)
        pat = re.search(r'b"((?:[^"\\]|\\.)*)"', body)
        take = re.search(r"\.take\((\d+)\)", body)
        n = int(take.group(1)) if take else None
        if n is None:                                  # the two cases that spell the length as an expression
            if re.search(r"pattern_len = 100 \* 1024", body):
                n = 100 * 1024
            elif re.search(r"\.take\(1600 \+ 7\)", body):
                n = 1607
        if not (lvl and pat and n is not None):
            cases.append({"name": name, "unparsed": True})
            continue
        raw = bytes(pat.group(1), "utf-8").decode("unicode_escape").encode("latin-1")
        line = src[:m.start()].count("\n") + 1
        cases.append({"name": name, "level": int(lvl.group(1)), "pattern_hex": raw.hex(),
                      "take": n, "cite": f"tests/offset_tests.rs:{line}"})
    with open(os.path.join(HERE, "reference_offset_cases.json"), "w") as f:
        json.dump(cases, f, indent=0)
    print(len(cases), "cases,", sum(1 for c in cases if c.get("unparsed")), "unparsed")


if __name__ == "__main__":
    main()
