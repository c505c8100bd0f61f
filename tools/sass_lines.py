#!/usr/bin/env python3
"""Static per-source-line SASS instruction counts of one kernel (nvdisasm -g output): how many
instructions a loop body costs, before spending GPU time.
usage: tools/sass_lines.py <cubin> <kernel substring> <source file> [first_line [last_line]]"""
import collections
import re
import subprocess
import sys

cubin, kern, srcname = sys.argv[1:4]
lo = int(sys.argv[4]) if len(sys.argv) > 4 else 0
hi = int(sys.argv[5]) if len(sys.argv) > 5 else 1 << 30
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
cnt = collections.Counter()
inside = False
cur = None
for line in txt:
    if line.startswith(".text."):
        inside = kern in line
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", line) and cur:
        cnt[cur] += 1
src = {}
total = 0
print("kernel instructions:", sum(cnt.values()))
for (f, l), v in sorted(cnt.items(), key=lambda x: (x[0][0] != srcname, x[0][0], x[0][1])):
    if f == srcname and lo <= l <= hi:
        total += v
        if f not in src:
            try:
                src[f] = open(next(p for p in sys.argv[6:] + ["libdeflate_rsx_b200/csrc/" + f])).read().split("\n")
            except Exception:
                src[f] = []
        text = src[f][l - 1].strip()[:100] if l - 1 < len(src[f]) else ""
        print(f"{l:5d} {v:4d}  {text}")
print(f"instructions attributed to {srcname}:{lo}-{hi}: {total}")
