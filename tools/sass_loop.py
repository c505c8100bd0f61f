#!/usr/bin/env python3
"""Counts the SASS instructions of the address range spanned by the source lines
[first_line, last_line] of one kernel (inlined callees included) and lists the opcodes.
usage: tools/sass_loop.py <cubin> <kernel substring> <source file> <first_line> [last_line]"""
import collections
import re
import subprocess
import sys

cubin, kern, srcname, L0 = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
L1 = int(sys.argv[5]) if len(sys.argv) > 5 else 1 << 30
txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
inside, cur, rows = False, None, []
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
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        rows.append((int(m.group(1), 16), cur, m.group(2)))
addrs = [a for a, c, _ in rows if c[0] == srcname and L0 <= c[1] <= L1]
lo, hi = min(addrs), max(addrs)
body = [r for r in rows if lo <= r[0] <= hi]
print(f"address range {lo:x}-{hi:x}: {len(body)} instructions")
by = collections.Counter(c for a, c, _ in body)
for (f, l), v in sorted(by.items()):
    if v >= 8:
        print(f"  {f}:{l}  {v}")
ops = collections.Counter((i.split()[1] if i.startswith("@") else i.split()[0]) for a, c, i in body)
print(ops.most_common(16))
