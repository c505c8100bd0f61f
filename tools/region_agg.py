#!/usr/bin/env python
"""Aggregates tools/ncu_lines.py output by code region of inflate.cuh.  usage: region_agg.py LINES.txt NSTREAMS"""
import re, sys
out = open(sys.argv[1]).read()
nstreams = int(sys.argv[2])
src = open("libdeflate_rsx_b200/csrc/inflate.cuh").read().splitlines()
def find(s):
    for i, l in enumerate(src):
        if s in l: return i + 1
    return 10**9
marks = [("bitreader", find("struct BitReader")), ("build", find("// ------------------------------------------------------------ table building")),
         ("outstate+zfill", find("// ------------------------------------------------------------------- output")),
         ("copy", find("// ---- match copy")), ("decode", find("// ------------------------------------------------------------ block decoding")),
         ("header", find("// read_dynamic_huffman_header")), ("static", find("__device__ void load_static_codes")),
         ("stream", find("// Raw DEFLATE stream [p, p+len)")), ("crc/adler-finish", find("// Group-wide CRC-32")), ("kernel", find("struct InflateArgs"))]
tot = int(re.search(r"(\d+) warp-instructions", out).group(1))
agg = {}
for l in out.splitlines()[1:]:
    m = re.match(r"\s*([\d.]+)% inst\s+([\d.]+)% smp\s+(\S+):(\d+)", l)
    if not m: continue
    pi, ps, f, ln = float(m.group(1)), float(m.group(2)), m.group(3), int(m.group(4))
    key = f
    if f == "inflate.cuh":
        key = "pre"
        for name, start in marks:
            if ln >= start: key = name
    a = agg.setdefault(key, [0, 0]); a[0] += pi; a[1] += ps
print(f"total {tot} warp-instr = {tot/nstreams:.0f} per stream")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if v[0] >= 0.1: print(f"{k:28s} inst {v[0]:5.1f}% = {v[0]/100*tot/nstreams:8.0f}/stream   samples {v[1]:5.1f}%")
