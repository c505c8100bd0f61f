#!/usr/bin/env python3
"""Per-source-line dynamic profile of one kernel of an ncu report: executed warp instructions, stall
samples and active lanes per line (SASS rows of the report zipped with nvdisasm's line table of the
cubin that was profiled).  usage: tools/ncu_hot.py REPORT.ncu-rep LIB.so KERNEL_SUBSTR SRC_FILE [topN]"""
import collections, csv, glob, io, os, re, subprocess, sys, tempfile
rep, so, kern, srcfile = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cub = glob.glob(tmp + "/*.cubin")[0]
txt = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.split("\n")
inside, cur, secs = False, None, {}
for line in txt:
    if line.startswith(".text."):
        inside = kern.split("<")[0].split("IL")[0] in line
        name = line.strip()
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        secs.setdefault(name, []).append((cur, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, b = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        b = {"name": r[1], "rows": []}
        blocks.append(b)
    elif b is not None:
        b["rows"].append(r)
blk = next(x for x in blocks if kern.split("IL")[0].replace("_ZN3bdf", "").lstrip("0123456789") in x["name"] or kern in x["name"])
hdr, prof = blk["rows"][0], blk["rows"][1:]
ie, it, iss = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
isrc = hdr.index("Source")
prof = [r for r in prof if len(r) > ie]
# the template instance that was profiled: same length, same opcodes
dis = next((d for d in secs.values() if len(d) == len(prof) and
            all(x[1].split()[0] == p[isrc].split()[0] for x, p in zip(d, prof))), None)
assert dis is not None, ("no section of the .so matches the profiled SASS", len(prof), {k: len(v) for k, v in secs.items()})
agg, smp, thr = collections.Counter(), collections.Counter(), collections.Counter()
for d, p in zip(dis, prof):
    agg[d[0]] += int(p[ie]); smp[d[0]] += int(p[iss]); thr[d[0]] += int(p[it])
tot, ts = sum(agg.values()), sum(smp.values())
src = open(srcfile).read().split("\n")
name = os.path.basename(srcfile)
print(f"{blk['name']}: {tot} warp instructions, {ts} stall samples")
# regions: consecutive lines of the source file summed in buckets of the comment markers '// ----' / '// ===='
marks = [i + 1 for i, l in enumerate(src) if re.search(r"// (----|====)", l)]
reg = collections.Counter(); regs = collections.Counter()
for (f, l), v in agg.items():
    if f == name:
        m = max([x for x in marks if x <= l], default=0)
        reg[m] += v; regs[m] += smp[(f, l)]
print("-- by region (comment markers)")
for m, v in sorted(reg.items()):
    print(f"{100*v/tot:5.1f}% inst {100*regs[m]/ts:5.1f}% smp  line {m}: {src[m-1].strip()[:90] if m else '(top)'}")
other = sum(v for (f, l), v in agg.items() if f != name)
print(f"{100*other/tot:5.1f}% inst in inlined code of other files")
print("-- hottest lines")
for k, v in sorted(agg.items(), key=lambda x: -smp[x[0]])[:top]:
    t = src[k[1] - 1].strip()[:90] if k[0] == name else ""
    print(f"{100*v/tot:5.1f}% inst {100*smp[k]/ts:5.1f}% smp lanes {thr[k]/max(v,1):4.1f}  {k[0]}:{k[1]}  {t}")
