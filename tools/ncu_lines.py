#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals for one kernel of an ncu report.

ncu's CSV source page is SASS-only, so the SASS offsets are joined with the
line table nvdisasm prints for the same cubin (the .so must be the build that
was profiled).   usage: ncu_lines.py REPORT.ncu-rep LIB.so KERNEL_SUBSTR [topN]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            blocks.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    for b in blocks:
        if kernel in b["name"]:
            hdr = b["rows"][0]
            return hdr, b["rows"][1:]
    raise SystemExit(f"kernel {kernel} not in report: {[b['name'] for b in blocks]}")


def line_table(so, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    table = {}
    for cb in cubins:
        txt = subprocess.run(["nvdisasm", "-g", "-c", cb], capture_output=True, text=True).stdout
        func, line, inl = None, None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                func = m.group(1)
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
            if m and func and kernel_match(func, kernel):
                table[int(m.group(1), 16)] = line
    return table


def kernel_match(mangled, kernel):
    k = re.sub(r"[^A-Za-z0-9_]", "", kernel.split("<")[0].split("::")[-1])
    if k not in mangled:
        return False
    ints = re.findall(r"\(int\)(\d+)", kernel)
    if ints:
        return "I" + "".join(f"Li{v}E" for v in ints) + "E" in mangled
    return True


def main():
    rep, so, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    hdr, rows = sass_rows(rep, kernel)
    ia, ie, ism = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    base = min(int(r[ia], 16) for r in rows if r and r[ia].startswith("0x"))
    # full kernel name for template matching
    table = line_table(so, kernel)
    agg, tot_i, tot_s = {}, 0, 0
    for r in rows:
        if not r or not r[ia].startswith("0x"):
            continue
        off = int(r[ia], 16) - base
        key = table.get(off, ("?", 0))
        i, s = int(r[ie] or 0), int(r[ism] or 0)
        a = agg.setdefault(key, [0, 0])
        a[0] += i
        a[1] += s
        tot_i += i
        tot_s += s
    srcs = {}
    print(f"kernel {kernel}: {tot_i} warp-instructions, {tot_s} stall samples, {len(table)} SASS lines mapped")
    for key, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        f, l = key
        text = ""
        for root in ("libdeflate_rsx_b200/csrc", "."):
            pth = os.path.join(root, f)
            if os.path.exists(pth):
                if pth not in srcs:
                    srcs[pth] = open(pth).read().splitlines()
                if 0 < l <= len(srcs[pth]):
                    text = srcs[pth][l - 1].strip()[:90]
                break
        print(f"{100 * i / max(tot_i, 1):5.1f}% inst {100 * s / max(tot_s, 1):5.1f}% smp  {f}:{l:<4} {text}")


if __name__ == "__main__":
    main()
