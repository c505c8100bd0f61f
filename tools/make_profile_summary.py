#!/usr/bin/env python
"""Writes profiles/<tag>_summary.md (+ traffic json) from an ncu report.
usage: make_profile_summary.py REPORT.ncu-rep KERNEL_NAME_SUBSTR TAG NSTREAMS [LAUNCHES.csv]"""
import csv, io, json, os, subprocess, sys

rep, kernel, tag, nstreams = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
launches = sys.argv[5] if len(sys.argv) > 5 else None
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
row = next(r for r in rows[2:] if kernel.split("<")[0] in r[hdr.index("Kernel Name")])
m = {h: (v, u) for h, u, v in zip(hdr, units, row)}
def g(k):
    return m.get(k, ("", ""))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
def to_bytes(v, u):
    f = float(v)
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
rd, wr = to_bytes(*g("dram__bytes_read.sum")), to_bytes(*g("dram__bytes_write.sum"))
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
out = [f"# ncu summary `{tag}` — kernel `{kernel}`", "",
       f"Source report: `{os.path.basename(rep)}` (`ncu --set full --clock-control none --import-source on`, one launch, "
       f"{nstreams} streams).  Cold-cache, serialised replay: compare shares, not absolutes.", "", "| metric | value | unit |", "|---|---|---|"]
for k in keys:
    v, u = g(k)
    if v != "":
        out.append(f"| `{k}` | {v} | {u} |")
out += ["", f"DRAM traffic per launch: read {rd/1e6:.1f} MB + write {wr/1e6:.1f} MB = {(rd+wr)/1e6:.1f} MB "
        f"({(rd+wr)/nstreams:.0f} B per stream; algorithmic C+U = 65936 B per stream for config 2)."]
lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep,
                        os.path.join(ROOT, "libdeflate_rsx_b200", "libbdeflate.so"), kernel, "25"],
                       capture_output=True, text=True).stdout
out += ["", "## Warp instructions and stall samples by source line (top 25)", "", "```", lines.rstrip(), "```"]
if launches and os.path.exists(launches):
    ls = [r for r in csv.reader(open(launches)) if len(r) > 14 and r[0].isdigit()]
    out += ["", f"## Launch list (`{os.path.basename(launches)}`, `--metrics gpu__time_duration.sum`)", "",
            "| # | kernel | grid | block | ns |", "|---|---|---|---|---|"]
    tot = sum(float(r[14].replace(",", "")) for r in ls) or 1
    for r in ls:
        out.append(f"| {r[0]} | `{r[4][:70]}` | {r[8]} | {r[7]} | {r[14]} |")
    share = sum(float(r[14].replace(",", "")) for r in ls if "inflate_kernel" in r[4]) / tot
    out.append("")
    out.append(f"inflate_kernel share of listed GPU time: {100*share:.1f} %")
open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
json.dump({"kernel": kernel, "streams_in_capture": nstreams, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_launch": rd + wr, "dram_bytes_per_stream": (rd + wr) / nstreams,
           "report": os.path.basename(rep)},
          open(os.path.join(ROOT, "profiles", "inflate_traffic.json"), "w"), indent=1)
print("\n".join(out[:40]))
