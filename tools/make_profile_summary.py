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
    out += ["", f"## Launch list (`{os.path.basename(launches)}`, `--metrics gpu__time_duration.sum`), by kernel", "",
            "The command is the default `bench.py` run: warm-up + timed steps of the headline (one",
            "`inflate_kernel<1, 16>` launch per step, nothing else inside the timed region), the end-to-end arm,",
            "then the mixed-corpus pipeline section (deflate_hc / gather / inflate_kernel<2, 16>).  torch's own",
            "`at::` kernels are the parity checks between the sections.", "",
            "| kernel | launches | total ns | mean ns | share of all listed |", "|---|---|---|---|---|"]
    tot = sum(float(r[14].replace(",", "")) for r in ls) or 1
    agg = {}
    for r in ls:
        name = r[4].split("(")[0][:60]
        a_ = agg.setdefault(name, [0, 0.0, r[8], r[7]])
        a_[0] += 1
        a_[1] += float(r[14].replace(",", ""))
    for name, (cnt, ns, grid, block) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{name}` | {cnt} | {ns:.0f} | {ns / cnt:.0f} | {100 * ns / tot:.1f} % |")
    ours = sum(v[1] for k, v in agg.items() if "bdf::" in k)
    out.append("")
    out.append(f"Share of this repo's kernels (`bdf::*`) in the listed GPU time: {100 * ours / tot:.1f} %.  "
               f"Headline step = one `inflate_kernel<1, 16>` launch = 100 % of its timed region.")
open(os.path.join(ROOT, "profiles", f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
json.dump({"kernel": kernel, "streams_in_capture": nstreams, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_launch": rd + wr, "dram_bytes_per_stream": (rd + wr) / nstreams,
           "report": os.path.basename(rep)},
          open(os.path.join(ROOT, "profiles", "inflate_traffic.json"), "w"), indent=1)
print("\n".join(out[:40]))
