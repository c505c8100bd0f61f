#!/usr/bin/env python3
"""profiles/<tag>_summary.md + an entry of profiles/traffic.json from one ncu report.
usage: profile_summary.py REPORT.ncu-rep KERNEL_SUBSTR TAG NSTREAMS ALG_BYTES_PER_STREAM TRAFFIC_KEY SRC_FILE [note]"""
import csv, io, json, os, subprocess, sys

rep, kernel, tag, nstreams, alg, key, src = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), float(sys.argv[5]), sys.argv[6], sys.argv[7]
note = sys.argv[8] if len(sys.argv) > 8 else ""
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
row = next(r for r in rows[2:] if kernel in r[hdr.index("Kernel Name")])
m = {h: (v, u) for h, u, v in zip(hdr, units, row)}
g = lambda k: m.get(k, ("", ""))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
rd, wr = to_bytes(*g("dram__bytes_read.sum")), to_bytes(*g("dram__bytes_write.sum"))
inst = float(g("smsp__inst_executed.sum")[0] or 0)
out = [f"# ncu summary `{tag}` — kernel `{row[hdr.index('Kernel Name')]}`", "",
       f"Source report: `{os.path.basename(rep)}` (`ncu --set full --clock-control none --import-source on`, one launch, "
       f"{nstreams} streams).  Cold-cache, serialised replay: compare shares, not absolutes.  {note}", "",
       "| metric | value | unit |", "|---|---|---|"]
for k in keys:
    v, u = g(k)
    if v != "":
        out.append(f"| `{k}` | {v} | {u} |")
out += ["", f"DRAM traffic per launch: read {rd/1e6:.1f} MB + write {wr/1e6:.1f} MB = {(rd+wr)/1e6:.1f} MB = "
        f"{(rd+wr)/nstreams:.0f} B per stream against {alg:.0f} algorithmic bytes per stream "
        f"(**{(rd+wr)/nstreams/alg:.2f}x**).  Warp instructions per uncompressed byte: {inst/nstreams/65536:.2f}."]
hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep,
                      os.path.join(ROOT, "libdeflate_rsx_b200", "libbdeflate.so"), kernel, src, "22"],
                     capture_output=True, text=True)
out += ["", "## Warp instructions, stall samples and active lanes by region and by source line", "", "```",
        (hot.stdout or hot.stderr).rstrip(), "```"]
OUT = os.environ.get("PROFILE_OUT", os.path.join(ROOT, "profiles"))
os.makedirs(OUT, exist_ok=True)
open(os.path.join(OUT, f"{tag}_summary.md"), "w").write("\n".join(out) + "\n")
tp = os.path.join(OUT, "traffic.json")
try:
    tj = json.load(open(tp))
except Exception:
    tj = {}
tj[key] = {"kernel": row[hdr.index("Kernel Name")], "streams_in_capture": nstreams, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_stream": (rd + wr) / nstreams, "algorithmic_bytes_per_stream": alg, "report": os.path.basename(rep)}
json.dump(tj, open(tp, "w"), indent=1)
print("\n".join(out[:36]))
