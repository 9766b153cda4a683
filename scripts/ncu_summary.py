"""Summarise ncu --set full reports of the cross-attention kernels (read here, no GPU needed):
python scripts/ncu_summary.py r02 gpurun_out/r02_xattn_B2.ncu-rep gpurun_out/r02_xattn_B16.ncu-rep
writes profiles/<tag>_xattn_ncu_full_summary.txt and profiles/<tag>_xattn_traffic.json."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


lines = ["ncu --set full --clock-control none; scripts/profile_xattn.py B biased (N=4096,H=8,D=40,T=77), rotating buffers"]
traffic = {}
TAG = sys.argv[1]
for rep in sys.argv[2:]:
    tag = re.search(r"_B(\d+)", rep).group(1)
    hdr, units, rows = rows_of(rep)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows:
        name = r[ix["Kernel Name"]]
        kind = "stats" if "stats" in name else ("fused" if "fused" in name else "fwd")
        lines.append(f"--- B={tag} {kind}: {name[:110]}")
        vals = {}
        for m in METRICS:
            if m in ix:
                vals[m] = (r[ix[m]], units[ix[m]])
                lines.append(f"{m} = {r[ix[m]]} {units[ix[m]]}")
        rd = float(vals["dram__bytes_read.sum"][0].replace(",", "")) * UNIT.get(vals["dram__bytes_read.sum"][1], 1.0)
        wr = float(vals["dram__bytes_write.sum"][0].replace(",", "")) * UNIT.get(vals["dram__bytes_write.sum"][1], 1.0)
        dur = float(vals["gpu__time_duration.sum"][0].replace(",", ""))
        if vals["gpu__time_duration.sum"][1] in ("ns", "nsecond"):
            dur /= 1e3
        traffic[f"B{tag}_{kind}"] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr,
                                     "duration_us_under_ncu": dur}
open(os.path.join(ROOT, "profiles", f"{TAG}_xattn_ncu_full_summary.txt"), "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(os.path.join(ROOT, "profiles", f"{TAG}_xattn_traffic.json"), "w"), indent=1)
print("\n".join(lines))
