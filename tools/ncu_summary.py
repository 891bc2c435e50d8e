#!/usr/bin/env python
"""tools/ncu_summary.py -- turn an `ncu --set full` report into the tracked summaries under profiles/.

    python tools/ncu_summary.py gpurun_out/prof_r01.ncu-rep profiles/r01_ncu_full --sites 67108864

Writes <out>.md (key metrics, pipe utilisation, stall reasons per launch) and updates
profiles/roofline_traffic.json (DRAM bytes per launch of the dominant kernel, read by bench.py)."""
from __future__ import annotations

import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    sites = int(sys.argv[sys.argv.index("--sites") + 1]) if "--sites" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu --set full summary: {os.path.basename(rep)}", ""]
    traffic = []
    for n, r in enumerate(data):
        name = r[col["Kernel Name"]].split("(")[0]
        lines += [f"## launch {n}: `{name}`", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in col:
                lines.append(f"| {k} | {r[col[k]]} | {units[col[k]]} |")
        rd = float(r[col["dram__bytes_read.sum"]]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[col["dram__bytes_read.sum"]]]
        wr = float(r[col["dram__bytes_write.sum"]]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units[col["dram__bytes_write.sum"]]]
        dur_us = float(r[col["gpu__time_duration.sum"]]) * {"us": 1, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(
            units[col["gpu__time_duration.sum"]], 1)
        traffic.append(rd + wr)
        lines.append(f"| DRAM traffic (read+write) | {(rd + wr) / 1e9:.4f} | GB |")
        lines.append(f"| DRAM GB/s under ncu (cold, serialised) | {(rd + wr) / 1e3 / dur_us:.0f} | GB/s |")
        if sites:
            lines.append(f"| algorithmic bytes (193 B x {sites} sites) | {193 * sites / 1e9:.4f} | GB |")
            lines.append(f"| traffic / algorithmic | {(rd + wr) / (193 * sites):.4f} | |")
        stalls = sorted(((float(r[i]), h) for h, i in col.items()
                         if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")),
                        reverse=True)
        lines += ["", "warp stall reasons (warps per issue-active cycle):", ""]
        for v, h in stalls[:8]:
            lines.append(f"* {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]}: {v:.3f}")
        lines.append("")
    with open(out + ".md", "w") as f:
        f.write("\n".join(lines) + "\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if sites and traffic:
        with open(os.path.join(root, "profiles", "roofline_traffic.json"), "w") as f:
            json.dump({"dram_bytes_per_launch": sum(traffic) / len(traffic), "sites_per_launch": sites,
                       "dram_bytes_per_site": sum(traffic) / len(traffic) / sites,
                       "note": f"dram__bytes_read.sum + dram__bytes_write.sum, mean of {len(traffic)} launch(es), "
                               f"{os.path.basename(rep)}; writes still resident in L2 at kernel end are not counted",
                       "source": os.path.basename(out) + ".md"}, f, indent=1)
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
