#!/usr/bin/env python
"""Condense `ncu --page raw --csv` exports (tools/capture_profiles.sh) into one small JSON:
per kernel, the metrics DESIGN.md and bench.py's `roofline.traffic` quote.
    python profiles/ncu_summary.py gpurun_out/raw_*_k3.csv > profiles/ncu_full_r1_summary.json"""
import csv
import json
import re
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed.sum",
]


def main(paths):
    out = {}
    for path in paths:
        rows = list(csv.reader(open(path)))
        if len(rows) < 3:
            continue
        hdr, units, data = rows[0], rows[1], rows[2]
        name = data[hdr.index("Kernel Name")]
        op = re.search(r"::(\w+Op)<", name).group(1)
        k = re.search(r"_k(\d+)\.csv$", path)
        d = {"kernel": name, "k": int(k.group(1)) if k else None}
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                d[key] = f"{data[i]} {units[i]}".strip()
        out[op if not k or k.group(1) == "3" else f"{op}_k{k.group(1)}"] = d
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1:])
