#!/usr/bin/env python
"""Reads `ncu --set full` reports of scripts/sweep_kernels.py runs and writes profiles/traffic.json (DRAM bytes per IMAGE
and kernel, which bench.py scales to its batch for `roofline.traffic`) and a one-table summary of the counters the
design discussion quotes (profiles/r02_ncu_full_summary.csv).
Usage: python scripts/ncu_traffic.py <config>:<batch>:<report.ncu-rep> [...]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [("k1_idct", "k1_dequant_idct8x8"), ("k2_", "k2_upsample_color"), ("kf_fused", "kf_fused")]
METRICS = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum",
           "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
           "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
UNIT = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}

traffic = {}
columns = []   # (title, {metric: (value, unit)})
for arg in sys.argv[1:]:
    cfg, batch, rep = arg.split(":", 2)
    batch = int(batch)
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"]
        key = next((k for sub, k in KEYS if sub in name), None)
        if not key:
            continue

        def val(field):
            return float(d[field]) * UNIT.get(units[hdr.index(field)].lower(), 1.0)
        bytes_ = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        traffic.setdefault(cfg, {}).setdefault(key, []).append(bytes_ / batch)
        columns.append(("%s %s x%d" % (cfg, name.split("(")[0].replace("void b200jpg::", ""), batch),
                        {m: (d.get(m, ""), units[hdr.index(m)] if m in hdr else "") for m in METRICS}))
out = {cfg: {k: sum(v) / len(v) for k, v in ks.items()} for cfg, ks in traffic.items()}
out["unit"] = "DRAM bytes (read + write) per image, ncu --set full; bench.py multiplies by its batch"
path = os.path.join(ROOT, "profiles", "traffic.json")
json.dump(out, open(path, "w"), indent=1, sort_keys=True)
with open(os.path.join(ROOT, "profiles", "r02_ncu_full_summary.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [c[0] for c in columns])
    for m in METRICS:
        w.writerow([m, next((c[1][m][1] for c in columns if c[1][m][1]), "")] + [c[1][m][0] for c in columns])
print(json.dumps(out))
