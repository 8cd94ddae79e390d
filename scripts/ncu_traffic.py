#!/usr/bin/env python
"""Reads an `ncu --set full` report of `bench.py` and writes profiles/traffic.json:
{"<config>:<batch>": {"k1_dequant_idct8x8": dram bytes per launch, "k2_upsample_color": ...}}.
Usage: python scripts/ncu_traffic.py <report.ncu-rep> <config> <batch>"""
import csv
import io
import json
import os
import subprocess
import sys

rep, cfg, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
out = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"]
    key = "k1_dequant_idct8x8" if "k1_" in name else ("k2_upsample_color" if "k2_" in name else None)
    if not key:
        continue

    def gb(field):
        v = float(d[field])
        unit = rows[1][hdr.index(field)].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)
    out.setdefault(key, []).append(gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"))
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
try:
    allv = json.load(open(path))
except Exception:
    allv = {}
allv["%s:%d" % (cfg, batch)] = {k: sum(v) / len(v) for k, v in out.items()}
json.dump(allv, open(path, "w"), indent=1, sort_keys=True)
print(json.dumps(allv))
