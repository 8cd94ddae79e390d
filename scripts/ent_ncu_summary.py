#!/usr/bin/env python
"""Turns an `ncu --metrics ... --csv` log of scripts/files_run.py into a one-table summary of the entropy-path kernels of one
full group (profiles/r02_entropy_ncu_summary.csv).  Usage: python scripts/ent_ncu_summary.py <log.csv> <out.csv>"""
import collections
import csv
import sys

log, out = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(log)) if len(r) > 10]
hdr = rows[0]
ik, iv, ig, im, iu, ib = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Grid Size", "Metric Name", "Metric Unit", "Block Size"))
launches = collections.OrderedDict()
for r in rows[1:]:
    e = launches.setdefault(int(r[0]), {"name": r[ik], "grid": r[ig], "block": r[ib], "m": collections.OrderedDict()})
    e["m"][r[im]] = (r[iv], r[iu])
ids = list(launches)
# the largest group captured (grid.y of k0_zero_headers = its images)
def group_images(e):
    return int(e["grid"].strip("()").split(",")[1])


start = max((i for i in ids if "zero_headers" in launches[i]["name"]), key=lambda i: group_images(launches[i]))
print("group of", group_images(launches[start]), "images")
group = []
for i in ids[ids.index(start):]:
    if group and "zero_headers" in launches[i]["name"]:
        break
    group.append(launches[i])
# the idle synchronisation launches are all alike: keep the first
kept, seen_idle = [], False
for e in group:
    t = float(e["m"]["gpu__time_duration.sum"][0].replace(",", ""))
    if "ent_sync" in e["name"] and t < 10000:
        if seen_idle:
            continue
        seen_idle = True
    kept.append(e)
metrics = list(kept[0]["m"])
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + ["%s grid %s block %s" % (e["name"].split("(")[0].replace("void ", "").replace("b200jpg::<unnamed>::", ""), e["grid"], e["block"]) for e in kept])
    for m in metrics:
        w.writerow([m, kept[0]["m"][m][1]] + [e["m"].get(m, ("", ""))[0] for e in kept])
print("wrote", out, len(kept), "kernels")
