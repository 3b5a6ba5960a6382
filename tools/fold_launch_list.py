#!/usr/bin/env python3
"""Fold an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (total, launches, ms per launch, share):
    python tools/fold_launch_list.py gpurun_out/launches.csv "header comment" > profiles/rNN_launch_list.csv"""
import collections, csv, re, sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")) if r]
hdr = rows[0]
ki, mi, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = re.sub(r"\(.*", "", r[ki])
    tot[name] += float(r[vi].replace(",", "")) * scale
    cnt[name] += 1
total = sum(tot.values())
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print("total_ms,launches,ms_per_launch,share,kernel")
for k, v in tot.most_common():
    print(f"{v:.3f},{cnt[k]},{v / cnt[k]:.3f},{v / total:.3f},{k}")
