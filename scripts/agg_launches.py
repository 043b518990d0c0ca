"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel: python scripts/agg_launches.py file.csv [...]"""
import collections
import csv
import re
import sys

for f in sys.argv[1:]:
    with open(f) as fh:
        lines = [l for l in fh if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rd:
        name = re.sub(r"\(.*", "", r[ki])
        m = re.search(r"update_mlp_tc_kernel<\(?(?:int\))?(\d)", r[ki])
        if m:
            name = "update_mlp_tc_kernel<%s>" % m.group(1)
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0, r[gi]])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f, "total us", round(tot, 1))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
        print(f"  {k[:60]:60s} n={a[0]:4d} total={a[1]:10.1f}us avg={a[1] / a[0]:9.1f} share={100 * a[1] / tot:5.1f}% grid={a[2]}")
