"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel.
usage: python profiles/summarize_launches.py profiles/r01/launches_cfg4_v1.csv"""
import collections
import csv
import re
import sys

lines = [ln for ln in open(sys.argv[1]) if not ln.startswith("==")]
agg = collections.OrderedDict()
for r in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", r["Kernel Name"])[:64]
    v = float(r["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r["Metric Unit"], v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':66s} {'n':>5s} {'total us':>10s} {'avg us':>9s} {'share':>6s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:66s} {n:5d} {t:10.1f} {t / n:9.1f} {100 * t / tot:5.1f}%")
print(f"{'total':66s} {'':5s} {tot:10.1f}")
