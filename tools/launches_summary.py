#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (`--metrics gpu__time_duration.sum --csv`): launches, total, share, mean.

    python tools/launches_summary.py profiles/r02_launches.csv
"""
import collections
import csv
import sys

path = sys.argv[1]
with open(path) as fh:
    lines = [l for l in fh if not l.startswith("==")]
tot, cnt = collections.Counter(), collections.Counter()
for row in csv.DictReader(lines):
    if row["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    us = v / 1e3 if row["Metric Unit"] in ("ns", "nsecond") else (v * 1e3 if row["Metric Unit"] in ("ms", "msecond") else v)
    tot[row["Kernel Name"]] += us
    cnt[row["Kernel Name"]] += 1
allus = sum(tot.values())
print(f"{sum(cnt.values())} launches, {allus:.1f} us in total\n")
print("launches   total us  share  mean us  kernel")
for k, us in tot.most_common(14):
    print(f"{cnt[k]:8d} {us:10.1f} {100 * us / allus:5.1f}% {us / cnt[k]:8.2f}  {k[:110]}")
print("\nkernels of this library in the window:")
for k, us in tot.most_common():
    if "ds::" in k or "loss_kernel" in k or "vote_kernel" in k:
        print(f"{cnt[k]:8d} launches, mean {us / cnt[k]:8.2f} us  {k[:110]}")
