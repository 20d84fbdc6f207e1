#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launch count,
total and mean device time, share of the listed launches.  Usage:
    python scripts/summarize_launches.py gpurun_out/launches.csv [skip_first_n] > profiles/<name>.md
Times under ncu are cold-cache and serialised: compare SHARES, not absolutes (B200_PROFILING.md)."""
import collections
import csv
import re
import sys

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr, data = rows[0], rows[1:][skip:]
ki, vi, gi, bi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
agg = collections.OrderedDict()
for r in data:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").strip()
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print(f"launches listed: {len(data)} (skipped first {skip}); total device time {tot / 1e3:.1f} us\n")
print("| kernel | launches | total us | mean us | share |")
print("|---|---:|---:|---:|---:|")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{name}` | {n} | {t / 1e3:.1f} | {t / n / 1e3:.2f} | {100 * t / tot:.1f}% |")
