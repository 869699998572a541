"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (name, launches, total, mean)."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"^(void )?povar::<unnamed>::", "", re.sub(r"\(.*", "", r[ki]))
    t = float(r[vi].replace(",", ""))
    if r[ui] == "ns":
        t /= 1000.0
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += t
total = sum(v[1] for v in agg.values())
print(f"{sum(v[0] for v in agg.values())} launches, {total / 1000.0:.2f} ms of kernel time")
print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {name[:90]} | {n} | {t:.1f} | {t / n:.1f} | {100.0 * t / total:.1f} % |")
