"""Aggregate an ncu gpu__time_duration launch list by kernel (and grid) -> markdown."""
import csv
import re
import sys
from collections import OrderedDict

src = sys.argv[1]
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.reader(lines))
hdr = rows[0]
ki, gi, vi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
agg = OrderedDict()
order = []
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
    v = float(r[vi].replace(",", ""))
    if r[ui] in ("nsecond", "ns"):
        v /= 1e3
    elif r[ui] in ("msecond", "ms"):
        v *= 1e3
    order.append((name, r[gi], v))
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"total {len(order)} launches, {tot/1e3:.3f} ms of kernel time (serialised, cold)")
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {k[:70]} | {n} | {t:.1f} | {100*t/tot:.1f}% |")
if len(sys.argv) > 2:
    print("\nfirst launches in order:")
    for name, g, v in order[: int(sys.argv[2])]:
        print(f"  {v:8.1f} us  grid {g:>14}  {name[:60]}")
