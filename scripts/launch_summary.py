"""python scripts/launch_summary.py <ncu launch csv> "<title>"  ->  markdown table: kernel, launches, total us, share."""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    n, t = agg.get(r[ki], (0, 0.0))
    agg[r[ki]] = (n + 1, t + v)
tot = sum(t for _, t in agg.values())
print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}\n")
print("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare shares)\n")
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k[:100]}` | {n} | {t:.1f} | {100 * t / tot:.2f}% |")
