"""ncu launch list (gpu__time_duration.sum CSV) -> markdown table per kernel.  usage: launch_summary.py list.csv > summary.md"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
agg = defaultdict(lambda: [0, 0.0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("pgmm::<unnamed>::", "")
    name = re.sub(r"^cub::(CUB_\w+::)?", "cub::", name)[:70]
    ms = float(r[iv].replace(",", "")) / 1e6
    a = agg[name]
    a[0] += 1
    a[1] += ms
    a[2] = max(a[2], ms)
tot = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms | share | max ms |\n|---|---:|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {a[0]} | {a[1]:.3f} | {100 * a[1] / tot:.1f}% | {a[2]:.3f} |")
print(f"\nTotal {tot:.1f} ms over {len(rows)} launches.")
