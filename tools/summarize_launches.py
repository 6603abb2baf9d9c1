"""Share of every kernel in an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import csv
import sys
from collections import defaultdict

tot = defaultdict(lambda: [0, 0.0])
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    name = r[ki].split("(")[0][-90:]
    tot[name][0] += 1
    tot[name][1] += float(r[vi].replace(",", ""))
all_ns = sum(v[1] for v in tot.values())
print(f"| kernel | launches | total ms | share |\n|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1])[:15]:
    print(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {100 * v[1] / all_ns:.1f} % |")
