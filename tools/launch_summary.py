"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
    python tools/launch_summary.py gpurun_out/launches.csv [skip_first_n] > profiles/rNN_..._summary.txt"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10]
hdr = rows[0]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size") if "Grid Size" in hdr else None
ui = hdr.index("Metric Unit")
tot = defaultdict(lambda: [0.0, 0])
n = 0
for r in rows[1:]:
    if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    n += 1
    if n <= skip:
        continue
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1e3)
    tot[name][0] += v
    tot[name][1] += 1
total = sum(t for t, _ in tot.values())
print(f"{path}: {n - skip} launches, sum {total / 1e3:.2f} ms (ncu per-launch times: cold caches, serialised)")
for name, (t, c) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"  {t / 1e3:8.3f} ms {100 * t / total:5.1f}% n={c:4d} avg={t / c:8.1f} us  {name[:110]}")
