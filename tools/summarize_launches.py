"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: the last analyze_cu pass (default) or,
with a second argument `all`, every launch of the file."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
L = [(r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:36], float(r[vi]) / 1e3) for r in rows[1:]]
finals = [i for i, (k, t) in enumerate(L) if k.startswith("k_cu_final")]
seg = L[finals[-2] + 1:finals[-1] + 1] if len(finals) >= 2 and len(sys.argv) < 3 else L
agg = collections.OrderedDict()
for k, t in seg:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(v[1] for v in agg.values())
for k, v in agg.items():
    print(f"{k:38s} n={v[0]:3d} {v[1]:9.1f} us {100 * v[1] / tot:5.1f}%")
print(f"total {tot:.1f} us")
