"""Summarise an ncu launch list (gpu__time_duration) by kernel: python tools/launch_summary.py file.csv [skip_launches]"""
import csv, collections, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else -1
if skip < 0:  # default: the last training step = from the last prologue_kernel launch on
    skip = max([int(r[rows[0].index("ID")]) for r in rows[1:] if "prologue_kernel" in r[rows[0].index("Kernel Name")]] or [0])
h = rows[0]; ki, vi, ii = h.index("Kernel Name"), h.index("Metric Value"), h.index("ID")
d = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if int(r[ii]) < skip:
        continue
    n = r[ki].replace("void ", "").replace("<unnamed>::", "")[:70]
    d[n][0] += 1; d[n][1] += float(r[vi])
tot = sum(v[1] for v in d.values())
for n, v in sorted(d.items(), key=lambda x: -x[1][1])[:28]:
    print(f"{v[1] / 1e6:9.3f} ms {v[0]:5d} {v[1] / tot * 100:5.1f}%  {n}")
print(f"total {tot / 1e6:.3f} ms")
