"""Aggregate the warp-stall samples of an .ncu-rep by source line: python tools/ncu_lines.py rep [topN] [kernel regex]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 3:
    cmd += ["--kernel-name", "regex:" + sys.argv[3]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
h = rows[hi]
stall = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
si, ie = h.index("# Samples"), h.index("Instructions Executed")
tot, lines, cur = 0, [], None
for r in rows[hi + 1:]:
    if len(r) < len(h) or r[0] == "":
        continue
    try:
        s = int(r[si])
    except ValueError:
        continue
    tot += s
    st = {h[i][6:]: int(r[i]) for i in stall if r[i] not in ("", "0")}
    lines.append((s, r[0], r[1].strip()[:90], int(r[ie] or 0), st))
print("total samples", tot)
for s, ln, src, n, st in sorted(lines, key=lambda x: -x[0])[:top]:
    t3 = sorted(st.items(), key=lambda x: -x[1])[:3]
    print(f"{s:6d} {s / tot * 100:5.1f}% L{ln:>4} inst={n:9d} {src} {t3}")
