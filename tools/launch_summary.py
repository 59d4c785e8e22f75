"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count / total / share.
    python tools/launch_summary.py gpurun_out/x.csv [first_fraction_to_skip]"""
import csv, re, sys
rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v *= {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3}.get(u, 1.0)
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("icon::", "").replace("void ", "")
    rows.append((name, v))
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
rows = rows[int(len(rows) * skip):] if skip < 1 else rows[-int(skip):]
agg = {}
for n, v in rows:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v for _, v in rows)
print(f"{len(rows)} launches, {tot / 1e3:.3f} ms total")
for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{v / 1e3:9.3f} ms {100 * v / tot:5.1f}%  x{c:4d}  {v / c:9.1f} us  {n[:110]}")
