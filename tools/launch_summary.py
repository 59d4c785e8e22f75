"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count / total / share.

    python tools/launch_summary.py gpurun_out/x.csv                 # every launch
    python tools/launch_summary.py gpurun_out/x.csv 0.5             # skip the first half of the launches
    python tools/launch_summary.py gpurun_out/x.csv icon:658        # the tail that starts 658 icon launches from the end
                                                                     # (= the last filter() of tools/filter_once.py)"""
import csv
import re
import sys


def load(path):
    rows = []
    lines = [l for l in open(path) if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3}.get(r["Metric Unit"], 1.0)
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("icon::", "").replace("void ", "")
        rows.append((name, v, r.get("Grid Size", "")))
    return rows


def main():
    rows = load(sys.argv[1])
    sel = sys.argv[2] if len(sys.argv) > 2 else "0"
    if sel.startswith("icon:"):
        icon = [i for i, (n, _, _) in enumerate(rows) if n.startswith("k_")]
        rows = rows[icon[-int(sel[5:])]:]
    else:
        rows = rows[int(len(rows) * float(sel)):]
    agg = {}
    for n, v, _ in rows:
        a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v for _, v, _ in rows)
    print(f"{len(rows)} launches, {tot / 1e3:.3f} ms total (serialised under ncu)")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
        print(f"{v / 1e3:9.3f} ms {100 * v / tot:5.1f}%  x{c:4d}  {v / c:9.1f} us  {n[:100]}")
    print("largest single launches:")
    for n, v, g in sorted(rows, key=lambda t: -t[1])[:12]:
        print(f"{v:9.1f} us  {n[:70]:70s} grid {g}")


if __name__ == "__main__":
    main()
