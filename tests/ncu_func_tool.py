"""Warp instructions / stall samples of an ncu source page aggregated by function (line ranges given on the command line):
ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv
python ncu_func_tool.py f.csv file:lo-hi=name ...   (unmatched lines are grouped by file)"""
import csv, os, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
ranges = []
for a in sys.argv[2:]:
    loc, name = a.split("=")
    f, r = loc.split(":")
    lo, hi = r.split("-")
    ranges.append((f, int(lo), int(hi), name))
fname, col = "?", None
b = defaultdict(lambda: [0, 0, 0])
for r in rows:
    if not r: continue
    if r[0] in ("File Path", "File Name"): fname = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        col = {}
        for i, n in enumerate(r): col.setdefault(n, i)
        continue
    if col is None or r[0] == "" or "Instructions Executed" not in col: continue
    try:
        ln = int(r[0]); ie = int(float(r[col["Instructions Executed"]] or 0)); te = int(float(r[col["Thread Instructions Executed"]] or 0))
        st = int(float(r[col["Warp Stall Sampling (All Samples)"]] or 0))
    except (ValueError, IndexError):
        continue
    key = fname
    for f, lo, hi, name in ranges:
        if f == fname and lo <= ln <= hi: key = name; break
    b[key][0] += ie; b[key][1] += te; b[key][2] += st
tot = sum(v[0] for v in b.values()); tst = sum(v[2] for v in b.values())
print(f"total warp inst {tot}  stall samples {tst}")
for k, v in sorted(b.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:44s} inst {100 * v[0] / max(tot, 1):5.1f}%  lanes {v[1] / max(v[0], 1):5.1f}  stall {100 * v[2] / max(tst, 1):5.1f}%")
