"""Top source lines of an ncu source page:
ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv; python ncu_source_tool.py f.csv [N]"""
import csv, os, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
data, fname, col = [], "?", None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        col = {}
        for i, n in enumerate(r): col.setdefault(n, i)
        continue
    if col is None or r[0] == "" or "Instructions Executed" not in col or len(r) <= col["Instructions Executed"]: continue
    try:
        data.append((fname, int(r[0]), r[1], int(float(r[col["Instructions Executed"]] or 0)),
                     int(float(r[col["Warp Stall Sampling (All Samples)"]] or 0)),
                     int(float(r[col["Thread Instructions Executed"]] or 0))))
    except ValueError:
        pass
ti = sum(d[3] for d in data); ts = sum(d[4] for d in data)
print(f"total warp-inst {ti}  samples {ts}")
def show(d):
    print(f"{d[0][:16]:16s}{d[1]:5d} {100*d[3]/ti:5.1f}% inst {100*d[4]/max(ts,1):5.1f}% stall lanes {d[5]/max(d[3],1):4.1f} | {d[2].strip()[:100]}")
print("--- by instructions")
for d in sorted(data, key=lambda d: -d[3])[:topn]: show(d)
print("--- by stall samples")
for d in sorted(data, key=lambda d: -d[4])[:topn // 2]: show(d)
