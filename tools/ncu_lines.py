"""Summarise an ncu source page (cuda,sass) export: % of executed warp instructions and stall samples per CUDA line.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:K > f.csv ; python tools/ncu_lines.py f.csv [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
fname = ""
agg = {}
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed")
        si = hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ii:
        continue
    if r[0].isdigit() and r[ii].isdigit():
        key = (fname, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1].strip()[:100]])
        a[0] += int(r[ii])
        a[1] += int(r[si]) if r[si].isdigit() else 0
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
print("total warp-instructions", tot, "samples", tots)
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if 100 * a[0] / tot >= minpct or 100 * a[1] / tots >= minpct:
        print(f"{100*a[0]/tot:5.1f}% inst {100*a[1]/tots:5.1f}% smp  {f}:{l}  {a[2]}")
