"""Sum the warp-stall sample columns of an ncu source-page CSV (see ncu_lines.py for the export command), overall and
for the top lines: python tools/ncu_stalls.py f.csv [n_lines]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hdr = None
fname = ""
tot = {}
per = {}
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or not r or not r[0].isdigit():
        continue
    for i, h in cols:
        if i < len(r) and r[i].isdigit():
            v = int(r[i])
            tot[h] = tot.get(h, 0) + v
            d = per.setdefault((fname, int(r[0])), {})
            d[h] = d.get(h, 0) + v
s = sum(tot.values()) or 1
print("stall samples by reason:")
for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"  {100*v/s:5.1f}%  {h}")
print("top lines:")
for k, d in sorted(per.items(), key=lambda kv: -sum(kv[1].values()))[:nl]:
    t = sum(d.values())
    top = ", ".join(f"{h[6:]} {100*v/t:.0f}%" for h, v in sorted(d.items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"  {100*t/s:5.1f}%  {k[0]}:{k[1]}  {top}")
