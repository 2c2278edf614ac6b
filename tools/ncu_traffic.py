"""profiles/r02_traffic.json from ncu --set full reports: dram__bytes_read.sum / dram__bytes_write.sum per kernel (one launch on the
bench batch); bench.py reads it for roofline.traffic.
usage: python tools/ncu_traffic.py <report.ncu-rep> [more reports...] > profiles/r02_traffic.json"""
import csv
import io
import json
import subprocess
import sys


def unit_scale(u):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


out = {"source": "ncu --set full --clock-control none (one launch per kernel, cold cache) of tools/prof_encode.py 5 256 / tools/prof_decode.py 4096 131072: " + ", ".join(sys.argv[1:]),
       "kernels": {}}
for path in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ki, ri, wi, ti = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("fb::", "").split("<")[0]
        out["kernels"][name] = {"dram_read_bytes": int(float(r[ri].replace(",", "")) * unit_scale(units[ri])),
                                "dram_write_bytes": int(float(r[wi].replace(",", "")) * unit_scale(units[wi])),
                                "ncu_duration_ms": float(r[ti].replace(",", "")) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "msecond": 1, "usecond": 1e-3, "nsecond": 1e-6}.get(units[ti], 1), "report": path.split("/")[-1]}
print(json.dumps(out, indent=1))
