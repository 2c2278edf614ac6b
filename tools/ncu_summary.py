"""Turn ncu reports / launch lists into the small text summaries committed under profiles/.
usage: python tools/ncu_summary.py <report.ncu-rep> > profiles/<name>.txt
       python tools/ncu_summary.py --launches <launches.csv> > profiles/<name>.txt"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"# {path}: ncu --set full --clock-control none, one capture per kernel (cold cache, serialised)")
    for r in rows[2:]:
        print(f"\n## {r[ki][:100]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:75s} {r[i]:>18s} {units[i]}")


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ni, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        name = r[ni].split("(")[0].replace("void fb::", "").replace("void ", "")
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        if v != v:                              # "nan": the launch the capture was cut at
            continue
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: ncu --metrics gpu__time_duration.sum --clock-control none (per-launch, cold-cache, serialised): compare SHARES, not absolutes")
    print(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n[:60]:60s} {a[0]:8d} {a[1] / 1e6:10.3f} {100 * a[1] / tot:6.1f}%")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[1])
