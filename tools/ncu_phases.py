"""Per-phase share of executed warp instructions and of stall samples for the fused encode kernel, from an ncu source-page
export: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:fused > f.csv ; python tools/ncu_phases.py f.csv [n_frames]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nfr = int(sys.argv[2]) if len(sys.argv) > 2 else 30208
fname, hdr, agg = "", None, {}
for r in rows:
    if len(r) >= 2 and r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ii:
        continue
    if r[0].isdigit() and r[ii].isdigit():
        a = agg.setdefault((fname, int(r[0])), [0, 0])
        a[0] += int(r[ii])
        a[1] += int(r[si]) if r[si].isdigit() else 0
tot = sum(a[0] for a in agg.values())
tots = sum(a[1] for a in agg.values())
src = open("pyflac_b200/csrc/enc_fused.cu").read().split("\n")


def find(pat):
    return next(i + 1 for i, l in enumerate(src) if pat in l)


def rng(f, lo, hi):
    return (sum(a[0] for (fn, l), a in agg.items() if fn == f and lo <= l <= hi),
            sum(a[1] for (fn, l), a in agg.items() if fn == f and lo <= l <= hi))


marks = [("autoc", "void fu_autoc_item", "------ fixed predictors"), ("fixed_sums", "void fu_fixed_sums", "------ LPC residual"),
         ("lpc residual", "void fu_lpc_psums_vec", "------ pack"), ("pack body", "------ pack", "------ the kernel"),
         ("kernel: stage+bits", "fused_encode_kernel(", "=================== queue A"), ("kernel: queue A (fixed task)", "=================== queue A", "=================== queue B"),
         ("kernel: queue B (lpc task)", "=================== queue B", "=================== selection"), ("kernel: select", "=================== selection", "=================== pack (row E12)"),
         ("kernel: pack headers", "=================== pack (row E12)", "=================== CRC-16 over"), ("kernel: crc+store", "=================== CRC-16 over", "------ host side")]
print("total warp-instructions %.3f G, %.0f per frame; stall samples %d" % (tot / 1e9, tot / nfr, tots))
for n, a, b in marks:
    i, s = rng("enc_fused.cu", find(a), find(b) - 1)
    print(f"{n:30s} instr {100 * i / tot:5.1f}% ({i / nfr:7.0f}/frame)  samples {100 * s / tots:5.1f}%")
for f in sorted(set(fn for fn, _ in agg)):
    if f != "enc_fused.cu":
        i, s = rng(f, 0, 10 ** 9)
        if i / tot > 0.002:
            print(f"{f:30s} instr {100 * i / tot:5.1f}% ({i / nfr:7.0f}/frame)  samples {100 * s / tots:5.1f}%")
