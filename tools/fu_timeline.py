"""Summarise a fused-kernel timeline dump (FLACB200_FU_TIMELINE=<file> python tools/prof_encode.py 5 256):
median cycles a CTA spends in each phase, plus how the autocorrelation warps spread over SM sub-partitions."""
import sys

import numpy as np

a = np.fromfile(sys.argv[1], np.uint64).reshape(-1, 16).astype(np.int64)
a = a[a[:, 10] > 0]
names = ["stage (TMA)", "OR/AND", "need list", "queue A (fixed)", "queue B (LPC)", "select", "zero image", "pack", "CRC", "store"]
tot = a[:, 10] - a[:, 0]
print(f"{len(a)} CTAs; CTA lifetime median {np.median(tot):.0f} cycles (p10 {np.percentile(tot, 10):.0f}, p90 {np.percentile(tot, 90):.0f})")
for k, n in enumerate(names):
    d = a[:, k + 1] - a[:, k]
    print(f"  {n:28s} median {np.median(d):8.0f}  p90 {np.percentile(d, 90):8.0f}  ({100 * np.median(d) / np.median(tot):4.1f}%)")
span = a[:, 10].max() - a[:, 0].min()
print(f"  worker kernel span {span} cycles")
