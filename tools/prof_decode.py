"""Decode-only workload for ncu captures: python tools/prof_decode.py [n_streams=128] [samples_per_stream=480000]
(128 of the bench streams by default; `4096 131072` is the BASELINE configs[3] shape of bench.py's decode leg)."""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench
from pyflac_b200 import _native as nat

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 128
n = int(sys.argv[2]) if len(sys.argv) > 2 else bench.N_SAMPLES
pcm = bench.make_pcm(0, ns) if n == bench.N_SAMPLES else bench.make_pcm_short(0, ns, n)
eng = nat.Engine(0)
blobs, _ = nat.encode_streams(eng, [pcm[s] for s in range(ns)], 48000, 16, 5, 4096)
for _ in range(2):
    out, infos = nat.decode_streams(eng, blobs)
print("ok", all(i.status == 0 for i in infos), eng.decode_kernel_times())
