"""Small decode-only workload for ncu captures (128 of the bench streams): python tools/prof_decode.py"""
import sys

import numpy as np

sys.path.insert(0, ".")
import bench
from pyflac_b200 import _native as nat

pcm = bench.make_pcm(0, 128)
eng = nat.Engine(0)
blobs, _ = nat.encode_streams(eng, [pcm[s] for s in range(128)], 48000, 16, 5, 4096)
for _ in range(2):
    out, infos = nat.decode_streams(eng, blobs)
print("ok", all(i.status == 0 for i in infos), eng.decode_kernel_times())
