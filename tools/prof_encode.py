"""Encode-only workload for ncu captures (n of the bench streams): python tools/prof_encode.py [level] [n_streams]"""
import sys

sys.path.insert(0, ".")
import bench
from pyflac_b200 import _native as nat

level = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
pcm = bench.make_pcm(0, n)
eng = nat.Engine(0)
eng.set_profiling(True)
for _ in range(2):
    blobs, out = nat.encode_streams(eng, [pcm[s] for s in range(n)], 48000, 16, level, 4096)
print("ok", out["total_bytes"], eng.launch_count, {k: (round(v, 4) if isinstance(v, float) else v) for k, v in eng.kernel_times().items()})
