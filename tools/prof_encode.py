"""Small encode-only workload for ncu captures (64 of the bench streams, level from argv): python tools/prof_encode.py [level]"""
import sys

sys.path.insert(0, ".")
import bench
from pyflac_b200 import _native as nat

level = int(sys.argv[1]) if len(sys.argv) > 1 else 5
pcm = bench.make_pcm(0, 64)
eng = nat.Engine(0)
for _ in range(2):
    blobs, out = nat.encode_streams(eng, [pcm[s] for s in range(64)], 48000, 16, level, 4096)
print("ok", out["total_bytes"], eng.launch_count)
