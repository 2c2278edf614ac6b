"""Device-resident throughput of the other BASELINE configs (parity-test shapes, not bench lines):
   C3 = 24-bit mono 192 kHz level 8 (scaled stream count), C4 = 4096-stream decode.  python tools/bench_configs.py [n_streams_c3] [n_streams_c4]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from pyflac_b200 import _native as nat
from pyflac_b200.synth import music_like

n3 = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n4 = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
eng = nat.Engine(0)
eng.set_profiling(True)
dev = torch.device("cuda", 0)


def run_encode(name, base, n_streams, sr, bps, level, bs, ch):
    n = base[0].shape[0]
    dt = np.int16 if bps <= 16 else np.int32
    pcm = np.empty((n_streams, n, ch), dt)
    for s in range(n_streams):
        pcm[s] = np.roll(base[s % len(base)], 977 * (s // len(base)), axis=0)
    d = torch.from_numpy(pcm.reshape(-1)).to(dev)
    off = np.arange(n_streams, dtype=np.uint64) * np.uint64(n * ch)
    cnt = np.full(n_streams, n, np.uint64)
    cfg = nat.Engine.make_config(sr, ch, bps, level, bs, container_bytes=dt().itemsize)
    for _ in range(2):
        eng.encode_device(cfg, d.data_ptr(), d.numel(), off, cnt)
    torch.cuda.synchronize()
    eng.join(); eng.sync()
    t0 = time.perf_counter()
    steps = 3
    for _ in range(steps):
        eng.encode_device(cfg, d.data_ptr(), d.numel(), off, cnt)
    eng.join(); eng.sync()
    dt_s = (time.perf_counter() - t0) / steps
    res = eng.result()
    print(name, f"{pcm.size / dt_s / 1e6:.0f} MSamples/s", f"{dt_s * 1e3:.2f} ms/step", "ratio %.3f" % (res.total_bytes / pcm.nbytes * (dt().itemsize * 8 / bps)),
          {k: (round(v, 3) if isinstance(v, float) else v) for k, v in eng.kernel_times().items()}, "guard", res.log_guard_hits)
    return pcm


base3 = [music_like(262144, 1, 192000, 24, seed=300 + s) for s in range(16)]
run_encode("C3 24-bit mono 192k L8 x%d" % n3, base3, n3, 192000, 24, 8, 4096, 1)
base2 = [music_like(131072, 2, 48000, 16, seed=400 + s) for s in range(32)]
pcm = run_encode("C5-shape s16 stereo L5 x%d (131072 samples)" % n4, base2, n4, 48000, 16, 5, 4096, 2)
# C4: decode what was just encoded
res = eng.result()
out = eng.fetch()
blob = np.ascontiguousarray(out["arena"][:int(res.total_bytes) + 16])
so = np.array([si.byte_off for si in out["streams"]], np.uint64)
sl = np.array([si.byte_len for si in out["streams"]], np.uint64)
d_blob = torch.from_numpy(blob).to(dev)
for _ in range(2):
    eng.decode_device(d_blob.data_ptr(), int(res.total_bytes), so, sl, 2)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    eng.decode_device(d_blob.data_ptr(), int(res.total_bytes), so, sl, 2)
torch.cuda.synchronize()
dt_s = (time.perf_counter() - t0) / 3
print("C4 decode x%d" % n4, f"{pcm.size / dt_s / 1e6:.0f} MSamples/s", f"{dt_s * 1e3:.2f} ms/step", eng.decode_kernel_times())
