"""Device-resident encode step time against the number of steps per run (is there a per-run tail?): python tools/steps_probe.py"""
import sys
import time

sys.path.insert(0, ".")
import numpy as np
import torch
import bench
from pyflac_b200 import _native as nat

n = 256
pcm = bench.make_pcm(0, n)
d_pcm = torch.from_numpy(pcm.reshape(-1)).cuda()
E = bench.N_SAMPLES * bench.CHANNELS
off = np.arange(n, dtype=np.uint64) * np.uint64(E)
smp = np.full(n, bench.N_SAMPLES, np.uint64)
eng = nat.Engine(0)
for md5 in (True, False):
    cfg = nat.Engine.make_config(48000, 2, 16, 5, 4096, container_bytes=2, do_md5=md5)
    for steps in (5, 10, 20, 40, 80):
        for _ in range(3):
            eng.encode_device(cfg, d_pcm.data_ptr(), d_pcm.numel(), off, smp)
        eng.join(); eng.sync(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            eng.encode_device(cfg, d_pcm.data_ptr(), d_pcm.numel(), off, smp)
        eng.join(); eng.sync(); torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        print(f"md5={md5} steps={steps}: {dt:.2f} ms total, {dt / steps:.3f} ms/step")
