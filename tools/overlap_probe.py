"""How much of the encode step is serialisation between kernels?  python tools/overlap_probe.py
(a) the bench loop with and without MD5; (b) the same work as two engines on two CUDA streams, each taking half of the streams per step:
their kernels overlap freely (the tail of one batch under the head of the other)."""
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
steps = 20


def loop(engines, parts, md5):
    cfg = nat.Engine.make_config(48000, 2, 16, 5, 4096, container_bytes=2, do_md5=md5)
    offs = []
    for (lo, hi) in parts:
        k = hi - lo
        offs.append((np.arange(k, dtype=np.uint64) * np.uint64(E), np.full(k, bench.N_SAMPLES, np.uint64), d_pcm.data_ptr() + lo * E * 2, k * E))
    def step():
        for e, (o, s, ptr, ne) in zip(engines, offs):
            e.encode_device(cfg, ptr, ne, o, s)
    for _ in range(5):
        step()
    for e in engines:
        e.join(); e.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    for e in engines:
        e.join(); e.sync()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


e1 = nat.Engine(0)
print("one engine, 256 streams per batch: md5 on %.3f ms/step, md5 off %.3f ms/step" % (loop([e1], [(0, n)], True), loop([e1], [(0, n)], False)))
e2 = nat.Engine(0)
print("two engines x 128 streams on two streams: md5 on %.3f ms/step, md5 off %.3f ms/step" % (loop([e1, e2], [(0, n // 2), (n // 2, n)], True), loop([e1, e2], [(0, n // 2), (n // 2, n)], False)))
e3 = nat.Engine(0); e4 = nat.Engine(0)
q = n // 4
print("four engines x 64 streams: md5 on %.3f ms/step, md5 off %.3f ms/step" % (loop([e1, e2, e3, e4], [(0, q), (q, 2 * q), (2 * q, 3 * q), (3 * q, n)], True), loop([e1, e2, e3, e4], [(0, q), (q, 2 * q), (2 * q, 3 * q), (3 * q, n)], False)))
