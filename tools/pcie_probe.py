"""Probe host<->device copy bandwidth and the host-path breakdown with / without MD5 (run on the GPU box)."""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench
from pyflac_b200 import _native as nat

dev = torch.device("cuda", 0)
h = torch.empty(491520000, dtype=torch.uint8).pin_memory()
d = torch.empty_like(h, device=dev)
for name, fn in [("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))]:
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 5
    print(name, "pinned GB/s", h.numel() / dt / 1e9)
pcm = bench.make_pcm(0)
hp = torch.from_numpy(pcm.reshape(-1)).pin_memory()
eng = nat.Engine(0)
L = nat.lib()
arena = torch.empty(hp.numel() * 2 + (64 << 20), dtype=torch.uint8).pin_memory()
off = (np.arange(256, dtype=np.uint64) * np.uint64(480000 * 2))
smp = np.full(256, 480000, np.uint64)
for md5 in (1, 0):
    cfg = nat.Engine.make_config(48000, 2, 16, 5, 4096, do_md5=bool(md5))
    tot = C.c_uint64(0)
    for it in range(4):
        t = time.perf_counter()
        rc = L.flacb200_encode_batch_host(eng._h, C.byref(cfg), hp.data_ptr(), hp.numel(), 256, off.ctypes.data, smp.ctypes.data,
                                          arena.data_ptr(), arena.numel(), C.byref(tot), None, None, None)
        dt = time.perf_counter() - t
        b = (C.c_double * 6)()
        L.flacb200_host_path_times(eng._h, b)
        print("md5", md5, "iter", it, "rc", rc, "ms", round(dt * 1e3, 2), [round(v, 2) for v in b])
