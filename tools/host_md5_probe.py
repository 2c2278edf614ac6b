"""Host -> host call with the MD5 split forced: python tools/host_md5_probe.py  (prints flacb200_host_path_info per setting)"""
import os
import sys

sys.path.insert(0, ".")
import numpy as np
import torch
import bench
from pyflac_b200 import _native as nat

n = 256
pcm = bench.make_pcm(0, n)
h = torch.from_numpy(pcm.reshape(-1)).pin_memory().numpy()
off = np.arange(n, dtype=np.uint64) * np.uint64(bench.N_SAMPLES * bench.CHANNELS)
smp = np.full(n, bench.N_SAMPLES, np.uint64)
arena = torch.empty(h.nbytes + (64 << 20), dtype=torch.uint8).pin_memory().numpy()
eng = nat.Engine(0)
cfg = nat.Engine.make_config(48000, 2, 16, 5, 4096, container_bytes=2)
for env in [dict(), dict(FLACB200_MD5_GPU_CHUNKS="4"), dict(FLACB200_MD5_GPU_CHUNKS="12"), dict(FLACB200_MD5_THREADS="2"), dict(FLACB200_MD5_THREADS="1")]:
    for k in ("FLACB200_MD5_GPU_CHUNKS", "FLACB200_CHUNKS", "FLACB200_MD5_THREADS", "FLACB200_MD5_NOGATE", "FLACB200_MD5_GATE_MODE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    for rep in range(6):
        out = eng.encode_host_to_host(cfg, h, off, smp, arena=arena)
    pi = out["path_info"]
    print(env, {k: (round(v, 2) if isinstance(v, float) else v) for k, v in pi.items()})
