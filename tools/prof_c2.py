"""configs[2]-shaped encode for ncu captures: 24-bit mono 192 kHz level 8, n streams of 65536 samples: python tools/prof_c2.py [n_streams]"""
import sys

sys.path.insert(0, ".")
import numpy as np
import torch
from pyflac_b200 import _native as nat
from pyflac_b200.synth import music_like

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
base = [music_like(65536, 1, 192000, 24, seed=300 + s) for s in range(16)]
pcm = np.stack([np.roll(base[s % 16], 977 * (s // 16), axis=0) for s in range(n)]).astype(np.int32)
d = torch.from_numpy(pcm.reshape(-1)).cuda()
off = np.arange(n, dtype=np.uint64) * np.uint64(65536)
cnt = np.full(n, 65536, np.uint64)
eng = nat.Engine(0)
eng.set_profiling(True)
cfg = nat.Engine.make_config(192000, 1, 24, 8, 4096, container_bytes=4)
for _ in range(2):
    eng.encode_device(cfg, d.data_ptr(), d.numel(), off, cnt)
eng.join(); eng.sync()
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in eng.kernel_times().items()})
