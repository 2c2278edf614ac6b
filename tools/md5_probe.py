"""Where does md5_kernel's time go?  python tools/md5_probe.py [n_streams]
Times the MD5 chain of one batch (a) alone on the GPU (the batch's other kernels long done), (b) as the bench runs it: the
MD5 kernels of the previous batches still in flight next to it."""
import sys

sys.path.insert(0, ".")
import numpy as np
import torch
import bench
from pyflac_b200 import _native as nat

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
pcm = bench.make_pcm(0, n)
d_pcm = torch.from_numpy(pcm.reshape(-1)).cuda()
off = np.arange(n, dtype=np.uint64) * np.uint64(bench.N_SAMPLES * bench.CHANNELS)
smp = np.full(n, bench.N_SAMPLES, np.uint64)
eng = nat.Engine(0)
eng.set_profiling(True)
cfg = nat.Engine.make_config(48000, 2, 16, 5, 4096, container_bytes=2)
steps_per_stream = bench.N_SAMPLES * bench.CHANNELS * 2 // 64 * 64
for label, reps in (("alone", 1), ("steady (5 sets in flight)", 8)):
    for _ in range(3):
        eng.encode_device(cfg, d_pcm.data_ptr(), d_pcm.numel(), off, smp)
        eng.join(); eng.sync()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0.record(torch.cuda.current_stream())
    for _ in range(reps):
        eng.encode_device(cfg, d_pcm.data_ptr(), d_pcm.numel(), off, smp)
    eng.join(); eng.sync()
    kt = eng.kernel_times()
    print(f"{label}: md5 {kt['md5']:.3f} ms = {kt['md5'] * 1e-3 * 1.965e9 / steps_per_stream:.1f} cycles/step @1965 MHz;"
          f" autoc {kt['autoc']:.3f} analyze {kt['analyze']:.3f} pack {kt['pack']:.3f}")
