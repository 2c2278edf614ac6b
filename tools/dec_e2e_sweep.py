"""Host->host decode of the BASELINE configs[3] shape with different chunk counts: python tools/dec_e2e_sweep.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from pyflac_b200 import _native as nat
eng = nat.Engine(0)
dev = torch.device("cuda", 0)
D, NS = 4096, 131072
pcm4 = bench.make_pcm_short(0, D, NS)
d4 = torch.from_numpy(pcm4.reshape(-1)).to(dev)
cfg = nat.Engine.make_config(48000, 2, 16, 5, 4096, container_bytes=2)
eng.encode_device(cfg, d4.data_ptr(), d4.numel(), np.arange(D, dtype=np.uint64) * np.uint64(NS * 2), np.full(D, NS, np.uint64))
enc = eng.fetch(); del d4
tot = int(enc["total_bytes"])
hb = torch.empty(tot + 16, dtype=torch.uint8).pin_memory(); hb.numpy()[:tot] = enc["arena"][:tot]; hb.numpy()[tot:] = 0
so = np.array([si.byte_off for si in enc["streams"]], np.uint64); sl = np.array([si.byte_len for si in enc["streams"]], np.uint64)
out = torch.empty(pcm4.size, dtype=torch.int16).pin_memory()
for ch in (1, 2, 3, 4, 6, 8):
    os.environ["FLACB200_DEC_CHUNKS"] = str(ch)
    eng.decode_host_pipelined(hb.numpy(), so, sl, out.numpy(), 2)
    t0 = time.perf_counter()
    for _ in range(3):
        n, infos = eng.decode_host_pipelined(hb.numpy(), so, sl, out.numpy(), 2)
    dt = (time.perf_counter() - t0) / 3
    ok = np.array_equal(out.numpy(), pcm4.reshape(-1)) and all(infos[s].status == 0 for s in range(D))
    print("chunks", ch, f"{dt*1e3:.1f} ms", f"{pcm4.size/dt/1e6:.0f} MSamples/s", "ok" if ok else "MISMATCH")
