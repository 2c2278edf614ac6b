"""GPU diagnostic: fused kernel vs multi-kernel path vs oracle on the inputs of test_fused_and_multi_kernel_paths_agree;
prints which streams / frames differ (test tooling, uses oracle/)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _checkers as ck  # noqa: E402
from pyflac_b200 import _native as nat  # noqa: E402
from pyflac_b200.synth import corpus_signal, CORPUS_KINDS, music_like  # noqa: E402

xs = [corpus_signal(kind, 4096 * 2 + 1000 + 3 * i, 2, 16, seed=40 + i) for i, kind in enumerate(CORPUS_KINDS)]
xs += [music_like(n, 2, 48000, 16, seed=n) for n in (1, 3, 17, 4097, 4096 * 3)]
names = list(CORPUS_KINDS) + ["m1", "m3", "m17", "m4097", "m12288"]
for level, bs in [(0, 0), (2, 0), (3, 0), (5, 0), (5, 1000), (6, 0), (8, 0), (8, 1152), (7, 4608), (5, 16)]:
    os.environ.pop("FLACB200_NO_FUSED", None)
    e1 = nat.Engine(0)
    a, oa = nat.encode_streams(e1, xs, 44100, 16, level, bs)
    e1.close()
    bad = 0
    for s, x in enumerate(xs):
        ref, roff, rlen = ck.oracle_encode(x, 44100, 16, level, bs, with_index=True)
        if a[s] != ref:
            bad += 1
            fo = [int(v) for v, st in zip(oa["frame_off"], oa["frame_stream"]) if st == s]
            fl = [int(v) for v, st in zip(oa["frame_len"], oa["frame_stream"]) if st == s]
            fs = [int(v) for v, st in zip(oa["frame_samples"], oa["frame_stream"]) if st == s]
            base = int(oa["streams"][s].byte_off)
            print(f"level {level} bs {bs} stream {s} ({names[s]}, n={len(x)}): fused differs from oracle; frames {len(fo)} vs {len(rlen)}")
            for f in range(min(len(fo), len(rlen))):
                g = a[s][fo[f] - base: fo[f] - base + fl[f]]
                r = ref[int(roff[f]): int(roff[f]) + int(rlen[f])]
                if g != r:
                    k = next((i for i in range(min(len(g), len(r))) if g[i] != r[i]), min(len(g), len(r)))
                    print(f"   frame {f} N={fs[f]} len gpu {len(g)} oracle {len(r)} first diff at byte {k}: gpu {g[max(0,k-2):k+6].hex()} oracle {r[max(0,k-2):k+6].hex()} hdr {r[:8].hex()}")
                    break
    print(f"level {level} bs {bs}: {bad} streams differ")
