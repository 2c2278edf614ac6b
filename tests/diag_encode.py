"""GPU parity diagnostic: encode the golden cases (and optional extra corpus) on the GPU and, for every
mismatch against libFLAC's bytes, locate the first diverging decision by comparing the kernel's
analysis trace with the oracle's trace.  Test tooling (uses oracle/); run on the GPU box:
    python tests/diag_encode.py [--extra]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _checkers as ck  # noqa: E402
from pyflac_b200 import _native as nat  # noqa: E402
from pyflac_b200.synth import corpus_signal, CORPUS_KINDS  # noqa: E402


def compare_traces(case, x, eng, max_report=3):
    ch, bps, sr, level, bs = case["channels"], case["bps"], case["sample_rate"], case["level"], case["blocksize"]
    cfg = nat.Engine.make_config(sr, ch, bps, level, bs, debug_trace=True)
    dt = np.int16 if bps <= 16 else np.int32
    pcm = np.ascontiguousarray(x, dt).reshape(-1)
    eng.encode_host(cfg, pcm, [0], [x.shape[0]])
    out = eng.fetch()
    nsig = ch + (2 if (ch == 2 and level in (1, 2, 4, 5, 6, 7, 8)) else 0)
    plans, ca, dbg = eng.fetch_trace(nsig)
    _, traces = ck.oracle_encode(x, sr, bps, level, bs, with_trace=True)
    nrep = 0
    for f in range(len(ca)):
        t = traces[f]
        msgs = []
        if int(ca[f]) != t.channel_assignment:
            msgs.append(f"  ca gpu {ca[f]} oracle {t.channel_assignment}")
        for s in range(nsig):
            g, d, o = plans[f * nsig + s], dbg[f * nsig + s], t.sig[s]
            ob = o.best
            def diff(name, a, b):
                if a != b:
                    msgs.append(f"  sig{s} {name}: gpu {a} oracle {b}")
            diff("wasted", g.wasted, o.wasted); diff("sbps", g.sbps, o.sbps)
            diff("fixed_err", list(d.fixed_err), list(o.fixed_err)); diff("fixed_order", d.fixed_order, o.fixed_order)
            diff("is_constant", d.is_constant, o.is_constant)
            diff("fixed_bits", d.fixed_bits, o.fixed_bits)
            diff("n_apod", d.n_apod, o.n_apod)
            for st in range(min(d.n_apod, o.n_apod, 9)):
                nl = 13
                ga, oa = list(d.autoc[st])[:nl], list(o.autoc[st])[:nl]
                if o.lpc_order[st] or d.lpc_order[st]:
                    if ga[:9] != oa[:9]:
                        bad = [i for i in range(nl) if ga[i] != oa[i]]
                        msgs.append(f"  sig{s} step{st} autoc differs at lags {bad}: gpu {[ga[i] for i in bad[:3]]} oracle {[oa[i] for i in bad[:3]]}")
                    ge, oe = list(d.lpc_err[st]), list(o.lpc_err[st])[:12]
                    mo = max(1, o.lpc_order[st])
                    if ge[:mo] != oe[:mo]:
                        msgs.append(f"  sig{s} step{st} lpc_err differs: gpu {ge[:mo]} oracle {oe[:mo]}")
                diff(f"step{st} lpc_order", d.lpc_order[st], o.lpc_order[st])
                diff(f"step{st} lpc_bits", d.lpc_bits[st], o.lpc_bits[st])
            diff("best.type", g.type, ob.type); diff("best.order", g.order, ob.order)
            diff("best.bits", g.bits_est, ob.bits_est)
            if ob.type == 3:
                diff("best.precision", g.precision, ob.precision); diff("best.shift", g.shift, ob.shift)
                diff("best.qlp", list(g.qlp)[:ob.order], list(ob.qlp)[:ob.order])
            if ob.type >= 2:
                diff("best.part_order", g.part_order, ob.partition_order)
                diff("best.rice", list(g.rice)[:1 << ob.partition_order], list(ob.rice)[:1 << ob.partition_order])
                diff("best.rice2", g.rice2, ob.rice2)
        if msgs:
            print(f" frame {f}: decision mismatch")
            for m in msgs[:24]:
                print(m)
            nrep += 1
            if nrep >= max_report:
                break
    if nrep == 0:
        print(" all decisions match the oracle -> the divergence is in packing")
    return out


def main():
    ck.build_checkers()
    with open(os.path.join(ROOT, "tests", "golden", "manifest.json")) as f:
        cases = json.load(f)["cases"]
    eng = nat.Engine(0)
    nbad = 0
    for case in cases:
        if case["level"] in (1, 4) and case["channels"] == 2:
            continue
        x = np.load(os.path.join(ROOT, "tests", "golden", case["name"] + ".pcm.npy"))
        flac = open(os.path.join(ROOT, "tests", "golden", case["name"] + ".flac"), "rb").read()
        try:
            got, out = nat.encode_streams(eng, [x], case["sample_rate"], case["bps"], case["level"], case["blocksize"])
        except Exception as e:  # noqa: BLE001
            print("CASE", case["name"], "ERROR", e)
            nbad += 1
            continue
        ok = got[0] == flac
        print("CASE", case["name"], "OK" if ok else "MISMATCH", "guard_hits", out["log_guard_hits"])
        if not ok:
            nbad += 1
            a, b = got[0], flac
            d = next((i for i in range(min(len(a), len(b))) if a[i] != b[i]), min(len(a), len(b)))
            fr = next((i for i, (o, l) in enumerate(zip(case["frame_off"], case["frame_len"])) if o + l > d), -1)
            print(f" len gpu {len(a)} ref {len(b)} first diff byte {d} (frame {fr}, frame starts at {case['frame_off'][fr] if fr >= 0 else -1})")
            print(f" gpu frame_len {list(out['frame_len'])[:6]} ref {case['frame_len'][:6]}")
            compare_traces(case, x, eng)
    if "--extra" in sys.argv:
        for level in [0, 2, 3, 5, 6, 7, 8]:
            for kind in CORPUS_KINDS:
                for ch, bps, n, bs, sr in [(2, 16, 4096 * 2 + 768, 0, 48000), (1, 24, 4096 + 100, 4096, 192000), (1, 16, 3000, 1000, 44100)]:
                    x = corpus_signal(kind, n, ch, bps, seed=level * 7 + ch)
                    ref = ck.oracle_encode(x, sr, bps, level, bs)
                    got, out = nat.encode_streams(eng, [x], sr, bps, level, bs)
                    if got[0] != ref:
                        nbad += 1
                        print("EXTRA MISMATCH", level, kind, ch, bps, n, bs)
                        compare_traces(dict(channels=ch, bps=bps, sample_rate=sr, level=level, blocksize=bs), x, eng, 1)
    print("TOTAL BAD", nbad)
    return 1 if nbad else 0


if __name__ == "__main__":
    sys.exit(main())
