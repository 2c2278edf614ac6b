"""Decode fixtures the level presets never produce ("foreign" streams), so that the decoder's generic paths are pinned too.

Run in the build container (needs oracle/_ref, i.e. /root/reference):  python tests/golden/make_foreign.py
Part A -- written by the bundled libFLAC 1.4.3 with settings away from the presets (oracle/ref_harness.c reads them from
          REF_* environment variables): predictor orders up to 32, exhaustive order search, partition orders 8 and
          partitions of 9 samples, tiny blocks.
Part B -- written by the small bit-writer below: escape-coded partitions (raw residuals, 0..17 bits), the 5-bit Rice
          parameter method, a stream that mixes CONSTANT / VERBATIM / FIXED 0..4 subframes, wasted bits and all four
          channel assignments.  Every crafted stream is accepted by the bundled libFLAC decoder and by the oracle, with the
          PCM the writer started from.
Stored: the .flac files; expectation = the STREAMINFO MD5 inside each file (checked here against the source PCM).
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "foreign")
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _checkers import build_checkers, oracle_decode, ref_decode, ref_encode  # noqa: E402
from pyflac_b200.synth import corpus_signal, music_like  # noqa: E402


def pcm_md5(x, bps):
    w = (bps + 7) // 8
    raw = np.ascontiguousarray(np.asarray(x).astype("<i4")).view(np.uint8).reshape(-1, 4)
    return hashlib.md5(np.ascontiguousarray(raw[:, :w]).tobytes()).digest()


# ------------------------------------------------------------------ part B: a minimal FLAC writer (format.h / RFC 9639) ----
class Bits:
    def __init__(self):
        self.v, self.n = 0, 0

    def put(self, val, nbits):
        if nbits:
            self.v = (self.v << nbits) | (int(val) & ((1 << nbits) - 1))
            self.n += nbits

    def put_signed(self, val, nbits):
        self.put(int(val) & ((1 << nbits) - 1), nbits)

    def unary(self, q):
        self.put(1, q + 1)

    def align(self):
        if self.n % 8:
            self.put(0, 8 - self.n % 8)

    def bytes(self):
        assert self.n % 8 == 0
        return self.v.to_bytes(self.n // 8, "big") if self.n else b""


def crc(data, poly, width):
    c, top, mask = 0, 1 << (width - 1), (1 << width) - 1
    for b in data:
        c ^= b << (width - 8)
        for _ in range(8):
            c = ((c << 1) ^ poly) & mask if c & top else (c << 1) & mask
    return c


def utf8_num(v):
    if v < 0x80:
        return bytes([v])
    out, n = [], 0
    while v >= (0x40 >> n):
        out.append(0x80 | (v & 0x3F)); v >>= 6; n += 1
    return bytes([((0xFF << (7 - n)) & 0xFF) | v] + out[::-1])


FIXED = {0: [], 1: [1], 2: [2, -1], 3: [3, -3, 1], 4: [4, -6, 4, -1]}


def subframe(bw, x, bps, spec):
    """spec: dict(type='constant'|'verbatim'|'fixed', order, po, parts=[('rice', k) | ('esc', nbits)], method, wasted)"""
    wasted = spec.get("wasted", 0)
    x = [int(v) >> wasted for v in x]
    bps -= wasted
    t = {"constant": 0, "verbatim": 1}.get(spec["type"], 8 + spec.get("order", 0))
    bw.put(0, 1); bw.put(t, 6); bw.put(1 if wasted else 0, 1)
    if wasted:
        bw.unary(wasted - 1)
    if spec["type"] == "constant":
        bw.put_signed(x[0], bps); return
    if spec["type"] == "verbatim":
        for v in x:
            bw.put_signed(v, bps)
        return
    order, po, method = spec["order"], spec["po"], spec.get("method", 0)
    for v in x[:order]:
        bw.put_signed(v, bps)
    res = [x[i] - sum(c * x[i - 1 - j] for j, c in enumerate(FIXED[order])) for i in range(order, len(x))]
    bw.put(method, 2); bw.put(po, 4)
    plen, esc = (5, 31) if method else (4, 15)
    psize, pos = len(x) >> po, 0
    for p in range(1 << po):
        n = psize - (order if p == 0 else 0)
        kind, arg = spec["parts"][p % len(spec["parts"])]
        r = res[pos:pos + n]; pos += n
        if kind == "esc":
            need = max([0] + [(int(v) if v >= 0 else ~int(v)).bit_length() + 1 for v in r]) if any(r) else 0
            nb = max(arg, need)
            bw.put(esc, plen); bw.put(nb, 5)
            for v in r:
                bw.put_signed(v, nb)
        else:
            bw.put(arg, plen)
            for v in r:
                u = (v << 1) if v >= 0 else ((-v) << 1) - 1
                bw.unary(u >> arg); bw.put(u & ((1 << arg) - 1), arg)


def frame(number, chans, bps, sample_rate_code, ca, specs):
    """chans: list of per-channel sample lists AFTER decorrelation; ca: 0 independent(n ch), 8 L/S, 9 S/R, 10 M/S"""
    n = len(chans[0])
    bs_code, bs_tail = {192: (1, b""), 576: (2, b""), 1152: (3, b""), 256: (8, b""), 512: (9, b""), 1024: (10, b""), 4096: (12, b"")}.get(
        n, (6, bytes([n - 1])) if n <= 256 else (7, (n - 1).to_bytes(2, "big")))
    ss = {8: 1, 12: 2, 16: 4, 20: 5, 24: 6, 32: 7}[bps]
    cab = (len(chans) - 1) if ca == 0 else ca
    hdr = bytes([0xFF, 0xF8, (bs_code << 4) | sample_rate_code, (cab << 4) | (ss << 1)]) + utf8_num(number) + bs_tail
    hdr += bytes([crc(hdr, 0x07, 8)])
    bw = Bits()
    for c, (x, spec) in enumerate(zip(chans, specs)):
        side = (ca == 8 and c == 1) or (ca == 9 and c == 0) or (ca == 10 and c == 1)
        subframe(bw, x, bps + (1 if side else 0), spec)
    bw.align()
    body = hdr + bw.bytes()
    return body + crc(body, 0x8005, 16).to_bytes(2, "big")


def stream(pcm, bps, sample_rate, sr_code, blocksize, plan):
    """pcm (n, ch) ints; plan(frame_index, block) -> (ca, specs).  Returns a complete .flac byte string."""
    n, ch = pcm.shape
    frames = []
    for f, s in enumerate(range(0, n, blocksize)):
        blk = pcm[s:s + blocksize].astype(np.int64)
        ca, specs = plan(f, blk)
        if ca == 0:
            chans = [blk[:, c].tolist() for c in range(ch)]
        else:
            L, R = blk[:, 0], blk[:, 1]
            M, S = (L + R) >> 1, L - R
            chans = {8: [L, S], 9: [S, R], 10: [M, S]}[ca]
            chans = [c.tolist() for c in chans]
        frames.append(frame(f, chans, bps, sr_code, ca, specs))
    sizes = [len(f) for f in frames]
    si = Bits()
    si.put(blocksize, 16); si.put(blocksize, 16); si.put(min(sizes), 24); si.put(max(sizes), 24)
    si.put(sample_rate, 20); si.put(ch - 1, 3); si.put(bps - 1, 5); si.put(n, 36)
    body = si.bytes() + pcm_md5(pcm, bps)
    return b"fLaC" + bytes([0x80]) + len(body).to_bytes(3, "big") + body + b"".join(frames)


def crafted():
    rng = np.random.default_rng(77)
    out = {}
    # 1: mono 16-bit, blocks of 64: FIXED orders 0..4, partition order 2 mixing Rice and escape partitions (incl. 0-bit escapes)
    x = (music_like(64 * 12, 1, 44100, 16, seed=5).astype(np.int64))
    x[64 * 3:64 * 4] = 1234                                            # a block whose residual is zero: 0-bit escape partitions
    def plan1(f, blk):
        order = f % 5
        parts = [[("rice", 9), ("esc", 0), ("rice", 12), ("esc", 13)], [("esc", 0)], [("esc", 17), ("rice", 14)]][f % 3]
        return 0, [dict(type="fixed", order=order, po=2, parts=parts)]
    out["crafted_escape_s16_mono_bs64"] = (stream(x, 16, 44100, 9, 64, plan1), x, 16)
    # 2: stereo 16-bit, blocks of 4096: every channel assignment, 5-bit parameters (method 1) with k up to 17 and escapes, wasted bits
    y = music_like(4096 * 5, 2, 48000, 16, seed=9).astype(np.int64)
    y[4096 * 3:4096 * 4] &= ~7                                          # three wasted bits in block 3
    def plan2(f, blk):
        ca = [10, 8, 9, 0, 10][f]
        w = 3 if f == 3 else 0
        a = dict(type="fixed", order=2 + f % 3, po=[0, 3, 6, 4, 5][f], method=1, parts=[("rice", 10), ("esc", 0), ("rice", 17), ("rice", 11)], wasted=w)
        b = dict(type="fixed", order=(f + 1) % 5, po=[4, 0, 2, 6, 1][f], method=f % 2, parts=[("rice", 11), ("rice", 13), ("esc", 15)], wasted=w)
        return ca, [a, b]
    out["crafted_method1_s16_st_bs4096"] = (stream(y, 16, 48000, 10, 4096, plan2), y, 16)
    # 3: 3 channels 24-bit, blocks of 192: CONSTANT / VERBATIM / FIXED side by side, escapes wider than 16 bits
    z = np.stack([np.full(192 * 4, -70000), rng.integers(-2**23, 2**23, 192 * 4), music_like(192 * 4, 1, 96000, 24, seed=3)[:, 0]], axis=1).astype(np.int64)
    def plan3(f, blk):
        return 0, [dict(type="constant"), dict(type="verbatim"), dict(type="fixed", order=4 - f, po=f % 3, method=1, parts=[("esc", 20 + f), ("rice", 16)])]
    out["crafted_mixed_s24_3ch_bs192"] = (stream(z, 24, 96000, 11, 192, plan3), z, 24)
    return out


def tuned():
    """part A: (name, pcm, sample_rate, bps, level, blocksize, subset, env)"""
    a = music_like(4096 * 3 + 500, 2, 96000, 16, seed=41)
    def tonal(n, ch, sr, bps, seed, partials=14):               # many close partials: only long predictors null them
        rng = np.random.default_rng(seed)
        t = np.arange(n) / sr
        x = sum(rng.uniform(0.02, 0.06) * np.sin(2 * np.pi * rng.uniform(200, 0.4 * sr / 2) * t + rng.uniform(0, 6.28)) for _ in range(partials))
        x = np.stack([x * (1 - 0.1 * c) + rng.normal(0, 2e-6, n) for c in range(ch)], axis=1)
        return np.rint(x * 2 ** (bps - 1)).astype(np.int32)
    b = tonal(4096 * 2, 1, 192000, 24, 42)
    a20 = tonal(4096 * 2, 2, 44100, 16, 46, partials=9)
    c = corpus_signal("mixed", 4096 * 3, 2, 16, seed=43)
    d = music_like(576 * 6 + 100, 2, 44100, 16, seed=44)
    e = music_like(16 * 40 + 7, 2, 8000, 16, seed=45)
    return [
        ("tuned_lpc32_s16_st", a, 96000, 16, 8, 4096, 1, dict(REF_MAX_LPC_ORDER="32", REF_EXHAUSTIVE="1", REF_QLP_PRECISION="15")),
        ("tuned_lpc32_s24_mono", b, 192000, 24, 8, 4096, 1, dict(REF_MAX_LPC_ORDER="32", REF_EXHAUSTIVE="1")),
        ("tuned_lpc20_s16_st_nonsubset", a20, 44100, 16, 8, 4096, 0, dict(REF_MAX_LPC_ORDER="20")),
        ("tuned_po8_s16_st", c, 48000, 16, 5, 4096, 1, dict(REF_MIN_PART_ORDER="8", REF_MAX_PART_ORDER="8")),
        ("tuned_bs576_po6_s16_st", d, 44100, 16, 5, 576, 1, dict(REF_MIN_PART_ORDER="6", REF_MAX_PART_ORDER="6")),
        ("tuned_bs16_s16_st", e, 8000, 16, 5, 16, 1, dict()),
    ]


def main():
    build_checkers()
    os.makedirs(OUT, exist_ok=True)
    manifest = []

    def check_and_store(name, data, pcm, bps, note):
        pcm = np.asarray(pcm).reshape(len(pcm), -1)
        assert data[26:42] == pcm_md5(pcm, bps), name
        got, info = ref_decode(data)
        assert info["errors"] == 0 and np.array_equal(got.astype(np.int64), pcm.astype(np.int64)), (name, "libFLAC decode")
        got, _ = oracle_decode(data)
        assert np.array_equal(got.astype(np.int64), pcm.astype(np.int64)), (name, "oracle decode")
        with open(os.path.join(OUT, name + ".flac"), "wb") as f:
            f.write(data)
        manifest.append(dict(name=name, channels=int(pcm.shape[1]), bps=bps, samples=int(pcm.shape[0]), flac_bytes=len(data),
                             pcm_md5=data[26:42].hex(), note=note))

    for name, pcm, sr, bps, level, bs, subset, env in tuned():
        plain = ref_encode(pcm, sr, bps, level, bs, seekable=True, streamable_subset=bool(subset))
        os.environ.update(env)
        try:
            data = ref_encode(pcm, sr, bps, level, bs, seekable=True, streamable_subset=bool(subset))
        finally:
            for k in env:
                del os.environ[k]
        check_and_store(name, data, pcm, bps, f"libFLAC 1.4.3 level {level} with {env} ({len(plain)} bytes without the tuning)")
    for name, (data, pcm, bps) in crafted().items():
        check_and_store(name, data, pcm, bps, "hand-written bitstream (make_foreign.py)")
    with open(os.path.join(OUT, "foreign.json"), "w") as f:
        json.dump(dict(cases=manifest), f, indent=1)
    for m in manifest:
        print(m["name"], m["flac_bytes"], m["note"][:90])


if __name__ == "__main__":
    main()
