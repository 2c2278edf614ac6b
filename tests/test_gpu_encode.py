"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(libflacb200.so via pyflac_b200/_native.py); the oracle / reference binary are only the checkers."""
import hashlib
import os

import numpy as np
import pytest

from conftest import golden_cases, load_golden
from pyflac_b200.synth import corpus_signal, CORPUS_KINDS, music_like

pytestmark = pytest.mark.gpu
CASES = golden_cases()


@pytest.fixture(scope="module")
def eng():
    from pyflac_b200 import _native as nat
    return nat.Engine(0)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_encode_matches_golden(eng, case):
    """bit-exact against bytes produced by the reference's libFLAC 1.4.3 (tests/golden)."""
    from pyflac_b200 import _native as nat
    x, flac = load_golden(case)
    got, out = nat.encode_streams(eng, [x], case["sample_rate"], case["bps"], case["level"], case["blocksize"])
    assert got[0] == flac
    assert [int(v) for v in out["frame_len"]] == case["frame_len"]
    assert out["log_guard_hits"] == 0


@pytest.mark.parametrize("level", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_encode_corpus_vs_oracle(eng, checkers, level):
    """every corpus kind (each subframe type / branch), 16-bit stereo + 24-bit mono + odd blocksize, one batch per level"""
    from pyflac_b200 import _native as nat
    for ch, bps, n, bs, sr in [(2, 16, 4096 * 2 + 768, 0, 48000), (1, 24, 4096 + 100, 4096, 192000), (1, 16, 3000, 1000, 44100),
                               (2, 16, 700, 192, 22050)]:
        xs = [corpus_signal(kind, n, ch, bps, seed=level * 7 + ch) for kind in CORPUS_KINDS]
        got, out = nat.encode_streams(eng, xs, sr, bps, level, bs)
        for kind, x, g in zip(CORPUS_KINDS, xs, got):
            assert g == checkers.oracle_encode(x, sr, bps, level, bs), (level, kind, ch, bps, n, bs)
        assert out["log_guard_hits"] == 0


def test_encode_loose_mid_side(eng, checkers):
    """levels 1 and 4 on stereo: one L/R-vs-M/S decision every round(0.4 s) of frames, followers analyse only that pair
    (up: process_subframes_ loose_mid_side_stereo).  Signals whose best assignment changes along the stream, several
    decision periods (blocksize / sample rate), a stream shorter than one period, and a mono stream in the same level."""
    from pyflac_b200 import _native as nat
    rng = np.random.default_rng(5)
    n = 4096 * 14 + 300
    a = corpus_signal("music", n, 2, 16, seed=31)
    b = corpus_signal("lr_uncorr", n, 2, 16, seed=32)
    sw = a.copy()
    sw[n // 3: 2 * n // 3] = b[n // 3: 2 * n // 3]                       # correlated -> uncorrelated -> correlated
    wide = a.copy(); wide[:, 1] = -wide[:, 0]                            # side-heavy
    for level in (1, 4):
        for sr, bs in [(44100, 0), (8000, 1152), (96000, 256), (48000, 4608)]:
            xs = [sw, wide, a[:3000], b]
            got, out = nat.encode_streams(eng, xs, sr, 16, level, bs)
            for i, (x, g) in enumerate(zip(xs, got)):
                assert g == checkers.oracle_encode(x, sr, 16, level, bs), (level, sr, bs, i)
            assert out["log_guard_hits"] == 0
        m = corpus_signal("music", 9000, 1, 16, seed=3)
        got, _ = nat.encode_streams(eng, [m], 44100, 16, level, 0)
        assert got[0] == checkers.oracle_encode(m, 44100, 16, level, 0)
    x24 = corpus_signal("music", 4096 * 6, 2, 24, seed=9)
    got, _ = nat.encode_streams(eng, [x24], 96000, 24, 4, 4096)
    assert got[0] == checkers.oracle_encode(x24, 96000, 24, 4, 4096)


def test_encode_32bit(eng, checkers):
    """bits_per_sample = 32 (pyFLAC's int32 input, reference encoder.py:109): the _limit_residual fixed-predictor search
    (orders whose residual leaves int32 are invalid; all-zero blocks are CONSTANT, other constant blocks FIXED order 1),
    LPC with the int32 residual check, full-scale noise (VERBATIM), INT32_MIN samples, wasted bits that bring the
    subframe below 28 bits.  Mono, 3 channels, and stereo at the levels without mid/side; lengths 0 or 1 mod 4."""
    from pyflac_b200 import _native as nat
    rng = np.random.default_rng(3)
    n = 4096 * 2 + 300
    m = music_like(n, 1, 48000, 24, seed=3).astype(np.int64)
    def c32(a):
        return np.clip(a, -2**31, 2**31 - 1).astype(np.int32)
    spikes = m * 100
    spikes[rng.integers(0, n, 5)] = -2**31
    sig = {"zeros": np.zeros((n, 1), np.int64), "dc_odd": np.full((n, 1), 123456789), "dc_min": np.full((n, 1), -2**31),
           "noise_full": rng.integers(-2**31, 2**31, (n, 1)), "music24_shl8": m * 256, "music_x200": m * 200 + rng.integers(-3, 4, (n, 1)),
           "music_clip": m * 400, "walk": np.cumsum(rng.integers(-2**24, 2**24, (n, 1)), axis=0), "spikes": spikes,
           "minmax": np.where(np.arange(n)[:, None] % 2 == 0, -2**31, 2**31 - 1), "lowbits": m // 64}
    names = list(sig)
    for level in (0, 2, 5, 8):
        for bs in (0, 1000):
            xs = [c32(sig[k]) for k in names]
            got, out = nat.encode_streams(eng, xs, 48000, 32, level, bs)
            for k, x, g in zip(names, xs, got):
                assert g == checkers.oracle_encode(x, 48000, 32, level, bs), (k, level, bs)
            assert out["log_guard_hits"] == 0
    multi = [c32(np.concatenate([sig[k], np.roll(sig[k], 17, axis=0) // 2, np.roll(sig[k], 40, axis=0) // 3], axis=1)) for k in ("music_x200", "walk", "zeros")]
    got, _ = nat.encode_streams(eng, multi, 44100, 32, 5, 0)
    for x, g in zip(multi, got):
        assert g == checkers.oracle_encode(x, 44100, 32, 5, 0)
    stereo = [c32(np.concatenate([sig[k], np.roll(sig[k], 9, axis=0) // 2], axis=1)) for k in ("music_x200", "noise_full", "zeros", "dc_odd")]
    for level in (0, 3):
        got, _ = nat.encode_streams(eng, stereo, 48000, 32, level, 0)
        for x, g in zip(stereo, got):
            assert g == checkers.oracle_encode(x, 48000, 32, level, 0), level
    if checkers.ref_available():
        g, _ = nat.encode_streams(eng, [c32(sig["music_x200"])], 48000, 32, 5, 0)
        assert g[0] == checkers.ref_encode(c32(sig["music_x200"]), 48000, 32, 5, 0)
    # 32-bit stereo with mid/side analysis: the side channel has 33 bits (64-bit predictor / residual arithmetic, 33-bit
    # warm-up and verbatim samples; an all-zero side reports one wasted bit, get_wasted_bits_wide_)
    L = m[:, 0] * 200
    R0 = np.roll(m[:, 0], 5) * 150
    pairs = {"same": (L, L.copy()), "anti_noisy": (L, -L + rng.integers(-3, 4, n)), "anti": (L, -L), "anti_odd": (L, (-L) // 2 * 2 + 1),
             "noise": (rng.integers(-2**31, 2**31, n), rng.integers(-2**31, 2**31, n)), "neg_m1": (sig["noise_full"][:, 0], -sig["noise_full"][:, 0] - 1),
             "near": (L, L + rng.integers(-1000, 1000, n)), "rails": (np.full(n, 2**31 - 1), np.full(n, -2**31)), "indep": (L, R0),
             "sparse": (np.where(rng.random(n) < 0.001, 2**31 - 1, 0), np.where(rng.random(n) < 0.001, -2**31, 0))}
    xs = [c32(np.stack([a, b], axis=1)) for a, b in pairs.values()]
    for level in (1, 2, 4, 5, 8):
        for bs in (0, 1152):
            got, out = nat.encode_streams(eng, xs, 48000, 32, level, bs)
            for k, x, g in zip(pairs, xs, got):
                assert g == checkers.oracle_encode(x, 48000, 32, level, bs), (k, level, bs)
            assert out["log_guard_hits"] == 0
    if checkers.ref_available():
        got, _ = nat.encode_streams(eng, xs, 48000, 32, 5, 0)
        for k, x, g in zip(pairs, xs, got):
            assert g == checkers.ref_encode(x, 48000, 32, 5, 0), k


def test_encode_limit_min_bitrate(eng, checkers):
    """FLAC__stream_encoder_set_limit_min_bitrate: a frame may not consist of constant subframes only
    (up: process_subframes_; ref: stream_encoder.h:1105-1115)"""
    from pyflac_b200 import _native as nat
    n = 4096 * 3 + 100
    sil = np.zeros((n, 2), np.int16)
    dc = np.full((n, 2), 1234, np.int16)
    half = corpus_signal("music", n, 2, 16, seed=4); half[:, 0] = 77               # first channel constant, second not
    half2 = corpus_signal("music", n, 2, 16, seed=5); half2[:, 1] = -5             # second channel constant
    part = corpus_signal("music", n, 2, 16, seed=6); part[4096:8192] = 0           # one silent frame in the middle
    for level in (0, 1, 2, 5, 8):
        xs = [sil, dc, half, half2, part]
        got, _ = nat.encode_streams(eng, xs, 48000, 16, level, 0, limit_min_bitrate=True)
        for i, (x, g) in enumerate(zip(xs, got)):
            assert g == checkers.oracle_encode(x, 48000, 16, level, 0, limit_min_bitrate=True), (level, i)
        got0, _ = nat.encode_streams(eng, xs, 48000, 16, level, 0)
        assert got0[0] != got[0]
    for ch in (1, 3):
        z = np.zeros((5000, ch), np.int16)
        got, _ = nat.encode_streams(eng, [z], 44100, 16, 5, 0, limit_min_bitrate=True)
        assert got[0] == checkers.oracle_encode(z, 44100, 16, 5, 0, limit_min_bitrate=True), ch
    ref = checkers.ref_encode(sil, 48000, 16, 5, 0, limit_min_bitrate=True) if checkers.ref_available() else None
    if ref is not None:
        got, _ = nat.encode_streams(eng, [sil], 48000, 16, 5, 0, limit_min_bitrate=True)
        assert got[0] == ref


def test_encode_edge_lengths(eng, checkers):
    """ragged batch: streams of different lengths incl. 1..5 samples, N*k, N*k+1, N*k+768"""
    from pyflac_b200 import _native as nat
    lens = [1, 2, 4, 5, 16, 4095, 4096, 4097, 8192, 8193, 4096 * 2 + 768, 100]
    xs = [corpus_signal("music", n, 2, 16, seed=n) for n in lens]
    got, out = nat.encode_streams(eng, xs, 44100, 16, 5, 0)
    for n, x, g in zip(lens, xs, got):
        assert g == checkers.oracle_encode(x, 44100, 16, 5, 0), n


def test_encode_multichannel_and_depths(eng, checkers):
    from pyflac_b200 import _native as nat
    for ch in [1, 3, 6, 8]:
        x = corpus_signal("lr_uncorr", 5000, ch, 16, seed=ch)
        got, _ = nat.encode_streams(eng, [x], 48000, 16, 5, 0)
        assert got[0] == checkers.oracle_encode(x, 48000, 16, 5, 0), ch
    for bps in [8, 12, 20, 24]:
        for kind in ["music", "wasted"]:
            x = corpus_signal(kind if bps >= 12 else "music", 6000, 2, bps, seed=bps)
            got, _ = nat.encode_streams(eng, [x], 96000, bps, 5, 0)
            assert got[0] == checkers.oracle_encode(x, 96000, bps, 5, 0), (bps, kind)


def test_encode_low_bit_depths(eng, checkers):
    """bits_per_sample 4..7 (libFLAC's lower limit is 4; none is in the streamable subset): qlp precision max(5, 2 + bps/2),
    every level family, mono / stereo (mid/side: the side channel has bps + 1 bits) / 3 channels, odd and default blocksizes."""
    from pyflac_b200 import _native as nat
    rng = np.random.default_rng(11)
    for bps in (4, 5, 6, 7):
        lo, hi = -(1 << (bps - 1)), (1 << (bps - 1)) - 1
        n = 4096 + 1152 + 77
        t = np.arange(n)
        for ch in (1, 2, 3):
            tone = np.stack([np.round((hi - 1) * 0.8 * np.sin(2 * np.pi * (220 + 30 * c) * t / 48000.0 + c)) for c in range(ch)], axis=1)
            xs = [np.clip(tone + rng.integers(-1, 2, (n, ch)), lo, hi).astype(np.int16),
                  rng.integers(lo, hi + 1, (n, ch)).astype(np.int16),                         # full-scale noise -> VERBATIM
                  np.full((n, ch), lo, np.int16),                                             # most negative value, constant
                  (np.clip(tone, lo, hi).astype(np.int16) >> 1) << 1]                         # one wasted bit
            for level, bs in [(0, 0), (2, 0), (5, 0), (8, 1000), (4, 576)]:
                got, out = nat.encode_streams(eng, xs, 48000, bps, level, bs, streamable_subset=False)
                for i, (x, g) in enumerate(zip(xs, got)):
                    assert g == checkers.oracle_encode(x, 48000, bps, level, bs, streamable_subset=False), (bps, ch, level, bs, i)
                assert out["log_guard_hits"] == 0
        if checkers.ref_available():
            x = np.clip(np.round(hi * 0.7 * np.sin(t / 9.0))[:, None] + rng.integers(-1, 2, (n, 2)), lo, hi).astype(np.int16)
            got, _ = nat.encode_streams(eng, [x], 44100, bps, 5, 0, streamable_subset=False)
            assert got[0] == checkers.ref_encode(x, 44100, bps, 5, 0, streamable_subset=False), bps


def test_encode_sample_rates_and_frame_numbers(eng, checkers):
    from pyflac_b200 import _native as nat
    for sr in [8000, 12345, 50001, 65535, 96000, 176400, 655350]:
        x = corpus_signal("music", 3000, 2, 16, seed=1)
        got, _ = nat.encode_streams(eng, [x], sr, 16, 5, 1024, streamable_subset=False)
        assert got[0] == checkers.oracle_encode(x, sr, 16, 5, 1024, streamable_subset=False), sr
    x = corpus_signal("music", 16 * 70000, 1, 16, seed=5)       # frame numbers up to 3 UTF-8 bytes
    got, _ = nat.encode_streams(eng, [x], 48000, 16, 0, 16)
    assert got[0] == checkers.oracle_encode(x, 48000, 16, 0, 16)


def test_result_frames_then_md5(eng, checkers):
    """flacb200_encode_result_frames / flacb200_encode_fetch_md5: the frames of a batch are final (and fetchable) before
    its MD5 chain ends; the digests arrive later and are patched into the stream images.  Everything but the 16 MD5 bytes
    must already equal libFLAC's output at the first moment, everything at the second."""
    from pyflac_b200 import _native as nat
    xs = [music_like(4096 * 40 + 123 * s, 2, 48000, 16, seed=300 + s) for s in range(24)]
    flat = np.concatenate([x.reshape(-1) for x in xs])
    sizes = np.array([x.size for x in xs], np.uint64)
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
    cfg = nat.Engine.make_config(48000, 2, 16, 5, 0)
    eng.encode_host(cfg, flat, offs, sizes // 2)
    r = eng.result(wait_md5=False)
    assert r.n_streams == 24 and r.total_bytes > 0
    dig = eng.fetch_md5()
    out = eng.fetch()
    for s, x in enumerate(xs):
        si = out["streams"][s]
        blob = out["arena"][int(si.byte_off): int(si.byte_off + si.byte_len)].tobytes()
        assert blob == checkers.oracle_encode(x, 48000, 16, 5, 0), s
        assert bytes(dig[s]) == hashlib.md5(x.tobytes()).digest() == bytes(si.md5) == blob[26:42]
    # do_md5 = 0: digests are zero, as libFLAC's set_do_md5(false)
    cfg0 = nat.Engine.make_config(48000, 2, 16, 5, 0, do_md5=False)
    eng.encode_host(cfg0, flat, offs, sizes // 2)
    eng.result(wait_md5=False)
    assert not eng.fetch_md5().any()
    out0 = eng.fetch()
    si = out0["streams"][3]
    assert bytes(out0["arena"][int(si.byte_off) + 26: int(si.byte_off) + 42]) == bytes(16)


def test_fused_and_multi_kernel_paths_agree(checkers, monkeypatch):
    """16-bit stereo runs the fused kernel (csrc/enc_fused.cu), FLACB200_NO_FUSED=1 the multi-kernel path the other layouts
    use: same bytes from both, for aligned and unaligned frame bases (the fused kernel stages with TMA only when the frame
    starts on a 16-byte boundary), odd blocksizes and every level without loose mid/side."""
    from pyflac_b200 import _native as nat
    xs = [corpus_signal(kind, 4096 * 2 + 1000 + 3 * i, 2, 16, seed=40 + i) for i, kind in enumerate(CORPUS_KINDS)]
    xs += [music_like(n, 2, 48000, 16, seed=n) for n in (1, 3, 17, 4097, 4096 * 3)]
    for level, bs in [(0, 0), (2, 0), (3, 0), (5, 0), (5, 1000), (6, 0), (8, 0), (8, 1152), (7, 4608), (5, 16)]:
        monkeypatch.delenv("FLACB200_NO_FUSED", raising=False)
        e1 = nat.Engine(0)
        a, oa = nat.encode_streams(e1, xs, 44100, 16, level, bs)
        launches_fused = e1.launch_count
        e1.close()
        monkeypatch.setenv("FLACB200_NO_FUSED", "1")
        e2 = nat.Engine(0)
        b, ob = nat.encode_streams(e2, xs, 44100, 16, level, bs)
        assert e2.launch_count > launches_fused            # the other path really ran
        e2.close()
        assert a == b, (level, bs)
        assert oa["log_guard_hits"] == 0 and ob["log_guard_hits"] == 0
    monkeypatch.delenv("FLACB200_NO_FUSED", raising=False)


def test_encode_vs_reference_binary_live(eng, checkers):
    """same-run comparison with the reference binary itself when oracle/_ref travelled to this box"""
    if not checkers.ref_available():
        pytest.skip("oracle/_ref not present")
    from pyflac_b200 import _native as nat
    xs = [music_like(4096 * 5 + 333, 2, 48000, 16, seed=100 + s) for s in range(16)]
    for level in (5, 8):
        got, _ = nat.encode_streams(eng, xs, 48000, 16, level, 4096)
        for x, g in zip(xs, got):
            assert g == checkers.ref_encode(x, 48000, 16, level, 4096)


def test_full_size_properties(eng, checkers):
    """BASELINE configs[1] shape (256 x 480000 x 2 int16, L5): size-independent properties on the whole batch +
    oracle decode round trip and byte equality on a sample of streams."""
    from pyflac_b200 import _native as nat
    import bench
    pcm = bench.make_pcm(0)
    cfg = nat.Engine.make_config(48000, 2, 16, 5, 4096)
    off = np.arange(256, dtype=np.uint64) * np.uint64(480000 * 2)
    eng.encode_host(cfg, pcm.reshape(-1), off, np.full(256, 480000, np.uint64))
    out = eng.fetch()
    assert out["log_guard_hits"] == 0
    assert len(out["frame_len"]) == 256 * 118
    arena = out["arena"]
    for s, si in enumerate(out["streams"]):
        blob = arena[int(si.byte_off): int(si.byte_off + si.byte_len)]
        assert bytes(blob[:4]) == b"fLaC"
        assert si.n_frames == 118 and si.total_samples == 480000
        # checksum of checksums: STREAMINFO MD5 == MD5 of the little-endian PCM
        assert bytes(si.md5) == hashlib.md5(pcm[s].tobytes()).digest()
        assert bytes(blob[26:42]) == bytes(si.md5)
    fo, fl = out["frame_off"], out["frame_len"]
    assert np.all(fo[1:] >= fo[:-1] + fl[:-1])                      # frames laid out in order without overlap
    assert np.all(arena[fo.astype(np.int64)] == 0xFF)               # every frame starts with the sync code
    blobs = [arena[int(si.byte_off): int(si.byte_off + si.byte_len)].tobytes() for si in out["streams"]]
    # byte equality of EVERY stream (all 256 are distinct): against the reference binary when it travelled to this box
    # (pthreads, seconds), else against the oracle port
    if checkers.ref_available():
        _, _, ref_blobs = checkers.ref_encode_mt(pcm, 48000, 16, 5, 4096, min(32, os.cpu_count() or 8), keep_bytes=True)
        for s in range(256):
            assert blobs[s] == ref_blobs[s].tobytes(), s
    else:
        for s in range(256):
            assert blobs[s] == checkers.oracle_encode(pcm[s], 48000, 16, 5, 4096), s
    for s in (0, 17, 255):                                          # encode -> oracle decode round trip + the oracle port's bytes
        dec, info = checkers.oracle_decode(blobs[s])                # validates every CRC-8/CRC-16 and the MD5
        assert np.array_equal(dec, pcm[s].astype(np.int32))
        assert blobs[s] == checkers.oracle_encode(pcm[s], 48000, 16, 5, 4096)


def test_full_size_config3_24bit_mono_level8(eng, checkers):
    """BASELINE configs[2] shape: 4096 streams x 262144 mono 24-bit (int32 container), 192 kHz, level 8.
    64 distinct streams tiled 64x (identical streams must give identical bytes); properties on the whole batch,
    byte equality with the oracle + decode round trip on samples."""
    from pyflac_b200 import _native as nat
    uniq = [music_like(262144, 1, 192000, 24, seed=900 + s) for s in range(64)]
    pcm = np.concatenate([u.reshape(-1) for u in uniq] * 64)
    n_streams = 4096
    off = np.arange(n_streams, dtype=np.uint64) * np.uint64(262144)
    cfg = nat.Engine.make_config(192000, 1, 24, 8, 4096, container_bytes=4)
    eng.encode_host(cfg, pcm, off, np.full(n_streams, 262144, np.uint64))
    out = eng.fetch()
    assert out["log_guard_hits"] == 0
    assert len(out["frame_len"]) == n_streams * 64
    arena = out["arena"]
    blobs = [arena[int(si.byte_off): int(si.byte_off + si.byte_len)] for si in out["streams"]]
    for s in range(64):                                                   # all 64 tiles of each distinct stream are identical
        for k in range(1, 64):
            assert np.array_equal(blobs[s], blobs[s + 64 * k]), (s, k)
    for s in range(64):                                                   # every distinct stream byte for byte against the oracle
        b = blobs[s].tobytes()
        assert b == checkers.oracle_encode(uniq[s], 192000, 24, 8, 4096), s
    for s in (0, 31, 63):
        dec, info = checkers.oracle_decode(blobs[s].tobytes())
        assert info["bps"] == 24 and np.array_equal(dec, uniq[s].astype(np.int32))
    import hashlib
    for s in (5, 4095):
        x = uniq[s % 64].astype("<i4").tobytes()
        le24 = b"".join(x[i:i + 3] for i in range(0, len(x), 4))
        assert bytes(out["streams"][s].md5) == hashlib.md5(le24).digest()


@pytest.mark.parametrize("bps,ch", [(16, 2), (24, 1), (24, 2), (32, 2), (8, 1), (12, 2), (20, 1)])
def test_md5_ragged_batch(eng, bps, ch):
    """md5_kernel: 40 streams of one batch with lengths around every boundary of its staging (0 whole 256-byte pieces, one,
    many; tails of 0..63 bytes; more than a warp of streams), odd element offsets that break the 16-byte alignment of some
    streams, each sample size in its container.  The digest must be MD5 of the samples as little-endian (bps+7)/8-byte values."""
    from pyflac_b200 import _native as nat
    rng = np.random.default_rng(bps * 10 + ch)
    dt = np.int16 if bps <= 16 else np.int32
    lens = [1, 2, 15, 16, 21, 22, 31, 32, 33, 63, 64, 65, 127, 128, 129, 255, 256, 257, 1000, 4096, 4097, 5000, 8192, 12345,
            20000, 3, 64, 640, 6400, 30001, 17, 170, 1700, 17000, 99, 999, 9999, 4095, 8191, 16383]
    xs = [rng.integers(-(1 << (bps - 1)), 1 << (bps - 1), size=(n, ch), dtype=np.int64).astype(dt) for n in lens]
    pad = [0, 1, 3, 0, 5, 0, 0, 7] * 5                           # elements of slack before each stream: some starts are misaligned
    parts, offs, pos = [], [], 0
    for x, g in zip(xs, pad):
        parts.append(np.zeros(g, dt)); pos += g
        offs.append(pos); parts.append(x.reshape(-1)); pos += x.size
    flat = np.concatenate(parts)
    cfg = nat.Engine.make_config(48000, ch, bps, 2, 0)
    eng.encode_host(cfg, flat, np.array(offs, np.uint64), np.array(lens, np.uint64))
    out = eng.fetch()
    nb = (bps + 7) // 8
    for s, x in enumerate(xs):
        le = x.reshape(-1).astype("<i4").view(np.uint8).reshape(-1, 4)[:, :nb].tobytes()
        assert bytes(out["streams"][s].md5) == hashlib.md5(le).digest(), (s, lens[s])


@pytest.mark.parametrize("bps", [16, 24])
def test_host_path_md5_placement(eng, checkers, monkeypatch, bps):
    """flacb200_encode_batch_host: whoever hashes a stream -- host threads, md5_kernel as the chunk lands in HBM, or both -- the
    images (frames, index, STREAMINFO with its MD5) equal the reference's.  The split is forced through the environment knobs."""
    from pyflac_b200 import _native as nat
    ch = 2 if bps == 16 else 1
    xs = [music_like(4096 * 3 + 517 * s, ch, 48000, bps, seed=700 + s) for s in range(30)]
    dt = np.int16 if bps == 16 else np.int32
    flat = np.concatenate([x.reshape(-1) for x in xs]).astype(dt)
    sizes = np.array([x.size for x in xs], np.uint64)
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
    cfg = nat.Engine.make_config(48000, ch, bps, 5, 4096)
    want = [checkers.oracle_encode(x, 48000, bps, 5, 4096) for x in xs]
    for gpu_chunks, threads in [("0", "1"), ("0", "3"), ("5", "2"), ("12", "1"), (None, None)]:
        for k, v in (("FLACB200_MD5_GPU_CHUNKS", gpu_chunks), ("FLACB200_MD5_THREADS", threads)):
            if v is None:
                monkeypatch.delenv(k, raising=False)
            else:
                monkeypatch.setenv(k, v)
        for rep in range(2):
            out = eng.encode_host_to_host(cfg, flat, offs, sizes // ch)
            pi = out["path_info"]
            if gpu_chunks == "0":
                assert pi["streams_hashed_on_gpu"] == 0
            if gpu_chunks == "12":
                assert pi["streams_hashed_on_gpu"] == 30 and pi["gpu_md5_done_ms"] > 0
            if gpu_chunks == "5":
                assert 0 < pi["streams_hashed_on_gpu"] < 30
            for s, x in enumerate(xs):
                si = out["streams"][s]
                blob = out["arena"][int(si.byte_off): int(si.byte_off + si.byte_len)].tobytes()
                assert blob == want[s], (gpu_chunks, threads, s)
                nb = (bps + 7) // 8
                le = x.reshape(-1).astype("<i4").view(np.uint8).reshape(-1, 4)[:, :nb].tobytes()
                assert bytes(si.md5) == hashlib.md5(le).digest()


@pytest.mark.parametrize("bps", [16, 24, 12])
def test_host_pipelined_submit_collect(eng, checkers, bps):
    """flacb200_encode_host_submit / _collect: seven batches of one layout and different audio, three in flight; every collected
    batch (images, index, digests from md5_kernel; host-hashed for the 12-bit-in-int16 case) equals the reference's output."""
    from pyflac_b200 import _native as nat
    ch = 2 if bps != 24 else 1
    dt = np.int16 if bps <= 16 else np.int32
    lens = [4096 * 2 + 311 * s for s in range(20)]
    sizes = np.array([n * ch for n in lens], np.uint64)
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
    cfg = nat.Engine.make_config(48000, ch, bps, 5, 4096)
    batches = [[music_like(n, ch, 48000, bps, seed=1000 * b + s) for s, n in enumerate(lens)] for b in range(7)]
    flats = [np.concatenate([x.reshape(-1) for x in xs]).astype(dt) for xs in batches]
    inflight, done = [], []
    for b in range(7):
        if len(inflight) == 3:
            done.append(eng.collect_host(inflight.pop(0)))
        inflight.append(eng.submit_host(cfg, flats[b], offs, sizes // ch))
    with pytest.raises(nat.NativeError):                               # a synchronous host call must not run over batches in flight
        eng.encode_host_to_host(cfg, flats[0], offs, sizes // ch)
    while inflight:
        done.append(eng.collect_host(inflight.pop(0)))
    nb = (bps + 7) // 8
    for b, out in enumerate(done):
        for s, x in enumerate(batches[b]):
            si = out["streams"][s]
            blob = out["arena"][int(si.byte_off): int(si.byte_off + si.byte_len)].tobytes()
            assert blob == checkers.oracle_encode(x, 48000, bps, 5, 4096), (b, s)
            le = x.reshape(-1).astype("<i4").view(np.uint8).reshape(-1, 4)[:, :nb].tobytes()
            assert bytes(si.md5) == hashlib.md5(le).digest()
    # and the synchronous call works again afterwards
    out = eng.encode_host_to_host(cfg, flats[1], offs, sizes // ch)
    si = out["streams"][3]
    assert out["arena"][int(si.byte_off): int(si.byte_off + si.byte_len)].tobytes() == checkers.oracle_encode(batches[1][3], 48000, bps, 5, 4096)


def test_fetch_md5_of_the_previous_batch(eng):
    """flacb200_encode_fetch_md5_back: while batch B's chain may still run, the digests of batch A (same layout) are available"""
    from pyflac_b200 import _native as nat
    lens = [4096 * 6 + 17 * s for s in range(12)]
    sizes = np.array([2 * n for n in lens], np.uint64)
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
    cfg = nat.Engine.make_config(48000, 2, 16, 5, 0)
    A = [music_like(n, 2, 48000, 16, seed=40 + s) for s, n in enumerate(lens)]
    B = [music_like(n, 2, 48000, 16, seed=90 + s) for s, n in enumerate(lens)]
    eng.encode_host(cfg, np.concatenate([x.reshape(-1) for x in A]), offs, sizes // 2)
    eng.result(wait_md5=False)
    eng.encode_host(cfg, np.concatenate([x.reshape(-1) for x in B]), offs, sizes // 2)
    eng.result(wait_md5=False)
    da, db = eng.fetch_md5_back(1), eng.fetch_md5()
    for s in range(len(lens)):
        assert bytes(da[s]) == hashlib.md5(A[s].tobytes()).digest()
        assert bytes(db[s]) == hashlib.md5(B[s].tobytes()).digest()


@pytest.mark.parametrize("path", ["device", "host_call", "submit_collect"])
@pytest.mark.parametrize("level,bps,ch", [(5, 16, 2), (8, 24, 1)])
def test_log_guard_host_redecision(checkers, path, level, bps, ch):
    """decisions that depend on libm log() inside the guard band are repeated on the host.  With the band widened to
    1e-2 many decisions are logged and the host confirms them (same bytes as the reference); with the kernels told to take the
    runner-up inside the band the host overrides every wrong decision and the second pass still yields the reference's bytes."""
    from pyflac_b200 import _native as nat
    eng = nat.Engine(0)
    xs = [music_like(4096 * 3 + 100 * s, ch, 48000, bps, seed=500 + s) for s in range(6)]
    dt = np.int16 if bps <= 16 else np.int32
    flat = np.concatenate([x.reshape(-1) for x in xs]).astype(dt)
    sizes = np.array([x.size for x in xs], np.uint64)
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
    cfg = nat.Engine.make_config(48000, ch, bps, level, 4096)
    want = [checkers.oracle_encode(x, 48000, bps, level, 4096) for x in xs]

    def run():
        if path == "device":
            eng.encode_host(cfg, flat, offs, sizes // ch)
            out = eng.fetch()
        elif path == "host_call":
            out = eng.encode_host_to_host(cfg, flat, offs, sizes // ch)
        else:
            out = eng.collect_host(eng.submit_host(cfg, flat, offs, sizes // ch))
        return [out["arena"][int(si.byte_off): int(si.byte_off + si.byte_len)].tobytes() for si in out["streams"]]

    assert run() == want and eng.log_guard_info()["in_band"] == 0       # the real band: nothing inside it
    eng.set_log_guard(1e-2, flip=False)
    assert run() == want
    gi = eng.log_guard_info()
    assert gi["in_band"] > 0 and gi["confirmed"] == gi["in_band"] - gi["unchecked"] and gi["overridden"] == 0, gi
    eng.set_log_guard(1e-2, flip=True)
    assert run() == want
    gi = eng.log_guard_info()
    assert gi["overridden"] > 0 and gi["unchecked"] == 0, gi
    eng.set_log_guard(1e-12, flip=False)
    assert run() == want
