"""CPU tests: pin the plain-C restatement (oracle/flac_oracle.c) against
 (1) the committed golden vectors produced by the reference's libFLAC 1.4.3 binary, and
 (2) the reference binary itself, live, when oracle/_ref is present,
 (3) the reference's own decode fixtures tests/data/*.flac when /root/reference exists."""
import glob
import hashlib
import os

import numpy as np
import pytest

from conftest import golden_cases, load_golden
from pyflac_b200.synth import corpus_signal, CORPUS_KINDS

CASES = golden_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_encode_matches_golden(checkers, case):
    x, flac = load_golden(case)
    got, off, ln = checkers.oracle_encode(x, case["sample_rate"], case["bps"], case["level"], case["blocksize"],
                                          with_index=True)
    assert got == flac
    assert [int(v) for v in off] == case["frame_off"]
    assert [int(v) for v in ln] == case["frame_len"]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_decode_matches_golden(checkers, case):
    x, flac = load_golden(case)
    pcm, info = checkers.oracle_decode(flac)
    assert info["channels"] == case["channels"] and info["bps"] == case["bps"]
    assert info["sample_rate"] == case["sample_rate"]
    assert np.array_equal(pcm, x.astype(np.int32).reshape(pcm.shape))


def test_streaminfo_md5_is_md5_of_le_samples(checkers):
    case = CASES[0]
    x, flac = load_golden(case)
    assert flac[26:42] == hashlib.md5(np.ascontiguousarray(x).astype("<i2").tobytes()).digest()


def test_stream_mode_leaves_streaminfo_zero(checkers):
    x, flac = load_golden(CASES[0])
    got = checkers.oracle_encode(x, CASES[0]["sample_rate"], 16, 5, 0, seekable=False)
    assert got[:12] == flac[:12] and got[12:18] == bytes(6) and got[18:21] == flac[18:21]
    assert got[21] & 0x0F == 0 and got[22:42] == bytes(20)
    assert got[42:] == flac[42:]


def test_crc_and_md5_helpers(checkers):
    L = checkers.oracle_lib()
    data = np.frombuffer(b"123456789", np.uint8)
    assert L.fo_crc8(data.ctypes.data, 9) == 0xF4          # CRC-8 poly 0x07 check value
    assert L.fo_crc16(data.ctypes.data, 9) == 0xFEE8       # CRC-16/BUYPASS (poly 0x8005, init 0) check value
    d = np.zeros(16, np.uint8)
    L.fo_md5(data.ctypes.data, 9, d.ctypes.data)
    assert d.tobytes() == hashlib.md5(b"123456789").digest()


needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(__file__), "..", "oracle", "_ref",
                                                               "libflacref.so")), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("level", [0, 1, 2, 3, 4, 5, 6, 7, 8])
def test_oracle_vs_reference_binary_levels(checkers, level):
    for kind in CORPUS_KINDS:
        for ch, bps, n, bs, sr in [(2, 16, 4096 * 2 + 768, 0, 48000), (1, 24, 4096 + 100, 4096, 192000)]:
            x = corpus_signal(kind, n, ch, bps, seed=level * 7 + ch)
            assert checkers.oracle_encode(x, sr, bps, level, bs) == checkers.ref_encode(x, sr, bps, level, bs), \
                (level, kind, ch, bps)


@needs_ref
def test_oracle_vs_reference_binary_edges(checkers):
    for bs in [16, 17, 100, 192, 576, 1000, 4608]:
        for n in [bs * 2 + 3, bs + 1, 5, 1, 4]:
            x = corpus_signal("music", n, 2, 16, seed=n)
            assert checkers.oracle_encode(x, 44100, 16, 5, bs) == checkers.ref_encode(x, 44100, 16, 5, bs), (bs, n)
    for sr in [0, 100, 12345, 50001, 65535, 65536, 255000, 256000, 655350, 655351, 1048575]:
        x = corpus_signal("music", 3000, 2, 16, seed=1)
        assert checkers.oracle_encode(x, sr, 16, 5, 1024, streamable_subset=False) == \
            checkers.ref_encode(x, sr, 16, 5, 1024, streamable_subset=False), sr
    for ch in [3, 6, 8]:
        x = corpus_signal("lr_uncorr", 5000, ch, 16, seed=ch)
        assert checkers.oracle_encode(x, 48000, 16, 5, 0) == checkers.ref_encode(x, 48000, 16, 5, 0), ch
    for kind in ["silence", "dc", "mixed"]:
        x = corpus_signal(kind, 9000, 2, 16, seed=3)
        assert checkers.oracle_encode(x, 48000, 16, 5, 0, limit_min_bitrate=True) == \
            checkers.ref_encode(x, 48000, 16, 5, 0, limit_min_bitrate=True), kind
    x = corpus_signal("music", 20000, 2, 16, seed=3)
    assert checkers.oracle_encode(x, 48000, 16, 5, 0, seekable=False) == \
        checkers.ref_encode(x, 48000, 16, 5, 0, seekable=False, chunk=777)


@needs_ref
def test_oracle_vs_reference_binary_random_settings(checkers):
    """seeded random walk over depth / channels / level / blocksize / length / signal kind / flags: the restatement and the
    bundled binary must agree byte for byte (encode) and the oracle must decode the binary's bytes back to the input"""
    rng = np.random.default_rng(20240607)
    for it in range(48):
        bps = int(rng.choice([4, 8, 12, 13, 16, 16, 16, 20, 24, 24]))
        ch = int(rng.choice([1, 1, 2, 2, 2, 3, 5, 8]))
        level = int(rng.integers(0, 9))
        bs = int(rng.choice([0, 0, 16, 64, 192, 255, 256, 1000, 1152, 2048, 4096, 4608]))
        n = int(rng.choice([1, 7, 100, 1153, 3000, 4096, 4097, 9000, 12288]))
        kind = str(rng.choice(CORPUS_KINDS))
        lmb = bool(rng.integers(0, 2))
        sr = int(rng.choice([8000, 22050, 44100, 48000]))
        subset = bps in (8, 12, 16, 20, 24)
        x = corpus_signal(kind, n, ch, bps, seed=1000 + it, sample_rate=sr)
        a = checkers.oracle_encode(x, sr, bps, level, bs, limit_min_bitrate=lmb, streamable_subset=subset)
        b = checkers.ref_encode(x, sr, bps, level, bs, limit_min_bitrate=lmb, streamable_subset=subset)
        assert a == b, (it, bps, ch, level, bs, n, kind, lmb, sr)
        dec, info = checkers.oracle_decode(b)
        assert info["bps"] == bps and np.array_equal(dec.reshape(n, ch), np.asarray(x).reshape(n, ch)), (it, "decode")


@needs_ref
def test_oracle_vs_reference_binary_32bit(checkers):
    """32-bit input: the _limit_residual predictor search as the shipped binary runs it (all-zero block CONSTANT, other
    constant blocks FIXED order 1, invalid orders when a residual leaves int32) and the 33-bit side channel of stereo
    (all-zero side reports one wasted bit).  Lengths 0 or 1 mod 4."""
    if not checkers.ref_available():
        pytest.skip("oracle/_ref not present")
    from pyflac_b200.synth import music_like
    rng = np.random.default_rng(12)
    n = 4096 + 400
    m = music_like(n, 2, 48000, 24, seed=6).astype(np.int64)
    L, R = m[:, 0] * 200, m[:, 1] * 180
    spikes = L.copy(); spikes[rng.integers(0, n, 4)] = -2**31
    mono = {"zeros": np.zeros(n, np.int64), "dc_odd": np.full(n, 7654321), "dc_min": np.full(n, -2**31), "music": L, "shl8": m[:, 0] * 256,
            "noise": rng.integers(-2**31, 2**31, n), "spikes": spikes, "minmax": np.where(np.arange(n) % 2 == 0, -2**31, 2**31 - 1)}
    def c32(a):
        return np.clip(a, -2**31, 2**31 - 1).astype(np.int32)
    for name, v in mono.items():
        x = c32(v)[:, None]
        for level in (0, 5, 8):
            assert checkers.oracle_encode(x, 48000, 32, level, 0) == checkers.ref_encode(x, 48000, 32, level, 0), (name, level)
    stereo = {"same": (L, L), "anti": (L, -L), "anti_noisy": (L, -L + rng.integers(-3, 4, n)), "indep": (L, R), "rails": (np.full(n, 2**31 - 1), np.full(n, -2**31)),
              "noise": (rng.integers(-2**31, 2**31, n), rng.integers(-2**31, 2**31, n)), "near": (L, L + rng.integers(-500, 500, n))}
    for name, (a, b) in stereo.items():
        x = c32(np.stack([a, b], axis=1))
        for level in (1, 3, 5, 8):
            assert checkers.oracle_encode(x, 44100, 32, level, 1152 if level == 1 else 0) == checkers.ref_encode(x, 44100, 32, level, 1152 if level == 1 else 0), (name, level)
        dec, _ = checkers.oracle_decode(checkers.ref_encode(x, 44100, 32, 5, 0))
        assert np.array_equal(dec, x)


def test_init_status_matches_reference(checkers):
    """Init validation order (reference tests/test_encoder.py:139-164,202-207)."""
    import ctypes as C
    L = checkers.oracle_lib()
    probes = [dict(sample_rate=2000000), dict(blocksize=1000000), dict(blocksize=65535), dict(channels=9),
              dict(bps=3), dict(bps=33), dict(bps=17), dict(blocksize=15), dict(blocksize=4609),
              dict(sample_rate=96000, blocksize=16385), dict(blocksize=8, level=0), dict()]
    for p in probes:
        for subset in (0, 1):
            kw = dict(sample_rate=48000, channels=2, bps=16, level=5, blocksize=0)
            kw.update(p)
            cfg = checkers.FoEncCfg(kw["sample_rate"], kw["channels"], kw["bps"], kw["level"], kw["blocksize"], 0, 0, subset)
            got = L.fo_encoder_init_status(C.byref(cfg), 1, 0, 0)
            rcfg = checkers.RefEncCfg(kw["sample_rate"], kw["channels"], kw["bps"], kw["level"], kw["blocksize"], 0, 0, subset, 1)
            out = np.zeros(1 << 16, np.uint8)
            x = np.zeros(kw["channels"] * 4, np.int32)
            r = checkers.ref_lib().ref_encode_stream(C.byref(rcfg), x.ctypes.data, 0, 0, out.ctypes.data, out.size,
                                                     None, None, None, 0, None)
            assert (got == 0) == (r >= 0), (p, subset, got, r)


def test_oracle_on_imported_reference_fixtures(checkers):
    """The committed copies of pyFLAC's tests/data/*.flac (tests/golden/fixtures): the oracle's decode hashes to the
    STREAMINFO MD5 == MD5 of the matching .wav, and its level-5 encode of that PCM is the file the bundled libFLAC wrote."""
    from conftest import fixture_cases, fixture_path, pcm_md5
    for c in fixture_cases():
        data = open(fixture_path(c["name"] + ".flac"), "rb").read()
        assert len(data) == c["flac_bytes"] and data[26:42].hex() == c["streaminfo_md5"]
        pcm, info = checkers.oracle_decode(data)
        assert pcm.shape == (c["samples"], c["channels"]) and info["bps"] == c["bps"] and info["sample_rate"] == c["sample_rate"]
        assert pcm_md5(pcm, c["bps"]) == c["streaminfo_md5"] == c["pcm_md5"]
        if c["level5"]:
            assert c["wav_md5"] == c["pcm_md5"]
            want = open(fixture_path(c["level5"]["file"]), "rb").read()
            x = pcm.astype(np.int16 if c["bps"] <= 16 else np.int32)
            assert checkers.oracle_encode(x, c["sample_rate"], c["bps"], 5, 0) == want, c["name"]


def test_oracle_decodes_foreign_streams(checkers):
    """tests/golden/foreign: predictor orders up to 32, partition order 8, 9-sample partitions, 16-sample blocks (libFLAC with
    tuned settings) and hand-written streams with escape partitions / 5-bit Rice parameters -- each was accepted by the
    bundled libFLAC decoder when it was made; the oracle must hash to the STREAMINFO MD5 and, when the binary is here, equal it."""
    from conftest import foreign_cases, foreign_path, pcm_md5
    for c in foreign_cases():
        data = open(foreign_path(c["name"] + ".flac"), "rb").read()
        pcm, info = checkers.oracle_decode(data)
        assert pcm.shape == (c["samples"], c["channels"]) and info["bps"] == c["bps"]
        assert pcm_md5(pcm, c["bps"]) == c["pcm_md5"] == data[26:42].hex(), c["name"]
        if checkers.ref_available():
            ref, rinfo = checkers.ref_decode(data)
            assert rinfo["errors"] == 0 and np.array_equal(ref, pcm), c["name"]


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests/data"), reason="reference fixtures not present")
def test_oracle_decodes_reference_fixtures(checkers):
    """tests/data/{mono,stereo,surround,32bit}.flac <-> .wav pairs: STREAMINFO MD5 == MD5 of decoded PCM."""
    for f in sorted(glob.glob("/root/reference/tests/data/*.flac")):
        data = open(f, "rb").read()
        pcm, info = checkers.oracle_decode(data)        # raises on CRC / MD5 mismatch
        width = (info["bps"] + 7) // 8
        raw = pcm.astype("<i4").tobytes()
        le = b"".join(raw[i:i + width] for i in range(0, len(raw), 4)) if width != 4 else raw
        assert hashlib.md5(le).digest() == data[26:42], f


def test_oracle_frames_of_256_kib_and_more(checkers):
    """the stream behind tests/test_gpu_decode.py::test_decode_frames_of_256_kib_and_more: 16384 samples x 8 channels of
    incompressible noise go out as VERBATIM subframes, 262 158 bytes per frame; the restatement round-trips it and (where the
    reference binary is present) writes the binary's bytes, and the binary decodes them."""
    rng = np.random.default_rng(3)
    x = rng.integers(-32768, 32768, (16384 * 2 + 100, 8)).astype(np.int16)
    blob = checkers.oracle_encode(x, 96000, 16, 5, 16384)
    assert len(blob) > 2 * 262144
    dec, info = checkers.oracle_decode(blob)
    assert info["bps"] == 16 and np.array_equal(dec.reshape(x.shape), x)
    if checkers.ref_available():
        assert checkers.ref_encode(x, 96000, 16, 5, 16384) == blob
        out = checkers.ref_decode(blob)
        pcm = out[0] if isinstance(out, tuple) else out
        assert np.array_equal(np.asarray(pcm).reshape(x.shape), x)
