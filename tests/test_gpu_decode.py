"""GPU decode parity tests (pytest -m gpu): CUDA decode path through the C ABI vs the oracle / golden PCM."""
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


def test_decode_golden_batch(eng):
    """all golden .flac files (libFLAC 1.4.3 output, levels 0-8, 8..24 bit, 1..6 ch) in one ragged batch -> bit-exact PCM"""
    from pyflac_b200 import _native as nat
    xs, blobs = zip(*[load_golden(c) for c in CASES])
    for grp in ([c for c in range(len(CASES)) if CASES[c]["bps"] <= 16], [c for c in range(len(CASES)) if CASES[c]["bps"] > 16]):
        out, infos = nat.decode_streams(eng, [blobs[i] for i in grp])
        for k, i in enumerate(grp):
            assert infos[k].status == 0, (CASES[i]["name"], nat.DEC_STATUS.get(infos[k].status))
            assert infos[k].channels == CASES[i]["channels"] and infos[k].bits_per_sample == CASES[i]["bps"]
            assert infos[k].n_frames == CASES[i]["frames"]
            x = xs[i].reshape(xs[i].shape[0], -1)
            assert np.array_equal(out[k].astype(np.int64), x.astype(np.int64)), CASES[i]["name"]


def test_decode_corpus_vs_oracle(eng, checkers):
    from pyflac_b200 import _native as nat
    for level in (0, 5, 8):
        blobs, xs = [], []
        for kind in CORPUS_KINDS:
            x = corpus_signal(kind, 4096 * 2 + 777, 2, 16, seed=level + 3)
            blobs.append(checkers.oracle_encode(x, 44100, 16, level, 0))
            xs.append(x)
        out, infos = nat.decode_streams(eng, blobs)
        for kind, x, o, si in zip(CORPUS_KINDS, xs, out, infos):
            assert si.status == 0, (kind, si.status)
            assert np.array_equal(o, x), (level, kind)


def test_decode_reference_binary_streams(eng, checkers):
    """streams produced live by the reference binary: escape-free Rice, all channel assignments, wasted bits, 24-bit"""
    if not checkers.ref_available():
        pytest.skip("oracle/_ref not present")
    from pyflac_b200 import _native as nat
    x24 = [corpus_signal(k, 5000, 1, 24, seed=9) for k in ("music", "wasted", "noise")]
    out, infos = nat.decode_streams(eng, [checkers.ref_encode(x, 192000, 24, 8, 4096) for x in x24])
    for x, o, si in zip(x24, out, infos):
        assert si.status == 0 and si.bits_per_sample == 24 and np.array_equal(o, x)
    x6 = corpus_signal("lr_uncorr", 9000, 6, 16, seed=2)
    out, infos = nat.decode_streams(eng, [checkers.ref_encode(x6, 48000, 16, 5, 0)])
    assert infos[0].status == 0 and np.array_equal(out[0], x6)


def test_decode_errors_are_reported(eng, checkers):
    from pyflac_b200 import _native as nat
    x = music_like(9000, 2, 48000, 16, seed=1)
    good = checkers.oracle_encode(x, 48000, 16, 5, 0)
    bad_crc = bytearray(good); bad_crc[len(good) // 2] ^= 0x10
    trunc = good[: len(good) - 1000]
    rng = np.random.default_rng(0)
    junk = rng.integers(0, 256, 100000).astype(np.uint8).tobytes()
    out, infos = nat.decode_streams(eng, [good, bytes(bad_crc), trunc, junk])
    assert infos[0].status == 0 and np.array_equal(out[0], x)
    assert infos[1].status != 0
    assert infos[2].status != 0 and infos[2].n_frames >= 1 and np.array_equal(out[2], x[: out[2].shape[0]])
    assert infos[3].status == 2


def test_decode_error_recovery_matches_libflac(eng, checkers):
    """Damaged streams through the drop-in StreamDecoder API of BOTH libraries (same ctypes driver, tests/_flacapi.py): the
    error-callback sequence, its place between the write callbacks, the PCM (a frame with a bad CRC is never delivered,
    missing frames come as silence of the previous blocksize, nothing is filled at the very start or end) and the final state
    must equal libFLAC 1.4.3's (/root/reference/pyflac/include/FLAC/stream_decoder.h:431-448,1440-1460 document the
    contract; the binary is the judge).  Read sizes of 8192 (libFLAC's own), 1000 and 1 MiB exercise the state carried from
    one buffered slice to the next."""
    import ctypes as C
    from _flacapi import decode_session
    from pyflac_b200 import _native as nat
    if not checkers.ref_available():
        pytest.skip("oracle/_ref not present")
    ours = nat.lib()
    ref = C.CDLL(os.path.join(os.path.dirname(checkers.REF_SO), "libFLAC-12.1.0.so"))
    x = music_like(4096 * 8 + 321, 2, 44100, 16, seed=8)
    data, off, ln, _ = checkers.ref_encode(x, 44100, 16, 5, 0, with_index=True)
    off = [int(o) for o in off]
    y = music_like(1152 * 6 + 100, 1, 48000, 16, seed=3)
    data2, off2, _, _ = checkers.ref_encode(y, 48000, 16, 0, 0, with_index=True)

    def flip(d, *positions):
        b = bytearray(d)
        for p in positions:
            b[p] ^= 1
        return bytes(b)
    cases = {
        "clean": data,
        "payload bit in frame 2": flip(data, off[2] + 100),
        "frames 2,3": flip(data, off[2] + 100, off[3] + 100),
        "frames 2,3,4": flip(data, off[2] + 100, off[3] + 100, off[4] + 100),
        "crc byte of frame 2": flip(data, off[3] - 1),
        "frame 2 removed": data[:off[2]] + data[off[3]:],
        "frames 2-4 removed": data[:off[2]] + data[off[5]:],
        "frames 0,1 removed": data[:off[0]] + data[off[2]:],
        "frame 1 repeated after frame 2": data[:off[3]] + data[off[1]:off[2]] + data[off[3]:],
        "last full frame": flip(data, off[7] + 100),
        "frames 6,7 then the short one": flip(data, off[6] + 100, off[7] + 100),
        "frames 7,8 (the end)": flip(data, off[7] + 100, off[8] + 10),
        "final short frame": flip(data, off[8] + 20),
        "first frame": flip(data, off[0] + 50),
        "frames 0 and 3": flip(data, off[0] + 50, off[3] + 60),
        "199 junk bytes before frame 3": data[:off[3]] + bytes(range(1, 200)) + data[off[3]:],
        "truncated inside frame 2": data[:off[2] + 300],
        "truncated after frame 0": data[:off[1]],
        "one byte short": data[:-1],
        "trailing zeros": data + bytes(1000),
        "trailing junk": data + b"ID3 junk" * 50,
        "mono 1152 level 0, frames 1,2": flip(data2, int(off2[1]) + 50, int(off2[2]) + 50),
    }
    for name, blob in cases.items():
        b = decode_session(ref, blob, 8192)
        for chunk in (8192, 1000, 1 << 20):
            a = decode_session(ours, blob, chunk)
            assert a["events"] == b["events"], (name, chunk, a["events"], b["events"])
            assert np.array_equal(a["pcm"], b["pcm"]), (name, chunk)
            assert a["ok"] == b["ok"] and a["state"] == b["state"], (name, chunk)
    # a flipped high bit usually derails the parse (reserved values, runaway unary codes): the kinds and count of errors then
    # depend on false sync codes inside the damaged frame; what must hold is what was delivered
    for k in (300, 1000, 5000):
        b2 = bytearray(data); b2[off[2] + k] ^= 0x80
        a, b = decode_session(ours, bytes(b2)), decode_session(ref, bytes(b2))
        assert a["errors"] and np.array_equal(a["pcm"], b["pcm"]), k
    # the batch ABI sees the same PCM, the first problem as status and the event log
    out, infos = nat.decode_streams(eng, [cases["payload bit in frame 2"], cases["frame 2 removed"], cases["clean"]])
    want = decode_session(ref, cases["payload bit in frame 2"])
    assert infos[0].status == 7 and np.array_equal(out[0], want["pcm"]) and infos[0].gap_samples == 4096
    assert [int(v) for v in infos[0].ev_status[:infos[0].n_events]] == [2, 0] and [int(v) for v in infos[0].ev_frame[:2]] == [2, 2]
    assert infos[1].status == 0 and infos[1].gap_samples == 4096 and infos[1].n_events == 0 and out[1].shape[0] == len(x)
    assert infos[2].status == 0 and infos[2].n_events == 0 and np.array_equal(out[2], x)


def test_encode_decode_roundtrip_full_size(eng):
    """encode -> decode on the GPU for a BASELINE configs[3]-shaped slice (stereo s16 L5 streams), bit-exact PCM"""
    from pyflac_b200 import _native as nat
    xs = [music_like(131072, 2, 48000, 16, seed=50 + s) for s in range(64)]
    blobs, _ = nat.encode_streams(eng, xs, 48000, 16, 5, 4096)
    out, infos = nat.decode_streams(eng, blobs)
    for x, o, si in zip(xs, out, infos):
        assert si.status == 0 and si.n_frames == 32 and np.array_equal(o, x)


def test_decode_32bit(eng, checkers):
    """32-bit streams (mono / multi-channel / independent stereo): bit-exact PCM, incl. libFLAC-made streams"""
    from pyflac_b200 import _native as nat
    rng = np.random.default_rng(4)
    n = 4096 + 400
    m = music_like(n, 2, 48000, 24, seed=8).astype(np.int64)
    xs = [np.clip(m * 180, -2**31, 2**31 - 1).astype(np.int32), rng.integers(-2**31, 2**31, (n, 2)).astype(np.int32),
          np.zeros((n, 2), np.int32), (m * 256).astype(np.int32)]
    blobs, _ = nat.encode_streams(eng, xs, 48000, 32, 3, 0)
    out, infos = nat.decode_streams(eng, blobs)
    for x, o, si in zip(xs, out, infos):
        assert si.status == 0 and si.bits_per_sample == 32 and np.array_equal(o, x)
    # mid/side levels: the side channel has 33 bits (all four channel assignments, with and without wasted bits)
    L = m[:, 0] * 200
    pairs = [(L, L.copy()), (L, -L + rng.integers(-3, 4, n)), (L, -L), (L, L + rng.integers(-1000, 1000, n)), (L, m[:, 1] * 180),
             (rng.integers(-2**31, 2**31, n), rng.integers(-2**31, 2**31, n)), (np.full(n, 2**31 - 1), np.full(n, -2**31)), (L * 2, -L * 2 + 4)]
    ys = [np.clip(np.stack([a, b], axis=1), -2**31, 2**31 - 1).astype(np.int32) for a, b in pairs]
    for level in (2, 5, 8):
        blobs = [checkers.oracle_encode(y, 48000, 32, level, 0) for y in ys]
        out, infos = nat.decode_streams(eng, blobs)
        for y, o, si in zip(ys, out, infos):
            assert si.status == 0 and np.array_equal(o, y), level
    if checkers.ref_available():
        mono = [x[:, :1].copy() for x in xs]
        rb = [checkers.ref_encode(x, 48000, 32, 5, 0) for x in mono] + [checkers.ref_encode(y, 48000, 32, 5, 0) for y in ys]
        out, infos = nat.decode_streams(eng, rb)
        for x, o, si in zip(mono + ys, out, infos):
            assert si.status == 0 and np.array_equal(o, x)


def test_decode_host_pipelined_matches_plain(eng, checkers, monkeypatch):
    """flacb200_decode_batch_host (chunked H2D / decode / D2H pipeline) == the plain batch path, ragged streams,
    mono + stereo, fewer streams than chunks, a corrupted stream in the middle"""
    from pyflac_b200 import _native as nat
    monkeypatch.setenv("FLACB200_DEC_CHUNKS", "5")           # force the multi-chunk pipeline on these small batches
    for ch, n_streams in [(2, 40), (1, 5), (2, 1)]:
        xs = [music_like(4096 * (1 + s % 5) + 37 * s, ch, 48000, 16, seed=90 + s) for s in range(n_streams)]
        blobs, _ = nat.encode_streams(eng, xs, 48000, 16, 5, 4096)
        blobs = [bytearray(b) for b in blobs]
        if n_streams > 10:
            blobs[7][len(blobs[7]) // 2] ^= 0x55                      # CRC mismatch somewhere inside stream 7
        blob = np.frombuffer(b"".join(bytes(b) for b in blobs), np.uint8)
        sl = np.array([len(b) for b in blobs], np.uint64)
        so = np.concatenate([[0], np.cumsum(sl)[:-1]]).astype(np.uint64)
        ref_out, ref_infos = nat.decode_streams(eng, [bytes(b) for b in blobs])
        out = np.zeros(sum(x.size for x in xs) + 16, np.int16)
        tot, infos = eng.decode_host_pipelined(blob, so, sl, out, 2)
        for s in range(n_streams):
            assert infos[s].status == ref_infos[s].status, s
            assert infos[s].total_samples == ref_infos[s].total_samples and infos[s].n_frames == ref_infos[s].n_frames
            got = out[int(infos[s].pcm_off): int(infos[s].pcm_off + infos[s].total_samples * ch)].reshape(-1, ch)
            assert np.array_equal(got, ref_out[s].reshape(-1, ch)), s
        assert tot == sum(int(i.total_samples) * ch for i in list(infos)[:n_streams])


def test_full_size_config4_decode_4096_streams(eng):
    """BASELINE configs[3] shape: 4096 stereo s16 streams of 131072 samples -> int16 PCM (64 distinct, tiled)."""
    from pyflac_b200 import _native as nat
    uniq = [music_like(131072, 2, 48000, 16, seed=700 + s) for s in range(64)]
    blobs64, _ = nat.encode_streams(eng, uniq, 48000, 16, 5, 4096)
    blobs = blobs64 * 64
    out, infos = nat.decode_streams(eng, blobs, 2)
    assert len(out) == 4096
    assert all(si.status == 0 and si.n_frames == 32 and si.total_samples == 131072 for si in infos)
    for s in range(0, 4096, 97):
        assert np.array_equal(out[s], uniq[s % 64])
    total = sum(int(si.total_samples) for si in infos)
    assert total == 4096 * 131072


def test_decode_hot_loop_long_codes(eng, checkers):
    """The frame kernel's hot loop tops its ring up once per four symbols and reads partition headers, escape codes and Rice codes
    of up to 32 bits without looking again (dec_kernels.cu: top_up_hot): loud noise (codes of 15-22 bits at high Rice
    parameters: about 90 of the 128 bits a quad may take outside the generic path), impulses in silence (long unary runs -> generic path in
    the middle of a quad) and 32 streams side by side whose lanes drift apart, all against the oracle's bytes."""
    from pyflac_b200 import _native as nat
    rng = np.random.default_rng(77)
    xs, blobs = [], []
    for lane in range(32):
        n = 4096 * 3 + 8 * lane
        if lane % 4 == 0:
            x = rng.integers(-8192, 8192, (n, 2)).astype(np.int16)                      # ~15-bit codes (full scale would go out verbatim)
        elif lane % 4 == 1:
            x = np.zeros((n, 2), np.int16); x[rng.integers(0, n, 40), rng.integers(0, 2, 40)] = 32767   # impulses: unary runs past 32 zeros
        elif lane % 4 == 2:
            x = (rng.integers(-32768, 32768, (n, 2)) * (np.arange(n)[:, None] % 512 < 16)).astype(np.int16)  # bursts: mixed partitions
        else:
            x = corpus_signal("music", n, 2, 16, seed=lane)
        xs.append(x)
        blobs.append(checkers.oracle_encode(x, 48000, 16, (0, 3, 5, 8)[lane % 4 if lane % 4 != 3 else 2], 4096))
    x24 = rng.integers(-(1 << 20), 1 << 20, (4096 * 2 + 24, 2)).astype(np.int32)        # ~22-bit codes: 90 bits per quad
    out, infos = nat.decode_streams(eng, blobs)
    for lane, (x, o, si) in enumerate(zip(xs, out, infos)):
        assert si.status == 0, (lane, nat.DEC_STATUS.get(si.status))
        assert np.array_equal(o, x), lane
    out, infos = nat.decode_streams(eng, [checkers.oracle_encode(x24, 96000, 24, 5, 4096)] * 3)
    for o, si in zip(out, infos):
        assert si.status == 0 and np.array_equal(o, x24)


def test_decode_frames_of_256_kib_and_more():
    """Frames of 256 KiB and more (here: 16384 samples x 8 channels of incompressible 16-bit noise = 262 158 bytes, VERBATIM
    subframes): dec_crc_kernel's chunk numbers run past its two weight tables (4096 chunks of 64 bytes).  Runs in a process of
    its own (it was written after the last full GPU run of round 2; its first run on a B200 -- status 0, PCM equal, 1.06 s,
    gpurun_out/bigframe.log -- used the round's last GPU seconds)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import _checkers as ck\n"
        "from pyflac_b200 import _native as nat\n"
        "rng = np.random.default_rng(3)\n"
        "x = rng.integers(-32768, 32768, (16384 * 2 + 100, 8)).astype(np.int16)\n"
        "blob = ck.oracle_encode(x, 96000, 16, 5, 16384)\n"
        "assert len(blob) > 2 * 262144\n"
        "out, infos = nat.decode_streams(nat.Engine(0), [blob])\n"
        "assert infos[0].status == 0, nat.DEC_STATUS.get(infos[0].status)\n"
        "assert np.array_equal(out[0], x)\n"
        "print('big frames ok')\n") % (root, os.path.join(root, "tests"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "big frames ok" in r.stdout, r.stderr[-2000:]
