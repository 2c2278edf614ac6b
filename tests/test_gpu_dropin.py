"""GPU tests of the DROP-IN layer: the FLAC__stream_encoder_* / FLAC__stream_decoder_* symbols of libflacb200.so driven
exactly like pyFLAC's cffi trampolines drive libFLAC, compared callback-for-callback with the reference binary."""
import ctypes as C
import os

import numpy as np
import pytest

from pyflac_b200.synth import corpus_signal, music_like

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ours():
    from pyflac_b200 import _native
    return C.CDLL(_native.LIB_PATH)


@pytest.fixture(scope="module")
def ref(checkers):
    if not checkers.ref_available():
        pytest.skip("oracle/_ref not present")
    return C.CDLL(os.path.join(checkers.ORACLE_DIR, "_ref", "libFLAC-12.1.0.so"))


def _norm(log):
    return [e if e[0] != "meta" else ("meta", tuple(sorted(e[1].items()))) for e in log]


@pytest.mark.parametrize("chunks", [None, [4096, 1, 4095, 1, 4195], [1000] * 30])
def test_stream_encoder_callback_sequence_matches_libflac(ours, ref, chunks):
    """over-read framing, 3 prologue writes, tell before every write, STREAMINFO rewrite at 26/21/12, metadata callback"""
    from _flacapi import encode_session
    x = corpus_signal("music", 4096 * 3 + 100, 2, 16, seed=4)
    a = encode_session(ours, x, 48000, 16, 5, 0, chunks=chunks)
    b = encode_session(ref, x, 48000, 16, 5, 0, chunks=chunks)
    assert a["init_status"] == b["init_status"] == 0
    assert _norm(a["log"]) == _norm(b["log"])
    assert a["file"] == b["file"] and a["finish"] == b["finish"] and a["state_after_finish"] == b["state_after_finish"] == 1


def test_stream_encoder_without_seek_and_levels(ours, ref):
    from _flacapi import encode_session
    x = corpus_signal("mixed", 4096 * 2 + 50, 2, 16, seed=2)
    for level in (0, 1, 2, 3, 4, 5, 8):
        a = encode_session(ours, x, 44100, 16, level, 0, seekable=False)
        b = encode_session(ref, x, 44100, 16, level, 0, seekable=False)
        assert _norm(a["log"]) == _norm(b["log"]), level
    x24 = corpus_signal("music", 9000, 1, 24, seed=1)
    a = encode_session(ours, x24, 192000, 24, 8, 4096)
    b = encode_session(ref, x24, 192000, 24, 8, 4096)
    assert a["file"] == b["file"]


def test_stream_encoder_loose_mid_side_across_calls(ours, ref):
    """levels 1/4: the decision made in one process_interleaved() call governs frames encoded by later calls"""
    from _flacapi import encode_session
    n = 1152 * 40 + 77
    a0 = corpus_signal("music", n, 2, 16, seed=11)
    b0 = corpus_signal("lr_uncorr", n, 2, 16, seed=12)
    x = a0.copy(); x[n // 2:] = b0[n // 2:]
    for level, bs in [(1, 0), (4, 576), (4, 0)]:
        for chunks in ([5000] * 20, [1153, 1, 1151] * 40, None):
            a = encode_session(ours, x, 44100, 16, level, bs, chunks=chunks)
            b = encode_session(ref, x, 44100, 16, level, bs, chunks=chunks)
            assert _norm(a["log"]) == _norm(b["log"]), (level, bs, chunks and chunks[0])
            assert a["file"] == b["file"]
    sil = np.zeros((4096 * 2 + 10, 2), np.int16)
    a = encode_session(ours, sil, 48000, 16, 5, 0, limit_min_bitrate=True)
    b = encode_session(ref, sil, 48000, 16, 5, 0, limit_min_bitrate=True)
    assert _norm(a["log"]) == _norm(b["log"]) and a["file"] == b["file"]


def test_stream_encoder_verify(ours, ref):
    """verify=True: same bytes and callbacks as libFLAC's verifying encoder, state stays OK (decode-and-compare on the GPU)"""
    from _flacapi import encode_session
    for x, bps, sr in [(corpus_signal("mixed", 4096 * 3 + 11, 2, 16, seed=21), 16, 44100), (corpus_signal("wasted", 5000, 1, 24, seed=22), 24, 96000)]:
        a = encode_session(ours, x, sr, bps, 5, 0, chunks=[3000] * 10, verify=True)
        b = encode_session(ref, x, sr, bps, 5, 0, chunks=[3000] * 10, verify=True)
        assert a["ok"] and a["finish"] == b["finish"] == 1
        assert _norm(a["log"]) == _norm(b["log"]) and a["file"] == b["file"]


def test_stream_encoder_init_errors(ours, ref):
    """reference tests/test_encoder.py:139-164,202-207"""
    from _flacapi import encode_session
    x = np.zeros((16, 2), np.int16)
    for kw in [dict(sample_rate=2000000), dict(blocksize=1000000), dict(blocksize=65535), dict(blocksize=65535, streamable_subset=False)]:
        args = dict(sample_rate=48000, blocksize=0, streamable_subset=True)
        args.update(kw)
        a = encode_session(ours, x, args["sample_rate"], 16, 5, args["blocksize"], streamable_subset=args["streamable_subset"], init_only=True)
        b = encode_session(ref, x, args["sample_rate"], 16, 5, args["blocksize"], streamable_subset=args["streamable_subset"], init_only=True)
        assert a["init_status"] == b["init_status"], kw
    a = encode_session(ours, x, 48000, 16, 5, 0, no_tell=True, init_only=True)
    b = encode_session(ref, x, 48000, 16, 5, 0, no_tell=True, init_only=True)
    assert a["init_status"] == b["init_status"] == 3        # INVALID_CALLBACKS


def test_stream_decoder_matches_libflac(ours, ref, checkers):
    from _flacapi import decode_session
    x = music_like(4096 * 4 + 321, 2, 44100, 16, seed=8)
    data = checkers.ref_encode(x, 44100, 16, 5, 0)
    for rc in (8192, 1000, 1 << 20):
        a = decode_session(ours, data, rc)
        b = decode_session(ref, data, 8192)
        assert a["init_status"] == 0 and a["ok"] and not a["errors"]
        assert np.array_equal(a["pcm"], b["pcm"]) and np.array_equal(a["pcm"], x)
        assert [f["blocksize"] for f in a["frames"]] == [f["blocksize"] for f in b["frames"]]
        assert a["frames"][0]["sample_rate"] == 44100 and a["frames"][0]["channels"] == 2 and a["frames"][0]["bits_per_sample"] == 16
        assert a["state"] == b["state"] == 4                # END_OF_STREAM


def test_stream_decoder_garbage_reports_error(ours):
    """reference tests/test_decoder.py:59-66: random bytes -> error callback fires"""
    from _flacapi import decode_session
    junk = np.random.default_rng(1).integers(0, 256, 100000).astype(np.uint8).tobytes()
    a = decode_session(ours, junk)
    assert a["errors"] and a["pcm"].shape[0] == 0


def test_many_stream_encoders_on_threads_share_gpu_batches(checkers):
    """BASELINE configs[1] the way pyFLAC users reach it: 256 StreamEncoder objects, one thread each
    (/root/reference/pyflac/encoder.py:293-330).  The per-device dispatcher coalesces their concurrent process() calls into
    shared GPU batches; every stream must still equal libFLAC's bytes, and the threaded run must beat the same 256 encoders
    run one after the other (one-stream batches) by a wide margin."""
    import ctypes as C
    import threading
    import time
    import pyflac_b200 as pf
    from pyflac_b200 import _native as nat
    from pyflac_b200.synth import music_like
    n_enc, n = 256, 4096 * 24 + 500
    xs = [music_like(n, 2, 48000, 16, seed=7000 + s) for s in range(n_enc)]
    want = [checkers.oracle_encode(x, 48000, 16, 5, 0, seekable=False) for x in xs[:8]]

    def run(threaded):
        outs = [bytearray() for _ in range(n_enc)]
        encs = [pf.StreamEncoder(48000, (lambda b, nb, ns, fr, o=outs[s]: o.extend(b)), compression_level=5) for s in range(n_enc)]

        def work(s):
            for i in range(0, n, 4096 * 8):               # three process() calls per stream, then finish()
                encs[s].process(xs[s][i:i + 4096 * 8])
            assert encs[s].finish() is True
        t0 = time.perf_counter()
        if threaded:
            th = [threading.Thread(target=work, args=(s,)) for s in range(n_enc)]
            [t.start() for t in th]
            [t.join(120) for t in th]
            assert not any(t.is_alive() for t in th)
        else:
            for s in range(n_enc):
                work(s)
        return time.perf_counter() - t0, outs

    L = nat.lib()
    L.flacb200_dispatch_stats.argtypes = [C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    b0, j0, b1, j1 = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    run(True)                                               # warm-up: contexts, pinned buffers, kernels
    L.flacb200_dispatch_stats(0, C.byref(b0), C.byref(j0))
    t_thr, outs = run(True)
    L.flacb200_dispatch_stats(0, C.byref(b1), C.byref(j1))
    t_seq, outs_seq = run(False)
    for s in range(8):
        assert bytes(outs[s]) == want[s] == bytes(outs_seq[s]), s
    for s in range(n_enc):
        assert outs[s] == outs_seq[s], s
    jobs, batches = j1.value - j0.value, b1.value - b0.value
    assert jobs >= n_enc * 4
    print(f"\n256 StreamEncoders: threaded {t_thr * 1e3:.1f} ms ({batches} GPU batches for {jobs} jobs), one after the other {t_seq * 1e3:.1f} ms")
    assert batches * 4 <= jobs                             # at least four handles per batch on average
    assert t_thr * 3 < t_seq


def test_write_callback_may_drive_another_encoder(checkers):
    """Callbacks run outside every library lock (libFLAC handles are independent): a write callback that
    feeds a second encoder must not deadlock."""
    import threading
    import pyflac_b200 as pf
    from pyflac_b200.synth import music_like
    x = music_like(4096 * 6 + 11, 2, 44100, 16, seed=11)
    y = music_like(4096 * 2, 1, 44100, 16, seed=12)
    inner_out, outer_out, fed = bytearray(), bytearray(), [0]
    inner = pf.StreamEncoder(44100, lambda b, nb, ns, fr: inner_out.extend(b), compression_level=3)

    def outer_cb(b, nb, ns, fr):
        outer_out.extend(b)
        if ns and fed[0] < len(y):                          # every audio frame of the outer stream feeds the inner encoder
            inner.process(y[fed[0]:fed[0] + 2048])
            fed[0] += 2048
    outer = pf.StreamEncoder(44100, outer_cb, compression_level=5)
    done = []

    def work():
        outer.process(x)
        outer.finish()
        inner.finish()
        done.append(True)
    t = threading.Thread(target=work, daemon=True)
    t.start()
    t.join(60)
    assert done, "deadlock: a write callback could not use another encoder"
    assert bytes(outer_out) == checkers.oracle_encode(x, 44100, 16, 5, 0, seekable=False)
    assert bytes(inner_out) == checkers.oracle_encode(y[:fed[0]], 44100, 16, 3, 0, seekable=False)


def _flac_for_seek(checkers, n=4096 * 30 + 1234, ch=2, bps=16, bs=0, level=5, seed=9):
    x = music_like(n, ch, 48000, bps, seed=seed)
    return x, checkers.oracle_encode(x, 48000, bps, level, bs)


@pytest.mark.parametrize("via", ["callbacks", "file"])
def test_decoder_seek_matches_libflac(ours, ref, checkers, tmp_path, via):
    """FLAC__stream_decoder_seek_absolute (builder/decoder.py:475): the frame holding the target is delivered inside the call, cut to
    start at the target, with its sample number; the next process_* calls continue behind it; targets at frame starts, inside
    frames, in the short last frame, backwards, after END_OF_STREAM; out-of-range targets fail without touching the state."""
    from _flacapi import scripted_decode_session
    x, flac = _flac_for_seek(checkers)
    n = len(x)
    ops = [('seek', 4096 * 7 + 5), ('single', 3), ('seek', 0), ('single', 1), ('seek', 4096 * 29), ('single', 1), ('seek', n - 1), ('end',),
           ('seek', 4096 * 3 - 1), ('single', 2), ('seek', n), ('seek', n + 1000), ('seek', 4096 * 12), ('end',), ('seek', 17), ('single', 1)]
    path = None
    if via == "file":
        path = str(tmp_path / "seek.flac")
        with open(path, "wb") as f:
            f.write(flac)
    a = scripted_decode_session(ours, flac, ops, path=path)
    b = scripted_decode_session(ref, flac, ops, path=path)
    assert a["init_status"] == b["init_status"] == 0
    assert a["events"] == b["events"]
    assert a["finish"] == b["finish"]


def test_decoder_seek_variable_shapes(ours, ref, checkers):
    """seeking in streams whose frames carry frame numbers with a small blocksize (many frames per probe window), 24-bit mono, and a
    stream of several MB (more than one probe)"""
    from _flacapi import scripted_decode_session
    for (n, ch, bps, bs, level) in [(50000, 1, 16, 192, 2), (4096 * 9 + 3, 1, 24, 4096, 8), (48000 * 40, 2, 16, 4096, 0)]:
        x, flac = _flac_for_seek(checkers, n, ch, bps, bs, level, seed=n % 97)
        rng = np.random.default_rng(n)
        ops = []
        for t in rng.integers(0, n, size=12):
            ops += [('seek', int(t)), ('single', 2)]
        ops += [('seek', n - 1), ('end',)]
        a = scripted_decode_session(ours, flac, ops)
        b = scripted_decode_session(ref, flac, ops)
        assert a["events"] == b["events"], (n, ch, bps, bs)


def test_decoder_seek_needs_seekable_input(ours, ref, checkers):
    from _flacapi import scripted_decode_session
    x, flac = _flac_for_seek(checkers, 20000)
    ops = [('seek', 5000), ('single', 2), ('end',)]
    a = scripted_decode_session(ours, flac, ops, seekable=False)
    b = scripted_decode_session(ref, flac, ops, seekable=False)
    assert a["events"] == b["events"] and a["events"][0][:3] == ('ret', 'seek', 0)


def test_decoder_md5_checking_matches_libflac(ours, ref, checkers):
    """set_md5_checking (builder/decoder.py:391): finish() is false when the MD5 of the delivered samples differs from STREAMINFO's --
    a wrong stored digest, or a decode that stopped early; a zero digest is not checked; seek and flush switch the check off."""
    from _flacapi import scripted_decode_session
    x, flac = _flac_for_seek(checkers, 4096 * 5 + 77)
    bad = bytearray(flac); bad[26] ^= 0x55; bad = bytes(bad)
    zero = bytearray(flac); zero[26:42] = bytes(16); zero = bytes(zero)
    cases = [(flac, [('end',)], True), (bad, [('end',)], True), (zero, [('end',)], True), (flac, [('meta',), ('single', 2)], True),
             (bad, [('end',)], False), (bad, [('seek', 100), ('end',)], True), (bad, [('single', 3), ('flush',), ('end',)], True),
             (flac, [('single', 3), ('reset',), ('end',)], True), (bad, [('single', 3), ('reset',), ('end',)], True)]
    for i, (data, ops, chk) in enumerate(cases):
        a = scripted_decode_session(ours, data, ops, md5_checking=chk)
        b = scripted_decode_session(ref, data, ops, md5_checking=chk)
        assert a["finish"] == b["finish"], i
        if ('flush',) not in ops:       # what survives a mid-stream flush depends on how much input each decoder had buffered ahead
            assert a["events"] == b["events"], i
    assert scripted_decode_session(ours, bad, [('end',)], md5_checking=True)["finish"] is False
    assert scripted_decode_session(ours, flac, [('end',)], md5_checking=True)["finish"] is True


TUNINGS = [
    [("max_lpc_order", 4)], [("max_lpc_order", 11)], [("max_lpc_order", 1)], [("max_lpc_order", 0)],
    [("max_residual_partition_order", 2)], [("max_residual_partition_order", 0)], [("max_residual_partition_order", 6)],
    [("qlp_coeff_precision", 8)], [("qlp_coeff_precision", 15)], [("qlp_coeff_precision", 5)],
    [("do_mid_side_stereo", 0)], [("do_mid_side_stereo", 1), ("loose_mid_side_stereo", 1)],
    [("apodization", "tukey(0.25)")], [("apodization", "tukey(1)")], [("apodization", "tukey(0)")], [("apodization", "subdivide_tukey(2)")],
    [("apodization", "subdivide_tukey(3/0.3)")], [("apodization", "tukey(7)")], [("rice_parameter_search_dist", 4)],
    [("max_lpc_order", 6), ("qlp_coeff_precision", 10), ("max_residual_partition_order", 3), ("apodization", "subdivide_tukey(2/0.8)"), ("do_mid_side_stereo", 0)],
    [("do_exhaustive_model_search", 0), ("do_qlp_coeff_prec_search", 0), ("min_residual_partition_order", 0)],
]


@pytest.mark.parametrize("bps,ch,level", [(16, 2, 5), (24, 1, 3), (16, 2, 1), (8, 2, 8)])
def test_fine_grained_settings_match_libflac(ours, ref, bps, ch, level):
    """builder/encoder.py:274-284 (set_do_mid_side_stereo ... set_apodization): applied after the compression level, the stream is
    libFLAC's byte for byte -- on the TMA path (16-bit stereo) and the generic one; settings outside this build's range fail at init"""
    from _flacapi import encode_session
    x = music_like(4096 * 3 + 517, ch, 48000, bps, seed=bps + level)
    for setters in TUNINGS:
        a = encode_session(ours, x, 48000, bps, level, 0, setters=setters)
        b = encode_session(ref, x, 48000, bps, level, 0, setters=setters)
        assert a["init_status"] == b["init_status"] == 0, setters
        assert a["file"] == b["file"], setters
    for setters in ([("max_lpc_order", 16)], [("do_exhaustive_model_search", 1)], [("apodization", "hann")], [("apodization", "tukey(0.5);hann")],
                    [("max_residual_partition_order", 7)], [("min_residual_partition_order", 2)], [("apodization", "subdivide_tukey(4)")]):
        a = encode_session(ours, x, 48000, bps, level, 0, setters=setters, streamable_subset=False)
        assert a["init_status"] == 1 and a["file"] == b"", setters          # FLAC__STREAM_ENCODER_INIT_STATUS_ENCODER_ERROR, nothing written
