"""CPU tests of the drop-in layer's CONTROL surface (no GPU needed: nothing is initialised): the FLAC__stream_encoder_* /
FLAC__stream_decoder_* setters, getters, defaults, preset table and string tables of libflacb200.so compared call by call
with the reference's libFLAC 1.4.3 (oracle/_ref) -- the part of the boundary pyFLAC's properties read
(pyflac/encoder.py:141-231, pyflac/decoder.py:108; stream_encoder.h:803-864 for the presets)."""
import ctypes as C
import os

import pytest


@pytest.fixture(scope="module")
def libs(checkers):
    if not checkers.ref_available():
        pytest.skip("oracle/_ref not present")
    from pyflac_b200 import _native
    ours = C.CDLL(_native.LIB_PATH)
    ref = C.CDLL(os.path.join(checkers.ORACLE_DIR, "_ref", "libFLAC-12.1.0.so"))
    for L in (ours, ref):
        L.FLAC__stream_encoder_new.restype = C.c_void_p
        L.FLAC__stream_decoder_new.restype = C.c_void_p
        L.FLAC__stream_encoder_get_resolved_state_string.restype = C.c_char_p
        L.FLAC__stream_decoder_get_resolved_state_string.restype = C.c_char_p
    return ours, ref


ENC_GETTERS = ["state", "verify", "streamable_subset", "channels", "bits_per_sample", "sample_rate", "blocksize", "do_mid_side_stereo",
               "loose_mid_side_stereo", "max_lpc_order", "qlp_coeff_precision", "do_qlp_coeff_prec_search", "do_escape_coding",
               "do_exhaustive_model_search", "min_residual_partition_order", "max_residual_partition_order",
               "rice_parameter_search_dist", "limit_min_bitrate"]


def _enc_snapshot(L, e):
    out = {}
    for g in ENC_GETTERS:
        f = getattr(L, "FLAC__stream_encoder_get_" + g)
        f.argtypes = [C.c_void_p]
        f.restype = C.c_uint32
        out[g] = f(e)
    f = L.FLAC__stream_encoder_get_total_samples_estimate
    f.argtypes = [C.c_void_p]
    f.restype = C.c_uint64
    out["total_samples_estimate"] = f(e)
    L.FLAC__stream_encoder_get_resolved_state_string.argtypes = [C.c_void_p]
    out["state_string"] = L.FLAC__stream_encoder_get_resolved_state_string(e)
    return out


def _set(L, e, name, v):
    f = getattr(L, "FLAC__stream_encoder_set_" + name)
    f.argtypes = [C.c_void_p, C.c_uint64 if name == "total_samples_estimate" else C.c_uint32]
    f.restype = C.c_int
    return f(e, v)


def test_encoder_defaults_and_presets_match_libflac(libs):
    ours, ref = libs
    snaps = []
    for L in libs:
        L.FLAC__stream_encoder_delete.argtypes = [C.c_void_p]
        e = L.FLAC__stream_encoder_new()
        s = [("new", _enc_snapshot(L, e))]
        for level in list(range(9)) + [12, 5]:                      # 12: above the table (libFLAC clamps to 8)
            r = _set(L, e, "compression_level", level)
            s.append((f"level {level}", r, _enc_snapshot(L, e)))
        L.FLAC__stream_encoder_delete(e)
        snaps.append(s)
    assert snaps[0] == snaps[1]


def test_encoder_setters_round_trip_like_libflac(libs):
    """the setters pyFLAC uses (pyflac/encoder.py:145-231) plus blocksize / sample-rate extremes: same return values and the same
    values read back -- libFLAC stores what it is given and validates at init (stream_encoder.h:1471-1532)"""
    script = [("channels", 2), ("channels", 0), ("channels", 9), ("bits_per_sample", 16), ("bits_per_sample", 33), ("bits_per_sample", 3),
              ("sample_rate", 48000), ("sample_rate", 0), ("sample_rate", 1048576), ("blocksize", 4096), ("blocksize", 15), ("blocksize", 65536),
              ("verify", 1), ("verify", 0), ("streamable_subset", 0), ("limit_min_bitrate", 1), ("total_samples_estimate", 123456789012),
              ("compression_level", 3), ("blocksize", 0)]
    snaps = []
    for L in libs:
        e = L.FLAC__stream_encoder_new()
        s = []
        for name, v in script:
            s.append((name, v, _set(L, e, name, v), _enc_snapshot(L, e)))
        L.FLAC__stream_encoder_delete(e)
        snaps.append(s)
    for a, b in zip(*snaps):
        assert a == b


def test_fine_grained_setters_round_trip_like_libflac(libs):
    """builder/encoder.py:274-284: the setters pyFLAC declares but never calls store what they are given (validation happens at init)
    and a later set_compression_level writes the level's presets over them"""
    script = [("do_mid_side_stereo", 0), ("loose_mid_side_stereo", 1), ("max_lpc_order", 3), ("max_lpc_order", 40), ("qlp_coeff_precision", 9),
              ("qlp_coeff_precision", 2), ("do_qlp_coeff_prec_search", 1), ("do_exhaustive_model_search", 1), ("min_residual_partition_order", 2),
              ("max_residual_partition_order", 4), ("max_residual_partition_order", 20), ("rice_parameter_search_dist", 3),
              ("compression_level", 7), ("max_lpc_order", 11), ("compression_level", 2), ("do_mid_side_stereo", 1)]
    snaps = []
    for L in libs:
        e = L.FLAC__stream_encoder_new()
        s = []
        for name, v in script:
            s.append((name, v, _set(L, e, name, v), _enc_snapshot(L, e)))
        L.FLAC__stream_encoder_delete(e)
        snaps.append(s)
    for a, b in zip(*snaps):
        assert a == b


def test_fine_grained_init_validation_matches_libflac(libs):
    """values libFLAC rejects at init are rejected with its status; values it accepts are OK -- or ENCODER_ERROR here (no device / outside
    this build's range, which fails loudly instead of encoding something else)"""
    import numpy as np
    import _flacapi as fa
    ours, ref = libs
    x = np.zeros((4, 2), np.int16)
    cases = [[("max_lpc_order", 33)], [("max_lpc_order", 32)], [("max_lpc_order", 13)], [("max_lpc_order", 5)], [("qlp_coeff_precision", 4)],
             [("qlp_coeff_precision", 5)], [("qlp_coeff_precision", 16)], [("max_residual_partition_order", 9)], [("max_residual_partition_order", 16)],
             [("max_lpc_order", 20), ("qlp_coeff_precision", 3)], [("apodization", "hann")], [("apodization", "tukey(0.25)")],
             [("do_exhaustive_model_search", 1)], [("min_residual_partition_order", 3)]]
    for setters in cases:
        for subset in (True, False):
            for bs in (0, 16, 4096):
                kw = dict(sample_rate=48000, bps=16, level=5, blocksize=bs, streamable_subset=subset, init_only=True, setters=setters)
                a = fa.encode_session(ours, x, **kw)["init_status"]
                b = fa.encode_session(ref, x, **kw)["init_status"]
                assert (a == b) if b != 0 else (a in (0, 1)), (setters, subset, bs, a, b)


def test_uninitialised_calls_behave_like_libflac(libs):
    """process / finish on a handle that was never initialised, and the decoder's defaults"""
    res = []
    for L in libs:
        e = L.FLAC__stream_encoder_new()
        L.FLAC__stream_encoder_finish.argtypes = [C.c_void_p]
        L.FLAC__stream_encoder_finish.restype = C.c_int
        r = [L.FLAC__stream_encoder_finish(e), _enc_snapshot(L, e)["state"]]
        L.FLAC__stream_encoder_delete(e)
        d = L.FLAC__stream_decoder_new()
        for g in ["state", "md5_checking", "channels", "bits_per_sample", "sample_rate", "blocksize"]:
            f = getattr(L, "FLAC__stream_decoder_get_" + g)
            f.argtypes = [C.c_void_p]
            f.restype = C.c_uint32
            r.append((g, f(d)))
        L.FLAC__stream_decoder_get_resolved_state_string.argtypes = [C.c_void_p]
        r.append(L.FLAC__stream_decoder_get_resolved_state_string(d))
        for name in ["finish", "flush", "reset", "process_single", "process_until_end_of_stream"]:
            f = getattr(L, "FLAC__stream_decoder_" + name)
            f.argtypes = [C.c_void_p]
            f.restype = C.c_int
            r.append((name, f(d)))
        f = L.FLAC__stream_decoder_set_md5_checking
        f.argtypes = [C.c_void_p, C.c_int]
        f.restype = C.c_int
        r.append(("set_md5_checking", f(d, 1), L.FLAC__stream_decoder_get_md5_checking(d)))
        # the metadata filter (builder/decoder.py:392-397): settable on an uninitialised handle, type codes up to 126
        for name in ["respond_all", "ignore_all"]:
            f = getattr(L, "FLAC__stream_decoder_set_metadata_" + name)
            f.argtypes = [C.c_void_p]
            f.restype = C.c_int
            r.append((name, f(d)))
        for name in ["respond", "ignore"]:
            f = getattr(L, "FLAC__stream_decoder_set_metadata_" + name)
            f.argtypes = [C.c_void_p, C.c_int]
            f.restype = C.c_int
            r.append((name, [f(d, t) for t in (0, 1, 6, 7, 126, 127)]))      # (libFLAC asserts on larger codes)
        for name in ["respond_application", "ignore_application"]:
            f = getattr(L, "FLAC__stream_decoder_set_metadata_" + name)
            f.argtypes = [C.c_void_p, C.c_char_p]
            f.restype = C.c_int
            r.append((name, f(d, b"abcd")))
        f = L.FLAC__stream_decoder_get_decode_position
        f.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        f.restype = C.c_int
        r.append(("get_decode_position", f(d, C.byref(C.c_uint64(0)))))
        L.FLAC__stream_decoder_delete.argtypes = [C.c_void_p]
        L.FLAC__stream_decoder_delete(d)
        res.append(r)
    assert res[0] == res[1]


def test_string_tables_match_libflac(libs):
    tables = {"FLAC__StreamEncoderStateString": 9, "FLAC__StreamEncoderInitStatusString": 14, "FLAC__StreamDecoderStateString": 10,
              "FLAC__StreamDecoderInitStatusString": 6, "FLAC__StreamDecoderErrorStatusString": 5,
              "FLAC__StreamEncoderWriteStatusString": 2, "FLAC__StreamDecoderWriteStatusString": 2}
    for sym, n in tables.items():
        got = []
        for L in libs:
            try:
                arr = (C.c_char_p * n).in_dll(L, sym)
            except ValueError:
                got.append(None)
                continue
            got.append([arr[i] for i in range(n)])
        if got[0] is None:                                    # not every table is part of pyFLAC's cdef; the ones it reads must be there
            assert sym in ("FLAC__StreamEncoderWriteStatusString", "FLAC__StreamDecoderWriteStatusString"), sym
            continue
        assert got[0] == got[1], sym


def test_init_validation_matches_libflac_without_a_device(libs):
    """init_stream validates the settings before it touches the GPU, in libFLAC's order: every rejected
    configuration returns libFLAC's status; an accepted one returns OK -- or ENCODER_ERROR here, where there is no device"""
    import itertools
    import numpy as np
    import _flacapi as fa
    ours, ref = libs
    x = np.zeros((4, 2), np.int16)
    grid = itertools.product([0, 1, 8000, 48000, 96000, 655350, 655351, 1048575], [3, 4, 8, 16, 24, 32, 33], [0, 15, 16, 1152, 4608, 4609, 16384, 16385, 65535],
                             [0, 5, 8], [True, False])
    n_bad = 0
    for sr, bps, bs, level, subset in grid:
        kw = dict(sample_rate=sr, bps=bps, level=level, blocksize=bs, streamable_subset=subset, init_only=True)
        a = fa.encode_session(ours, x, **kw)["init_status"]
        b = fa.encode_session(ref, x, **kw)["init_status"]
        if b == 0:
            assert a in (0, 1), (kw, a, b)
        else:
            n_bad += 1
            assert a == b, (kw, a, b)
    assert n_bad > 100
    for ch in (0, 1, 8, 9):                                                 # channel count; missing write callback
        xc = np.zeros((4, max(ch, 1)), np.int16)
        a, b = [fa.encode_session(L, xc, 48000, 16, init_only=True)["init_status"] if ch else None for L in (ours, ref)]
        assert (a == b) or (b == 0 and a == 1), (ch, a, b)


def test_init_callback_and_file_errors_match_libflac(libs):
    """missing callbacks, seek without tell, files that cannot be opened: same init status and same state afterwards
    (reference tests: tests/test_decoder.py:107-111 `test_process_invalid_file`, tests/test_encoder.py:233-252)"""
    import _flacapi as fa
    res = []
    for L in libs:
        r = []
        L.FLAC__stream_decoder_init_stream.argtypes = [C.c_void_p] * 10
        L.FLAC__stream_decoder_init_stream.restype = C.c_int
        L.FLAC__stream_decoder_init_file.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.FLAC__stream_decoder_init_file.restype = C.c_int
        L.FLAC__stream_decoder_get_state.argtypes = [C.c_void_p]
        L.FLAC__stream_decoder_delete.argtypes = [C.c_void_p]
        rd, wr, er = fa.DEC_READ_CB(lambda *a: 0), fa.DEC_WRITE_CB(lambda *a: 0), fa.DEC_ERROR_CB(lambda *a: None)
        p = lambda f: C.cast(f, C.c_void_p)                                  # noqa: E731
        d = L.FLAC__stream_decoder_new()
        r.append(L.FLAC__stream_decoder_init_stream(d, None, None, None, None, None, p(wr), None, p(er), None))      # no read callback
        r.append(L.FLAC__stream_decoder_init_stream(d, p(rd), None, None, None, None, None, None, p(er), None))      # no write callback
        r.append(L.FLAC__stream_decoder_init_stream(d, p(rd), None, None, None, None, p(wr), None, None, None))      # no error callback
        r.append(L.FLAC__stream_decoder_init_stream(d, p(rd), p(rd), None, None, None, p(wr), None, p(er), None))    # seek without tell/length/eof
        r.append(L.FLAC__stream_decoder_init_file(d, b"/nonexistent/dir/x.flac", p(wr), None, p(er), None))
        r.append(L.FLAC__stream_decoder_init_file(d, b"/nonexistent/dir/x.flac", None, None, p(er), None))
        r.append(L.FLAC__stream_decoder_get_state(d))
        L.FLAC__stream_decoder_delete(d)
        fa._proto(L)
        L.FLAC__stream_encoder_init_stream.argtypes = [C.c_void_p] * 6
        w, s, t = fa.WRITE_CB(lambda *a: 0), fa.SEEK_CB(lambda *a: 0), fa.TELL_CB(lambda *a: 0)
        e = L.FLAC__stream_encoder_new()
        r.append(L.FLAC__stream_encoder_init_stream(e, None, None, None, None, None))                                 # no write callback
        r.append(L.FLAC__stream_encoder_init_stream(e, p(w), p(s), None, None, None))                                 # seek without tell
        r.append(L.FLAC__stream_encoder_get_state(e))
        r.append(L.FLAC__stream_encoder_init_file(e, b"/nonexistent/dir/x.flac", None, None))
        r.append(L.FLAC__stream_encoder_get_state(e))
        L.FLAC__stream_encoder_delete(e)
        res.append(r)
    assert res[0] == res[1], res
