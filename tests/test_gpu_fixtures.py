"""GPU parity on the reference's OWN fixtures (pyFLAC tests/data, imported under tests/golden/fixtures): real audio,
files written by libFLAC 1.3.3 / 1.4.2 with SEEKTABLE + VORBIS_COMMENT + 8 KiB PADDING in front of the frames.
Mirrors pyFLAC's tests/test_decoder.py and tests/test_encoder.py cases that use these files."""
import os
import struct
import tempfile

import numpy as np
import pytest

from conftest import fixture_cases, fixture_path, foreign_cases, foreign_path, pcm_md5

pytestmark = pytest.mark.gpu
FIX = fixture_cases()


def _blob(name):
    with open(fixture_path(name), "rb") as f:
        return f.read()


def _write_wav(path, x, sr, bits):
    raw = np.ascontiguousarray(x).astype("<i2" if bits == 16 else "<i4").tobytes()
    ch = x.shape[1]
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVEfmt " +
                struct.pack("<IHHIIHH", 16, 1, ch, sr, sr * ch * bits // 8, ch * bits // 8, bits) + b"data" + struct.pack("<I", len(raw)) + raw)


@pytest.fixture(scope="module")
def eng():
    from pyflac_b200 import _native as nat
    return nat.Engine(0)


def test_batch_decode_fixtures_md5(eng, checkers):
    """one ragged batch per container width: PCM hashes to the STREAMINFO MD5 (== the .wav data) and equals the oracle's"""
    from pyflac_b200 import _native as nat
    for grp in ([c for c in FIX if c["bps"] <= 16], [c for c in FIX if c["bps"] > 16]):
        out, infos = nat.decode_streams(eng, [_blob(c["name"] + ".flac") for c in grp])
        for c, o, si in zip(grp, out, infos):
            assert si.status == 0, (c["name"], nat.DEC_STATUS.get(si.status))
            assert (si.channels, si.bits_per_sample, si.sample_rate, si.total_samples) == (c["channels"], c["bps"], c["sample_rate"], c["samples"])
            assert o.shape == (c["samples"], c["channels"])
            assert pcm_md5(o, c["bps"]) == c["streaminfo_md5"], c["name"]
            want, _ = checkers.oracle_decode(_blob(c["name"] + ".flac"))
            assert np.array_equal(o.astype(np.int64), want.astype(np.int64))


def test_batch_encode_fixtures_match_libflac(eng, checkers):
    """real audio through the CUDA encoder == the file the bundled libFLAC 1.4.3 wrote (levels 5 committed; 0/8 vs oracle)"""
    from pyflac_b200 import _native as nat
    for c in FIX:
        if not c["level5"]:
            continue
        x, _ = checkers.oracle_decode(_blob(c["name"] + ".flac"))
        x = np.ascontiguousarray(x.astype(np.int16 if c["bps"] <= 16 else np.int32))
        got, _ = nat.encode_streams(eng, [x], c["sample_rate"], c["bps"], 5, 0)
        assert got[0] == _blob(c["level5"]["file"]), c["name"]
        for level in (0, 8):
            got, _ = nat.encode_streams(eng, [x], c["sample_rate"], c["bps"], level, 0)
            assert got[0] == checkers.oracle_encode(x, c["sample_rate"], c["bps"], level, 0), (c["name"], level)


def test_stream_decoder_on_fixture_like_reference_tests():
    """tests/test_decoder.py: test_process (whole file) and test_process_blocks (1024-byte pieces) on stereo.flac"""
    import pyflac_b200 as pf
    c = next(c for c in FIX if c["name"] == "stereo")
    data = _blob("stereo.flac")
    for step in (len(data), 1024):
        got = []

        def cb(audio, sr, ch, n):
            assert isinstance(audio, np.ndarray) and isinstance(sr, int) and isinstance(ch, int) and isinstance(n, int)
            got.append(audio.copy())
        dec = pf.StreamDecoder(write_callback=cb)
        for i in range(0, len(data), step):
            dec.process(data[i:i + step])
        dec.finish()
        assert not dec._thread.is_alive()
        pcm = np.concatenate(got)
        assert pcm.dtype == np.int16 and pcm_md5(pcm, 16) == c["streaminfo_md5"]


def test_file_decoder_on_fixtures_like_reference_tests():
    """tests/test_decoder.py TestFileDecoder: mono / stereo / 32-bit files decode to a WAV, the 8-bit file raises"""
    import pyflac_b200 as pf
    from pyflac_b200 import wav
    with pytest.raises(pf.DecoderProcessException):
        pf.FileDecoder(fixture_path("8bit.flac")).process()
    for c in FIX:
        if c["name"] == "8bit":
            continue
        with tempfile.TemporaryDirectory() as d:
            outp = os.path.join(d, "out.wav")
            samples, sr = pf.FileDecoder(fixture_path(c["name"] + ".flac"), outp).process()
            assert sr == c["sample_rate"] and samples.shape == (c["samples"], c["channels"]) and samples.dtype == np.float64
            assert wav.info(outp).subtype == "PCM_16"
            if c["bps"] == 16:
                assert pcm_md5(np.rint(samples * 32768.0).astype(np.int32), 16) == c["streaminfo_md5"]


def test_file_encoder_on_fixture_audio(checkers):
    """tests/test_encoder.py TestFileEncoder: WAV in -> FLAC file out, bytes identical to libFLAC's for the same WAV"""
    import pyflac_b200 as pf
    for c in FIX:
        if not c["level5"]:
            continue
        x, _ = checkers.oracle_decode(_blob(c["name"] + ".flac"))
        with tempfile.TemporaryDirectory() as d:
            wavp, flacp = os.path.join(d, "in.wav"), os.path.join(d, "out.flac")
            _write_wav(wavp, x, c["sample_rate"], c["bps"])
            data = pf.FileEncoder(wavp, flacp, compression_level=5).process()
            assert data == _blob(c["level5"]["file"]) == open(flacp, "rb").read(), c["name"]


def test_batch_decode_foreign_streams(eng, checkers):
    """streams the presets never produce (tests/golden/foreign): predictor orders up to 32 (the decoder's generic path),
    partition order 8, partitions of 9 samples, 16-sample blocks, escape-coded partitions with 0..22 raw bits, 5-bit Rice
    parameters, CONSTANT / VERBATIM / FIXED side by side, wasted bits, every channel assignment"""
    from pyflac_b200 import _native as nat
    cases = foreign_cases()
    blobs = {c["name"]: open(foreign_path(c["name"] + ".flac"), "rb").read() for c in cases}
    for grp in ([c for c in cases if c["bps"] <= 16], [c for c in cases if c["bps"] > 16]):
        out, infos = nat.decode_streams(eng, [blobs[c["name"]] for c in grp])
        for c, o, si in zip(grp, out, infos):
            assert si.status == 0, (c["name"], nat.DEC_STATUS.get(si.status))
            assert o.shape == (c["samples"], c["channels"]) and si.bits_per_sample == c["bps"]
            assert pcm_md5(o, c["bps"]) == c["pcm_md5"], c["name"]
            want, _ = checkers.oracle_decode(blobs[c["name"]])
            assert np.array_equal(o.astype(np.int64), want.astype(np.int64)), c["name"]
    # and one of them through the drop-in stream decoder, fed in small pieces
    import pyflac_b200 as pf
    c = next(c for c in cases if c["name"] == "crafted_method1_s16_st_bs4096")
    got = []
    dec = pf.StreamDecoder(write_callback=lambda a, sr, ch, n: got.append(a.copy()))
    data = blobs[c["name"]]
    for i in range(0, len(data), 3000):
        dec.process(data[i:i + 3000])
    dec.finish()
    assert pcm_md5(np.concatenate(got), 16) == c["pcm_md5"]
