"""GPU tests of the pyFLAC-compatible Python classes (the reference's own test scenarios, tests/test_encoder.py and
tests/test_decoder.py, re-stated against pyflac_b200) plus byte parity of what the callbacks deliver."""
import os
import struct
import tempfile

import numpy as np
import pytest

from pyflac_b200.synth import music_like

pytestmark = pytest.mark.gpu


def write_wav(path, x, sr, bits):
    x = np.ascontiguousarray(x)
    ch = x.shape[1]
    raw = x.astype("<i2" if bits == 16 else "<i4").tobytes()
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(raw)) + b"WAVEfmt " +
                struct.pack("<IHHIIHH", 16, 1, ch, sr, sr * ch * bits // 8, ch * bits // 8, bits) + b"data" + struct.pack("<I", len(raw)) + raw)


def test_stream_encoder_callbacks_and_parity(checkers):
    import pyflac_b200 as pf
    x = music_like(4096 * 3 + 500, 2, 44100, 16, seed=3)
    chunks, pos, meta = [], [0], []
    image = bytearray()

    def write_cb(buf, nbytes, nsamples, frame):
        assert isinstance(buf, bytes) and isinstance(nbytes, int) and isinstance(nsamples, int) and isinstance(frame, int)
        chunks.append((nbytes, nsamples, frame))
        image[pos[0]:pos[0] + nbytes] = buf
        pos[0] += nbytes

    def seek_cb(off):
        pos[0] = off

    enc = pf.StreamEncoder(44100, write_cb, seek_cb, lambda: pos[0], lambda m: meta.append((m.data.stream_info.total_samples, bytes(m.data.stream_info.md5sum))),
                           compression_level=5, blocksize=0, verify=True)
    assert enc.state == pf.EncoderState.UNINITIALIZED and str(enc.state) == "FLAC__STREAM_ENCODER_UNINITIALIZED"
    for i in range(0, len(x), 1024):                      # reference tests feed 1024-sample blocks
        enc.process(x[i:i + 1024])
    assert enc.state == pf.EncoderState.OK
    assert enc.finish() is True
    assert enc.state == pf.EncoderState.UNINITIALIZED
    assert chunks[0][:2] == (4, 0) and chunks[1][:2] == (38, 0) and chunks[2][:2] == (44, 0)
    assert [c[1] for c in chunks if c[1]] == [4096, 4096, 4096, 500]
    assert bytes(image) == checkers.oracle_encode(x, 44100, 16, 5, 0)
    assert meta and meta[0][0] == len(x) and meta[0][1] == bytes(image[26:42])


def test_stream_encoder_errors():
    import pyflac_b200 as pf
    e = pf.StreamEncoder(48000, lambda *a: None)
    with pytest.raises(TypeError):
        e.process([1, 2, 3, 4])
    for kw, code in [(dict(sample_rate=2000000), 6), (dict(blocksize=1000000), 7), (dict(blocksize=65535), 11)]:
        args = dict(sample_rate=48000, blocksize=0)
        args.update(kw)
        e = pf.StreamEncoder(args["sample_rate"], lambda *a: None, blocksize=args["blocksize"])
        with pytest.raises(pf.EncoderInitException) as ei:
            e.process(np.zeros((1024, 2), np.int16))
        assert ei.value.code == code and str(ei.value).startswith("FLAC__STREAM_ENCODER_INIT_STATUS_")
    e = pf.StreamEncoder(48000, lambda *a: None, blocksize=65535, streamable_subset=False)
    e.process(np.zeros((1024, 2), np.int16))
    e.finish()
    e = pf.StreamEncoder(48000, lambda *a: None, seek_callback=lambda o: None)          # seek without tell
    with pytest.raises(pf.EncoderInitException) as ei:
        e.process(np.zeros((1024, 2), np.int16))
    assert ei.value.code == 3

    def boom(*a):
        raise RuntimeError("client failure")
    e = pf.StreamEncoder(48000, boom)
    with pytest.raises((pf.EncoderInitException, pf.EncoderProcessException)):
        e.process(np.zeros((1024, 2), np.int16))


def test_encoder_property_roundtrip():
    """reference tests/test_encoder.py:32-93 (setter -> getter on the raw handle)"""
    import pyflac_b200 as pf
    from pyflac_b200.encoder import _Encoder
    e = _Encoder()
    for name, val in [("_verify", True), ("_channels", 1), ("_bits_per_sample", 24), ("_sample_rate", 48000), ("_blocksize", 128),
                      ("_streamable_subset", False), ("_limit_min_bitrate", True)]:
        setattr(e, name, val)
        assert getattr(e, name) == val
    with pytest.raises(NotImplementedError):
        e._compression_level
    assert e.state == pf.EncoderState.UNINITIALIZED


def test_file_encoder_and_decoder_roundtrip(checkers):
    import pyflac_b200 as pf
    for bits, ch in [(16, 1), (16, 2), (32, 1), (32, 2)]:
        x = music_like(4096 * 2 + 124, ch, 44100, 16, seed=bits + ch)
        xs = x if bits == 16 else (x.astype(np.int32) << 8)       # 32-bit container, 24 significant bits: wasted-bits path
        with tempfile.TemporaryDirectory() as d:
            wavp, flacp, outp = os.path.join(d, "a.wav"), os.path.join(d, "a.flac"), os.path.join(d, "b.wav")
            write_wav(wavp, xs, 44100, bits)
            data = pf.FileEncoder(wavp, flacp, compression_level=5).process()
            assert data == open(flacp, "rb").read() == checkers.oracle_encode(xs, 44100, bits, 5, 0)
            pcm, sr = pf.FileDecoder(flacp, outp).process()
            assert sr == 44100 and pcm.dtype == np.float64 and pcm.shape == (len(x), ch)
            if bits == 16:
                assert np.array_equal(np.rint(pcm * 32768.0).astype(np.int16), x)
            else:       # the reference writes PCM_16 whatever the stream's depth (decoder.py:300-313): the high 16 bits survive
                assert np.max(np.abs(pcm - xs.astype(np.float64) / 2147483648.0)) <= 2.0 ** -15
    with pytest.raises(pf.DecoderInitException):
        pf.FileDecoder("/nonexistent/file.flac")


def test_stream_and_oneshot_decoder(checkers):
    import pyflac_b200 as pf
    x = music_like(4096 * 5 + 77, 2, 48000, 16, seed=21)
    data = checkers.oracle_encode(x, 48000, 16, 5, 0)
    got = []

    def cb(audio, sr, ch, n):
        assert audio.dtype == np.int16 and audio.shape == (n, ch) and sr == 48000
        got.append(audio.copy())
    dec = pf.StreamDecoder(cb)
    for i in range(0, len(data), 1024):                   # reference test_process_blocks
        dec.process(data[i:i + 1024])
    dec.finish()
    assert np.array_equal(np.concatenate(got), x)
    got.clear()
    pf.OneShotDecoder(cb, data)
    assert np.array_equal(np.concatenate(got), x)
    dec = pf.StreamDecoder(cb)                             # reference test_process_invalid_data
    dec.process(np.random.default_rng(0).integers(0, 256, 100000).astype(np.uint8).tobytes())
    with pytest.raises(pf.DecoderProcessException):
        dec.finish()
    data8 = checkers.oracle_encode(music_like(5000, 1, 22050, 8, seed=1), 22050, 8, 5, 0)      # 8-bit stream: reference rejects it
    with tempfile.NamedTemporaryFile(suffix=".flac") as f:
        f.write(data8)
        f.flush()
        with pytest.raises(pf.DecoderProcessException):
            pf.FileDecoder(f.name).process()


def test_batch_frontend(checkers):
    import pyflac_b200 as pf
    xs = [music_like(10000 + 100 * s, 2, 48000, 16, seed=s) for s in range(8)]
    blobs, info = pf.encode_batch(xs, 48000, compression_level=8)
    for x, b in zip(xs, blobs):
        assert b == checkers.oracle_encode(x, 48000, 16, 8, 0)
    out, infos = pf.decode_batch(blobs)
    for x, o in zip(xs, out):
        assert np.array_equal(o, x)
    x24 = [music_like(9000, 1, 192000, 24, seed=7)]
    blobs, _ = pf.encode_batch(x24, 192000, compression_level=8, blocksize=4096, bits_per_sample=24)
    assert blobs[0] == checkers.oracle_encode(x24[0], 192000, 24, 8, 4096)


def test_config_c1_file_encoder_10s_stereo(checkers, tmp_path):
    """BASELINE configs[0], the exact shape: one 10 s 48 kHz stereo int16 WAV through FileEncoder at level 5, default
    blocksize (117 frames of 4096 + one of 768).  The file must equal libFLAC's byte for byte."""
    import pyflac_b200 as pf
    x = music_like(480000, 2, 48000, 16, seed=2024)
    wav, flac = str(tmp_path / "c1.wav"), str(tmp_path / "c1.flac")
    write_wav(wav, x, 48000, 16)
    data = pf.FileEncoder(wav, flac, compression_level=5).process()
    want = checkers.ref_encode(x, 48000, 16, 5, 0, seekable=True) if checkers.ref_available() else checkers.oracle_encode(x, 48000, 16, 5, 0)
    assert data == want and open(flac, "rb").read() == want
    assert data == checkers.oracle_encode(x, 48000, 16, 5, 0)
    dec, info = checkers.oracle_decode(data)
    assert np.array_equal(dec, x.astype(np.int32))
    out_wav = str(tmp_path / "c1_back.wav")
    y, sr = pf.FileDecoder(flac, out_wav).process()
    assert sr == 48000 and y.shape == x.shape and np.array_equal(np.round(y * 32768.0).astype(np.int32), x.astype(np.int32))


def test_cli_round_trip(checkers, tmp_path):
    """`python -m pyflac_b200 in.wav -o out.flac` then `... out.flac -o back.wav` (reference pyflac/__main__.py:20-57):
    the .flac equals libFLAC's for the same level / blocksize, the WAV that comes back holds the same samples; a file
    that is neither RIFF nor fLaC is refused."""
    import subprocess
    import sys
    from pyflac_b200 import wav as pwav
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    x = music_like(4096 * 4 + 321, 2, 44100, 16, seed=77)
    wav, flac, back = str(tmp_path / "a.wav"), str(tmp_path / "a.flac"), str(tmp_path / "b.wav")
    write_wav(wav, x, 44100, 16)
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run([sys.executable, "-m", "pyflac_b200", wav, "-o", flac, "-c", "8", "-b", "1152"], capture_output=True, text=True, env=env, cwd=root)
    assert r.returncode == 0, r.stderr
    assert open(flac, "rb").read() == checkers.oracle_encode(x, 44100, 16, 8, 1152)
    r = subprocess.run([sys.executable, "-m", "pyflac_b200", flac, "-o", back], capture_output=True, text=True, env=env, cwd=root)
    assert r.returncode == 0, r.stderr
    y, sr = pwav.read_pcm(back)
    assert sr == 44100 and np.array_equal(y, x)
    # default output name: input with the suffix swapped
    r = subprocess.run([sys.executable, "-m", "pyflac_b200", wav], capture_output=True, text=True, env=env, cwd=root)
    assert r.returncode == 0 and os.path.exists(str(tmp_path / "a.flac"))
    assert open(str(tmp_path / "a.flac"), "rb").read() == checkers.oracle_encode(x, 44100, 16, 5, 0)
    junk = str(tmp_path / "junk.bin")
    open(junk, "wb").write(b"not audio at all")
    r = subprocess.run([sys.executable, "-m", "pyflac_b200", junk], capture_output=True, text=True, env=env, cwd=root)
    assert r.returncode != 0 and "WAV or a FLAC" in r.stderr
