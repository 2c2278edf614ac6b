"""CPU tests of host-side logic: WAV I/O, stream sharding + rank reductions over gloo (world_size 2), package surface."""
import os
import struct
import subprocess
import sys
import tempfile

import numpy as np

from conftest import ROOT


def test_wav_reader_writer_roundtrip():
    from pyflac_b200 import wav
    rng = np.random.default_rng(0)
    x = rng.integers(-30000, 30000, (1000, 2)).astype(np.int16)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "a.wav")
        w = wav.Pcm16Writer(p, 44100, 2)
        w.write(x[:400]); w.write(x[400:]); w.close()
        i = wav.info(p)
        assert (i.samplerate, i.channels, i.frames, i.subtype) == (44100, 2, 1000, "PCM_16")
        y, sr = wav.read_pcm(p)
        assert sr == 44100 and np.array_equal(x, y)
        f, _ = wav.read_float64(p)
        assert f.dtype == np.float64 and np.array_equal(np.rint(f * 32768).astype(np.int16), x)
        # WAVE_FORMAT_EXTENSIBLE + extra chunk before data, 32-bit
        x32 = rng.integers(-2**31, 2**31 - 1, (50, 1)).astype(np.int32)
        fmt = struct.pack("<HHIIHH", 0xFFFE, 1, 48000, 48000 * 4, 4, 32) + struct.pack("<HHI", 22, 32, 4) + \
            struct.pack("<H", 1) + b"\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71"
        raw = x32.astype("<i4").tobytes()
        body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"LIST" + struct.pack("<I", 4) + b"abcd" + b"data" + struct.pack("<I", len(raw)) + raw
        with open(p, "wb") as fh:
            fh.write(b"RIFF" + struct.pack("<I", len(body)) + body)
        y, sr = wav.read_pcm(p)
        assert sr == 48000 and np.array_equal(y, x32)
        # 8-bit input is rejected like the reference (encoder.py:372-378)
        fmt8 = struct.pack("<HHIIHH", 1, 1, 8000, 8000, 1, 8)
        body = b"WAVE" + b"fmt " + struct.pack("<I", 16) + fmt8 + b"data" + struct.pack("<I", 4) + b"\x80\x80\x80\x80"
        with open(p, "wb") as fh:
            fh.write(b"RIFF" + struct.pack("<I", len(body)) + body)
        try:
            wav.read_pcm(p)
            assert False
        except ValueError:
            pass


def test_shard_range_covers_everything():
    from pyflac_b200.dist import shard_range
    for n in (0, 1, 7, 256, 32768):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import torch, torch.distributed as dist
from pyflac_b200.dist import shard_range, reduce_max, reduce_sum
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
lo, hi = shard_range(257, r, w)
mx = reduce_max([10.0 + r, float(hi - lo)])
sm = reduce_sum([float(hi - lo)])
dist.barrier()
if r == 0:
    print("RESULT", mx[0], mx[1], sm[0])
dist.destroy_process_group()
'''


def test_rank_reductions_gloo_world2():
    """the N>1 path of bench.py (barrier, max-over-ranks timing, summed work) on the gloo backend"""
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(_WORKER)
        path = f.name
    try:
        env = dict(os.environ, MASTER_ADDR="127.0.0.1")
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                              "--master-port", "29577", path, ROOT], capture_output=True, text=True, timeout=240, env=env)
        line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
        assert line, out.stdout + out.stderr
        _, a, b, c = line[0].split()
        assert float(a) == 11.0 and float(b) == 129.0 and float(c) == 257.0
    finally:
        os.unlink(path)


_WORKER_SG = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch, torch.distributed as dist
from pyflac_b200.dist import scatter_streams, gather_packed, shard_range
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
n, e = 7, 50                                            # 7 streams over 2 ranks: blocks of 4 and 3
full = torch.arange(n * e, dtype=torch.int16).reshape(n, e) if r == 0 else None
blk, (lo, hi) = scatter_streams(full, n, e, torch.int16, "cpu")
ok = bool((blk == torch.arange(n * e, dtype=torch.int16).reshape(n, e)[lo:hi]).all()) and (lo, hi) == shard_range(n, r, w)
payload = (blk.reshape(-1).to(torch.int32) % 251).to(torch.uint8)[: 100 + 37 * r]      # "packed bytes", different length per rank
bufs, sizes = gather_packed(torch.cat([payload, torch.zeros(64, dtype=torch.uint8)]), payload.numel(), "cpu")
if r == 0:
    exp1 = (torch.arange(n * e, dtype=torch.int16).reshape(n, e)[4:7].reshape(-1).to(torch.int32) % 251).to(torch.uint8)[:137]
    ok = ok and sizes == [100, 137] and bool((bufs[0] == payload).all()) and bool((bufs[1] == exp1).all())
flag = torch.tensor([1.0 if ok else 0.0]); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if r == 0:
    print("RESULT", float(flag))
dist.destroy_process_group()
'''


def test_scatter_gather_gloo_world2():
    """pyflac_b200.dist.scatter_streams / gather_packed (the NCCL scatter of PCM born on one GPU and the gather of the
    packed bytes) on the gloo backend with uneven blocks and ragged payloads"""
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(_WORKER_SG)
        path = f.name
    try:
        env = dict(os.environ, MASTER_ADDR="127.0.0.1")
        out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                              "--master-port", "29578", path, ROOT], capture_output=True, text=True, timeout=240, env=env)
        line = [l for l in out.stdout.splitlines() if l.startswith("RESULT")]
        assert line, out.stdout + out.stderr
        assert float(line[0].split()[1]) == 1.0
    finally:
        os.unlink(path)


def test_package_surface_matches_reference_names():
    import pyflac_b200 as pf
    for name in ["StreamEncoder", "FileEncoder", "EncoderState", "EncoderInitException", "EncoderProcessException",
                 "StreamDecoder", "FileDecoder", "OneShotDecoder", "DecoderState", "DecoderInitException", "DecoderProcessException"]:
        assert hasattr(pf, name)
    assert str(pf.EncoderState.UNINITIALIZED) == "FLAC__STREAM_ENCODER_UNINITIALIZED"
    assert str(pf.DecoderState.UNINITIALIZED) == "FLAC__STREAM_DECODER_UNINITIALIZED"
    assert str(pf.EncoderInitException(3)) == "FLAC__STREAM_ENCODER_INIT_STATUS_INVALID_CALLBACKS"
    assert str(pf.DecoderInitException(4)) == "FLAC__STREAM_DECODER_INIT_STATUS_ERROR_OPENING_FILE"


def test_encoder_and_decoder_properties_without_a_device():
    """reference tests/test_encoder.py:32-93 and tests/test_decoder.py:37-40: property setters / getters and the state of a
    handle that was never initialised need no GPU; initialising one without a device fails loudly (no CPU fallback)"""
    import numpy as np
    import pytest
    import pyflac_b200 as pf
    from pyflac_b200.encoder import _Encoder
    from pyflac_b200.decoder import _Decoder
    e = _Encoder()
    assert e._limit_min_bitrate is False or e._limit_min_bitrate == 0
    for name, val in [("_verify", True), ("_channels", 2), ("_bits_per_sample", 24), ("_sample_rate", 48000), ("_blocksize", 128),
                      ("_streamable_subset", False), ("_limit_min_bitrate", True)]:
        setattr(e, name, val)
        assert getattr(e, name) == val
    e._compression_level = 8
    assert e.state == pf.EncoderState.UNINITIALIZED and str(e.state) == "FLAC__STREAM_ENCODER_UNINITIALIZED"
    with pytest.raises(TypeError):
        e.process([1, 2, 3, 4])
    d = _Decoder()
    assert d.state == pf.DecoderState.UNINITIALIZED and str(d.state) == "FLAC__STREAM_DECODER_UNINITIALIZED"
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if not has_gpu:
        with pytest.raises(pf.EncoderInitException) as ei:
            pf.StreamEncoder(48000, lambda *a: None).process(np.zeros((4096, 2), np.int16))
        assert "ENCODER_ERROR" in str(ei.value)


def test_wav_reader_on_reference_wavs_when_present():
    """pyFLAC's tests/data/*.wav (PCM_16 mono/stereo, WAVE_FORMAT_EXTENSIBLE 5.1, PCM_32) through the soundfile-free reader:
    the samples hash to the MD5 recorded in tests/golden/fixtures (== STREAMINFO MD5 of the matching .flac)"""
    import hashlib
    import numpy as np
    import pytest
    from conftest import fixture_cases
    from pyflac_b200 import wav
    src = "/root/reference/tests/data"
    if not os.path.isdir(src):
        pytest.skip("reference WAV files not present")
    for c in fixture_cases():
        if not c["wav_md5"]:
            continue
        x, sr = wav.read_pcm(os.path.join(src, c["name"] + ".wav"))
        assert sr == c["sample_rate"] and x.shape == (c["samples"], c["channels"])
        assert x.dtype == (np.int16 if c["bps"] == 16 else np.int32)
        assert hashlib.md5(np.ascontiguousarray(x).tobytes()).hexdigest() == c["wav_md5"], c["name"]
        y, sr2 = wav.read_float64(os.path.join(src, c["name"] + ".wav"))
        assert sr2 == sr and y.dtype == np.float64 and y.shape == x.shape and float(np.max(np.abs(y))) <= 1.0


def test_public_signatures_match_reference_source_when_present():
    """constructor and method signatures (names, order, defaults) of the five public classes, read from the reference's
    source with `ast` (no GPU here; on the GPU box tests/test_gpu_reference_suite.py builds the reference's cffi modules against
    libflacb200.so and runs its unmodified test-suite)"""
    import ast
    import inspect
    import pytest
    import pyflac_b200 as pf
    files = {"/root/reference/pyflac/encoder.py": ["StreamEncoder", "FileEncoder"],
             "/root/reference/pyflac/decoder.py": ["StreamDecoder", "FileDecoder", "OneShotDecoder"]}
    if not all(os.path.exists(f) for f in files):
        pytest.skip("reference source not present")
    for path, classes in files.items():
        tree = ast.parse(open(path).read())
        for node in tree.body:
            if not (isinstance(node, ast.ClassDef) and node.name in classes):
                continue
            ours = getattr(pf, node.name)
            for fn in node.body:
                if not isinstance(fn, ast.FunctionDef) or (fn.name.startswith("_") and fn.name != "__init__"):
                    continue
                assert hasattr(ours, fn.name), (node.name, fn.name)
                sig = inspect.signature(getattr(ours, fn.name))
                ref_names = [a.arg for a in fn.args.args]
                assert list(sig.parameters)[:len(ref_names)] == ref_names, (node.name, fn.name, list(sig.parameters), ref_names)
                n_def = len(fn.args.defaults)
                for a, d in zip(fn.args.args[len(fn.args.args) - n_def:], fn.args.defaults):
                    want = ast.literal_eval(d)
                    assert sig.parameters[a.arg].default == want, (node.name, fn.name, a.arg)
    # the private base classes carry the properties pyFLAC's own tests poke (tests/test_encoder.py:32-93)
    from pyflac_b200.encoder import _Encoder
    from pyflac_b200.decoder import _Decoder
    for path, cls, ours in [("/root/reference/pyflac/encoder.py", "_Encoder", _Encoder), ("/root/reference/pyflac/decoder.py", "_Decoder", _Decoder)]:
        node = next(n for n in ast.parse(open(path).read()).body if isinstance(n, ast.ClassDef) and n.name == cls)
        for fn in node.body:
            if isinstance(fn, ast.FunctionDef) and not fn.name.startswith("__"):
                assert hasattr(ours, fn.name), (cls, fn.name)
                is_prop = any(isinstance(d, ast.Name) and d.id == "property" for d in fn.decorator_list)
                if is_prop:
                    assert isinstance(inspect.getattr_static(ours, fn.name), property), (cls, fn.name)
