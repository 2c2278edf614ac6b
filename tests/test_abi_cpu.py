"""CPU tests of the boundary: libflacb200.so loads without a GPU, exports every symbol include/*.h declares,
validates settings like libFLAC, refuses to compute without a device (no CPU fallback), and the host build of the
decision arithmetic (fb_math.cuh) agrees with the oracle."""
import ctypes as C
import json
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden_cases, load_golden


@pytest.fixture(scope="module")
def lib():
    from pyflac_b200 import build, _native
    build.build()
    return _native.lib()


def declared_symbols():
    syms = set()
    for h in os.listdir(os.path.join(ROOT, "include")):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        syms |= set(re.findall(r"\b((?:flacb200|FLAC__)\w+)\s*\(", txt))
        syms |= set(re.findall(r"extern\s+[\w\s\*]+?\b((?:flacb200|FLAC__)\w+)\s*(?:\[\])?\s*;", txt))
    types = {"FLAC__bool", "FLAC__byte", "FLAC__int32", "FLAC__uint64"}
    return {s for s in syms if not s.endswith("Callback") and s not in types}


def test_library_exports_every_declared_symbol(lib):
    from pyflac_b200 import _native
    out = subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if l.strip()}
    missing = sorted(s for s in declared_symbols() if s not in exported)
    assert not missing, missing


def test_no_device_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pyflac_b200 import _native
    with pytest.raises(_native.NoCudaDevice):
        _native.Engine(0)


def test_validate_matches_oracle(lib, checkers):
    from pyflac_b200 import _native as nat
    L = checkers.oracle_lib()
    for sr, ch, bps, lvl, bs, sub in [(48000, 2, 16, 5, 0, 1), (2000000, 2, 16, 5, 0, 1), (48000, 2, 16, 5, 1000000, 1),
                                      (48000, 2, 16, 5, 65535, 1), (48000, 2, 16, 5, 65535, 0), (48000, 9, 16, 5, 0, 1),
                                      (48000, 2, 3, 5, 0, 1), (48000, 2, 17, 5, 0, 1), (48000, 2, 17, 5, 0, 0),
                                      (96000, 2, 16, 5, 16385, 1), (48000, 2, 16, 8, 8, 0), (48000, 2, 16, 0, 15, 0)]:
        cfg = nat.EncConfig(sr, ch, bps, lvl, bs, 2, 1, 1, sub, 0)
        ocfg = checkers.FoEncCfg(sr, ch, bps, lvl, bs, 0, 0, sub)
        assert lib.flacb200_enc_validate(C.byref(cfg)) == L.fo_encoder_init_status(C.byref(ocfg), 1, 0, 0), (sr, ch, bps, lvl, bs, sub)


@pytest.fixture(scope="module")
def fbmath():
    d = os.path.join(ROOT, "tests", "cpu_shim")
    so = os.path.join(d, "libfbmath_host.so")
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-std=c++17", "-o", so,
                    os.path.join(d, "fb_math_host.cpp")], check=True)
    L = C.CDLL(so)
    L.t_crc16_mulmod.restype = C.c_uint16
    L.t_crc16_byte.restype = C.c_uint16
    L.t_crc16_chunked.restype = C.c_uint16
    L.t_crc16_chunked.argtypes = [C.c_void_p, C.c_uint32]
    L.t_crc16_plain.restype = C.c_uint16
    L.t_crc16_plain.argtypes = [C.c_void_p, C.c_uint32]
    L.t_rice_parameter.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
    L.t_rice_bits.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
    return L


def test_fb_math_host_build_matches_oracle_traces(fbmath, checkers):
    """frame headers, Levinson errors and quantised coefficients of the kernel's scalar code (host build)"""
    maxlpc = {0: 0, 1: 0, 2: 0, 3: 6, 4: 8, 5: 8, 6: 8, 7: 12, 8: 12}
    nlev = nq = 0
    for case in golden_cases():
        x, flac = load_golden(case)
        _, traces = checkers.oracle_encode(x, case["sample_rate"], case["bps"], case["level"], case["blocksize"], with_trace=True)
        for f, o in enumerate(case["frame_off"]):
            t = traces[f]
            hdr = np.zeros(16, np.uint8)
            n = fbmath.t_frame_header(hdr.ctypes.data, case["channels"], case["bps"], case["sample_rate"], t.blocksize,
                                      t.frame_number, t.channel_assignment)
            assert hdr[:n].tobytes() == flac[o:o + n]
            for s in range(t.n_signals):
                sg = t.sig[s]
                b = sg.best
                first = next((k for k in range(sg.n_apod) if sg.lpc_bits[k] == b.bits_est), -1)
                for st in range(sg.n_apod):
                    if sg.lpc_order[st] == 0:
                        continue
                    mo = min(maxlpc[case["level"]], t.blocksize - 1)
                    ac = np.array(list(sg.autoc[st])[:mo + 1])
                    lp = np.zeros(144, np.float32)
                    err = np.zeros(12)
                    nm = fbmath.t_levinson(ac.ctypes.data, mo, lp.ctypes.data, err.ctypes.data)
                    assert np.array_equal(err[:nm], np.array(list(sg.lpc_err[st])[:nm]))
                    nlev += 1
                    if b.type == 3 and st == first:
                        q = np.zeros(16, np.int32)
                        sh = C.c_int(0)
                        rc = fbmath.t_quantize(lp[(b.order - 1) * 12:].ctypes.data, b.order, b.precision, q.ctypes.data, C.byref(sh))
                        assert rc == 0 and sh.value == b.shift and list(q[:b.order]) == list(b.qlp)[:b.order]
                        nq += 1
    assert nlev > 100 and nq > 30


def test_rice_parameter_fixed_point_quirk(fbmath):
    """n=4095, sum=65536 -> k=4 (truncated 18-bit reciprocal), not the exact ceil-log2's 5"""
    assert fbmath.t_rice_parameter(65536, 4095, 15) == 4
    assert fbmath.t_rice_parameter(0, 128, 15) == 0 and fbmath.t_rice_parameter(1, 128, 15) == 0
    assert fbmath.t_rice_parameter(1 << 40, 128, 15) == 14 and fbmath.t_rice_parameter(1 << 40, 128, 31) == 30
    assert fbmath.t_rice_bits(0, 100, 10) == 4 + 100 + 20 - 50


def test_crc16_chunk_combination(fbmath):
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, 256, 300).astype(np.uint8), rng.integers(0, 256, 77).astype(np.uint8)

    def crc(bs):
        c = 0
        for v in bs:
            c = fbmath.t_crc16_byte(c, int(v))
        return c
    xp, base, e = 1, 2, 8 * len(b)
    while e:
        if e & 1:
            xp = fbmath.t_crc16_mulmod(xp, base)
        base = fbmath.t_crc16_mulmod(base, base)
        e >>= 1
    assert crc(np.concatenate([a, b])) == (fbmath.t_crc16_mulmod(crc(a), xp) ^ crc(b))
    assert crc(b"123456789") == 0xFEE8


def test_host_multibuffer_md5_matches_hashlib():
    """the host->host encode path hashes the caller's PCM with 16-lane AVX-512 / 8-lane AVX2 / scalar MD5 (csrc/md5_mb.h,
    md5_host.h): every path this CPU has against hashlib, ragged lengths incl. 0, 55, 56, 63, 64, 65 and multi-block"""
    import hashlib
    d = os.path.join(ROOT, "tests", "cpu_shim")
    so = os.path.join(d, "libmd5_host_shim.so")
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-o", so, os.path.join(d, "md5_host_shim.cpp")], check=True)
    L = C.CDLL(so)
    L.t_md5_group.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    widest = L.t_md5_lanes()
    rng = np.random.default_rng(5)
    lens_sets = [[0, 1, 55, 56, 57, 63, 64, 65, 119, 120, 128, 1000, 4096, 4097, 100000, 7],
                 [1920000 // 8] * 16, [64 * 50] * 5 + [64 * 50 + 3] * 3, [200, 100, 300]]
    for lens in lens_sets:
        for n in sorted({len(lens), min(len(lens), 8), 1}):
            ls = np.array(lens[:n], np.uint64)
            off = np.concatenate([[0], np.cumsum(ls)[:-1]]).astype(np.uint64)
            blob = rng.integers(0, 256, int(ls.sum()) + 64, dtype=np.uint8)
            want = b"".join(hashlib.md5(blob[int(o):int(o + l)].tobytes()).digest() for o, l in zip(off, ls))
            for lanes in (16, 8, 1):
                if lanes > widest or (lanes == 8 and n > 8):
                    continue
                got = np.zeros(16 * n, np.uint8)
                assert L.t_md5_group(blob.ctypes.data, off.ctypes.data, ls.ctypes.data, n, lanes, got.ctypes.data) == 0
                assert got.tobytes() == want, (lens[:n], lanes)


def test_crc16_chunk_weights_any_frame_size(fbmath):
    """dec_crc_kernel's combination of per-chunk CRCs (pyflac_b200/csrc/fb_math.cuh: crc16_weigh_chunk) equals the bytewise CRC-16
    for every size, including frames of 256 KiB and more, where the chunk number runs past the two weight tables (a 65535-sample
    verbatim stereo frame is 262 160 bytes)."""
    rng = np.random.default_rng(5)
    for n in (1, 63, 64, 65, 4096, 16383, 16384 + 7, 262143, 262144, 262145, 262144 + 64 * 5 + 3, 600001, 1 << 20, (1 << 21) + 12345, 5000000):
        buf = rng.integers(0, 256, n).astype(np.uint8)
        assert fbmath.t_crc16_chunked(buf.ctypes.data, n) == fbmath.t_crc16_plain(buf.ctypes.data, n), n
