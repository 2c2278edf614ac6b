"""ctypes wrappers around the TEST-ONLY checker libraries.

* ``ref``    -> oracle/_ref/libflacref.so   : the reference's own libFLAC 1.4.3 binary behind a C harness
* ``oracle`` -> oracle/_build/libflac_oracle.so : the plain-C restatement

Nothing in the product package (pyflac_b200/) imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libflacref.so")
ORACLE_SO = os.path.join(ORACLE_DIR, "_build", "libflac_oracle.so")


def build_checkers():
    """(Re)build the checkers; `make ref` is a no-op copy when /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "port", "ref"], check=True,
                   stdout=subprocess.DEVNULL)


class RefEncCfg(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bps", C.c_uint32),
                ("level", C.c_uint32), ("blocksize", C.c_uint32), ("seekable", C.c_int),
                ("limit_min_bitrate", C.c_int), ("streamable_subset", C.c_int), ("do_md5", C.c_int)]


class FoEncCfg(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bps", C.c_uint32),
                ("level", C.c_uint32), ("blocksize", C.c_uint32), ("seekable", C.c_int32),
                ("limit_min_bitrate", C.c_int32), ("streamable_subset", C.c_int32)]


FO_MAX_CH, FO_MAX_LPC, FO_MAX_PO, FO_MAX_APOD = 8, 32, 8, 16


class FoSubframe(C.Structure):
    _fields_ = [("type", C.c_int32), ("order", C.c_int32), ("wasted", C.c_int32), ("sbps", C.c_int32),
                ("precision", C.c_int32), ("shift", C.c_int32), ("qlp", C.c_int32 * FO_MAX_LPC),
                ("partition_order", C.c_int32), ("rice2", C.c_int32), ("rice", C.c_uint32 * (1 << FO_MAX_PO)),
                ("bits_est", C.c_uint32)]


class FoSignalTrace(C.Structure):
    _fields_ = [("wasted", C.c_int32), ("sbps", C.c_int32), ("fixed_err", C.c_uint64 * 5),
                ("fixed_order", C.c_int32), ("fixed_bits", C.c_uint32), ("is_constant", C.c_int32),
                ("n_apod", C.c_int32),
                ("autoc", (C.c_double * (FO_MAX_LPC + 1)) * FO_MAX_APOD),
                ("lpc_err", (C.c_double * FO_MAX_LPC) * FO_MAX_APOD),
                ("lpc_order", C.c_int32 * FO_MAX_APOD), ("lpc_bits", C.c_uint32 * FO_MAX_APOD),
                ("best", FoSubframe)]


class FoFrameTrace(C.Structure):
    _fields_ = [("blocksize", C.c_uint32), ("frame_number", C.c_uint32), ("channel_assignment", C.c_int32),
                ("n_signals", C.c_int32), ("sig", FoSignalTrace * (FO_MAX_CH + 2))]


_ref = None
_oracle = None


def ref_available():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        L = C.CDLL(REF_SO)
        L.ref_vendor_string.restype = C.c_char_p
        L.ref_encode_stream.restype = C.c_long
        L.ref_encode_stream.argtypes = [C.POINTER(RefEncCfg), C.c_void_p, C.c_uint64, C.c_uint32,
                                        C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_uint32, C.POINTER(C.c_uint32)]
        L.ref_decode_stream.restype = C.c_long
        L.ref_decode_stream.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ref_encode_mt.restype = C.c_double
        L.ref_encode_mt.argtypes = [C.POINTER(RefEncCfg), C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32,
                                    C.c_uint32, C.POINTER(C.c_uint64), C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_decode_mt.restype = C.c_double
        L.ref_decode_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32,
                                    C.POINTER(C.c_uint64)]
        _ref = L
    return _ref


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build_checkers()
        L = C.CDLL(ORACLE_SO)
        L.fo_encode_stream.restype = C.c_long
        L.fo_encode_stream.argtypes = [C.POINTER(FoEncCfg), C.c_void_p, C.c_uint64, C.c_void_p, C.c_size_t,
                                       C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32),
                                       C.c_void_p, C.c_uint32]
        L.fo_encoder_init_status.restype = C.c_int
        L.fo_encoder_init_status.argtypes = [C.POINTER(FoEncCfg), C.c_int, C.c_int, C.c_int]
        L.fo_decode_stream.restype = C.c_long
        L.fo_decode_stream.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64, C.c_void_p]
        L.fo_window_tukey.argtypes = [C.c_void_p, C.c_int32, C.c_float]
        L.fo_md5.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.fo_crc8.restype = C.c_uint8
        L.fo_crc8.argtypes = [C.c_void_p, C.c_size_t]
        L.fo_crc16.restype = C.c_uint16
        L.fo_crc16.argtypes = [C.c_void_p, C.c_size_t]
        _oracle = L
    return _oracle


def _as_i32(pcm):
    pcm = np.ascontiguousarray(pcm)
    if pcm.ndim == 1:
        pcm = pcm[:, None]
    return np.ascontiguousarray(pcm.astype(np.int32)), pcm.shape[0], pcm.shape[1]


def ref_encode(pcm, sample_rate, bps, level=5, blocksize=0, seekable=True, chunk=0,
               limit_min_bitrate=False, streamable_subset=True, with_index=False):
    """libFLAC bytes for one stream. pcm: (n,) or (n, ch) integer array (values must fit `bps`)."""
    x, n, ch = _as_i32(pcm)
    cfg = RefEncCfg(sample_rate, ch, bps, level, blocksize, int(seekable), int(limit_min_bitrate),
                    int(streamable_subset), 1)
    cap = x.size * 5 + 65536
    out = np.empty(cap, np.uint8)
    maxf = n // 16 + 8
    off = np.zeros(maxf, np.uint64)
    ln = np.zeros(maxf, np.uint32)
    smp = np.zeros(maxf, np.uint32)
    nf = C.c_uint32(0)
    r = ref_lib().ref_encode_stream(C.byref(cfg), x.ctypes.data, n, chunk, out.ctypes.data, cap,
                                    off.ctypes.data, ln.ctypes.data, smp.ctypes.data, maxf, C.byref(nf))
    if r < 0:
        raise RuntimeError(f"ref_encode_stream failed: {r}")
    b = out[:r].tobytes()
    if with_index:
        k = nf.value
        return b, off[:k].copy(), ln[:k].copy(), smp[:k].copy()
    return b


def ref_decode(data, max_samples=None):
    """libFLAC decode of a .flac byte string -> (int32 array (n, ch), info dict)."""
    buf = np.frombuffer(data, np.uint8)
    info = np.zeros(4, np.uint32)
    if max_samples is None:
        n = ref_lib().ref_decode_stream(buf.ctypes.data, buf.size, None, 0, info.ctypes.data)
        if n < 0:
            raise RuntimeError(f"ref_decode_stream failed: {n}")
        max_samples = n
    ch = int(info[0]) or 8
    out = np.zeros((max_samples, ch), np.int32)
    n = ref_lib().ref_decode_stream(buf.ctypes.data, buf.size, out.ctypes.data, max_samples, info.ctypes.data)
    if n < 0:
        raise RuntimeError(f"ref_decode_stream failed: {n}")
    return out[:n], dict(channels=int(info[0]), bps=int(info[1]), sample_rate=int(info[2]), errors=int(info[3]))


def oracle_encode(pcm, sample_rate, bps, level=5, blocksize=0, seekable=True, limit_min_bitrate=False,
                  streamable_subset=True, with_index=False, with_trace=False):
    x, n, ch = _as_i32(pcm)
    cfg = FoEncCfg(sample_rate, ch, bps, level, blocksize, int(seekable), int(limit_min_bitrate),
                   int(streamable_subset))
    cap = x.size * 5 + 65536
    out = np.empty(cap, np.uint8)
    maxf = n // 16 + 8
    off = np.zeros(maxf, np.uint64)
    ln = np.zeros(maxf, np.uint32)
    nf = C.c_uint32(0)
    traces = None
    tcap = 0
    if with_trace:
        tcap = min(maxf, with_trace if isinstance(with_trace, int) and with_trace > 1 else maxf)
        traces = (FoFrameTrace * tcap)()
    r = oracle_lib().fo_encode_stream(C.byref(cfg), x.ctypes.data, n, out.ctypes.data, cap,
                                      off.ctypes.data, ln.ctypes.data, maxf, C.byref(nf),
                                      C.cast(traces, C.c_void_p) if traces is not None else None, tcap)
    if r < 0:
        raise RuntimeError(f"fo_encode_stream failed: {r}")
    b = out[:r].tobytes()
    res = [b]
    if with_index:
        res += [off[:nf.value].copy(), ln[:nf.value].copy()]
    if with_trace:
        res.append(traces)
    return res[0] if len(res) == 1 else tuple(res)


def oracle_decode(data):
    buf = np.frombuffer(data, np.uint8)
    info = np.zeros(4, np.uint32)
    n = oracle_lib().fo_decode_stream(buf.ctypes.data, buf.size, None, 0, info.ctypes.data)
    if n < 0:
        raise RuntimeError(f"fo_decode_stream failed: {n}")
    ch = int(info[0]) or 1
    out = np.zeros((n, ch), np.int32)
    n2 = oracle_lib().fo_decode_stream(buf.ctypes.data, buf.size, out.ctypes.data, n, info.ctypes.data)
    if n2 != n:
        raise RuntimeError(f"fo_decode_stream failed: {n2}")
    return out, dict(channels=int(info[0]), bps=int(info[1]), sample_rate=int(info[2]), blocksize=int(info[3]))


def ref_encode_mt(pcm, sample_rate, bps, level=5, blocksize=0, n_threads=1, keep_bytes=False):
    """Time libFLAC on `n_threads` pthreads over a batch pcm[(streams, n, ch)] (int16 or int32 container).
    Returns (seconds, total_bytes, list_of_bytes or None)."""
    pcm = np.ascontiguousarray(pcm)
    ns, n, ch = pcm.shape
    cfg = RefEncCfg(sample_rate, ch, bps, level, blocksize, 1, 0, 1, 1)
    total = C.c_uint64(0)
    p16 = pcm.ctypes.data if pcm.dtype == np.int16 else None
    p32 = pcm.ctypes.data if pcm.dtype == np.int32 else None
    assert p16 or p32
    out_all = lens = None
    stride = 0
    if keep_bytes:
        stride = n * ch * pcm.dtype.itemsize + n * ch // 2 + 65536
        out_all = np.empty(ns * stride, np.uint8)
        lens = np.zeros(ns, np.uint64)
    dt = ref_lib().ref_encode_mt(C.byref(cfg), p32, p16, n, ns, n_threads, C.byref(total),
                                 out_all.ctypes.data if keep_bytes else None, stride,
                                 lens.ctypes.data if keep_bytes else None)
    if dt < 0:
        raise RuntimeError("ref_encode_mt failed")
    blobs = None
    if keep_bytes:
        blobs = [out_all[s * stride: s * stride + int(lens[s])] for s in range(ns)]
    return dt, int(total.value), blobs


def ref_decode_mt(blob, offs, lens, n_threads):
    """Time libFLAC decode of many .flac byte strings (uint8 array + uint64 offsets/lengths) on n_threads pthreads.
    Returns (seconds, inter-channel samples decoded)."""
    blob = np.ascontiguousarray(blob, np.uint8)
    offs = np.ascontiguousarray(offs, np.uint64)
    lens = np.ascontiguousarray(lens, np.uint64)
    total = C.c_uint64(0)
    dt = ref_lib().ref_decode_mt(blob.ctypes.data, offs.ctypes.data, lens.ctypes.data, len(offs), n_threads, C.byref(total))
    if dt < 0:
        raise RuntimeError("ref_decode_mt failed")
    return dt, int(total.value)
