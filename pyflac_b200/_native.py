"""ctypes binding of libflacb200.so (include/flacb200.h).

The library is built in-tree by pyflac_b200/build.py (nvcc, sm_100a).  Importing this module never
touches CUDA; creating an engine (``Engine()``) does and raises ``NoCudaDevice`` when no GPU is
usable -- there is deliberately no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libflacb200.so")


class NoCudaDevice(RuntimeError):
    pass


class NativeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"flacb200 error {code}: {msg}")
        self.code = code


class EncConfig(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bits_per_sample", C.c_uint32),
                ("compression_level", C.c_uint32), ("blocksize", C.c_uint32), ("container_bytes", C.c_uint32),
                ("write_prologue", C.c_uint32), ("do_md5", C.c_uint32), ("streamable_subset", C.c_uint32),
                ("debug_trace", C.c_uint32), ("limit_min_bitrate", C.c_uint32),
                ("tune", C.c_uint32), ("do_mid_side", C.c_uint32), ("loose_mid_side", C.c_uint32), ("max_lpc_order", C.c_uint32),
                ("qlp_coeff_precision", C.c_uint32), ("max_residual_partition_order", C.c_uint32), ("apod_parts", C.c_uint32),
                ("apod_p", C.c_float)]


class StreamInfo(C.Structure):
    _fields_ = [("total_samples", C.c_uint64), ("byte_off", C.c_uint64), ("byte_len", C.c_uint64),
                ("min_framesize", C.c_uint32), ("max_framesize", C.c_uint32), ("n_frames", C.c_uint32),
                ("pad", C.c_uint32), ("md5", C.c_uint8 * 16)]


class EncResult(C.Structure):
    _fields_ = [("total_bytes", C.c_uint64), ("n_frames", C.c_uint32), ("n_streams", C.c_uint32),
                ("log_guard_hits", C.c_uint64), ("d_arena", C.c_void_p), ("d_frame_off", C.c_void_p),
                ("d_frame_len", C.c_void_p)]


K_MAX_ORDER, K_MAX_PARTS, K_MAX_APOD, K_MAX_LAGS = 12, 64, 9, 13


class SubframePlan(C.Structure):
    _fields_ = [("type", C.c_uint8), ("order", C.c_uint8), ("wasted", C.c_uint8), ("sbps", C.c_uint8),
                ("precision", C.c_uint8), ("part_order", C.c_uint8), ("rice2", C.c_uint8), ("pad0", C.c_uint8),
                ("shift", C.c_int32), ("bits_est", C.c_uint32), ("qlp", C.c_int32 * K_MAX_ORDER),
                ("rice", C.c_uint8 * K_MAX_PARTS)]


class SignalDebug(C.Structure):
    _fields_ = [("fixed_err", C.c_uint64 * 5), ("fixed_order", C.c_int32), ("fixed_bits", C.c_uint32),
                ("is_constant", C.c_int32), ("n_apod", C.c_int32),
                ("autoc", (C.c_double * (K_MAX_LAGS + 1)) * K_MAX_APOD),
                ("lpc_err", (C.c_double * K_MAX_ORDER) * K_MAX_APOD),
                ("lpc_order", C.c_int32 * K_MAX_APOD), ("lpc_bits", C.c_uint32 * K_MAX_APOD)]


assert C.sizeof(SubframePlan) == 128

_lib = None


def lib():
    """Load libflacb200.so (building it first if the sources are newer and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if _build.needs_build():
        # sources newer than the library (or no library): rebuild when nvcc is here; a snapshot shipped to a box
        # without nvcc keeps the library it came with
        if os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")) or not os.path.exists(LIB_PATH):
            _build.build()
    L = C.CDLL(LIB_PATH)
    L.flacb200_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    L.flacb200_destroy.argtypes = [C.c_void_p]
    L.flacb200_last_error.restype = C.c_char_p
    L.flacb200_last_error.argtypes = [C.c_void_p]
    L.flacb200_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.flacb200_sync.argtypes = [C.c_void_p]
    L.flacb200_join.argtypes = [C.c_void_p]
    L.flacb200_enc_validate.argtypes = [C.POINTER(EncConfig)]
    L.flacb200_encode_batch.argtypes = [C.c_void_p, C.POINTER(EncConfig), C.c_void_p, C.c_int, C.c_uint64, C.c_uint32,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
    L.flacb200_encode_result.argtypes = [C.c_void_p, C.POINTER(EncResult)]
    L.flacb200_encode_result_frames.argtypes = [C.c_void_p, C.POINTER(EncResult)]
    L.flacb200_encode_fetch_md5.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.flacb200_encode_set_prev_assignment.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.flacb200_encode_fetch_assignments.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.flacb200_encode_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
    L.flacb200_encode_fetch_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
    L.flacb200_encode_batch_host.argtypes = [C.c_void_p, C.POINTER(EncConfig), C.c_void_p, C.c_uint64, C.c_uint32,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint64),
                                             C.c_void_p, C.c_void_p, C.c_void_p]
    L.flacb200_host_path_info.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.flacb200_log_guard_info.argtypes = [C.c_void_p, C.c_void_p]
    L.flacb200_set_log_guard.argtypes = [C.c_void_p, C.c_double, C.c_int]
    L.flacb200_encode_fetch_md5_back.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    L.flacb200_encode_host_submit.argtypes = [C.c_void_p, C.POINTER(EncConfig), C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    L.flacb200_encode_host_collect.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64)]
    L.flacb200_set_profiling.argtypes = [C.c_void_p, C.c_int]
    L.flacb200_kernel_times.argtypes = [C.c_void_p, C.c_void_p]
    L.flacb200_launch_count.restype = C.c_uint64
    L.flacb200_launch_count.argtypes = [C.c_void_p]
    _lib = L
    return L


class Engine:
    """One engine == one CUDA device context of the batch encoder/decoder."""

    def __init__(self, device=0):
        self._L = lib()
        h = C.c_void_p()
        rc = self._L.flacb200_create(C.byref(h), device)
        if rc == 1:
            raise NoCudaDevice("libflacb200: no usable CUDA device (this package has no CPU fallback)")
        if rc != 0:
            raise NativeError(rc, "flacb200_create failed")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._L.flacb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise NativeError(rc, self._L.flacb200_last_error(self._h).decode())

    def set_stream(self, cuda_stream_ptr):
        self._check(self._L.flacb200_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        self._check(self._L.flacb200_sync(self._h))

    def join(self):
        """Make the engine stream wait for the side-stream work (MD5, finalize) of all in-flight batches."""
        self._check(self._L.flacb200_join(self._h))

    def set_log_guard(self, rel=1e-12, flip=False):
        """Test hook for the libm-log guard: band width and deliberately wrong device decisions (flacb200_set_log_guard)."""
        self._check(self._L.flacb200_set_log_guard(self._h, float(rel), int(flip)))

    def log_guard_info(self):
        v = np.zeros(4, np.uint64)
        self._check(self._L.flacb200_log_guard_info(self._h, v.ctypes.data))
        return dict(in_band=int(v[0]), confirmed=int(v[1]), overridden=int(v[2]), unchecked=int(v[3]))

    def set_profiling(self, on=True):
        self._check(self._L.flacb200_set_profiling(self._h, int(on)))

    def kernel_times(self):
        """Device ms of the last batch: analysis (its three kernels together), pack, scan, compact, finalize, md5, then
        the analysis split: frame_bits, autoc, analyze."""
        ms = np.zeros(10, np.float32)
        self._check(self._L.flacb200_kernel_times(self._h, ms.ctypes.data))
        d = dict(zip(["analysis", "pack", "scan", "compact", "finalize", "md5", "frame_bits", "autoc", "analyze"], [float(v) for v in ms[:9]]))
        if ms[9] > 0:       # the TMA-staged kernels of csrc/enc_fused.cu ran (16-bit stereo): there is no OR/AND pass
            d.pop("frame_bits")
            d["path"] = "tma"
        return d

    @property
    def launch_count(self):
        return int(self._L.flacb200_launch_count(self._h))

    # ------------------------------------------------------------ encode
    @staticmethod
    def make_config(sample_rate, channels, bits_per_sample, compression_level=5, blocksize=0, container_bytes=None,
                    write_prologue=True, do_md5=True, streamable_subset=True, debug_trace=False, limit_min_bitrate=False, tune=None):
        """tune: None (the level's presets) or a dict with any of do_mid_side, loose_mid_side, max_lpc_order, qlp_coeff_precision,
        max_residual_partition_order, apod_parts, apod_p -- the fields it leaves out keep the level's values."""
        if container_bytes is None:
            container_bytes = 2 if bits_per_sample <= 16 else 4
        cfg = EncConfig(sample_rate, channels, bits_per_sample, compression_level, blocksize, container_bytes,
                        int(write_prologue), int(do_md5), int(streamable_subset), int(debug_trace), int(limit_min_bitrate))
        if tune is not None:
            ms, loose, lpc, po, parts = [(0, 0, 0, 3, 1), (1, 1, 0, 3, 1), (1, 0, 0, 3, 1), (0, 0, 6, 4, 1), (1, 1, 8, 4, 1), (1, 0, 8, 5, 1), (1, 0, 8, 6, 2),
                                         (1, 0, 12, 6, 2), (1, 0, 12, 6, 3)][min(int(compression_level), 8)]
            cfg.tune = 1
            cfg.do_mid_side = int(tune.get("do_mid_side", ms)); cfg.loose_mid_side = int(tune.get("loose_mid_side", loose))
            cfg.max_lpc_order = int(tune.get("max_lpc_order", lpc)); cfg.qlp_coeff_precision = int(tune.get("qlp_coeff_precision", 0))
            cfg.max_residual_partition_order = int(tune.get("max_residual_partition_order", po))
            cfg.apod_parts = int(tune.get("apod_parts", parts)); cfg.apod_p = float(tune.get("apod_p", 0.5))
        return cfg

    def encode_device(self, cfg, pcm_ptr, pcm_elems, stream_off, stream_samples, first_frame_number=None):
        """Asynchronous batch encode of PCM resident in HBM. stream_off/stream_samples: uint64 numpy arrays."""
        so = np.ascontiguousarray(stream_off, np.uint64)
        ss = np.ascontiguousarray(stream_samples, np.uint64)
        ff = None if first_frame_number is None else np.ascontiguousarray(first_frame_number, np.uint32)
        self._keep = (so, ss, ff)
        self._check(self._L.flacb200_encode_batch(self._h, C.byref(cfg), C.c_void_p(pcm_ptr), 1, pcm_elems, len(so),
                                                  so.ctypes.data, ss.ctypes.data, ff.ctypes.data if ff is not None else None))

    def encode_host(self, cfg, pcm, stream_off, stream_samples, first_frame_number=None):
        pcm = np.ascontiguousarray(pcm)
        so = np.ascontiguousarray(stream_off, np.uint64)
        ss = np.ascontiguousarray(stream_samples, np.uint64)
        ff = None if first_frame_number is None else np.ascontiguousarray(first_frame_number, np.uint32)
        self._keep = (pcm, so, ss, ff)
        self._check(self._L.flacb200_encode_batch(self._h, C.byref(cfg), pcm.ctypes.data, 0, pcm.size, len(so),
                                                  so.ctypes.data, ss.ctypes.data, ff.ctypes.data if ff is not None else None))

    def result(self, wait_md5=True):
        """Sizes + device pointers of the last batch.  wait_md5=False returns once the frames, index and prologues are final
        (the MD5 fields of STREAMINFO still zero); fetch_md5() then waits for the digests."""
        r = EncResult()
        fn = self._L.flacb200_encode_result if wait_md5 else self._L.flacb200_encode_result_frames
        self._check(fn(self._h, C.byref(r)))
        self._last_n_streams = int(r.n_streams)
        return r

    def fetch_md5(self):
        """(n_streams, 16) uint8: the STREAMINFO MD5 of every stream of the last batch (waits for the MD5 chain)."""
        n = getattr(self, "_last_n_streams", None)
        if n is None:
            n = int(self.result(wait_md5=False).n_streams)
        out = np.zeros((max(n, 1), 16), np.uint8)
        self._check(self._L.flacb200_encode_fetch_md5(self._h, out.ctypes.data, out.nbytes))
        return out[:n]

    def fetch_md5_back(self, back=1):
        """Digests of the batch `back` batches before the last one (same layout); see flacb200_encode_fetch_md5_back."""
        n = self._last_n_streams
        out = np.zeros((max(n, 1), 16), np.uint8)
        self._check(self._L.flacb200_encode_fetch_md5_back(self._h, int(back), out.ctypes.data, out.nbytes))
        return out[:n]

    def fetch(self):
        """-> dict(arena=uint8 array, frame_off, frame_len, frame_samples, frame_stream, streams=[StreamInfo])"""
        r = self.result()
        arena = np.empty(max(int(r.total_bytes), 1), np.uint8)
        nf, ns = r.n_frames, r.n_streams
        off = np.zeros(max(nf, 1), np.uint64)
        ln = np.zeros(max(nf, 1), np.uint32)
        smp = np.zeros(max(nf, 1), np.uint32)
        stm = np.zeros(max(nf, 1), np.uint32)
        infos = (StreamInfo * max(ns, 1))()
        self._check(self._L.flacb200_encode_fetch(self._h, arena.ctypes.data, arena.size, off.ctypes.data, ln.ctypes.data,
                                                  smp.ctypes.data, stm.ctypes.data, C.cast(infos, C.c_void_p)))
        return dict(arena=arena[:int(r.total_bytes)], frame_off=off[:nf], frame_len=ln[:nf], frame_samples=smp[:nf],
                    frame_stream=stm[:nf], streams=[infos[i] for i in range(ns)], log_guard_hits=int(r.log_guard_hits),
                    total_bytes=int(r.total_bytes))

    def encode_host_to_host(self, cfg, pcm, stream_off, stream_samples, arena_cap=None, arena=None):
        """flacb200_encode_batch_host: PCM in host memory -> complete .flac images in host memory, one synchronous call
        (chunked H2D / kernels / D2H pipeline, MD5 on host threads and/or the GPU).  Returns the dict of fetch() plus
        path_info (flacb200_host_path_info)."""
        pcm = np.ascontiguousarray(pcm)
        so = np.ascontiguousarray(stream_off, np.uint64)
        ss = np.ascontiguousarray(stream_samples, np.uint64)
        ns = len(so)
        nf = int(sum((int(n) + self._blocksize_of(cfg) - 1) // self._blocksize_of(cfg) for n in ss))
        cap = int(arena_cap or (pcm.nbytes * 2 + nf * 64 + ns * 256 + (1 << 20)))
        if arena is None:
            arena = np.empty(cap, np.uint8)
        else:
            cap = arena.size                                       # caller's buffer (e.g. pinned memory)
        off = np.zeros(max(nf, 1), np.uint64)
        ln = np.zeros(max(nf, 1), np.uint32)
        infos = (StreamInfo * max(ns, 1))()
        tot = C.c_uint64(0)
        self._check(self._L.flacb200_encode_batch_host(self._h, C.byref(cfg), pcm.ctypes.data, pcm.size, ns, so.ctypes.data, ss.ctypes.data,
                                                       arena.ctypes.data, cap, C.byref(tot), off.ctypes.data, ln.ctypes.data,
                                                       C.cast(infos, C.c_void_p)))
        v = np.zeros(10, np.float64)
        self._check(self._L.flacb200_host_path_info(self._h, v.ctypes.data, 10))
        return dict(arena=arena[:tot.value], frame_off=off[:nf], frame_len=ln[:nf], streams=[infos[i] for i in range(ns)],
                    total_bytes=int(tot.value),
                    path_info=dict(host_md5_done_ms=v[0], kernels_done_ms=v[2], d2h_done_ms=v[3], total_ms=v[5], gpu_md5_done_ms=v[6],
                                   streams_hashed_on_gpu=int(v[7]), host_md5_threads=int(v[8]), chunks=int(v[9])))

    def submit_host(self, cfg, pcm, stream_off, stream_samples, arena_cap=None):
        """flacb200_encode_host_submit: enqueue one host -> host batch and return a ticket (up to 3 in flight)."""
        pcm = np.ascontiguousarray(pcm)
        so = np.ascontiguousarray(stream_off, np.uint64)
        ss = np.ascontiguousarray(stream_samples, np.uint64)
        ns = len(so)
        nf = int(sum((int(n) + self._blocksize_of(cfg) - 1) // self._blocksize_of(cfg) for n in ss))
        cap = int(arena_cap or (pcm.nbytes * 2 + nf * 64 + ns * 256 + (1 << 20)))
        job = dict(pcm=pcm, so=so, ss=ss, arena=np.empty(cap, np.uint8), off=np.zeros(max(nf, 1), np.uint64), ln=np.zeros(max(nf, 1), np.uint32),
                   infos=(StreamInfo * max(ns, 1))(), nf=nf, ns=ns)
        t = C.c_int(-1)
        self._check(self._L.flacb200_encode_host_submit(self._h, C.byref(cfg), pcm.ctypes.data, pcm.size, ns, so.ctypes.data, ss.ctypes.data,
                                                        job["arena"].ctypes.data, cap, job["off"].ctypes.data, job["ln"].ctypes.data,
                                                        C.cast(job["infos"], C.c_void_p), C.byref(t)))
        if not hasattr(self, "_jobs"):
            self._jobs = {}
        self._jobs[t.value] = job
        return t.value

    def collect_host(self, ticket):
        """flacb200_encode_host_collect: wait for a submitted batch; returns the dict of encode_host_to_host (without path_info)."""
        job = self._jobs.pop(ticket)
        tot = C.c_uint64(0)
        self._check(self._L.flacb200_encode_host_collect(self._h, ticket, C.byref(tot)))
        return dict(arena=job["arena"][:tot.value], frame_off=job["off"][:job["nf"]], frame_len=job["ln"][:job["nf"]],
                    streams=[job["infos"][i] for i in range(job["ns"])], total_bytes=int(tot.value))

    @staticmethod
    def _blocksize_of(cfg):
        if cfg.blocksize:
            return int(cfg.blocksize)
        return 1152 if cfg.compression_level < 3 else 4096

    def fetch_trace(self, n_signals, want_debug=True):
        r = self.result()
        n = r.n_frames * n_signals
        plans = (SubframePlan * max(n, 1))()
        ca = np.zeros(max(r.n_frames, 1), np.uint8)
        dbg = (SignalDebug * max(n, 1))() if want_debug else None
        self._check(self._L.flacb200_encode_fetch_trace(self._h, C.cast(plans, C.c_void_p), C.sizeof(plans), ca.ctypes.data,
                                                        C.cast(dbg, C.c_void_p) if dbg is not None else None,
                                                        C.sizeof(dbg) if dbg is not None else 0))
        return plans, ca[:r.n_frames], dbg


def encode_streams(engine, streams, sample_rate, bits_per_sample, compression_level=5, blocksize=0, **kw):
    """Convenience: encode a list of (n, ch) integer arrays -> list of bytes (complete .flac images)."""
    chs = {(s.shape[1] if s.ndim == 2 else 1) for s in streams}
    assert len(chs) == 1
    ch = chs.pop()
    dt = np.int16 if bits_per_sample <= 16 else np.int32
    flat = [np.ascontiguousarray(s, dt).reshape(-1) for s in streams]
    sizes = np.array([f.size for f in flat], np.uint64)
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64) if len(flat) else np.zeros(0, np.uint64)
    pcm = np.concatenate(flat) if flat else np.zeros(0, dt)
    cfg = Engine.make_config(sample_rate, ch, bits_per_sample, compression_level, blocksize,
                             container_bytes=pcm.dtype.itemsize, **kw)
    engine.encode_host(cfg, pcm, offs, sizes // ch)
    out = engine.fetch()
    res = []
    for si in out["streams"]:
        res.append(out["arena"][int(si.byte_off):int(si.byte_off + si.byte_len)].tobytes())
    return res, out


# ------------------------------------------------------------------ decode (batch ABI)
class DecStreamInfo(C.Structure):
    _fields_ = [("total_samples", C.c_uint64), ("pcm_off", C.c_uint64), ("consumed", C.c_uint64), ("n_frames", C.c_uint32),
                ("status", C.c_int32), ("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bits_per_sample", C.c_uint32),
                ("max_blocksize", C.c_uint32), ("n_events", C.c_uint32), ("gap_samples", C.c_uint32), ("ev_frame", C.c_uint32 * 16),
                ("ev_status", C.c_uint8 * 16), ("next_sample", C.c_uint64), ("last_blocksize", C.c_uint32), ("have_last", C.c_uint32)]


class DecResult(C.Structure):
    _fields_ = [("total_elems", C.c_uint64), ("n_streams", C.c_uint32), ("n_frames", C.c_uint32),
                ("out_container_bytes", C.c_uint32), ("n_candidates", C.c_uint32), ("d_pcm", C.c_void_p)]


class DecRawParams(C.Structure):
    _fields_ = [("sample_rate", C.c_uint32), ("channels", C.c_uint32), ("bits_per_sample", C.c_uint32), ("flags", C.c_uint32),
                ("next_sample", C.c_uint64), ("last_blocksize", C.c_uint32), ("fixed_blocksize", C.c_uint32)]


DEC_STATUS = {0: "ok", 2: "not FLAC", 3: "bad metadata", 4: "bad frame", 5: "incomplete frame", 6: "lost sync",
              7: "CRC-16 mismatch", 8: "unsupported", 9: "reserved values"}


def _dec_proto(L):
    if getattr(L, "_dec_proto_done", False):
        return
    L.flacb200_decode_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p,
                                        C.c_uint32, C.c_void_p]
    L.flacb200_decode_result.argtypes = [C.c_void_p, C.POINTER(DecResult)]
    L.flacb200_decode_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_uint32]
    L.flacb200_decode_kernel_times.argtypes = [C.c_void_p, C.c_void_p]
    L.flacb200_decode_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32,
                                             C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    L._dec_proto_done = True


def _engine_decode_device(self, blob_ptr, blob_bytes, stream_off, stream_len, out_container_bytes=0):
    """Asynchronous-ish batch decode of FLAC bytes resident in HBM (sizes force 3 internal syncs)."""
    _dec_proto(self._L)
    so = np.ascontiguousarray(stream_off, np.uint64)
    sl = np.ascontiguousarray(stream_len, np.uint64)
    self._keep_dec = (so, sl)
    self._check(self._L.flacb200_decode_batch(self._h, C.c_void_p(blob_ptr), 1, blob_bytes, len(so), so.ctypes.data,
                                              sl.ctypes.data, out_container_bytes, None))


def _engine_decode_host(self, blob, stream_off, stream_len, out_container_bytes=0, raw=None):
    _dec_proto(self._L)
    blob = np.ascontiguousarray(blob, np.uint8)
    so = np.ascontiguousarray(stream_off, np.uint64)
    sl = np.ascontiguousarray(stream_len, np.uint64)
    rp = DecRawParams(*raw) if raw else None
    self._keep_dec = (blob, so, sl, rp)
    self._check(self._L.flacb200_decode_batch(self._h, blob.ctypes.data, 0, blob.size, len(so), so.ctypes.data, sl.ctypes.data,
                                              out_container_bytes, C.byref(rp) if rp else None))


def _engine_decode_host_pipelined(self, blob, stream_off, stream_len, out, out_container_bytes=2, raw=None):
    """Host -> host decode in one pipelined call (H2D / kernels / D2H of different chunks overlap).
    `out`: preallocated (ideally pinned) int16/int32 array; returns (total_elems, [DecStreamInfo])."""
    _dec_proto(self._L)
    blob = np.ascontiguousarray(blob, np.uint8)
    so = np.ascontiguousarray(stream_off, np.uint64)
    sl = np.ascontiguousarray(stream_len, np.uint64)
    rp = DecRawParams(*raw) if raw else None
    infos = (DecStreamInfo * max(len(so), 1))()
    tot = C.c_uint64(0)
    self._check(self._L.flacb200_decode_batch_host(self._h, blob.ctypes.data, blob.size, len(so), so.ctypes.data, sl.ctypes.data,
                                                   out_container_bytes, C.byref(rp) if rp else None, out.ctypes.data, out.nbytes,
                                                   C.byref(tot), C.cast(infos, C.c_void_p)))
    return int(tot.value), infos


def _engine_decode_result(self):
    _dec_proto(self._L)
    r = DecResult()
    self._check(self._L.flacb200_decode_result(self._h, C.byref(r)))
    return r


def _engine_decode_fetch(self):
    """-> (pcm flat array in the output container, [DecStreamInfo], frame_samples)"""
    r = self.decode_result()
    dt = np.int16 if r.out_container_bytes == 2 else np.int32
    pcm = np.empty(max(int(r.total_elems), 1), dt)
    infos = (DecStreamInfo * max(r.n_streams, 1))()
    fs = np.zeros(max(r.n_frames, 1), np.uint32)
    self._check(self._L.flacb200_decode_fetch(self._h, pcm.ctypes.data, pcm.nbytes, C.cast(infos, C.c_void_p), fs.ctypes.data, r.n_frames))
    return pcm[:int(r.total_elems)], [infos[i] for i in range(r.n_streams)], fs[:r.n_frames]


def _engine_decode_kernel_times(self):
    _dec_proto(self._L)
    ms = np.zeros(6, np.float32)
    self._check(self._L.flacb200_decode_kernel_times(self._h, ms.ctypes.data))
    return dict(zip(["sync_scan", "frame_decode", "chain_layout", "post", "crc16"], [float(v) for v in ms[:5]]))


Engine.decode_device = _engine_decode_device
Engine.decode_host = _engine_decode_host
Engine.decode_host_pipelined = _engine_decode_host_pipelined
Engine.decode_result = _engine_decode_result
Engine.decode_fetch = _engine_decode_fetch
Engine.decode_kernel_times = _engine_decode_kernel_times


def decode_streams(engine, blobs, out_container_bytes=0, check=False):
    """Convenience: decode a list of .flac byte strings -> list of (n, ch) arrays + infos.
    check=True raises NativeError when any stream ends with a non-zero status (truncated, CRC mismatch, ...);
    otherwise the caller reads infos[s].status (DEC_STATUS)."""
    sizes = np.array([len(b) for b in blobs], np.uint64)
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64) if len(blobs) else np.zeros(0, np.uint64)
    blob = np.frombuffer(b"".join(blobs) + bytes(16), np.uint8)
    engine.decode_host(blob, offs, sizes, out_container_bytes)
    pcm, infos, _ = engine.decode_fetch()
    if check:
        bad = [(s, int(si.status)) for s, si in enumerate(infos) if si.status != 0]
        if bad:
            raise NativeError(6, "decode failed for streams " + ", ".join(f"{s}: {DEC_STATUS.get(st, st)}" for s, st in bad[:8]))
    out = []
    for si in infos:
        ch = max(int(si.channels), 1)
        n = int(si.total_samples)
        out.append(pcm[int(si.pcm_off): int(si.pcm_off) + n * ch].reshape(n, ch))
    return out, infos
