"""ctypes view of the DROP-IN layer of libflacb200.so (include/flacb200_flac_api.h): the FLAC__stream_encoder_* and
FLAC__stream_decoder_* entry points, callback types and payload structures pyFLAC binds through cffi
(reference: pyflac/builder/encoder.py:34-324, pyflac/builder/decoder.py:32-478)."""
import ctypes as C

from . import _native

# ---- callback types (builder/encoder.py:251-256, builder/decoder.py:368-375) ----
ENC_WRITE_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_ubyte), C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p)
ENC_SEEK_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_void_p)
ENC_TELL_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p)
ENC_META_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p)
ENC_PROGRESS_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p)
DEC_READ_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_ubyte), C.POINTER(C.c_size_t), C.c_void_p)
DEC_WRITE_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(C.c_int32)), C.c_void_p)
DEC_ERROR_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_void_p)


# ---- payload structures ----
class StreamInfo(C.Structure):          # builder/encoder.py:129-137
    _fields_ = [("min_blocksize", C.c_uint32), ("max_blocksize", C.c_uint32), ("min_framesize", C.c_uint32),
                ("max_framesize", C.c_uint32), ("sample_rate", C.c_uint32), ("channels", C.c_uint32),
                ("bits_per_sample", C.c_uint32), ("total_samples", C.c_uint64), ("md5sum", C.c_ubyte * 16)]


class _MetadataData(C.Union):
    _fields_ = [("stream_info", StreamInfo), ("pad_", C.c_uint64 * 24)]


class StreamMetadata(C.Structure):      # builder/encoder.py:234-248 (only STREAMINFO is ever delivered)
    _fields_ = [("type", C.c_int), ("is_last", C.c_int), ("length", C.c_uint32), ("data", _MetadataData)]


class _FrameNumber(C.Union):
    _fields_ = [("frame_number", C.c_uint32), ("sample_number", C.c_uint64)]


class FrameHeader(C.Structure):         # builder/decoder.py:146-158
    _fields_ = [("blocksize", C.c_uint32), ("sample_rate", C.c_uint32), ("channels", C.c_uint32),
                ("channel_assignment", C.c_int), ("bits_per_sample", C.c_uint32), ("number_type", C.c_int),
                ("number", _FrameNumber), ("crc", C.c_uint8)]


class Frame(C.Structure):
    _fields_ = [("header", FrameHeader)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = _native.lib()
    vp, u32, i = C.c_void_p, C.c_uint32, C.c_int
    L.FLAC__stream_encoder_new.restype = vp
    L.FLAC__stream_encoder_delete.argtypes = [vp]
    L.FLAC__stream_encoder_delete.restype = None
    for n in ("verify", "channels", "bits_per_sample", "sample_rate", "compression_level", "blocksize", "streamable_subset",
              "limit_min_bitrate"):
        f = getattr(L, "FLAC__stream_encoder_set_" + n)
        f.argtypes, f.restype = [vp, u32], i
    for n in ("verify", "channels", "bits_per_sample", "sample_rate", "blocksize", "streamable_subset", "limit_min_bitrate",
              "state"):
        f = getattr(L, "FLAC__stream_encoder_get_" + n)
        f.argtypes, f.restype = [vp], u32
    L.FLAC__stream_encoder_init_stream.argtypes = [vp, ENC_WRITE_CB, ENC_SEEK_CB, ENC_TELL_CB, ENC_META_CB, vp]
    L.FLAC__stream_encoder_init_stream.restype = i
    L.FLAC__stream_encoder_init_file.argtypes = [vp, C.c_char_p, ENC_PROGRESS_CB, vp]
    L.FLAC__stream_encoder_init_file.restype = i
    L.FLAC__stream_encoder_process_interleaved.argtypes = [vp, vp, u32]
    L.FLAC__stream_encoder_process_interleaved.restype = i
    L.FLAC__stream_encoder_finish.argtypes = [vp]
    L.FLAC__stream_encoder_finish.restype = i

    L.FLAC__stream_decoder_new.restype = vp
    L.FLAC__stream_decoder_delete.argtypes = [vp]
    L.FLAC__stream_decoder_delete.restype = None
    L.FLAC__stream_decoder_get_state.argtypes = [vp]
    L.FLAC__stream_decoder_get_state.restype = i
    L.FLAC__stream_decoder_init_stream.argtypes = [vp, DEC_READ_CB, vp, vp, vp, vp, DEC_WRITE_CB, vp, DEC_ERROR_CB, vp]
    L.FLAC__stream_decoder_init_stream.restype = i
    L.FLAC__stream_decoder_init_file.argtypes = [vp, C.c_char_p, DEC_WRITE_CB, vp, DEC_ERROR_CB, vp]
    L.FLAC__stream_decoder_init_file.restype = i
    for n in ("finish", "process_single", "process_until_end_of_stream"):
        f = getattr(L, "FLAC__stream_decoder_" + n)
        f.argtypes, f.restype = [vp], i
    _lib = L
    return L


def string_table(name, n):
    """Read one of the exported `const char * const X[]` tables (e.g. FLAC__StreamEncoderStateString)."""
    arr = (C.c_char_p * n).in_dll(lib(), name)
    return [a.decode() for a in arr]
