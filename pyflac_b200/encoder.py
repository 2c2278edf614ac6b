"""pyFLAC-compatible encoder classes on top of libflacb200.so.

Same public surface as the reference (`pyflac/encoder.py`): `StreamEncoder`, `FileEncoder`, `EncoderState`,
`EncoderInitException`, `EncoderProcessException`, same constructor arguments, same callback signatures
(reference pyflac/encoder.py:293-316 ctor, :429-494 trampolines), same lazy initialisation on the first
`process()` (:104-110) -- but the native calls go through ctypes into this repository's C ABI
(`FLAC__stream_encoder_*` of libflacb200.so, include/flacb200_flac_api.h) whose per-frame arithmetic runs in
CUDA kernels.  There is no CPU fallback: without a usable GPU `process()` raises `EncoderInitException`.

Additive (not in the reference): `encode_batch()` -- many independent streams per call, the only way to keep a
B200 busy; the per-object classes encode whatever complete frames one `process()` call holds as one batch.
"""
import ctypes as C
import logging
import tempfile
from enum import Enum
from pathlib import Path
from typing import Callable, Sequence

import numpy as np

from . import _capi, _native, wav

_STATE_NAMES = ["OK", "UNINITIALIZED", "OGG_ERROR", "VERIFY_DECODER_ERROR", "VERIFY_MISMATCH_IN_AUDIO_DATA", "CLIENT_ERROR",
                "IO_ERROR", "FRAMING_ERROR", "MEMORY_ALLOCATION_ERROR"]


class EncoderState(Enum):
    """The encoder state (values of FLAC__StreamEncoderState, reference builder/encoder.py:51-61)."""
    OK = 0
    UNINITIALIZED = 1
    OGG_ERROR = 2
    VERIFY_DECODER_ERROR = 3
    VERIFY_MISMATCH_IN_AUDIO_DATA = 4
    CLIENT_ERROR = 5
    IO_ERROR = 6
    FRAMING_ERROR = 7
    MEMORY_ALLOCATION_ERROR = 8

    def __str__(self):
        return _capi.string_table("FLAC__StreamEncoderStateString", 9)[self.value]


class EncoderInitException(Exception):
    """Raised when initialisation of a `StreamEncoder` / `FileEncoder` fails; `code` is the FLAC__StreamEncoderInitStatus."""

    def __init__(self, code):
        self.code = code

    def __str__(self):
        return _capi.string_table("FLAC__StreamEncoderInitStatusString", 14)[self.code]


class EncoderProcessException(Exception):
    """Raised when an error occurs while processing audio data."""


class _Handle:
    """Owns one FLAC__StreamEncoder and the ctypes callback thunks that must outlive it."""

    def __init__(self):
        self.lib = _capi.lib()
        self.ptr = self.lib.FLAC__stream_encoder_new()
        self.thunks = []

    def __del__(self):
        try:
            if self.ptr:
                self.lib.FLAC__stream_encoder_delete(self.ptr)
                self.ptr = None
        except Exception:  # interpreter shutdown
            pass


class _Encoder:
    """Settings, lazy init and `process()` shared by the stream and file encoders."""

    def __init__(self):
        self._initialised = False
        self._h = _Handle()
        self._lib = self._h.lib
        self._encoder = self._h.ptr
        self.logger = logging.getLogger(__name__)

    def _init(self):
        raise NotImplementedError

    # -- processing
    def process(self, samples: np.ndarray):
        """Process some samples (numpy int16 / int32, shape (n,) or (n, channels)).

        Raises:
            TypeError: if `samples` is not a numpy array
            EncoderInitException: first call only, if the settings are rejected
            EncoderProcessException: if encoding fails
        """
        if not isinstance(samples, np.ndarray):
            raise TypeError("Processing only supports numpy arrays")
        if not self._initialised:
            self._channels = samples.shape[1] if samples.ndim > 1 else 1
            self._bits_per_sample = samples.dtype.itemsize * 8
            self._init()
        block = np.ascontiguousarray(samples).astype(np.int32)
        ok = self._lib.FLAC__stream_encoder_process_interleaved(self._encoder, block.ctypes.data, len(block))
        if not ok:
            raise EncoderProcessException(str(self.state))

    def finish(self) -> bool:
        """Flush the last (short) frame, finalise STREAMINFO and reset the encoder. Returns True on success."""
        return bool(self._lib.FLAC__stream_encoder_finish(self._encoder))

    @property
    def state(self) -> EncoderState:
        return EncoderState(self._lib.FLAC__stream_encoder_get_state(self._encoder))

    # -- settings (same private property names the reference's tests poke at: tests/test_encoder.py:32-93)
    def _setting(name):   # noqa: N805
        def getter(self):
            return getattr(self._lib, "FLAC__stream_encoder_get_" + name)(self._encoder)

        def setter(self, value):
            getattr(self._lib, "FLAC__stream_encoder_set_" + name)(self._encoder, int(value))
        return property(getter, setter)

    _verify = _setting("verify")
    _channels = _setting("channels")
    _bits_per_sample = _setting("bits_per_sample")
    _sample_rate = _setting("sample_rate")
    _blocksize = _setting("blocksize")
    _streamable_subset = _setting("streamable_subset")
    _limit_min_bitrate = _setting("limit_min_bitrate")

    @property
    def _compression_level(self):
        raise NotImplementedError

    @_compression_level.setter
    def _compression_level(self, value):
        self._lib.FLAC__stream_encoder_set_compression_level(self._encoder, int(value))


class StreamEncoder(_Encoder):
    """Real-time style encoder: raw audio in through `process()`, FLAC bytes out through `write_callback`.

    Args mirror the reference (pyflac/encoder.py:293-316):
        sample_rate, write_callback(buffer: bytes, num_bytes, num_samples, current_frame),
        seek_callback(offset), tell_callback() -> int, metadata_callback(metadata),
        compression_level=5, blocksize=0, streamable_subset=True, verify=False, limit_min_bitrate=False
    `num_samples == 0` marks stream header / STREAMINFO rewrite data, otherwise it is the frame's blocksize.
    """

    def __init__(self, sample_rate: int, write_callback: Callable[[bytes, int, int, int], None],
                 seek_callback: Callable[[int], None] = None, tell_callback: Callable[[], int] = None,
                 metadata_callback: Callable[[object], None] = None, compression_level: int = 5, blocksize: int = 0,
                 streamable_subset: bool = True, verify: bool = False, limit_min_bitrate: bool = False):
        super().__init__()
        self.write_callback = write_callback
        self.seek_callback = seek_callback
        self.tell_callback = tell_callback
        self.metadata_callback = metadata_callback
        self._sample_rate = sample_rate
        self._blocksize = blocksize
        self._compression_level = compression_level
        self._streamable_subset = streamable_subset
        self._verify = verify
        self._limit_min_bitrate = limit_min_bitrate

    def _init(self):
        def on_write(_enc, buf, nbytes, samples, frame, _cd):      # exception => FATAL_ERROR, like cffi's onerror
            try:
                self.write_callback(C.string_at(buf, nbytes), nbytes, samples, frame)
                return 0
            except Exception:  # noqa: BLE001
                self.logger.exception("write_callback failed")
                return 1

        def on_seek(_enc, offset, _cd):
            try:
                self.seek_callback(offset)
                return 0
            except Exception:  # noqa: BLE001
                return 1

        def on_tell(_enc, poff, _cd):
            try:
                poff[0] = int(self.tell_callback())
                return 0
            except Exception:  # noqa: BLE001
                return 1

        def on_meta(_enc, md, _cd):
            self.metadata_callback(C.cast(md, C.POINTER(_capi.StreamMetadata)).contents)

        null = lambda T: C.cast(None, T)  # noqa: E731
        th = [_capi.ENC_WRITE_CB(on_write),
              _capi.ENC_SEEK_CB(on_seek) if self.seek_callback else null(_capi.ENC_SEEK_CB),
              _capi.ENC_TELL_CB(on_tell) if self.tell_callback else null(_capi.ENC_TELL_CB),
              _capi.ENC_META_CB(on_meta) if self.metadata_callback else null(_capi.ENC_META_CB)]
        self._h.thunks = th
        rc = self._lib.FLAC__stream_encoder_init_stream(self._encoder, th[0], th[1], th[2], th[3], None)
        if rc != 0:
            raise EncoderInitException(rc)
        self._initialised = True


class FileEncoder(_Encoder):
    """Encode a WAV file (PCM_16 or PCM_32) to a FLAC file; `process()` returns the FLAC bytes.

    Args mirror the reference (pyflac/encoder.py:363-391): input_file, output_file=None (temporary file),
    compression_level=5, blocksize=0, streamable_subset=True, verify=False.
    """

    def __init__(self, input_file: Path, output_file: Path = None, compression_level: int = 5, blocksize: int = 0,
                 streamable_subset: bool = True, verify: bool = False):
        super().__init__()
        self.__raw_audio, self._sample_rate = wav.read_pcm(str(input_file))      # ValueError unless PCM_16 / PCM_32
        if output_file:
            self.__output_file = output_file
        else:
            self.__temp = tempfile.NamedTemporaryFile(suffix=".flac")
            self.__output_file = Path(self.__temp.name)
        self._blocksize = blocksize
        self._compression_level = compression_level
        self._streamable_subset = streamable_subset
        self._verify = verify

    def _init(self):
        def on_progress(_enc, bytes_written, samples_written, frames_written, total_frames_estimate, _cd):
            self.logger.debug(f"{frames_written} frames written ({bytes_written} bytes, {samples_written} samples)")

        th = _capi.ENC_PROGRESS_CB(on_progress)
        self._h.thunks = [th]
        rc = self._lib.FLAC__stream_encoder_init_file(self._encoder, str(self.__output_file).encode("utf8"), th, None)
        if rc != 0:
            raise EncoderInitException(rc)
        self._initialised = True

    def process(self) -> bytes:
        """Encode the whole file; returns the bytes of the FLAC file that was written."""
        super().process(self.__raw_audio)
        self.finish()
        with open(self.__output_file, "rb") as f:
            return f.read()


# ---------------------------------------------------------------------------- additive batch front-end
_engines = {}


def _engine(device=0):
    if device not in _engines:
        _engines[device] = _native.Engine(device)
    return _engines[device]


def encode_batch(streams: Sequence[np.ndarray], sample_rate: int, compression_level: int = 5, blocksize: int = 0,
                 bits_per_sample: int = None, device: int = 0, streamable_subset: bool = True):
    """Encode many independent streams in one GPU batch.

    streams: sequence of int16 / int32 arrays of shape (n,) or (n, channels) (same channel count and dtype).
    bits_per_sample defaults to the dtype width (16 / 32) like the reference (encoder.py:109); pass 24 (with
    int32 arrays) for 24-bit audio, which the reference's Python layer cannot express.
    Returns (list of complete .flac byte strings, info dict with frame index / sizes).
    """
    arrs = [np.asarray(s) for s in streams]
    if bits_per_sample is None:
        bits_per_sample = arrs[0].dtype.itemsize * 8 if arrs else 16
    return _native.encode_streams(_engine(device), [a.reshape(a.shape[0], -1) for a in arrs], sample_rate, bits_per_sample,
                                  compression_level, blocksize, streamable_subset=streamable_subset)
