"""pyFLAC-compatible decoder classes on top of libflacb200.so.

Same public surface as the reference (`pyflac/decoder.py`): `StreamDecoder`, `FileDecoder`, `OneShotDecoder`,
`DecoderState`, `DecoderInitException`, `DecoderProcessException`; the write callback receives
`(audio: ndarray (blocksize, channels) int16|int32, sample_rate, num_channels, num_samples)` exactly as the
reference delivers it (pyflac/decoder.py:482-527).  Native calls go through ctypes into this repository's C ABI
(`FLAC__stream_decoder_*` of libflacb200.so), which batches every complete frame it holds into one GPU launch
sequence.  No CPU fallback: without a usable GPU the constructors raise `DecoderInitException`.

Additive: `decode_batch()` -- many independent .flac byte strings per call.
"""
import ctypes as C
import logging
import queue
import tempfile
import threading
from enum import Enum
from pathlib import Path
from typing import Callable, Sequence

import numpy as np

from . import _capi, _native, wav


class DecoderState(Enum):
    """The decoder state (values of FLAC__StreamDecoderState, reference builder/decoder.py:49-60)."""
    SEARCH_FOR_METADATA = 0
    READ_METADATA = 1
    SEARCH_FOR_FRAME_SYNC = 2
    READ_FRAME = 3
    END_OF_STREAM = 4
    OGG_ERROR = 5
    SEEK_ERROR = 6
    ABORTED = 7
    MEMORY_ALLOCATION_ERROR = 8
    UNINITIALIZED = 9

    def __str__(self):
        return _capi.string_table("FLAC__StreamDecoderStateString", 10)[self.value]


class DecoderInitException(Exception):
    """Raised when initialisation fails; `code` is the FLAC__StreamDecoderInitStatus."""

    def __init__(self, code):
        self.code = code

    def __str__(self):
        return _capi.string_table("FLAC__StreamDecoderInitStatusString", 6)[self.code]


class DecoderProcessException(Exception):
    """Raised when a fatal read, write or memory error (or a bitstream error) occurs while decoding."""


class _Decoder:
    def __init__(self):
        self._lib = _capi.lib()
        self._decoder = self._lib.FLAC__stream_decoder_new()
        self._thunks = []
        self._error = None
        self.write_callback = None
        self.logger = logging.getLogger(__name__)

    def __del__(self):
        try:
            if self._decoder:
                self._lib.FLAC__stream_decoder_delete(self._decoder)
                self._decoder = None
        except Exception:  # interpreter shutdown
            pass

    def finish(self):
        """Flush, release resources and return the decoder to `DecoderState.UNINITIALIZED`."""
        self._lib.FLAC__stream_decoder_finish(self._decoder)

    @property
    def state(self) -> DecoderState:
        return DecoderState(self._lib.FLAC__stream_decoder_get_state(self._decoder))

    def process(self):
        """Overridden by the stream / file decoders (reference pyflac/decoder.py:110-111)."""
        raise NotImplementedError

    # ---- trampolines shared by all decoder flavours
    def _on_write(self, _dec, frame, buffers, _cd):
        try:
            h = C.cast(frame, C.POINTER(_capi.Frame)).contents.header
            if h.bits_per_sample not in (16, 32):         # reference decoder.py:502-503 (=> ABORT)
                raise ValueError("FLAC decoder only supports 16-bit or 32-bit samples")
            dt = np.int16 if h.bits_per_sample == 16 else np.int32
            block = np.empty((h.blocksize, h.channels), dt)
            for c in range(h.channels):
                block[:, c] = np.ctypeslib.as_array(buffers[c], shape=(h.blocksize,))
            self.write_callback(block, int(h.sample_rate), int(h.channels), int(h.blocksize))
            return 0
        except Exception as e:  # noqa: BLE001
            self._error = self._error or str(e)
            return 1

    def _on_error(self, _dec, status, _cd):
        msg = _capi.string_table("FLAC__StreamDecoderErrorStatusString", 5)[status]
        self.logger.error(f"Error in libFLAC decoder: {msg}")
        self._error = msg
        self._wake()

    def _wake(self):
        pass


class StreamDecoder(_Decoder):
    """Push-style decoder: feed FLAC bytes with `process()`, receive PCM blocks through `write_callback` on a
    background thread; `finish()` must be called at the end (raises `DecoderProcessException` on errors)."""

    def __init__(self, write_callback: Callable[[np.ndarray, int, int, int], None]):
        super().__init__()
        self.write_callback = write_callback
        self._chunks = queue.Queue()
        self._partial = b""
        self._done = False
        th = [_capi.DEC_READ_CB(self._on_read), _capi.DEC_WRITE_CB(self._on_write), _capi.DEC_ERROR_CB(self._on_error)]
        self._thunks = th
        rc = self._lib.FLAC__stream_decoder_init_stream(self._decoder, th[0], None, None, None, None, th[1], None, th[2], None)
        if rc != 0:
            raise DecoderInitException(rc)
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        if not self._lib.FLAC__stream_decoder_process_until_end_of_stream(self._decoder):
            self._error = self._error or "A fatal read, write, or memory allocation error occurred"

    def _wake(self):
        self._chunks.put(None)

    def _on_read(self, _dec, buf, pbytes, _cd):
        """Blocks until data, end of stream, or an error (reference decoder.py:394-452)."""
        want = pbytes[0]
        while not self._partial:
            if self._error:
                pbytes[0] = 0
                return 2                                     # ABORT
            try:
                item = self._chunks.get(timeout=0.05)
            except queue.Empty:
                if self._done:
                    pbytes[0] = 0
                    return 1                                 # END_OF_STREAM
                continue
            if item:
                self._partial = item
        k = min(want, len(self._partial))
        C.memmove(buf, self._partial, k)
        self._partial = self._partial[k:]
        pbytes[0] = k
        return 0

    def process(self, data: bytes):
        """Queue some FLAC bytes (non-blocking; decoding happens on the background thread)."""
        self._chunks.put(bytes(data))

    def finish(self):
        self._done = True
        self._thread.join()
        super().finish()
        if self._error:
            raise DecoderProcessException(self._error)


class FileDecoder(_Decoder):
    """Decode a FLAC file to a (PCM_16) WAV file; `process()` returns `(float64 samples (n, ch), sample_rate)`
    like the reference's `sf.read(output, always_2d=True)` (pyflac/decoder.py:300-302)."""

    def __init__(self, input_file: Path, output_file: Path = None):
        super().__init__()
        self.__output = None
        self.write_callback = self._write_block
        if output_file:
            self.__output_file = output_file
        else:
            self.__temp = tempfile.NamedTemporaryFile(suffix=".wav")
            self.__output_file = Path(self.__temp.name)
        th = [_capi.DEC_WRITE_CB(self._on_write), _capi.DEC_ERROR_CB(self._on_error)]
        self._thunks = th
        rc = self._lib.FLAC__stream_decoder_init_file(self._decoder, str(input_file).encode("utf8"), th[0], None, th[1], None)
        if rc != 0:
            raise DecoderInitException(rc)

    def _write_block(self, block, sample_rate, channels, _n):
        if self.__output is None:
            self.__output = wav.Pcm16Writer(str(self.__output_file), sample_rate, channels)
        self.__output.write(block)

    def process(self):
        ok = self._lib.FLAC__stream_decoder_process_until_end_of_stream(self._decoder)
        if self.__output is not None:
            self.__output.close()
        super().finish()
        if not ok or self._error:
            raise DecoderProcessException(self._error or str(self.state))
        return wav.read_float64(str(self.__output_file))


class OneShotDecoder(_Decoder):
    """Decode a complete FLAC byte string synchronously in the constructor (no thread)."""

    def __init__(self, write_callback: Callable[[np.ndarray, int, int, int], None], buffer: bytes):
        super().__init__()
        self.write_callback = write_callback
        self._data = bytes(buffer)
        self._pos = 0
        th = [_capi.DEC_READ_CB(self._on_read), _capi.DEC_WRITE_CB(self._on_write), _capi.DEC_ERROR_CB(self._on_error)]
        self._thunks = th
        rc = self._lib.FLAC__stream_decoder_init_stream(self._decoder, th[0], None, None, None, None, th[1], None, th[2], None)
        if rc != 0:
            raise DecoderInitException(rc)
        ok = self._lib.FLAC__stream_decoder_process_until_end_of_stream(self._decoder)
        super().finish()
        if not ok or self._error:
            raise DecoderProcessException(self._error or "A fatal read, write, or memory allocation error occurred")

    def _on_read(self, _dec, buf, pbytes, _cd):
        k = min(pbytes[0], len(self._data) - self._pos)
        if k <= 0:
            pbytes[0] = 0
            return 1
        C.memmove(buf, self._data[self._pos:self._pos + k], k)
        self._pos += k
        pbytes[0] = k
        return 0


# ---------------------------------------------------------------------------- additive batch front-end
def decode_batch(blobs: Sequence[bytes], device: int = 0, strict: bool = True):
    """Decode many .flac byte strings in one GPU batch -> (list of (n, channels) int16/int32 arrays, list of info).
    strict (default): a stream that ends with an error (truncated, CRC mismatch, not FLAC ...) raises
    DecoderProcessException instead of coming back short; strict=False returns what decoded and info[s].status."""
    from .encoder import _engine
    out, infos = _native.decode_streams(_engine(device), list(blobs))
    if strict:
        bad = [(s, int(si.status)) for s, si in enumerate(infos) if si.status != 0]
        if bad:
            raise DecoderProcessException("; ".join(f"stream {s}: {_native.DEC_STATUS.get(st, st)}" for s, st in bad[:8]))
    return out, infos
