"""pyflac_b200 -- a B200-native FLAC encode/decode engine behind pyFLAC's Python API.

Drop-in names (reference `pyflac/__init__.py:14-26`): StreamEncoder, FileEncoder, EncoderState,
EncoderInitException, EncoderProcessException, StreamDecoder, FileDecoder, OneShotDecoder, DecoderState,
DecoderInitException, DecoderProcessException.  Additive: encode_batch / decode_batch (many streams per call).
Importing the package does not touch CUDA; the first encoder/decoder that needs the device does.
"""
__title__ = "pyflac_b200"
__version__ = "0.1.0"

from .encoder import (EncoderInitException, EncoderProcessException, EncoderState, FileEncoder,  # noqa: F401
                      StreamEncoder, encode_batch)
from .decoder import (DecoderInitException, DecoderProcessException, DecoderState, FileDecoder,  # noqa: F401
                      OneShotDecoder, StreamDecoder, decode_batch)

__all__ = ["StreamEncoder", "FileEncoder", "EncoderState", "EncoderInitException", "EncoderProcessException",
           "StreamDecoder", "FileDecoder", "OneShotDecoder", "DecoderState", "DecoderInitException",
           "DecoderProcessException", "encode_batch", "decode_batch"]
