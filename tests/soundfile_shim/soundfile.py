"""Test-only stand-in for the `soundfile` package (libsndfile is not in this image): just what the reference's
pyflac/encoder.py:372-380, pyflac/decoder.py:300-313, tests/ and examples/passthrough.py call, on top of
pyflac_b200.wav.  Put this directory on sys.path BEFORE importing the reference package."""
import numpy as np

from pyflac_b200 import wav as _wav


class _Info:
    def __init__(self, i):
        self.samplerate, self.channels, self.frames, self.subtype = i.samplerate, i.channels, i.frames, i.subtype
        self.format = "WAV"


def info(path):
    return _Info(_wav.info(str(path)))


def read(path, dtype="float64", always_2d=False):
    path = str(path)
    if str(dtype) in ("int16", "int32"):
        x, sr = _wav.read_pcm(path)
        x = x.astype(np.dtype(str(dtype)))
    else:
        x, sr = _wav.read_float64(path)
    if not always_2d and x.ndim == 2 and x.shape[1] == 1:
        x = x[:, 0]
    return x, sr


class SoundFile:
    """mode='w' only: PCM_16 WAV (libsndfile's default subtype for .wav), int16 / int32 blocks appended."""

    def __init__(self, path, mode="r", samplerate=None, channels=None, **_):
        if mode != "w":
            raise NotImplementedError("soundfile shim: only mode='w'")
        self._w = _wav.Pcm16Writer(str(path), int(samplerate), int(channels))

    def write(self, data):
        self._w.write(np.asarray(data))

    def close(self):
        self._w.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
