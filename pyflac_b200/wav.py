"""Minimal RIFF/WAVE reader and writer (PCM, incl. WAVE_FORMAT_EXTENSIBLE) so the File classes need no `soundfile`.

Only what pyFLAC's FileEncoder / FileDecoder observably do with libsndfile is reproduced
(reference pyflac/encoder.py:372-380, pyflac/decoder.py:300-313): read PCM_16 / PCM_32 as int16 / int32,
write PCM_16, and read a file back as float64 in [-1, 1).
"""
import struct

import numpy as np

_SUBTYPES = {8: "PCM_U8", 16: "PCM_16", 24: "PCM_24", 32: "PCM_32"}


class WavInfo:
    def __init__(self, samplerate, channels, frames, bits, subtype, data_offset, data_bytes):
        self.samplerate, self.channels, self.frames = samplerate, channels, frames
        self.bits, self.subtype = bits, subtype
        self.data_offset, self.data_bytes = data_offset, data_bytes


def info(path):
    with open(path, "rb") as f:
        head = f.read(12)
        if len(head) < 12 or head[:4] != b"RIFF" or head[8:12] != b"WAVE":
            raise ValueError(f"{path}: not a RIFF/WAVE file")
        fmt = None
        while True:
            ck = f.read(8)
            if len(ck) < 8:
                raise ValueError(f"{path}: no data chunk")
            cid, size = ck[:4], struct.unpack("<I", ck[4:])[0]
            if cid == b"fmt ":
                fmt = f.read(size)
                if size & 1:
                    f.read(1)
            elif cid == b"data":
                if fmt is None:
                    raise ValueError(f"{path}: data chunk before fmt chunk")
                tag, ch, sr, _, align, bits = struct.unpack("<HHIIHH", fmt[:16])
                if tag == 0xFFFE and len(fmt) >= 40:
                    tag = struct.unpack("<H", fmt[24:26])[0]
                subtype = _SUBTYPES.get(bits, f"PCM_{bits}") if tag == 1 else ("FLOAT" if tag == 3 else f"FORMAT_{tag}")
                off = f.tell()
                return WavInfo(sr, ch, size // max(align, 1), bits, subtype, off, size)
            else:
                f.seek(size + (size & 1), 1)


def read_pcm(path):
    """-> (int16 or int32 array (frames, channels), samplerate). Only PCM_16 / PCM_32 (what pyFLAC accepts)."""
    i = info(path)
    if i.subtype not in ("PCM_16", "PCM_32"):
        raise ValueError(f"WAV input data type must be either PCM_16 or PCM_32: Got {i.subtype}")
    with open(path, "rb") as f:
        f.seek(i.data_offset)
        raw = f.read(i.frames * i.channels * (i.bits // 8))
    dt = "<i2" if i.bits == 16 else "<i4"
    x = np.frombuffer(raw, dt).reshape(-1, i.channels)
    return x.astype(np.int16 if i.bits == 16 else np.int32), i.samplerate


class Pcm16Writer:
    """Append-only PCM_16 WAV writer (what sf.SoundFile(..., mode='w', format default) gives FileDecoder)."""

    def __init__(self, path, samplerate, channels):
        self.f = open(path, "wb")
        self.sr, self.ch, self.nbytes = samplerate, channels, 0
        self.f.write(b"\0" * 44)

    def write(self, block):
        b = np.ascontiguousarray(block)
        if b.dtype != np.int16:
            b = (b >> 16).astype(np.int16) if b.dtype == np.int32 else b.astype(np.int16)   # libsndfile scales int32 -> int16
        raw = b.astype("<i2").tobytes()
        self.f.write(raw)
        self.nbytes += len(raw)

    def close(self):
        if self.f.closed:
            return
        self.f.seek(0)
        self.f.write(b"RIFF" + struct.pack("<I", 36 + self.nbytes) + b"WAVEfmt " +
                     struct.pack("<IHHIIHH", 16, 1, self.ch, self.sr, self.sr * self.ch * 2, self.ch * 2, 16) +
                     b"data" + struct.pack("<I", self.nbytes))
        self.f.close()


def read_float64(path):
    """sf.read(path, always_2d=True) equivalent for the PCM files this package writes: float64 in [-1, 1), samplerate."""
    i = info(path)
    with open(path, "rb") as f:
        f.seek(i.data_offset)
        raw = f.read(i.frames * i.channels * (i.bits // 8))
    if i.bits == 16:
        x = np.frombuffer(raw, "<i2").astype(np.float64) / 32768.0
    elif i.bits == 32:
        x = np.frombuffer(raw, "<i4").astype(np.float64) / 2147483648.0
    else:
        raise ValueError(f"unsupported WAV subtype {i.subtype}")
    return x.reshape(-1, i.channels), i.samplerate
