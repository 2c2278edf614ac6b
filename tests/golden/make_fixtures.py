"""Import the reference's own decode fixtures (pyFLAC tests/data/*.flac, CC0 audio from freesound.org) as golden vectors.

Run in the build container (needs /root/reference and oracle/_ref):  python tests/golden/make_fixtures.py
For every fixture this stores
  * the .flac file itself (a data vector: libFLAC 1.3.3 / 1.4.2 output with SEEKTABLE, VORBIS_COMMENT and 8 KiB PADDING),
  * the MD5 of the PCM in the matching .wav (what pyFLAC's tests compare a decode against; equals the STREAMINFO MD5),
  * for the fixtures pyFLAC can encode (16- and 32-bit): the exact file the bundled libFLAC 1.4.3 writes for that PCM at
    level 5 through FileEncoder's seekable path (pyflac/encoder.py:393-426), `<name>_l5.flac`.
The PCM itself is not stored: tests obtain it by decoding the fixture with the oracle and check it against the MD5.
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "fixtures")
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _checkers import build_checkers, oracle_decode, ref_encode  # noqa: E402
from pyflac_b200 import wav  # noqa: E402

SRC = "/root/reference/tests/data"
NAMES = ["mono", "stereo", "surround", "32bit", "8bit"]


def main():
    build_checkers()
    os.makedirs(OUT, exist_ok=True)
    manifest = []
    for name in NAMES:
        src = os.path.join(SRC, name + ".flac")
        shutil.copyfile(src, os.path.join(OUT, name + ".flac"))
        os.chmod(os.path.join(OUT, name + ".flac"), 0o644)
        data = open(src, "rb").read()
        pcm, info = oracle_decode(data)
        width = (info["bps"] + 7) // 8
        raw = np.ascontiguousarray(pcm.astype("<i4")).tobytes()
        le = raw if width == 4 else b"".join(raw[i:i + width] for i in range(0, len(raw), 4))
        entry = dict(name=name, channels=int(pcm.shape[1]), bps=int(info["bps"]), sample_rate=int(info["sample_rate"]),
                     samples=int(pcm.shape[0]), flac_bytes=len(data), streaminfo_md5=data[26:42].hex(),
                     pcm_md5=hashlib.md5(le).hexdigest(), wav_md5=None, level5=None)
        wpath = os.path.join(SRC, name + ".wav")
        if os.path.exists(wpath):
            x, sr = wav.read_pcm(wpath)
            x = x.reshape(x.shape[0], -1)
            assert sr == info["sample_rate"] and np.array_equal(x.astype(np.int64), pcm.astype(np.int64)), name
            entry["wav_md5"] = hashlib.md5(np.ascontiguousarray(x).tobytes()).hexdigest()
            enc = ref_encode(np.ascontiguousarray(x), sr, info["bps"], 5, 0, seekable=True)
            with open(os.path.join(OUT, name + "_l5.flac"), "wb") as f:
                f.write(enc)
            entry["level5"] = dict(file=name + "_l5.flac", flac_bytes=len(enc))
        manifest.append(entry)
    with open(os.path.join(OUT, "fixtures.json"), "w") as f:
        json.dump(dict(source="pyFLAC tests/data (CC0, freesound.org)", cases=manifest), f, indent=1)
    print("wrote", len(manifest), "fixtures")


if __name__ == "__main__":
    main()
