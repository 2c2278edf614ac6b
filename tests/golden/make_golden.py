"""Generate the committed golden vectors from the REFERENCE binary (libFLAC 1.4.3 as bundled by pyFLAC).

Run in the build container (needs oracle/_ref, i.e. /root/reference):  python tests/golden/make_golden.py
Each case stores the input PCM and the exact bytes libFLAC produced (seekable mode == FileEncoder output,
pyflac/encoder.py:393-426), so the restatement and the CUDA path can be pinned on machines where the
reference is absent.  Cases are small (a few frames) to keep the repository light.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from _checkers import ref_encode, ref_lib  # noqa: E402
from pyflac_b200.synth import corpus_signal  # noqa: E402

CASES = [
    # name, kind, n, channels, bps, sample_rate, level, blocksize
    ("music_s16_st_l5", "music", 4096 * 2 + 768, 2, 16, 48000, 5, 0),
    ("music_s16_st_l8", "music", 4096 * 2 + 100, 2, 16, 48000, 8, 0),
    ("music_s16_st_l0", "music", 1152 * 3 + 5, 2, 16, 44100, 0, 0),
    ("music_s16_st_l2", "music", 1152 * 3, 2, 16, 44100, 2, 0),
    ("music_s16_st_l3", "music", 4096 + 9, 2, 16, 48000, 3, 0),
    ("music_s16_st_l6", "music", 4096 + 1, 2, 16, 48000, 6, 0),
    ("music_s16_st_l7", "music", 4096, 2, 16, 48000, 7, 0),
    ("music_s16_st_l1", "lr_uncorr", 4096 * 3, 2, 16, 48000, 1, 1024),
    ("music_s16_st_l4", "mixed", 4096 * 3, 2, 16, 48000, 4, 1024),
    ("music_s24_mono_l8", "music", 4096 * 2, 1, 24, 192000, 8, 4096),
    ("music_s24_mono_l5", "music", 4096 + 333, 1, 24, 192000, 5, 4096),
    ("mixed_s16_st_l5", "mixed", 4096 * 3, 2, 16, 48000, 5, 0),
    ("wasted_s16_st_l5", "wasted", 4096 + 512, 2, 16, 48000, 5, 0),
    ("silence_s16_st_l5", "silence", 4096 + 100, 2, 16, 48000, 5, 0),
    ("noise_s16_mono_l5", "noise", 4096, 1, 16, 48000, 5, 0),
    ("square_s16_st_l5", "square", 4096, 2, 16, 48000, 5, 0),
    ("surround_s16_6ch_l5", "lr_uncorr", 4096 + 50, 6, 16, 48000, 5, 0),
    ("music_s8_st_l5", "music", 4096, 2, 8, 22050, 5, 0),
    ("music_s20_st_l5", "music", 2304 + 17, 2, 20, 96000, 5, 2304),
    ("music_s16_bs100_l5", "music", 350, 2, 16, 12345, 5, 100),
    # 32-bit (pyFLAC's int32 input): _limit_residual predictor search, no mid/side; lengths 0 or 1 mod 4 (see
    # oracle/flac_oracle.c:fixed_best_predictor_limit for why other tails are not reproducible from the binary itself)
    ("music_s32_mono_l5", "music", 4096 * 2 + 768, 1, 32, 44100, 5, 0),
    ("mixed_s32_mono_l8", "mixed", 4096 + 401, 1, 32, 96000, 8, 0),
    ("silence_s32_mono_l5", "silence", 4096 + 100, 1, 32, 48000, 5, 0),
    ("wasted_s32_3ch_l5", "wasted", 4096 + 512, 3, 32, 48000, 5, 0),
    ("music_s32_st_l3", "music", 4096 * 2, 2, 32, 48000, 3, 0),
    ("mixed_s32_st_l0", "mixed", 1152 * 3 + 4, 2, 32, 48000, 0, 0),
    # 32-bit stereo with mid/side analysis: 33-bit side channel
    ("music_s32_st_l5", "music", 4096 * 2 + 768, 2, 32, 48000, 5, 0),
    ("mixed_s32_st_l8", "mixed", 4096 + 400, 2, 32, 48000, 8, 0),
    ("lr_uncorr_s32_st_l2", "lr_uncorr", 1152 * 2 + 4, 2, 32, 44100, 2, 0),
]


def main():
    manifest = []
    assert ref_lib().ref_vendor_string() == b"reference libFLAC 1.4.3 20230623"
    for name, kind, n, ch, bps, sr, level, bs in CASES:
        x = corpus_signal(kind, n, ch, bps, seed=len(name), sample_rate=sr)
        flac, off, ln, smp = ref_encode(x, sr, bps, level, bs, seekable=True, with_index=True)
        np.save(os.path.join(HERE, name + ".pcm.npy"), x)
        with open(os.path.join(HERE, name + ".flac"), "wb") as f:
            f.write(flac)
        manifest.append(dict(name=name, kind=kind, n=n, channels=ch, bps=bps, sample_rate=sr, level=level,
                             blocksize=bs, flac_bytes=len(flac), frames=len(off),
                             frame_off=[int(v) for v in off], frame_len=[int(v) for v in ln]))
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(dict(vendor="reference libFLAC 1.4.3 20230623", cases=manifest), f, indent=1)
    print("wrote", len(manifest), "cases")


if __name__ == "__main__":
    main()
