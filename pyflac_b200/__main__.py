"""`python -m pyflac_b200 <file>`: WAV -> FLAC or FLAC -> WAV by magic sniffing (reference pyflac/__main__.py:20-57)."""
import argparse
from pathlib import Path

from . import FileDecoder, FileEncoder


def main():
    p = argparse.ArgumentParser(prog="pyflac_b200", description="B200-native FLAC encoder/decoder with pyFLAC's interface")
    p.add_argument("input_file", type=Path)
    p.add_argument("-o", "--output-file", type=Path)
    p.add_argument("-c", "--compression-level", type=int, choices=range(9), default=5)
    p.add_argument("-b", "--block-size", type=int, default=0)
    # as in the reference: verification is ON unless -v is given (pyflac/__main__.py:31)
    p.add_argument("-v", "--verify", action="store_false", default=True)
    a = p.parse_args()
    with open(a.input_file, "rb") as f:
        magic = f.read(4)
    if magic == b"RIFF":
        out = a.output_file or a.input_file.with_suffix(".flac")
        FileEncoder(a.input_file, out, a.compression_level, a.block_size, verify=a.verify).process()
    elif magic == b"fLaC":
        out = a.output_file or a.input_file.with_suffix(".wav")
        FileDecoder(a.input_file, out).process()
    else:
        raise ValueError("Please provide either a WAV or a FLAC file")
    print(out)


if __name__ == "__main__":
    main()
