"""Run by tools/host_logic_check.sh: metadata flow of the drop-in decoder (scratch no-device build) against libFLAC 1.4.3 on the CPU."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import _checkers as ck                                   # noqa: E402
from _flacapi import scripted_decode_session              # noqa: E402
from pyflac_b200.synth import music_like                  # noqa: E402

ck.build_checkers()
ref = C.CDLL(os.path.join(ck.ORACLE_DIR, "_ref", "libFLAC-12.1.0.so"))
ours = C.CDLL(sys.argv[1])
x = music_like(4096 * 3 + 77, 2, 44100, 16, seed=21)
data = ck.ref_encode(x, 44100, 16, 5, 0)
bad = 0
n = 0
for padlen in (None, 0, 100, 70000, 300000, 3000000):      # None: the stream as libFLAC wrote it
    big = data if padlen is None else data[:42] + bytes([1]) + padlen.to_bytes(3, "big") + bytes(padlen) + data[42:]
    nblocks = 2 if padlen is None else 3
    for ops in ([('single', 1)] * nblocks, [('meta',)], [('single', 1), ('meta',)], [('meta',), ('reset',), ('meta',)]):
        for seek in (True, False):
            if not seek and ('reset',) in ops:
                continue                                   # reset needs seekable input to rewind
            for rc in (8192, 1000, None):
                a = scripted_decode_session(ours, big, ops, meta=True, seekable=seek, read_chunk=rc)
                b = scripted_decode_session(ref, big, ops, meta=True, seekable=seek, read_chunk=rc)
                n += 1
                if a["events"] != b["events"]:
                    bad += 1
                    print("DIFF", padlen, ops, seek, rc, "\n ours", a["events"][:8], "\n ref ", b["events"][:8])
# truncated metadata: the stream ends inside a block
meta_end = 46 + int.from_bytes(data[43:46], 'big')         # STREAMINFO ends at byte 42, the VORBIS_COMMENT block behind it here
assert data[42] == 0x84 and meta_end < 200
for cut in (0, 3, 4, 7, 20, 41, 42, 45, 46, 60, meta_end - 1):     # (from meta_end on the sessions reach audio frames: GPU tests)
    for ops in ([('meta',)], [('single', 1)], [('single', 1)] * 4, [('end',)], [('single', 1), ('end',)], [('meta',), ('single', 2)]):
        a = scripted_decode_session(ours, data[:cut], ops, meta=True, seekable=False)
        b = scripted_decode_session(ref, data[:cut], ops, meta=True, seekable=False)
        n += 1
        if a["events"] != b["events"]:
            bad += 1
            print("DIFF truncated at", cut, ops, "\n ours", a["events"], "\n ref ", b["events"])
print(f"{n} sessions, {bad} differ from libFLAC")
sys.exit(1 if bad else 0)
