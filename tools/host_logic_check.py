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
def control_session(L, data, seekable=True):
    """Getters, refused calls and return values around init / metadata / flush / reset / finish (no audio frame is reached)."""
    log = []
    P = "FLAC__stream_decoder_"
    def f(name, res=C.c_int, args=(C.c_void_p,)):
        fn = getattr(L, P + name); fn.restype = res; fn.argtypes = list(args); return fn
    new = f("new", C.c_void_p, ())
    getters = [("get_state", C.c_int), ("get_md5_checking", C.c_int), ("get_total_samples", C.c_uint64), ("get_channels", C.c_uint32),
               ("get_channel_assignment", C.c_int), ("get_bits_per_sample", C.c_uint32), ("get_sample_rate", C.c_uint32), ("get_blocksize", C.c_uint32)]
    def snap(tag, with_pos=True):
        vals = [int(f(n, r)(d)) for n, r in getters]
        s = f("get_resolved_state_string", C.c_char_p)(d)
        pos = C.c_uint64(12345)
        rv = f("get_decode_position", C.c_int, (C.c_void_p, C.POINTER(C.c_uint64)))(d, C.byref(pos))
        log.append((tag, vals, s, rv, pos.value if rv and with_pos else None))
    pos = [0]
    def r(dec, buf, pbytes, cd):
        k = min(pbytes[0], len(data) - pos[0])
        if k == 0:
            pbytes[0] = 0; return 1
        C.memmove(buf, data[pos[0]:pos[0]+k], k); pos[0] += k; pbytes[0] = k; return 0
    def sk(dec, off, cd): pos[0] = min(int(off), len(data)); return 0
    def tl(dec, poff, cd): poff[0] = pos[0]; return 0
    def ln(dec, plen, cd): plen[0] = len(data); return 0
    def ef(dec, cd): return int(pos[0] >= len(data))
    def w(dec, frame, buffers, cd): return 0
    def e(dec, status, cd): log.append(('e', status))
    cbs = (DEC_READ_CB(r), DEC_SEEK_CB(sk), DEC_TELL_CB(tl), DEC_LENGTH_CB(ln), DEC_EOF_CB(ef), DEC_WRITE_CB(w), DEC_ERROR_CB(e))
    null = lambda T: C.cast(None, T)
    d = new()
    snap("new")
    for name in ("process_single", "process_until_end_of_metadata", "process_until_end_of_stream", "skip_single_frame", "flush", "reset", "finish"):
        log.append((name + " uninit", f(name)(d)))
    log.append(("seek uninit", f("seek_absolute", C.c_int, (C.c_void_p, C.c_uint64))(d, 0)))
    log.append(("set_md5", f("set_md5_checking", C.c_int, (C.c_void_p, C.c_int))(d, 1)))
    snap("after set_md5")
    init = getattr(L, P + "init_stream"); init.restype = C.c_int
    init.argtypes = [C.c_void_p, DEC_READ_CB, DEC_SEEK_CB, DEC_TELL_CB, DEC_LENGTH_CB, DEC_EOF_CB, DEC_WRITE_CB, C.c_void_p, DEC_ERROR_CB, C.c_void_p]
    # invalid callback sets
    log.append(("init no read", init(d, null(DEC_READ_CB), cbs[1], cbs[2], cbs[3], cbs[4], cbs[5], None, cbs[6], None)))
    log.append(("init seek w/o tell", init(d, cbs[0], cbs[1], null(DEC_TELL_CB), cbs[3], cbs[4], cbs[5], None, cbs[6], None)))
    snap("after bad init")
    if seekable:
        log.append(("init", init(d, cbs[0], cbs[1], cbs[2], cbs[3], cbs[4], cbs[5], None, cbs[6], None)))
    else:
        log.append(("init", init(d, cbs[0], null(DEC_SEEK_CB), null(DEC_TELL_CB), null(DEC_LENGTH_CB), null(DEC_EOF_CB), cbs[5], None, cbs[6], None)))
    snap("after init")
    log.append(("init again", init(d, cbs[0], cbs[1], cbs[2], cbs[3], cbs[4], cbs[5], None, cbs[6], None)))
    log.append(("set_md5 after init", f("set_md5_checking", C.c_int, (C.c_void_p, C.c_int))(d, 0)))
    log.append(("respond_all after init", f("set_metadata_respond_all")(d)))
    log.append(("single", f("process_single")(d))); snap("after single")
    log.append(("single", f("process_single")(d))); snap("after single 2")
    log.append(("meta", f("process_until_end_of_metadata")(d))); snap("after meta")
    log.append(("meta again", f("process_until_end_of_metadata")(d))); snap("after meta again")
    log.append(("flush", f("flush")(d))); snap("after flush", with_pos=False)     # (how far each library had read ahead)
    if seekable:
        log.append(("reset", f("reset")(d))); snap("after reset")
        log.append(("meta", f("process_until_end_of_metadata")(d))); snap("after meta 2")
    log.append(("finish", f("finish")(d))); snap("after finish")
    log.append(("finish again", f("finish")(d)))
    f("delete", None)(d)
    return log


from _flacapi import DEC_READ_CB, DEC_SEEK_CB, DEC_TELL_CB, DEC_LENGTH_CB, DEC_EOF_CB, DEC_WRITE_CB, DEC_ERROR_CB   # noqa: E402
from _metablocks import rich_stream as _rich                 # noqa: E402
for dat in (data, _rich(data)[0]):
    for seekable in (True, False):
        a, b = control_session(ours, dat, seekable), control_session(ref, dat, seekable)
        n += 1
        if a != b:
            bad += 1
            for ea, eb in zip(a, b):
                if ea != eb:
                    print("DIFF control surface, seekable =", seekable, "\n ours", ea, "\n ref ", eb)

# every block type through the metadata callback under the respond / ignore filters (tests/_metablocks.py)
from _metablocks import rich_stream, block, vorbis_comment, picture, cuesheet   # noqa: E402
rich, nb = rich_stream(data)
FILTERS = [(), (('respond_all',),), (('respond', 4),), (('respond', 2), ('ignore_application', b"abcd")), (('respond_application', b"wxyz"),),
           (('respond_all',), ('ignore', 0)), (('respond_all',), ('ignore', 2), ('respond_application', b"abcd"), ('respond_application', b"none")),
           (('ignore_all',), ('respond', 3), ('respond', 6), ('respond', 50)), (('respond_all',), ('ignore_application', b"wxyz"), ('respond', 2)),
           (('respond', 127),), (('respond', 126), ('ignore', 1), ('respond', 5)), (('respond_all',), ('ignore_all',), ('respond', 1))]
for resp in FILTERS:
    for ops in ([('single', 1)] * nb, [('meta',)], [('single', 3), ('meta',)], [('meta',), ('reset',), ('single', 2)]):
        for rc in (1000, None):
            a = scripted_decode_session(ours, rich, ops, meta=True, seekable=True, read_chunk=rc, respond=resp)
            b = scripted_decode_session(ref, rich, ops, meta=True, seekable=True, read_chunk=rc, respond=resp)
            n += 1
            if a["events"] != b["events"]:
                bad += 1
                for i, (ea, eb) in enumerate(zip(a["events"], b["events"])):
                    if ea != eb:
                        print("DIFF rich", resp, ops, rc, "event", i, "\n ours", str(ea)[:300], "\n ref ", str(eb)[:300])
                        break
                else:
                    print("DIFF rich (length)", resp, ops, rc, len(a["events"]), len(b["events"]))
# filter calls after init are refused; finish() puts the filter back to its default
for L in (ours, ref):
    pass
# blocks whose content does not fit their length
odd = {
    "vorbis comment cut inside an entry": block(4, vorbis_comment(b"vendor", [b"A=1", b"B=2"])[:-2]),
    "vorbis comment with a huge count": block(4, vorbis_comment(b"vendor", [])[:-4] + (1 << 30).to_bytes(4, "little")),
    "vorbis comment of 4 bytes": block(4, bytes(4)),
    "picture cut inside its data": block(6, picture(3, b"image/png", b"d", 1, 1, 8, 0, bytes(100))[:-10]),
    "cuesheet cut inside a track": block(5, cuesheet(b"", 0, False, [(0, 1, b"", 0, 0, [(0, 1)])])[:-5]),
    "seektable of 20 bytes": block(3, bytes(20)),
    "application of 3 bytes": block(2, b"abc"),
}
import struct                                              # noqa: E402
odd.update({
    "vorbis vendor longer than the block": block(4, struct.pack("<I", 1000) + b"short" + struct.pack("<I", 0)),
    "vorbis entry longer than the block": block(4, vorbis_comment(b"v", [b"A=1"])[:-3] + b"" ) ,
    "vorbis entry length far too big": block(4, struct.pack("<I", 1) + b"v" + struct.pack("<I", 1) + struct.pack("<I", 5000) + b"abc"),
    "vorbis trailing bytes": block(4, vorbis_comment(b"v", [b"A=1"]) + b"extra"),
    "vorbis fewer entries than announced": block(4, vorbis_comment(b"v", [b"A=1"])[:5] + struct.pack("<I", 3) + struct.pack("<I", 3) + b"A=1"),
    "picture mime longer than the block": block(6, struct.pack(">II", 3, 5000) + b"image/png"),
    "picture trailing bytes": block(6, picture(3, b"image/png", b"d", 1, 1, 8, 0, bytes(10)) + b"extra"),
    "picture data_length too big": block(6, picture(3, b"image/png", b"d", 1, 1, 8, 0, bytes(10))[:-10 - 4] + struct.pack(">I", 500) + bytes(10)),
    "cuesheet trailing bytes": block(5, cuesheet(b"", 0, False, [(0, 1, b"", 0, 0, [(0, 1)])]) + b"xx"),
    "cuesheet of 10 bytes": block(5, bytes(10)),
    "streaminfo of 40 bytes then padding": block(1, bytes(7)),
    "unknown type, empty": block(77, b""),
    "vorbis two junk bytes where an entry should start": block(4, vorbis_comment(b"v", [b"A=1"])[:5] + struct.pack("<I", 3) + struct.pack("<I", 3) + b"A=1" + b"zz"),
    "vorbis count 100000": block(4, struct.pack("<I", 1) + b"v" + struct.pack("<I", 100000)),
    "vorbis count 100001": block(4, struct.pack("<I", 1) + b"v" + struct.pack("<I", 100001)),
    "vorbis vendor only, no count": block(4, struct.pack("<I", 4) + b"vend"),
    "application of 4 bytes": block(2, b"abcd"),
    "seektable of 17 bytes": block(3, bytes(17)),
    "seektable of 36 bytes": block(3, bytes(36)),
    "picture type 20": block(6, picture(20, b"image/png", b"d", 1, 1, 8, 0, bytes(4))),
    "picture type 21": block(6, picture(21, b"image/png", b"d", 1, 1, 8, 0, bytes(4))),
    "picture type 2^32 - 1": block(6, picture(0xffffffff, b"image/png", b"d", 1, 1, 8, 0, bytes(4))),
})
for name, blk in odd.items():
    strm = data[:42] + blk + data[42:]
    for ops in ([('single', 1)] * 2, [('meta',)], [('meta',), ('meta',)]):       # (behind a block reported as BAD_METADATA the sessions reach audio frames: GPU tests)
        a = scripted_decode_session(ours, strm, ops, meta=True, seekable=False, respond=(('respond_all',),))
        b = scripted_decode_session(ref, strm, ops, meta=True, seekable=False, respond=(('respond_all',),))
        n += 1
        if a["events"] != b["events"]:
            bad += 1
            print("DIFF odd block:", name, ops, "\n ours", str(a["events"])[:400], "\n ref ", str(b["events"])[:400])
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
# bytes in front of "fLaC": an ID3v2 tag (skipped), junk (LOST_SYNC once per run), both, tags of many input slices
def id3(nbytes, flags=0):
    return b"ID3\x03\x00" + bytes([flags]) + bytes([(nbytes >> 21) & 0x7f, (nbytes >> 14) & 0x7f, (nbytes >> 7) & 0x7f, nbytes & 0x7f]) + bytes(nbytes)


heads = {"id3 100": id3(100), "id3 5000": id3(5000), "id3 300000": id3(300000), "id3 3000000": id3(3000000), "junk 10": b"0123456789", "junk 3": b"abc",
         "two id3": id3(10) + id3(20), "id3 then junk": id3(10) + b"xy", "junk then id3": b"xy" + id3(10), "zeros 1000": bytes(1000),
         "zeros 3000000": bytes(3000000), "partial markers": b"fLxfLaxffLa.", "I, ID, IDx": b"I.ID.IDx", "ID3 cut short": b"ID3\x03\x00", "ff ff 00": b"\xff\xff\x00",
         "ff 00 ff 01": b"\xff\x00\xff\x01", "f": b"f", "fLa": b"fLa"}
for name, head in heads.items():
    for body in (data, b""):
        if body and name in ("f", "fLa", "partial markers"):
            continue    # the real marker is then lost ("ffLaC": the second f is not looked at twice) and libFLAC decodes the frames without STREAMINFO
        for ops in ([('single', 1)] * 2, [('meta',)], [('meta',), ('meta',)]):
            for rc in (1000, None):
                a = scripted_decode_session(ours, head + body, ops, meta=True, seekable=True, read_chunk=rc)
                b = scripted_decode_session(ref, head + body, ops, meta=True, seekable=True, read_chunk=rc)
                n += 1
                if a["events"] != b["events"]:
                    bad += 1
                    print("DIFF head:", name, len(body), ops, rc, "\n ours", str(a["events"])[:300], "\n ref ", str(b["events"])[:300])
    if name in ("two id3", "id3 then junk"):       # a seek as the first call fails with the failing search, and reports nothing
        for ops in ([('seek', 5000), ('meta',)], [('seek', 5000), ('single', 1), ('single', 1)]):
            a = scripted_decode_session(ours, head + data, ops, meta=True, seekable=True)
            b = scripted_decode_session(ref, head + data, ops, meta=True, seekable=True)
            n += 1
            if a["events"] != b["events"]:
                bad += 1
                print("DIFF head, seek first:", name, ops, "\n ours", str(a["events"])[:300], "\n ref ", str(b["events"])[:300])
    if name in ("f", "fLa", "partial markers", "ID3 cut short"):      # (the last: libFLAC's position at the end of input is that of its word-wise reader)
        continue
    a, b = control_session(ours, head + data, True), control_session(ref, head + data, True)
    n += 1
    if a != b:
        bad += 1
        for ea, eb in zip(a, b):
            if ea != eb:
                print("DIFF head, control surface:", name, "\n ours", ea, "\n ref ", eb)

# a read callback that aborts, hands over nothing, or claims the end of the stream while the metadata is being read (small reads, so
# that both libraries are still inside the metadata at that call); and the same sessions over a FILE
for k in range(0, 12):
    for what in ("abort", "empty", "eof"):
        for ops in ([('meta',)], [('single', 1)] * 4, [('meta',), ('meta',)] + ([('single', 1)] if what != 'empty' else [])):
            a = scripted_decode_session(ours, _rich(data)[0], ops, meta=True, seekable=False, read_chunk=64, respond=(('respond_all',),), read_script={k: what})
            b = scripted_decode_session(ref, _rich(data)[0], ops, meta=True, seekable=False, read_chunk=64, respond=(('respond_all',),), read_script={k: what})
            n += 1
            if a["events"] != b["events"]:
                bad += 1
                print("DIFF read callback", k, what, ops, "\n ours", str(a["events"])[-300:], "\n ref ", str(b["events"])[-300:])
import tempfile as _tf                                     # noqa: E402
with _tf.TemporaryDirectory() as tmp:
    for name, blob in (("plain", data), ("rich", _rich(data)[0]), ("id3", id3(5000) + data), ("junk", b"0123456789" + data), ("cut", data[:60]), ("empty", b"")):
        fn = os.path.join(tmp, name + ".flac")
        with open(fn, "wb") as f:
            f.write(blob)
        for ops in ([('meta',)], [('single', 1)] * 2, [('meta',), ('reset',), ('meta',)]):
            for resp in ((), (('respond_all',),)):
                a = scripted_decode_session(ours, blob, ops, meta=True, path=fn, respond=resp)
                b = scripted_decode_session(ref, blob, ops, meta=True, path=fn, respond=resp)
                n += 1
                if a != b:
                    bad += 1
                    print("DIFF file decode", name, ops, resp, "\n ours", str(a)[-400:], "\n ref ", str(b)[-400:])

# encoder: a stream without a single sample never reaches a kernel -- header writes, tell / seek traffic, the STREAMINFO rewrite at
# finish (min framesize 2^24 - 1: no frame ever lowered it), the metadata callback, and what happens when a callback fails
import numpy as np                                         # noqa: E402
from _flacapi import encode_session                        # noqa: E402
for ch, bps, sr, level, bs in ((2, 16, 44100, 5, 0), (1, 24, 96000, 8, 4096), (8, 8, 8000, 0, 1152), (2, 32, 192000, 3, 256), (3, 12, 22050, 5, 0)):
    x = np.zeros((0, ch), np.int32)
    for seekable in (True, False):
        for meta in (True, False):
            fails = [None] + [{k: i} for k in ("write", "seek", "tell") for i in range(7)] if seekable else [None] + [{"write": i} for i in range(4)]
            for fail in fails:
                a = encode_session(ours, x, sr, bps, level, bs, seekable=seekable, metadata=meta, fail=fail)
                b = encode_session(ref, x, sr, bps, level, bs, seekable=seekable, metadata=meta, fail=fail)
                n += 1
                if a != b:
                    bad += 1
                    print("DIFF empty encode", (ch, bps, sr, level, bs), seekable, meta, fail)
                    for k in a:
                        if a[k] != b.get(k):
                            print("  ", k, "\n    ours", str(a[k])[-400:], "\n    ref ", str(b.get(k))[-400:])
# the same through FLAC__stream_encoder_init_file (FileEncoder): statuses, states, the file on disk
import tempfile                                            # noqa: E402


def file_session(L, path, ch, bps, sr, est):
    e = L.FLAC__stream_encoder_new()
    L.FLAC__stream_encoder_set_channels(e, ch)
    L.FLAC__stream_encoder_set_bits_per_sample(e, bps)
    L.FLAC__stream_encoder_set_sample_rate(e, sr)
    if est is not None:
        L.FLAC__stream_encoder_set_total_samples_estimate.argtypes = [C.c_void_p, C.c_uint64]
        L.FLAC__stream_encoder_set_total_samples_estimate(e, est)
    st = L.FLAC__stream_encoder_init_file(e, path.encode(), None, None)
    s1 = L.FLAC__stream_encoder_get_state(e)
    st2 = L.FLAC__stream_encoder_init_file(e, path.encode(), None, None)         # ALREADY_INITIALIZED
    fin = L.FLAC__stream_encoder_finish(e)
    s2 = L.FLAC__stream_encoder_get_state(e)
    L.FLAC__stream_encoder_delete(e)
    return st, s1, st2, fin, s2, open(path, "rb").read() if os.path.exists(path) else None


with tempfile.TemporaryDirectory() as tmp:
    for ch, bps, sr, est, name in ((2, 16, 44100, None, "a.flac"), (1, 24, 96000, 1000, "a.flac"), (6, 20, 48000, 0, "a.flac"), (2, 16, 44100, None, "no/such/dir.flac")):
        a = file_session(ours, os.path.join(tmp, "ours_" + name), ch, bps, sr, est)
        b = file_session(ref, os.path.join(tmp, "ref_" + name), ch, bps, sr, est)
        n += 1
        if a != b:
            bad += 1
            print("DIFF file encode", (ch, bps, sr, est, name), "\n ours", a, "\n ref ", b)
# what an initialised encoder's getters report (the blocksize, precision and mid/side switches as init resolved them), that setters
# are refused from then on, and process() calls that do not fill a block
ENC_GETTERS = ["state", "verify", "streamable_subset", "channels", "bits_per_sample", "sample_rate", "blocksize", "do_mid_side_stereo",
               "loose_mid_side_stereo", "max_lpc_order", "qlp_coeff_precision", "do_qlp_coeff_prec_search", "do_escape_coding",
               "do_exhaustive_model_search", "min_residual_partition_order", "max_residual_partition_order", "rice_parameter_search_dist",
               "limit_min_bitrate"]


def getter_session(L, ch, bps, sr, level, bs, subset):
    from _flacapi import WRITE_CB as _W, SEEK_CB as _S, TELL_CB as _T, META_CB as _M, _proto as _p
    _p(L)

    def snap(e):
        out = []
        for g in ENC_GETTERS:
            f = getattr(L, "FLAC__stream_encoder_get_" + g)
            f.argtypes = [C.c_void_p]
            f.restype = C.c_uint32
            out.append(f(e))
        return out
    w = _W(lambda *a: 0)
    e = L.FLAC__stream_encoder_new()
    L.FLAC__stream_encoder_set_channels(e, ch)
    L.FLAC__stream_encoder_set_bits_per_sample(e, bps)
    L.FLAC__stream_encoder_set_sample_rate(e, sr)
    L.FLAC__stream_encoder_set_compression_level(e, level)
    L.FLAC__stream_encoder_set_blocksize(e, bs)
    L.FLAC__stream_encoder_set_streamable_subset(e, subset)
    null = lambda T: C.cast(None, T)  # noqa: E731
    st = L.FLAC__stream_encoder_init_stream(e, w, null(_S), null(_T), null(_M), None)
    a = snap(e) if st == 0 else None                     # (after a refused init libFLAC's getters show how far its init got)
    r = [L.FLAC__stream_encoder_set_channels(e, 1), L.FLAC__stream_encoder_set_blocksize(e, 1234), L.FLAC__stream_encoder_set_compression_level(e, 0)] if st == 0 else None
    x = np.zeros((10, ch), np.int32)
    p0 = L.FLAC__stream_encoder_process_interleaved(e, x.ctypes.data, 0) if st == 0 else None
    p1 = L.FLAC__stream_encoder_process_interleaved(e, x.ctypes.data, 10) if st == 0 else None
    b = snap(e) if st == 0 else None
    L.FLAC__stream_encoder_delete(e)
    return st, a, r, p0, p1, b


for ch in (1, 2, 8):
    for bps in (8, 16, 24, 32):
        for sr in (8000, 44100, 96000, 384000, 1048575):
            for level in (0, 2, 5, 8):
                for bs in (0, 16, 1152, 4608, 16384):
                    for subset in (1, 0):
                        a = getter_session(ours, ch, bps, sr, level, bs, subset)
                        b = getter_session(ref, ch, bps, sr, level, bs, subset)
                        n += 1
                        if a != b:
                            bad += 1
                            if bad < 20:
                                print("DIFF encoder getters", (ch, bps, sr, level, bs, subset), "\n ours", a, "\n ref ", b)

# an encoder deleted without finish(): torn down without a callback (samples that never filled a block are dropped, the file keeps
# the STREAMINFO written at init)
from _flacapi import WRITE_CB, SEEK_CB, TELL_CB, META_CB, _proto   # noqa: E402


def delete_session(L, x, path=None):
    _proto(L)
    log = []

    def w(enc, buf, nbytes, samples, frame, cd):
        log.append(('write', bytes(buf[:nbytes]), samples, frame))
        return 0

    def s_(enc, off, cd):
        log.append(('seek', off))
        return 0

    def t(enc, poff, cd):
        poff[0] = 0
        log.append(('tell',))
        return 0

    def m(enc, md, cd):
        log.append(('meta',))
    cbs = (WRITE_CB(w), SEEK_CB(s_), TELL_CB(t), META_CB(m))
    e = L.FLAC__stream_encoder_new()
    L.FLAC__stream_encoder_set_channels(e, x.shape[1])
    L.FLAC__stream_encoder_set_bits_per_sample(e, 16)
    L.FLAC__stream_encoder_set_sample_rate(e, 44100)
    st = L.FLAC__stream_encoder_init_file(e, path.encode(), None, None) if path else L.FLAC__stream_encoder_init_stream(e, cbs[0], cbs[1], cbs[2], cbs[3], None)
    x32 = np.ascontiguousarray(x.astype(np.int32))
    ok = L.FLAC__stream_encoder_process_interleaved(e, x32.ctypes.data, len(x32)) if len(x32) else 1
    log.append(('processed', st, ok, L.FLAC__stream_encoder_get_state(e)))
    L.FLAC__stream_encoder_delete(e)
    log.append(('deleted', open(path, "rb").read() if path else None))
    return log


with tempfile.TemporaryDirectory() as tmp:
    for nsamp in (0, 1, 100, 4096):                       # (4097 samples would fill a block: a frame, a kernel)
        for ch in (1, 2):
            xs = (np.arange(nsamp * ch, dtype=np.int32).reshape(nsamp, ch) * 37) % 2000 - 1000
            for path in (None, os.path.join(tmp, "d.flac")):
                a = delete_session(ours, xs, path)
                b = delete_session(ref, xs, path)
                n += 1
                if a != b:
                    bad += 1
                    print("DIFF delete without finish", nsamp, ch, bool(path), "\n ours", str(a)[-300:], "\n ref ", str(b)[-300:])
print(f"{n} sessions, {bad} differ from libFLAC")
sys.exit(1 if bad else 0)
