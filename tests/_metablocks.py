"""Hand-built FLAC metadata blocks (format: https://xiph.org/flac/format.html#metadata_block) for the drop-in decoder's
metadata-callback tests: every block type pyFLAC's cdef declares (builder/decoder.py:233-365)."""
import struct


def block(btype, payload, last=False):
    return bytes([(0x80 if last else 0) | btype]) + len(payload).to_bytes(3, "big") + payload


def application(app_id, data):
    return app_id + data


def seektable(points):
    return b"".join(struct.pack(">QQH", s, o, n) for s, o, n in points)


def vorbis_comment(vendor, comments):
    out = struct.pack("<I", len(vendor)) + vendor + struct.pack("<I", len(comments))
    for c in comments:
        out += struct.pack("<I", len(c)) + c
    return out


def cuesheet(mcn, lead_in, is_cd, tracks):
    """tracks: (offset, number, isrc(12 bytes), type, pre_emphasis, [(index_offset, index_number), ...])"""
    out = mcn.ljust(128, b"\0") + struct.pack(">Q", lead_in) + bytes([0x80 if is_cd else 0]) + bytes(258) + bytes([len(tracks)])
    for off, num, isrc, ttype, pre, idx in tracks:
        out += struct.pack(">QB", off, num) + isrc.ljust(12, b"\0") + bytes([(ttype << 7) | (pre << 6)]) + bytes(13) + bytes([len(idx)])
        for ioff, inum in idx:
            out += struct.pack(">QB", ioff, inum) + bytes(3)
    return out


def picture(ptype, mime, desc, w, h, depth, colors, data):
    return (struct.pack(">II", ptype, len(mime)) + mime + struct.pack(">I", len(desc)) + desc +
            struct.pack(">IIIII", w, h, depth, colors, len(data)) + data)


def rich_stream(flac):
    """flac = b'fLaC' + STREAMINFO + one more (last) block + frames, as libFLAC writes it: the same audio with one block of every
    type in between."""
    assert flac[:4] == b"fLaC" and flac[4] == 0 and flac[42] & 0x80
    tail_len = int.from_bytes(flac[43:46], "big")
    frames = flac[46 + tail_len:]
    blocks = [
        block(1, bytes(1000)),
        block(2, application(b"abcd", b"application payload \x00\x01\x02")),
        block(2, application(b"wxyz", b"")),
        block(3, seektable([(0, 0, 4096), (4096, 5000, 4096), (0xFFFFFFFFFFFFFFFF, 0, 0)])),
        block(4, vorbis_comment(b"a vendor", [b"TITLE=one", b"ARTIST=two\xc3\xa9", b""])),
        block(5, cuesheet(b"1234567890123", 88200, True,
                          [(0, 1, b"ABCDE1234567", 0, 1, [(0, 0), (588, 1)]), (441000, 2, b"", 1, 0, [(0, 1)]), (882000, 170, b"", 0, 0, [])])),
        block(6, picture(3, b"image/png", b"cover \xe2\x9c\x93", 32, 24, 8, 0, bytes(range(256)) * 3)),
        block(1, b""),
        block(3, b""),
        block(4, vorbis_comment(b"", [])),
        block(6, picture(0, b"", b"", 0, 0, 0, 0, b"")),
        block(50, b"a block of a type nobody knows"),
        block(2, application(b"abcd", b"second of its id"), last=True),
    ]
    return flac[:42] + b"".join(blocks) + frames, len(blocks) + 1
