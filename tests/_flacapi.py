"""Drive the libFLAC stream-encoder/decoder C API of ANY shared library through ctypes and log what the callbacks see.

Used by the tests to run the same session against (a) libflacb200.so's drop-in layer and (b) the reference's
libFLAC 1.4.3 (oracle/_ref) and compare the callback streams byte for byte -- the contract pyFLAC's cffi
trampolines rely on (pyflac/encoder.py:429-494, pyflac/decoder.py:394-549).
"""
import ctypes as C

import numpy as np

WRITE_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_ubyte), C.c_size_t, C.c_uint32, C.c_uint32, C.c_void_p)
SEEK_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_void_p)
TELL_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p)
META_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p)

DEC_READ_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_ubyte), C.POINTER(C.c_size_t), C.c_void_p)
DEC_WRITE_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.POINTER(C.c_int32)), C.c_void_p)
DEC_ERROR_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_void_p)


class StreamInfoView(C.Structure):
    _fields_ = [("type", C.c_int), ("is_last", C.c_int), ("length", C.c_uint32), ("pad", C.c_uint32),
                ("min_blocksize", C.c_uint32), ("max_blocksize", C.c_uint32), ("min_framesize", C.c_uint32),
                ("max_framesize", C.c_uint32), ("sample_rate", C.c_uint32), ("channels", C.c_uint32),
                ("bits_per_sample", C.c_uint32), ("pad2", C.c_uint32), ("total_samples", C.c_uint64), ("md5sum", C.c_ubyte * 16)]


class FrameHeaderView(C.Structure):
    _fields_ = [("blocksize", C.c_uint32), ("sample_rate", C.c_uint32), ("channels", C.c_uint32),
                ("channel_assignment", C.c_int), ("bits_per_sample", C.c_uint32), ("number_type", C.c_int),
                ("number", C.c_uint64), ("crc", C.c_uint8)]


def metadata_view(md):
    """What a metadata callback can see of a FLAC__StreamMetadata (builder/decoder.py:233-365), read through the C layout of
    libFLAC 1.4.3 on LP64: ('m', type, is_last, length, ...content of the block...)."""
    def u32(a): return C.c_uint32.from_address(a).value
    def u64(a): return C.c_uint64.from_address(a).value
    def ptr(a): return C.c_void_p.from_address(a).value or 0
    def raw(a, n): return C.string_at(a, n) if a and n else b""
    v = C.cast(md, C.POINTER(StreamInfoView)).contents
    t, base = int(v.type), (md if isinstance(md, int) else C.cast(md, C.c_void_p).value) + 16       # the union starts at byte 16
    head = ('m', t, int(v.is_last), int(v.length))
    if t == 0:
        return head + (int(v.sample_rate), int(v.channels), int(v.bits_per_sample), int(v.total_samples))
    if t == 1:
        return head
    if t == 2:      # id[4], data*
        return head + (raw(base, 4), raw(ptr(base + 8), v.length - 4))
    if t == 3:      # num_points, points* -> {u64 sample_number, u64 stream_offset, u32 frame_samples} (24 bytes each)
        n, pts = u32(base), ptr(base + 8)
        return head + (tuple((u64(pts + 24 * i), u64(pts + 24 * i + 8), u32(pts + 24 * i + 16)) for i in range(n)),)
    if t == 4:      # vendor {u32 length, entry*}, num_comments, comments* -> {u32 length, entry*} (16 bytes each); entries are NUL-terminated
        def entry(a): return raw(ptr(a + 8), u32(a) + 1)
        n, cs = u32(base + 16), ptr(base + 24)
        return head + (entry(base), tuple(entry(cs + 16 * i) for i in range(n)))
    if t == 5:      # char mcn[129], u64 lead_in @136, int is_cd @144, u32 num_tracks @148, tracks* @152
        nt, ts = u32(base + 148), ptr(base + 152)
        tracks = []
        for i in range(nt):     # u64 offset, u8 number @8, char isrc[13] @9, type:1 pre_emphasis:1 @22, u8 num_indices @23, indices* @24 (32 bytes)
            a = ts + 32 * i
            ni, ix = raw(a + 23, 1)[0], ptr(a + 24)
            tracks.append((u64(a), raw(a + 8, 1)[0], raw(a + 9, 13), raw(a + 22, 1)[0] & 3, ni,
                           tuple((u64(ix + 16 * k), raw(ix + 16 * k + 8, 1)[0]) for k in range(ni))))
        return head + (raw(base, 129), u64(base + 136), u32(base + 144), tuple(tracks))
    if t == 6:      # int type, char* mime @8, byte* description @16, u32 width @24, height, depth, colors, data_length @40, data* @48
        mime, desc = ptr(base + 8), ptr(base + 16)
        return head + (u32(base), C.string_at(mime) if mime else None, C.string_at(desc) if desc else None,
                       u32(base + 24), u32(base + 28), u32(base + 32), u32(base + 36), u32(base + 40), raw(ptr(base + 48), u32(base + 40)))
    return head + (raw(ptr(base), v.length),)


def _proto(L):
    L.FLAC__stream_encoder_new.restype = C.c_void_p
    for n in ["delete", "finish"]:
        getattr(L, "FLAC__stream_encoder_" + n).argtypes = [C.c_void_p]
    for n in ["set_verify", "set_channels", "set_bits_per_sample", "set_sample_rate", "set_compression_level", "set_blocksize",
              "set_streamable_subset", "set_limit_min_bitrate"]:
        f = getattr(L, "FLAC__stream_encoder_" + n)
        f.argtypes = [C.c_void_p, C.c_uint32]
        f.restype = C.c_int
    for n in ["get_state", "get_verify", "get_channels", "get_bits_per_sample", "get_sample_rate", "get_blocksize",
              "get_streamable_subset", "get_limit_min_bitrate"]:
        f = getattr(L, "FLAC__stream_encoder_" + n)
        f.argtypes = [C.c_void_p]
        f.restype = C.c_uint32
    L.FLAC__stream_encoder_init_stream.argtypes = [C.c_void_p, WRITE_CB, SEEK_CB, TELL_CB, META_CB, C.c_void_p]
    L.FLAC__stream_encoder_init_stream.restype = C.c_int
    L.FLAC__stream_encoder_init_file.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p]
    L.FLAC__stream_encoder_init_file.restype = C.c_int
    L.FLAC__stream_encoder_process_interleaved.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.FLAC__stream_encoder_process_interleaved.restype = C.c_int
    L.FLAC__stream_encoder_finish.restype = C.c_int
    return L


def encode_session(L, pcm, sample_rate, bps, level=5, blocksize=0, chunks=None, seekable=True, metadata=True,
                   verify=False, streamable_subset=True, init_only=False, no_tell=False, limit_min_bitrate=False, setters=(),
                   fail=None):
    """Run one StreamEncoder session; returns dict(init_status, log=[(event, ...)], file=bytes image, ok=bool).
    log events: ('write', bytes, samples, frame) | ('seek', off) | ('tell', off) | ('meta', dict).
    fail: {'write' | 'seek' | 'tell': k} makes the k-th call (from 0) of that callback report an error."""
    _proto(L)
    x = np.ascontiguousarray(pcm)
    if x.ndim == 1:
        x = x[:, None]
    n, ch = x.shape
    x32 = np.ascontiguousarray(x.astype(np.int32))
    log, image, pos = [], bytearray(), [0]
    calls = {"write": 0, "seek": 0, "tell": 0}

    def failing(kind):
        k = calls[kind]
        calls[kind] += 1
        return bool(fail) and fail.get(kind) == k

    def w(enc, buf, nbytes, samples, frame, cd):
        b = bytes(buf[:nbytes])
        log.append(("write", b, samples, frame))
        if failing("write"):
            return 1
        image[pos[0]:pos[0] + nbytes] = b
        pos[0] += nbytes
        return 0

    def s(enc, off, cd):
        log.append(("seek", off))
        if failing("seek"):
            return 1
        pos[0] = off
        return 0

    def t(enc, poff, cd):
        log.append(("tell", pos[0]))
        if failing("tell"):
            return 1
        poff[0] = pos[0]
        return 0

    def m(enc, md, cd):
        v = C.cast(md, C.POINTER(StreamInfoView)).contents
        log.append(("meta", dict(type=v.type, length=v.length, min_blocksize=v.min_blocksize, max_blocksize=v.max_blocksize,
                                 min_framesize=v.min_framesize, max_framesize=v.max_framesize, sample_rate=v.sample_rate,
                                 channels=v.channels, bits_per_sample=v.bits_per_sample, total_samples=v.total_samples,
                                 md5=bytes(v.md5sum))))

    wcb, scb, tcb, mcb = WRITE_CB(w), SEEK_CB(s), TELL_CB(t), META_CB(m)
    e = L.FLAC__stream_encoder_new()
    L.FLAC__stream_encoder_set_verify(e, int(verify))
    L.FLAC__stream_encoder_set_channels(e, ch)
    L.FLAC__stream_encoder_set_bits_per_sample(e, bps)
    L.FLAC__stream_encoder_set_sample_rate(e, sample_rate)
    L.FLAC__stream_encoder_set_compression_level(e, level)
    L.FLAC__stream_encoder_set_blocksize(e, blocksize)
    L.FLAC__stream_encoder_set_streamable_subset(e, int(streamable_subset))
    L.FLAC__stream_encoder_set_limit_min_bitrate(e, int(limit_min_bitrate))
    for name, v in setters:                                     # fine-grained settings, applied after the compression level like a libFLAC client
        f = getattr(L, "FLAC__stream_encoder_set_" + name)
        f.argtypes = [C.c_void_p, C.c_char_p if name == "apodization" else C.c_uint32]
        f.restype = C.c_int
        f(e, v.encode() if name == "apodization" else v)
    null = lambda T: C.cast(None, T)  # noqa: E731
    st = L.FLAC__stream_encoder_init_stream(e, wcb, scb if seekable else null(SEEK_CB),
                                            null(TELL_CB) if (not seekable or no_tell) else tcb,
                                            mcb if metadata else null(META_CB), None)
    res = dict(init_status=st, log=log, ok=True)
    res["state_after_init_call"] = L.FLAC__stream_encoder_get_state(e)
    if st == 0 and not init_only:
        res["state_after_init"] = L.FLAC__stream_encoder_get_state(e)
        if chunks is None:
            chunks = [n]
        done = 0
        for cnum in chunks:
            cnum = min(cnum, n - done)
            if cnum <= 0:
                break
            seg = x32[done:done + cnum]
            if not L.FLAC__stream_encoder_process_interleaved(e, seg.ctypes.data, cnum):
                res["ok"] = False
                break
            done += cnum
        res["finish"] = L.FLAC__stream_encoder_finish(e)
        res["state_after_finish"] = L.FLAC__stream_encoder_get_state(e)
    L.FLAC__stream_encoder_delete(e)
    res["file"] = bytes(image)
    return res


def decode_session(L, data, read_chunk=8192):
    """Run one StreamDecoder session over a .flac byte string; returns dict(pcm=(n,ch) int32, frames=[headers], errors=[...], ok)."""
    L.FLAC__stream_decoder_new.restype = C.c_void_p
    L.FLAC__stream_decoder_delete.argtypes = [C.c_void_p]
    L.FLAC__stream_decoder_finish.argtypes = [C.c_void_p]
    L.FLAC__stream_decoder_finish.restype = C.c_int
    L.FLAC__stream_decoder_get_state.argtypes = [C.c_void_p]
    L.FLAC__stream_decoder_process_until_end_of_stream.argtypes = [C.c_void_p]
    L.FLAC__stream_decoder_process_until_end_of_stream.restype = C.c_int
    L.FLAC__stream_decoder_init_stream.argtypes = [C.c_void_p, DEC_READ_CB, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   DEC_WRITE_CB, C.c_void_p, DEC_ERROR_CB, C.c_void_p]
    L.FLAC__stream_decoder_init_stream.restype = C.c_int
    pos = [0]
    blocks, frames, errors, events = [], [], [], []       # events: ('w', blocksize) / ('e', status) in callback order

    def r(dec, buf, pbytes, cd):
        want = min(pbytes[0], read_chunk)
        k = min(want, len(data) - pos[0])
        if k == 0:
            pbytes[0] = 0
            return 1
        C.memmove(buf, data[pos[0]:pos[0] + k], k)
        pos[0] += k
        pbytes[0] = k
        return 0

    def w(dec, frame, buffers, cd):
        h = C.cast(frame, C.POINTER(FrameHeaderView)).contents
        frames.append(dict(blocksize=h.blocksize, sample_rate=h.sample_rate, channels=h.channels,
                           bits_per_sample=h.bits_per_sample, channel_assignment=h.channel_assignment))
        blk = np.empty((h.blocksize, h.channels), np.int32)
        for c in range(h.channels):
            blk[:, c] = np.ctypeslib.as_array(buffers[c], shape=(h.blocksize,))
        blocks.append(blk)
        events.append(('w', int(h.blocksize)))
        return 0

    def e(dec, status, cd):
        errors.append(status)
        events.append(('e', int(status)))

    rcb, wcb, ecb = DEC_READ_CB(r), DEC_WRITE_CB(w), DEC_ERROR_CB(e)
    d = L.FLAC__stream_decoder_new()
    st = L.FLAC__stream_decoder_init_stream(d, rcb, None, None, None, None, wcb, None, ecb, None)
    res = dict(init_status=st, frames=frames, errors=errors, events=events)
    if st == 0:
        res["ok"] = bool(L.FLAC__stream_decoder_process_until_end_of_stream(d))
        res["state"] = L.FLAC__stream_decoder_get_state(d)
        L.FLAC__stream_decoder_finish(d)
    L.FLAC__stream_decoder_delete(d)
    res["pcm"] = np.concatenate(blocks) if blocks else np.zeros((0, 1), np.int32)
    return res


DEC_SEEK_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_void_p)
DEC_TELL_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p)
DEC_LENGTH_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p)
DEC_EOF_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)


def scripted_decode_session(L, data, ops, md5_checking=False, seekable=True, path=None, meta=False, read_chunk=None, respond=(),
                            read_script=None):
    """A StreamDecoder session driven by a script, over seekable callbacks (or a file when `path` is given).
    ops: ('seek', sample) | ('single', n) | ('end',) | ('flush',) | ('reset',) | ('meta',).
    Returns dict(events=[...], finish=bool): events are ('w', number_type, sample_number, blocksize, crc32 of the samples),
    ('e', status) and ('ret', op, return value, decoder state) in the order they happened; with meta=True a metadata callback is
    registered and logs ('m', type, is_last, length, sample_rate, channels, bits_per_sample, total_samples).  read_chunk caps what
    one read callback hands over (None: as much as is asked for).  respond: filter calls made before init, in order, e.g.
    ('respond_all',), ('ignore', 4), ('respond_application', b"abcd"); their return values are logged as ('set', name, value).
    read_script: {k: 'abort' | 'empty' | 'eof'} makes the k-th read callback (from 0) abort, hand over nothing with CONTINUE, or
    claim the end of the stream."""
    import zlib
    for n, at, rt in [("new", [], C.c_void_p), ("delete", [C.c_void_p], None), ("finish", [C.c_void_p], C.c_int),
                      ("get_state", [C.c_void_p], C.c_int), ("process_until_end_of_stream", [C.c_void_p], C.c_int),
                      ("process_until_end_of_metadata", [C.c_void_p], C.c_int), ("process_single", [C.c_void_p], C.c_int),
                      ("flush", [C.c_void_p], C.c_int), ("reset", [C.c_void_p], C.c_int),
                      ("seek_absolute", [C.c_void_p, C.c_uint64], C.c_int), ("set_md5_checking", [C.c_void_p, C.c_int], C.c_int),
                      ("get_total_samples", [C.c_void_p], C.c_uint64)]:
        f = getattr(L, "FLAC__stream_decoder_" + n)
        f.argtypes, f.restype = at, rt
    L.FLAC__stream_decoder_init_stream.argtypes = [C.c_void_p, DEC_READ_CB, DEC_SEEK_CB, DEC_TELL_CB, DEC_LENGTH_CB, DEC_EOF_CB,
                                                   DEC_WRITE_CB, C.c_void_p, DEC_ERROR_CB, C.c_void_p]
    L.FLAC__stream_decoder_init_stream.restype = C.c_int
    L.FLAC__stream_decoder_init_file.argtypes = [C.c_void_p, C.c_char_p, DEC_WRITE_CB, C.c_void_p, DEC_ERROR_CB, C.c_void_p]
    L.FLAC__stream_decoder_init_file.restype = C.c_int
    pos = [0]
    events = []

    reads = [0]

    def r(dec, buf, pbytes, cd):
        what = (read_script or {}).get(reads[0])
        reads[0] += 1
        if what == 'abort':
            return 2
        if what == 'empty':
            pbytes[0] = 0
            return 0
        if what == 'eof':
            pbytes[0] = 0
            return 1
        k = min(pbytes[0], len(data) - pos[0], read_chunk or (1 << 62))
        if k == 0:
            pbytes[0] = 0
            return 1
        C.memmove(buf, data[pos[0]:pos[0] + k], k)
        pos[0] += k
        pbytes[0] = k
        return 0

    def mcb(dec, md, cd):
        events.append(metadata_view(md))

    def sk(dec, off, cd):
        pos[0] = min(int(off), len(data))
        return 0

    def tl(dec, poff, cd):
        poff[0] = pos[0]
        return 0

    def ln(dec, plen, cd):
        plen[0] = len(data)
        return 0

    def ef(dec, cd):
        return int(pos[0] >= len(data))

    def w(dec, frame, buffers, cd):
        h = C.cast(frame, C.POINTER(FrameHeaderView)).contents
        crc = 0
        for c in range(h.channels):
            crc = zlib.crc32(np.ctypeslib.as_array(buffers[c], shape=(h.blocksize,)).tobytes(), crc)
        events.append(('w', int(h.number_type), int(h.number), int(h.blocksize), crc))
        return 0

    def e(dec, status, cd):
        events.append(('e', int(status)))

    cbs = (DEC_READ_CB(r), DEC_SEEK_CB(sk), DEC_TELL_CB(tl), DEC_LENGTH_CB(ln), DEC_EOF_CB(ef), DEC_WRITE_CB(w), DEC_ERROR_CB(e))
    mcb_c = META_CB(mcb)
    mptr = C.cast(mcb_c, C.c_void_p) if meta else None
    null = lambda T: C.cast(None, T)  # noqa: E731
    d = L.FLAC__stream_decoder_new()
    L.FLAC__stream_decoder_set_md5_checking(d, int(md5_checking))
    for call in respond:
        f = getattr(L, "FLAC__stream_decoder_set_metadata_" + call[0])
        f.restype = C.c_int
        if len(call) == 1:
            f.argtypes = [C.c_void_p]
            rv = f(d)
        elif isinstance(call[1], bytes):
            f.argtypes = [C.c_void_p, C.c_char_p]
            rv = f(d, call[1])
        else:
            f.argtypes = [C.c_void_p, C.c_int]
            rv = f(d, call[1])
        events.append(('set', call[0], int(rv)))
    if path is not None:
        st = L.FLAC__stream_decoder_init_file(d, path.encode(), cbs[5], mptr, cbs[6], None)
    elif seekable:
        st = L.FLAC__stream_decoder_init_stream(d, cbs[0], cbs[1], cbs[2], cbs[3], cbs[4], cbs[5], mptr, cbs[6], None)
    else:
        st = L.FLAC__stream_decoder_init_stream(d, cbs[0], null(DEC_SEEK_CB), null(DEC_TELL_CB), null(DEC_LENGTH_CB), null(DEC_EOF_CB),
                                                cbs[5], mptr, cbs[6], None)
    res = dict(init_status=st, events=events)
    if st == 0:
        for op in ops:
            if op[0] == 'seek':
                rv = L.FLAC__stream_decoder_seek_absolute(d, op[1])
            elif op[0] == 'single':
                rv = 1
                for _ in range(op[1]):
                    rv = L.FLAC__stream_decoder_process_single(d)
            elif op[0] == 'end':
                rv = L.FLAC__stream_decoder_process_until_end_of_stream(d)
            elif op[0] == 'meta':
                rv = L.FLAC__stream_decoder_process_until_end_of_metadata(d)
            elif op[0] == 'flush':
                rv = L.FLAC__stream_decoder_flush(d)
            elif op[0] == 'reset':
                rv = L.FLAC__stream_decoder_reset(d)
            events.append(('ret', op[0], int(rv), int(L.FLAC__stream_decoder_get_state(d))))
        res["finish"] = bool(L.FLAC__stream_decoder_finish(d))
    L.FLAC__stream_decoder_delete(d)
    return res
