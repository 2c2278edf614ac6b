// flac_api_dec.cu -- drop-in FLAC__stream_decoder_* layer (include/flacb200_flac_api.h) on top of the batch decoder.
//
// libFLAC pulls bytes through the read callback and parses sequentially; here the handle pulls a large slice,
// parses the metadata blocks on the host (a few dozen bytes), hands every complete frame it holds to the GPU in
// ONE flacb200_decode_batch call (headerless "raw" mode with the STREAMINFO parameters), and then fires the
// write callback once per frame, in order, with the same payload pyFLAC reads (decoder.py:482-527:
// frame.header.{blocksize,sample_rate,channels,bits_per_sample} and one int32 plane per channel).  Bytes of a
// frame that is still incomplete stay buffered until more input arrives.  Errors surface through the error
// callback with libFLAC's status values (LOST_SYNC, BAD_HEADER, FRAME_CRC_MISMATCH, UNPARSEABLE_STREAM).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <algorithm>
#include <deque>
#include <mutex>
#include <vector>

#include "../../include/flacb200.h"
#include "../../include/flacb200_flac_api.h"
#include "md5_host.h"

extern "C" {
// values read from the reference binary (pyflac/decoder.py:46,60,546)
const char *const FLAC__StreamDecoderStateString[] = {
    "FLAC__STREAM_DECODER_SEARCH_FOR_METADATA", "FLAC__STREAM_DECODER_READ_METADATA", "FLAC__STREAM_DECODER_SEARCH_FOR_FRAME_SYNC",
    "FLAC__STREAM_DECODER_READ_FRAME", "FLAC__STREAM_DECODER_END_OF_STREAM", "FLAC__STREAM_DECODER_OGG_ERROR", "FLAC__STREAM_DECODER_SEEK_ERROR",
    "FLAC__STREAM_DECODER_ABORTED", "FLAC__STREAM_DECODER_MEMORY_ALLOCATION_ERROR", "FLAC__STREAM_DECODER_UNINITIALIZED"};
const char *const FLAC__StreamDecoderInitStatusString[] = {
    "FLAC__STREAM_DECODER_INIT_STATUS_OK", "FLAC__STREAM_DECODER_INIT_STATUS_UNSUPPORTED_CONTAINER", "FLAC__STREAM_DECODER_INIT_STATUS_INVALID_CALLBACKS",
    "FLAC__STREAM_DECODER_INIT_STATUS_MEMORY_ALLOCATION_ERROR", "FLAC__STREAM_DECODER_INIT_STATUS_ERROR_OPENING_FILE",
    "FLAC__STREAM_DECODER_INIT_STATUS_ALREADY_INITIALIZED"};
const char *const FLAC__StreamDecoderErrorStatusString[] = {
    "FLAC__STREAM_DECODER_ERROR_STATUS_LOST_SYNC", "FLAC__STREAM_DECODER_ERROR_STATUS_BAD_HEADER", "FLAC__STREAM_DECODER_ERROR_STATUS_FRAME_CRC_MISMATCH",
    "FLAC__STREAM_DECODER_ERROR_STATUS_UNPARSEABLE_STREAM", "FLAC__STREAM_DECODER_ERROR_STATUS_BAD_METADATA"};
}

namespace {

enum { DS_SEARCH_FOR_METADATA = 0, DS_READ_METADATA = 1, DS_SEARCH_FOR_FRAME_SYNC = 2, DS_READ_FRAME = 3, DS_END_OF_STREAM = 4,
       DS_SEEK_ERROR = 6, DS_ABORTED = 7, DS_MEMORY_ALLOCATION_ERROR = 8, DS_UNINITIALIZED = 9 };
enum { DI_OK = 0, DI_UNSUPPORTED_CONTAINER = 1, DI_INVALID_CALLBACKS = 2, DI_MEMORY_ALLOCATION_ERROR = 3, DI_ERROR_OPENING_FILE = 4, DI_ALREADY_INITIALIZED = 5 };
enum { ERR_LOST_SYNC = 0, ERR_BAD_HEADER = 1, ERR_FRAME_CRC_MISMATCH = 2, ERR_UNPARSEABLE_STREAM = 3, ERR_BAD_METADATA = 4 };

// one engine context per CUDA device for the decoder handles; the device comes from FLACB200_DEVICE (default 0).
// The mutex covers GPU work only: callbacks are fired after it is released.
std::mutex g_dec_mu;
flacb200_ctx* g_dec_ctx[64] = {nullptr};
int g_dec_rc[64];
bool g_dec_init = false;
flacb200_ctx* dec_ctx() {
    int dev = 0;
    if (const char* l = getenv("FLACB200_DEVICE")) dev = atoi(l);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!g_dec_init) { for (int& r : g_dec_rc) r = -1; g_dec_init = true; }
    if (g_dec_rc[dev] == -1) g_dec_rc[dev] = flacb200_create(&g_dec_ctx[dev], dev);
    return g_dec_rc[dev] == 0 ? g_dec_ctx[dev] : nullptr;
}

// what the next callback delivers: a decoded frame, a frame of silence standing in for a missing one, or an error status
struct PendingFrame { uint32_t blocksize; std::vector<int32_t> planar; int error = -1; uint64_t first_sample = 0; };   // planar = [channel][sample]; error >= 0: FLAC__StreamDecoderErrorStatus

struct DecImpl {
    int state = DS_UNINITIALIZED;
    FLAC__bool md5_checking = 0;
    FLAC__StreamDecoderReadCallback read_cb = nullptr; FLAC__StreamDecoderWriteCallback write_cb = nullptr;
    FLAC__StreamDecoderErrorCallback error_cb = nullptr; FLAC__StreamDecoderMetadataCallback meta_cb = nullptr;
    FLAC__StreamDecoderSeekCallback seek_cb = nullptr; FLAC__StreamDecoderTellCallback tell_cb = nullptr;
    FLAC__StreamDecoderLengthCallback length_cb = nullptr; FLAC__StreamDecoderEofCallback eof_cb = nullptr;
    void* client = nullptr;
    // Metadata is read the way libFLAC reads it: the stream marker first (find_magic), then one block per process_single -- a block
    // is read (and its callback fires) in the call that finds all of it buffered; `in` keeps the metadata bytes until the last block
    // has been read.  meta_pos: offset in `in` of the next block header.  meta_skip: a block could not be used (BAD_METADATA): the
    // blocks behind it are walked without a word on the way to the first frame.
    struct MetaBlock { uint32_t type, len; bool last; size_t off; };     // off: first data byte in `in`
    size_t meta_pos = 0; bool have_si = false, meta_skip = false;
    void meta_reset() { meta_pos = 0; have_si = false; meta_skip = false; find_reset(); }
    // the search for "fLaC" in front of the blocks (find_magic): resumable, because the errors it reports must be reported once
    size_t find_scan = 0; int find_i = 0, find_id = 0, id3_need = 0; bool find_first = true; uint32_t id3_size = 0; uint64_t find_skip = 0;
    void find_reset() { find_scan = 0; find_i = find_id = id3_need = 0; find_first = true; id3_size = 0; find_skip = 0; }
    bool is_seeking = false;              // metadata read on behalf of a seek is not reported (stream_decoder.c: is_seeking)
    // FLAC__stream_decoder_set_metadata_respond* / _ignore*: one switch per block type (default: STREAMINFO only) and, for
    // APPLICATION blocks, the ids that are exceptions to their type's switch
    bool meta_filter[128]; std::vector<uint32_t> meta_ids;
    void filter_defaults() { for (bool& f : meta_filter) f = false; meta_filter[0] = true; meta_ids.clear(); }
    DecImpl() { filter_defaults(); }
    uint64_t first_frame_offset = 0;      // byte offset of the first audio frame (behind the metadata blocks)
    // MD5 of the delivered samples against STREAMINFO's (stream_decoder.h: set_md5_checking; off once a seek or flush happened)
    bool md5_active = false; fb::Md5 md5; uint8_t stored_md5[16] = {0}; std::vector<int32_t> md5_tmp;
    FILE* file = nullptr;
    std::vector<uint8_t> in;              // bytes not yet consumed
    bool eof = false, metadata_done = false;
    uint32_t sample_rate = 0, channels = 0, bps = 0, blocksize = 0;
    uint64_t total_samples = 0, frame_index = 0, bytes_consumed = 0;
    uint32_t min_blocksize = 0;
    bool have_last = false; uint64_t next_sample = 0; uint32_t last_blocksize = 0;     // stream position behind the last delivered frame
    std::deque<PendingFrame> ready;       // decoded frames not yet delivered
    FLAC__Frame frame;                    // callback payload (header filled per frame)
    // what FLAC__stream_decoder_get_channels / _bits_per_sample / _sample_rate / _blocksize report: the header of the last frame
    // that was delivered, as libFLAC (0 before the first frame, whatever STREAMINFO said)
    uint32_t hdr_channels = 0, hdr_bps = 0, hdr_sample_rate = 0, hdr_blocksize = 0;
};
struct DHandle { FLAC__StreamDecoder pub; DecImpl impl; };
inline DecImpl* D(const FLAC__StreamDecoder* d) { return d ? (DecImpl*)d->private_ : nullptr; }

void report(FLAC__StreamDecoder* d, int status) { DecImpl* m = D(d); if (m->error_cb && !m->is_seeking) m->error_cb(d, status, m->client); }   // (nothing met on behalf of a seek is reported)

// pull one slice of input; returns false on abort
bool pull(FLAC__StreamDecoder* d, size_t want) {
    DecImpl* m = D(d);
    if (m->eof) return true;
    const size_t old = m->in.size();
    m->in.resize(old + want);
    size_t got = want; int st;
    if (m->file) { got = fread(m->in.data() + old, 1, want, m->file); st = got == 0 ? 1 : 0; }
    else st = m->read_cb(d, m->in.data() + old, &got, m->client);
    if (st == 2) { m->in.resize(old); m->state = DS_ABORTED; return false; }
    if (st == 1) { m->eof = true; if (got > want) got = 0; }
    m->in.resize(old + (st == 1 && !m->file ? got : got));
    // (a read callback that hands over nothing and says CONTINUE is simply asked again, as libFLAC does)
    return true;
}

// STREAMINFO (34 bytes at q) into the decoder's fields and, if it is asked for, the metadata callback
void take_streaminfo(FLAC__StreamDecoder* d, const uint8_t* q, uint32_t len, bool last) {
    DecImpl* m = D(d);
    m->blocksize = (uint32_t)q[2] << 8 | q[3]; m->min_blocksize = (uint32_t)q[0] << 8 | q[1];
    m->sample_rate = (uint32_t)q[10] << 12 | (uint32_t)q[11] << 4 | (q[12] >> 4);
    m->channels = ((q[12] >> 1) & 7) + 1;
    m->bps = (((uint32_t)q[12] & 1) << 4 | (q[13] >> 4)) + 1;
    m->total_samples = ((uint64_t)(q[13] & 0xF) << 32) | (uint64_t)q[14] << 24 | (uint64_t)q[15] << 16 | (uint64_t)q[16] << 8 | q[17];
    memcpy(m->stored_md5, q + 18, 16);
    { bool any = false; for (int i = 0; i < 16; i++) any |= q[18 + i] != 0; if (!any) m->md5_active = false; }   // an unset MD5 is not checked
    if (m->meta_cb && m->meta_filter[0] && !m->is_seeking) {
        FLAC__StreamMetadata md; memset(&md, 0, sizeof md);
        md.type = 0; md.is_last = last; md.length = len;
        md.data.stream_info.min_blocksize = (uint32_t)q[0] << 8 | q[1]; md.data.stream_info.max_blocksize = m->blocksize;
        md.data.stream_info.min_framesize = (uint32_t)q[4] << 16 | (uint32_t)q[5] << 8 | q[6];
        md.data.stream_info.max_framesize = (uint32_t)q[7] << 16 | (uint32_t)q[8] << 8 | q[9];
        md.data.stream_info.sample_rate = m->sample_rate; md.data.stream_info.channels = m->channels;
        md.data.stream_info.bits_per_sample = m->bps; md.data.stream_info.total_samples = m->total_samples;
        memcpy(md.data.stream_info.md5sum, q + 18, 16);
        m->meta_cb(d, &md, m->client);
    }
}

inline uint32_t be32(const uint8_t* p) { return (uint32_t)p[0] << 24 | (uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]; }
inline uint64_t be64(const uint8_t* p) { return (uint64_t)be32(p) << 32 | be32(p + 4); }
inline uint32_t le32(const uint8_t* p) { return (uint32_t)p[3] << 24 | (uint32_t)p[2] << 16 | (uint32_t)p[1] << 8 | p[0]; }

// One metadata block to the metadata callback if the filter asks for it (up: read_metadata_, stream_decoder.c; layouts:
// builder/decoder.py:256-365).  Everything the block points to lives in the locals below for the duration of the callback.
// Returns 0 (reported, or not asked for) or what libFLAC 1.4.3 was seen to do with a block it cannot use (pinned on the binary by
// tools/host_logic_check.py): kMetaBad -- content that does not fit the block's length: BAD_METADATA is reported, metadata reading
// ends and the frame search starts inside the block; kMetaRefused -- a VORBIS_COMMENT that claims more than 100 000 comments: the
// call fails, the next one reads the next block; kMetaFatal -- an APPLICATION block shorter than its id: MEMORY_ALLOCATION_ERROR;
// kMetaStuck -- a SEEKTABLE whose length is not a whole number of points: this and every later call fail.
enum { kMetaOk = 0, kMetaBad = 1, kMetaRefused = 2, kMetaFatal = 3, kMetaStuck = 4 };
int deliver_meta_block(FLAC__StreamDecoder* d, const DecImpl::MetaBlock& b) {
    DecImpl* m = D(d);
    const uint8_t* q = m->in.data() + b.off;
    if (b.type == 0) { if (b.len >= 34) take_streaminfo(d, q, b.len, b.last); return kMetaOk; }
    bool want = b.type < 128 && m->meta_filter[b.type];
    if (b.type == 2) {                                                  // APPLICATION: the listed ids are the exceptions
        if (b.len < 4) return kMetaFatal;
        const uint32_t id = be32(q);
        if (!m->meta_ids.empty()) for (uint32_t v : m->meta_ids) if (v == id) { want = !want; break; }
    }
    if (b.type == 3 && b.len % 18 != 0) return kMetaStuck;               // (the seek table is read whether or not it is asked for)
    if (!want) return kMetaOk;                                          // blocks nobody asked for are skipped unread
    if (b.type == 3 && b.len == 0) return kMetaOk;                      // an empty SEEKTABLE is not reported
    FLAC__StreamMetadata md; memset(&md, 0, sizeof md);
    md.type = (int)b.type; md.is_last = b.last; md.length = b.len;
    const uint8_t* e = q + b.len;
    std::vector<uint8_t> bytes;                                         // NUL-terminated copies of the block's strings, packed
    std::vector<size_t> str_off;                                        // ... and where each starts (pointers are taken at the end)
    auto keep = [&](const uint8_t* p, size_t n) { str_off.push_back(bytes.size()); bytes.insert(bytes.end(), p, p + n); bytes.push_back(0); };
    std::vector<FLAC__StreamMetadata_SeekPoint> points;
    std::vector<FLAC__StreamMetadata_VorbisComment_Entry> comments;
    std::vector<FLAC__StreamMetadata_CueSheet_Track> tracks;
    std::vector<std::vector<FLAC__StreamMetadata_CueSheet_Index>> indices;
    switch (b.type) {
    case 1: break;                                                      // PADDING: nothing but its length
    case 2:
        memcpy(md.data.application.id, q, 4);
        md.data.application.data = b.len > 4 ? const_cast<FLAC__byte*>(q + 4) : nullptr;
        break;
    case 3: {
        const uint32_t n = b.len / 18;
        points.resize(n);
        for (uint32_t i = 0; i < n; i++) { const uint8_t* p = q + 18 * (size_t)i; points[i].sample_number = be64(p); points[i].stream_offset = be64(p + 8); points[i].frame_samples = (uint32_t)p[16] << 8 | p[17]; }
        md.data.seek_table.num_points = n; md.data.seek_table.points = n ? points.data() : nullptr;
        break;
    }
    case 4: {
        const uint8_t* p = q;
        if (e - p < 8) return kMetaBad;
        const uint32_t vl = le32(p); p += 4;
        if ((size_t)(e - p) < (size_t)vl + 4) return kMetaBad;
        keep(p, vl); p += vl;
        uint32_t nc = le32(p); p += 4;
        const uint32_t nc_said = nc;
        if (nc > 100000) return kMetaRefused;
        for (uint32_t i = 0; i < nc_said; i++) {
            if (e - p < 4) { if (p != e) return kMetaBad; nc = i; break; }   // the block ends behind an entry, earlier than announced: the entries that are there count
            const uint32_t cl = le32(p); p += 4;
            if ((size_t)(e - p) < cl) return kMetaBad;
            FLAC__StreamMetadata_VorbisComment_Entry c; c.length = cl; c.entry = nullptr; comments.push_back(c);
            keep(p, cl); p += cl;
        }
        if (p != e) return kMetaBad;
        md.data.vorbis_comment.vendor_string.length = vl; md.data.vorbis_comment.vendor_string.entry = bytes.data() + str_off[0];
        for (uint32_t i = 0; i < nc; i++) comments[i].entry = bytes.data() + str_off[1 + i];
        md.data.vorbis_comment.num_comments = nc; md.data.vorbis_comment.comments = nc ? comments.data() : nullptr;
        break;
    }
    case 5: {
        const uint8_t* p = q;
        if (e - p < 128 + 8 + 259 + 1) return kMetaBad;
        memcpy(md.data.cue_sheet.media_catalog_number, p, 128); md.data.cue_sheet.media_catalog_number[128] = 0; p += 128;
        md.data.cue_sheet.lead_in = be64(p); p += 8;
        md.data.cue_sheet.is_cd = p[0] >> 7; p += 259;
        const uint32_t nt = *p++;
        tracks.resize(nt); indices.resize(nt);
        for (uint32_t t = 0; t < nt; t++) {
            if (e - p < 8 + 1 + 12 + 14 + 1) return kMetaBad;
            FLAC__StreamMetadata_CueSheet_Track& T = tracks[t]; memset(&T, 0, sizeof T);
            T.offset = be64(p); p += 8; T.number = *p++;
            memcpy(T.isrc, p, 12); T.isrc[12] = 0; p += 12;
            T.type = p[0] >> 7; T.pre_emphasis = (p[0] >> 6) & 1; p += 14;
            T.num_indices = *p++;
            if ((size_t)(e - p) < (size_t)T.num_indices * 12) return kMetaBad;
            indices[t].resize(T.num_indices);
            for (uint32_t i = 0; i < T.num_indices; i++) { indices[t][i].offset = be64(p); indices[t][i].number = p[8]; p += 12; }
            T.indices = T.num_indices ? indices[t].data() : nullptr;
        }
        if (p != e) return kMetaBad;
        md.data.cue_sheet.num_tracks = nt; md.data.cue_sheet.tracks = nt ? tracks.data() : nullptr;
        break;
    }
    case 6: {
        const uint8_t* p = q;
        if (e - p < 8) return kMetaBad;
        { const uint32_t t = be32(p); md.data.picture.type = t <= 20 ? (int)t : 0; } p += 4;   // (a type beyond the defined ones is handed over as OTHER)
        const uint32_t ml = be32(p); p += 4;
        if ((size_t)(e - p) < (size_t)ml + 4) return kMetaBad;
        keep(p, ml); p += ml;
        const uint32_t dl = be32(p); p += 4;
        if ((size_t)(e - p) < (size_t)dl + 20) return kMetaBad;
        keep(p, dl); p += dl;
        md.data.picture.width = be32(p); md.data.picture.height = be32(p + 4); md.data.picture.depth = be32(p + 8); md.data.picture.colors = be32(p + 12);
        md.data.picture.data_length = be32(p + 16); p += 20;
        if ((size_t)(e - p) != md.data.picture.data_length) return kMetaBad;
        md.data.picture.data = md.data.picture.data_length ? const_cast<FLAC__byte*>(p) : nullptr;
        md.data.picture.mime_type = reinterpret_cast<char*>(bytes.data() + str_off[0]); md.data.picture.description = bytes.data() + str_off[1];
        break;
    }
    default:
        md.data.unknown.data = b.len ? const_cast<FLAC__byte*>(q) : nullptr;
        break;
    }
    if (m->meta_cb && !m->is_seeking) m->meta_cb(d, &md, m->client);
    return kMetaOk;
}

// The search for "fLaC" at the head of the input (up: find_metadata_, stream_decoder.c; pinned on the binary by
// tools/host_logic_check.py): an ID3v2 tag in front of it is skipped without a word; any other byte that is not part of the marker
// is reported as LOST_SYNC -- once per run of such bytes, a run ending wherever a byte continues the marker; bytes other than the
// marker behind an ID3v2 tag make the call fail once and the next call search afresh.  returns 1 found (find_scan stands behind
// the marker), 0 need more input, -1 a frame sync code came first (a stream without metadata: not decodable by this build),
// -2 the input ended, -3 this call fails and the next one searches on.
int find_magic(FLAC__StreamDecoder* d) {
    DecImpl* m = D(d);
    const std::vector<uint8_t>& in = m->in;
    static const uint8_t kMagic[4] = {'f', 'L', 'a', 'C'}, kId3[3] = {'I', 'D', '3'};
    while (m->find_i < 4) {
        if (m->find_skip) {                                             // inside an ID3v2 tag
            const uint64_t k = std::min<uint64_t>(m->find_skip, in.size() - m->find_scan);
            m->find_scan += (size_t)k; m->find_skip -= k;
            if (m->find_skip) return m->eof ? -2 : 0;
            continue;
        }
        if (m->find_scan >= in.size()) return m->eof ? -2 : 0;
        const uint8_t x = in[m->find_scan];
        if (m->id3_need) {                                              // the rest of an ID3v2 header: version (2), flags (1), size (4 x 7 bits)
            m->find_scan++;
            if (m->id3_need <= 4) m->id3_size = (m->id3_size << 7) | (x & 0x7fu);
            if (--m->id3_need == 0) m->find_skip = m->id3_size;
            continue;
        }
        const bool marker = x == kMagic[m->find_i];
        const bool id3 = !marker && m->find_id < 3 && x == kId3[m->find_id];
        if (!marker && !id3 && m->find_id < 3 && x == 0xff && m->find_scan + 1 >= in.size()) return m->eof ? -2 : 0;   // needs the byte behind it
        m->find_scan++;
        if (marker) { m->find_first = true; m->find_i++; m->find_id = 0; continue; }
        if (m->find_id >= 3) { m->find_i = m->find_id = 0; m->find_first = true; return -3; }
        if (id3) { m->find_i = 0; if (++m->find_id == 3) { m->id3_need = 7; m->id3_size = 0; } continue; }
        m->find_id = 0;
        if (x == 0xff) {
            const uint8_t y = in[m->find_scan];
            if (y != 0xff) {                                            // (a second 0xff is looked at again: it may start the sync code)
                m->find_scan++;
                if ((y >> 1) == 0x7c) return -1;
            }
        }
        m->find_i = 0;
        if (m->find_first) { report(d, ERR_LOST_SYNC); m->find_first = false; }
    }
    return 1;
}

// the metadata is behind us: the input buffer starts at the first frame
void finish_metadata(DecImpl* m) {
    m->in.erase(m->in.begin(), m->in.begin() + (long)m->meta_pos);
    m->bytes_consumed += m->meta_pos; m->meta_pos = 0;
    m->first_frame_offset = m->bytes_consumed;
    m->metadata_done = true;
}
// the input ended before the metadata did: everything was read
int end_inside_metadata(DecImpl* m) {
    m->bytes_consumed += m->in.size(); m->in.clear(); m->meta_pos = 0; m->find_scan = 0;
    m->state = DS_END_OF_STREAM;
    return -1;
}

// One step through the head of the stream: the marker, then ONE metadata block (up: find_metadata_ / read_metadata_,
// stream_decoder.c).  returns 1 a block was read (or, in the until-end calls, the metadata is done), 0 more input was pulled,
// -1 the call fails (the state says why), 2 go on with the frames.
int metadata_step(FLAC__StreamDecoder* d, bool until_end, size_t slice) {
    DecImpl* m = D(d);
    if (m->find_i < 4) {                                                // still looking for the marker
        const int r = find_magic(d);
        if (r == 0 || r == -3) {
            // a long stretch without the marker (or a large ID3v2 tag) is not kept: the input buffer holds what is still needed
            if (m->find_scan > (1u << 20) && m->find_i == 0 && m->id3_need == 0) {
                m->in.erase(m->in.begin(), m->in.begin() + (long)m->find_scan);
                m->bytes_consumed += m->find_scan; m->find_scan = 0;
            }
            if (r == -3) return -1;                                     // this call fails, the state stands, the next call searches on
            return pull(d, slice) ? 0 : -1;
        }
        if (r == -2) return end_inside_metadata(m);
        if (r == -1) {
            // libFLAC goes on to decode such frames without STREAMINFO; this build cannot (pyFLAC's tests expect an error for
            // arbitrary data, tests/test_decoder.py:59-66)
            if (m->find_first) report(d, ERR_LOST_SYNC);
            m->state = m->eof ? DS_END_OF_STREAM : DS_ABORTED;
            return -1;
        }
        m->meta_pos = m->find_scan;
    }
    for (;;) {
        // the next block, once all of it is buffered
        const size_t avail = m->in.size() - m->meta_pos;
        const uint8_t* p = m->in.data() + m->meta_pos;
        const bool have_hdr = avail >= 4;
        const uint32_t type = have_hdr ? (uint32_t)(p[0] & 0x7f) : 0u, len = have_hdr ? ((uint32_t)p[1] << 16 | (uint32_t)p[2] << 8 | p[3]) : 0u;
        if (!have_hdr || avail < 4 + (size_t)len) {
            const bool ended = m->eof;
            if (!ended && pull(d, slice)) return 0;
            // the input ended, or the read callback aborted, inside the metadata: a block that was being read for the client (neither
            // skipped nor one of the two libFLAC keeps for itself) is reported as BAD_METADATA on the way out
            if (m->meta_skip) { if (ended) report(d, ERR_LOST_SYNC); }
            else if (have_hdr && type != 0 && type != 3 && m->meta_filter[type]) report(d, ERR_BAD_METADATA);
            return ended ? end_inside_metadata(m) : -1;
        }
        const DecImpl::MetaBlock b = {type, len, (p[0] >> 7) != 0, m->meta_pos + 4};
        m->meta_pos += 4 + (size_t)len;
        if (m->meta_skip) {                                             // behind a block that could not be used: on to the first frame
            if (!b.last) continue;
            finish_metadata(m);
            PendingFrame pf; pf.blocksize = 0; pf.error = ERR_LOST_SYNC; m->ready.push_back(std::move(pf));   // what lay between that block and the first frame
            return 2;
        }
        m->have_si |= type == 0 && len >= 34;
        m->state = b.last ? DS_SEARCH_FOR_FRAME_SYNC : DS_READ_METADATA;
        const int rc = deliver_meta_block(d, b);                        // (reads the block where it lies in `in`)
        if (rc == kMetaStuck) { m->meta_pos -= 4 + (size_t)len; m->state = DS_READ_METADATA; return -1; }
        if (rc == kMetaFatal) { m->state = DS_MEMORY_ALLOCATION_ERROR; return -1; }
        if (rc == kMetaBad) {
            report(d, ERR_BAD_METADATA);
            m->state = DS_SEARCH_FOR_FRAME_SYNC;
            if (b.last) { finish_metadata(m); PendingFrame pf; pf.blocksize = 0; pf.error = ERR_LOST_SYNC; m->ready.push_back(std::move(pf)); }
            else m->meta_skip = true;
            return -1;
        }
        if (b.last) {
            if (!m->have_si) { report(d, ERR_BAD_METADATA); m->state = m->eof ? DS_END_OF_STREAM : DS_ABORTED; return -1; }   // nothing this build can decode without STREAMINFO
            finish_metadata(m);
        }
        if (rc == kMetaRefused) return -1;
        if (!until_end || b.last) return 1;
    }
}

// decode every complete frame currently buffered (one GPU batch); returns false on fatal error.  What libFLAC 1.4.3
// does around damaged input is reproduced from the engine's per-stream event log (dec_kernels.cu:dec_chain_kernel): error
// callbacks in their place between the frames, a frame with a bad CRC never delivered, missing frames delivered as silence.
bool decode_buffered(FLAC__StreamDecoder* d) {
    DecImpl* m = D(d);
    if (m->in.empty()) return true;
    std::vector<int32_t> pcm; std::vector<uint32_t> fs; std::vector<uint64_t> fo;
    flacb200_dec_stream_info si;
    {
        std::lock_guard<std::mutex> lk(g_dec_mu);                      // GPU work only; callbacks fire after the lock is gone
        flacb200_ctx* ctx = dec_ctx();
        if (!ctx) { m->state = DS_MEMORY_ALLOCATION_ERROR; return false; }
        flacb200_dec_raw_params raw; memset(&raw, 0, sizeof raw);
        raw.sample_rate = m->sample_rate; raw.channels = m->channels; raw.bits_per_sample = m->bps;
        raw.flags = (m->eof ? 0u : 1u) | (m->have_last ? 2u : 0u);
        raw.next_sample = m->next_sample; raw.last_blocksize = m->last_blocksize;
        raw.fixed_blocksize = (m->min_blocksize == m->blocksize) ? m->blocksize : 0u;
        const uint64_t off = 0, len = m->in.size();
        if (flacb200_decode_batch(ctx, m->in.data(), 0, len, 1, &off, &len, 4, &raw) != 0) { report(d, ERR_UNPARSEABLE_STREAM); m->state = DS_ABORTED; return false; }
        flacb200_dec_result r;
        if (flacb200_decode_result(ctx, &r) != 0) { m->state = DS_ABORTED; return false; }
        pcm.resize((size_t)r.total_elems + 1); fs.resize(r.n_frames + 1); fo.resize(r.n_frames + 1);
        if (flacb200_decode_fetch(ctx, pcm.data(), pcm.size() * 4, &si, fs.data(), r.n_frames) != 0 ||
            flacb200_decode_fetch_frame_offsets(ctx, fo.data(), r.n_frames) != 0) { m->state = DS_ABORTED; return false; }
    }
    const uint32_t ch = si.channels ? si.channels : m->channels;
    auto queue_error = [&](int status) { PendingFrame pf; pf.blocksize = 0; pf.error = status; m->ready.push_back(std::move(pf)); };
    auto queue_events = [&](uint32_t before_frame) {
        for (uint32_t k = 0; k < si.n_events && k < 16; k++) if (si.ev_frame[k] == before_frame) queue_error(si.ev_status[k]);
    };
    uint64_t at = 0;                                                   // sample position within this batch's PCM
    uint32_t silence_bs = m->last_blocksize;
    // stream position of this batch's first PCM sample, from the position behind its last frame (frame headers)
    const uint64_t extent = si.n_frames ? fo[si.n_frames - 1] + fs[si.n_frames - 1] : 0;
    const uint64_t base = si.next_sample >= extent ? si.next_sample - extent : 0;
    for (uint32_t f = 0; f < si.n_frames; f++) {
        queue_events(f);
        // frames missing in front of this one stand as silence, in units of the previous frame's blocksize
        for (uint64_t gap = fo[f] - at; gap > 0;) {
            const uint32_t n = (uint32_t)((silence_bs && gap > silence_bs) ? silence_bs : gap);
            PendingFrame pf; pf.blocksize = n; pf.planar.assign((size_t)n * ch, 0); pf.first_sample = base + at;
            m->ready.push_back(std::move(pf));
            gap -= n; at += n;
        }
        PendingFrame pf; pf.blocksize = fs[f]; pf.planar.resize((size_t)fs[f] * ch); pf.first_sample = base + fo[f];
        const int32_t* src = pcm.data() + (size_t)fo[f] * ch;
        for (uint32_t i = 0; i < fs[f]; i++) for (uint32_t c = 0; c < ch; c++) pf.planar[(size_t)c * fs[f] + i] = src[(size_t)i * ch + c];
        m->ready.push_back(std::move(pf));
        at = fo[f] + fs[f]; silence_bs = fs[f];
    }
    queue_events(si.n_frames);
    for (uint32_t k = 16; k < si.n_events; k++) queue_error(ERR_LOST_SYNC);       // beyond the log: reported, kind unknown
    if (si.n_frames) { m->channels = ch; if (si.sample_rate) m->sample_rate = si.sample_rate; if (si.bits_per_sample) m->bps = si.bits_per_sample; }
    m->have_last = si.have_last != 0; m->next_sample = si.next_sample; m->last_blocksize = si.last_blocksize;
    if (si.status == 8) { queue_error(ERR_UNPARSEABLE_STREAM); }                  // a stream this build cannot represent
    size_t drop = (size_t)si.consumed;
    if (m->eof || si.status == 8) drop = m->in.size();
    else if (si.n_frames == 0 && m->in.size() > (1u << 20)) {
        // a megabyte without a single frame: report the lost sync and keep only what a sync code could straddle
        queue_error(ERR_LOST_SYNC);
        drop = m->in.size() - 16;
    }
    m->in.erase(m->in.begin(), m->in.begin() + (long)drop);
    m->bytes_consumed += drop;
    return true;
}

// deliver one queued frame through the write callback; returns false if the client aborted
bool deliver_one(FLAC__StreamDecoder* d) {
    DecImpl* m = D(d);
    PendingFrame& pf = m->ready.front();
    if (pf.error >= 0) {                                                // an error event in its place between the frames
        const int status = pf.error;
        m->ready.pop_front();
        report(d, status);
        return m->state != DS_ABORTED;
    }
    const int32_t* planes[8] = {nullptr};
    for (uint32_t c = 0; c < m->channels && c < 8; c++) planes[c] = pf.planar.data() + (size_t)c * pf.blocksize;
    memset(&m->frame.header, 0, sizeof m->frame.header);
    m->frame.header.blocksize = pf.blocksize; m->frame.header.sample_rate = m->sample_rate; m->frame.header.channels = m->channels;
    // libFLAC hands every frame over with its sample number, whatever the header carried (stream_decoder.c read_frame_header_)
    m->frame.header.channel_assignment = 0; m->frame.header.bits_per_sample = m->bps; m->frame.header.number_type = 1;
    m->frame.header.number.sample_number = pf.first_sample;
    m->hdr_channels = m->channels; m->hdr_bps = m->bps; m->hdr_sample_rate = m->sample_rate; m->hdr_blocksize = pf.blocksize;
    if (m->md5_active) {
        m->md5_tmp.resize((size_t)pf.blocksize * m->channels);
        for (uint32_t c = 0; c < m->channels; c++) for (uint32_t i = 0; i < pf.blocksize; i++) m->md5_tmp[(size_t)i * m->channels + c] = planes[c][i];
        m->md5.update_samples(m->md5_tmp.data(), m->md5_tmp.size(), 4, (m->bps + 7) / 8);
    }
    m->state = DS_READ_FRAME;
    const int st = m->write_cb(d, &m->frame, planes, m->client);
    m->frame_index++;
    m->ready.pop_front();
    if (st != 0) { m->state = DS_ABORTED; return false; }
    m->state = DS_SEARCH_FOR_FRAME_SYNC;
    return true;
}

// make progress: returns 1 if something was delivered/parsed, 0 at end of stream, -1 on abort/fatal
int step(FLAC__StreamDecoder* d, bool until_end) {
    DecImpl* m = D(d);
    const size_t kSlice = until_end ? (1u << 20) : (1u << 16);
    for (;;) {
        if (m->state == DS_ABORTED || m->state == DS_END_OF_STREAM) return m->state == DS_ABORTED ? -1 : 0;
        if (!m->ready.empty()) {
            const bool was_error = m->ready.front().error >= 0;
            if (!deliver_one(d)) return -1;
            if (was_error && !until_end) continue;                       // an error callback does not end a process_single: it goes on to the next frame
            return 1;
        }
        if (!m->metadata_done) {
            const int r = metadata_step(d, until_end, kSlice);
            if (r == 0 || r == 2) continue;
            if (r == 1 && until_end) continue;                           // (the until-end calls go straight on with the frames)
            return r;
        }
        // need more frames: decode what is buffered once a slice (or the tail) is available
        if (!m->eof && m->in.size() < 16) { if (!pull(d, kSlice)) return -1; continue; }
        const size_t before = m->in.size();
        if (!decode_buffered(d)) return -1;
        if (!m->ready.empty()) continue;
        if (m->eof && (m->in.empty() || m->in.size() == before)) { m->state = DS_END_OF_STREAM; return 0; }
        if (!m->eof && m->in.size() == before) { if (!pull(d, kSlice)) return -1; }   // frame larger than what is buffered
    }
}

}  // namespace

extern "C" {

FLAC__StreamDecoder* FLAC__stream_decoder_new(void) {
    DHandle* h = new DHandle();
    h->pub.protected_ = nullptr; h->pub.private_ = &h->impl;
    return &h->pub;
}
void FLAC__stream_decoder_delete(FLAC__StreamDecoder* d) {
    if (!d) return;
    FLAC__stream_decoder_finish(d);
    delete reinterpret_cast<DHandle*>(d);
}
FLAC__bool FLAC__stream_decoder_set_md5_checking(FLAC__StreamDecoder* d, FLAC__bool v) { DecImpl* m = D(d); if (m->state != DS_UNINITIALIZED) return 0; m->md5_checking = v; return 1; }
// up: FLAC__stream_decoder_set_metadata_respond* / _ignore* (stream_decoder.h:847-948; builder/decoder.py:392-397): settable before
// init only; naming the APPLICATION type as a whole forgets the listed ids; an id named while its type's switch already says the
// same changes nothing
FLAC__bool FLAC__stream_decoder_set_metadata_respond(FLAC__StreamDecoder* d, int type) {
    DecImpl* m = D(d);
    if (m->state != DS_UNINITIALIZED || type < 0 || type > 126) return 0;
    m->meta_filter[type] = true; if (type == 2) m->meta_ids.clear();
    return 1;
}
FLAC__bool FLAC__stream_decoder_set_metadata_respond_application(FLAC__StreamDecoder* d, const FLAC__byte* id) {
    DecImpl* m = D(d);
    if (m->state != DS_UNINITIALIZED || !id) return 0;
    if (!m->meta_filter[2]) m->meta_ids.push_back(be32(id));
    return 1;
}
FLAC__bool FLAC__stream_decoder_set_metadata_respond_all(FLAC__StreamDecoder* d) {
    DecImpl* m = D(d);
    if (m->state != DS_UNINITIALIZED) return 0;
    for (bool& f : m->meta_filter) f = true;
    m->meta_ids.clear();
    return 1;
}
FLAC__bool FLAC__stream_decoder_set_metadata_ignore(FLAC__StreamDecoder* d, int type) {
    DecImpl* m = D(d);
    if (m->state != DS_UNINITIALIZED || type < 0 || type > 126) return 0;
    m->meta_filter[type] = false; if (type == 2) m->meta_ids.clear();
    return 1;
}
FLAC__bool FLAC__stream_decoder_set_metadata_ignore_application(FLAC__StreamDecoder* d, const FLAC__byte* id) {
    DecImpl* m = D(d);
    if (m->state != DS_UNINITIALIZED || !id) return 0;
    if (m->meta_filter[2]) m->meta_ids.push_back(be32(id));
    return 1;
}
FLAC__bool FLAC__stream_decoder_set_metadata_ignore_all(FLAC__StreamDecoder* d) {
    DecImpl* m = D(d);
    if (m->state != DS_UNINITIALIZED) return 0;
    for (bool& f : m->meta_filter) f = false;
    m->meta_ids.clear();
    return 1;
}
int FLAC__stream_decoder_get_state(const FLAC__StreamDecoder* d) { return D(d)->state; }
const char* FLAC__stream_decoder_get_resolved_state_string(const FLAC__StreamDecoder* d) { return FLAC__StreamDecoderStateString[D(d)->state]; }
FLAC__bool FLAC__stream_decoder_get_md5_checking(const FLAC__StreamDecoder* d) { return D(d)->md5_checking; }
FLAC__uint64 FLAC__stream_decoder_get_total_samples(const FLAC__StreamDecoder* d) { return D(d)->total_samples; }
uint32_t FLAC__stream_decoder_get_channels(const FLAC__StreamDecoder* d) { return D(d)->hdr_channels; }
int FLAC__stream_decoder_get_channel_assignment(const FLAC__StreamDecoder*) { return 0; }
uint32_t FLAC__stream_decoder_get_bits_per_sample(const FLAC__StreamDecoder* d) { return D(d)->hdr_bps; }
uint32_t FLAC__stream_decoder_get_sample_rate(const FLAC__StreamDecoder* d) { return D(d)->hdr_sample_rate; }
uint32_t FLAC__stream_decoder_get_blocksize(const FLAC__StreamDecoder* d) { return D(d)->hdr_blocksize; }
// stream_decoder.h:1083-1099: needs a tell callback (FILE input always has one)
FLAC__bool FLAC__stream_decoder_get_decode_position(const FLAC__StreamDecoder* d, FLAC__uint64* p) { const DecImpl* m = D(d); if (!p || (!m->file && !m->tell_cb)) return 0; *p = m->bytes_consumed + (m->metadata_done ? 0 : m->find_i == 4 ? m->meta_pos : m->find_scan);   /* (inside the head of the stream: what the marker search and the blocks have taken) */ return 1; }

static int init_common(FLAC__StreamDecoder* d) {
    DecImpl* m = D(d);
    m->in.clear(); m->eof = false; m->metadata_done = false; m->ready.clear(); m->frame_index = 0; m->bytes_consumed = 0;
    m->sample_rate = m->channels = m->bps = m->blocksize = 0; m->total_samples = 0; m->min_blocksize = 0;
    m->have_last = false; m->next_sample = 0; m->last_blocksize = 0; m->first_frame_offset = 0;
    m->meta_reset(); m->is_seeking = false;
    m->hdr_channels = m->hdr_bps = m->hdr_sample_rate = m->hdr_blocksize = 0;
    m->md5_active = m->md5_checking != 0; m->md5.init(); memset(m->stored_md5, 0, 16);
    {
        std::lock_guard<std::mutex> lk(g_dec_mu);
        if (!dec_ctx()) { m->state = DS_MEMORY_ALLOCATION_ERROR; return DI_MEMORY_ALLOCATION_ERROR; }     // no CUDA device: fail loudly, no CPU fallback
    }
    m->state = DS_SEARCH_FOR_METADATA;
    return DI_OK;
}
int FLAC__stream_decoder_init_stream(FLAC__StreamDecoder* d, FLAC__StreamDecoderReadCallback r, FLAC__StreamDecoderSeekCallback s, FLAC__StreamDecoderTellCallback t,
                                     FLAC__StreamDecoderLengthCallback l, FLAC__StreamDecoderEofCallback e, FLAC__StreamDecoderWriteCallback w,
                                     FLAC__StreamDecoderMetadataCallback mcb, FLAC__StreamDecoderErrorCallback ecb, void* client) {
    DecImpl* m = D(d);
    if (m->state != DS_UNINITIALIZED) return DI_ALREADY_INITIALIZED;
    if (!r || !w || !ecb || (s && (!t || !l || !e))) return DI_INVALID_CALLBACKS;
    m->read_cb = r; m->write_cb = w; m->error_cb = ecb; m->meta_cb = mcb; m->client = client; m->file = nullptr;
    m->seek_cb = s; m->tell_cb = t; m->length_cb = l; m->eof_cb = e;
    return init_common(d);
}
int FLAC__stream_decoder_init_FILE(FLAC__StreamDecoder* d, FILE* f, FLAC__StreamDecoderWriteCallback w, FLAC__StreamDecoderMetadataCallback mcb,
                                   FLAC__StreamDecoderErrorCallback ecb, void* client) {
    DecImpl* m = D(d);
    if (m->state != DS_UNINITIALIZED) return DI_ALREADY_INITIALIZED;
    if (!f || !w || !ecb) return DI_INVALID_CALLBACKS;
    m->read_cb = nullptr; m->write_cb = w; m->error_cb = ecb; m->meta_cb = mcb; m->client = client; m->file = f;
    m->seek_cb = nullptr; m->tell_cb = nullptr; m->length_cb = nullptr; m->eof_cb = nullptr;
    const int rc = init_common(d);
    // a failed init leaves the FILE with the caller (init_file closes the one it opened): finish()/delete must not close it again
    if (rc != DI_OK) { m->file = nullptr; m->state = DS_UNINITIALIZED; }
    return rc;
}
int FLAC__stream_decoder_init_file(FLAC__StreamDecoder* d, const char* filename, FLAC__StreamDecoderWriteCallback w, FLAC__StreamDecoderMetadataCallback mcb,
                                   FLAC__StreamDecoderErrorCallback ecb, void* client) {
    DecImpl* m = D(d);
    if (m->state != DS_UNINITIALIZED) return DI_ALREADY_INITIALIZED;
    if (!w || !ecb) return DI_INVALID_CALLBACKS;
    FILE* f = filename ? fopen(filename, "rb") : stdin;
    if (!f) return DI_ERROR_OPENING_FILE;                     // tests/test_decoder.py:107-111
    const int rc = FLAC__stream_decoder_init_FILE(d, f, w, mcb, ecb, client);
    if (rc != DI_OK && f != stdin) fclose(f);
    return rc;
}
int FLAC__stream_decoder_init_ogg_stream(FLAC__StreamDecoder*, FLAC__StreamDecoderReadCallback, FLAC__StreamDecoderSeekCallback, FLAC__StreamDecoderTellCallback,
                                         FLAC__StreamDecoderLengthCallback, FLAC__StreamDecoderEofCallback, FLAC__StreamDecoderWriteCallback,
                                         FLAC__StreamDecoderMetadataCallback, FLAC__StreamDecoderErrorCallback, void*) { return DI_UNSUPPORTED_CONTAINER; }
int FLAC__stream_decoder_init_ogg_FILE(FLAC__StreamDecoder*, FILE*, FLAC__StreamDecoderWriteCallback, FLAC__StreamDecoderMetadataCallback, FLAC__StreamDecoderErrorCallback, void*) { return DI_UNSUPPORTED_CONTAINER; }
int FLAC__stream_decoder_init_ogg_file(FLAC__StreamDecoder*, const char*, FLAC__StreamDecoderWriteCallback, FLAC__StreamDecoderMetadataCallback, FLAC__StreamDecoderErrorCallback, void*) { return DI_UNSUPPORTED_CONTAINER; }

FLAC__bool FLAC__stream_decoder_finish(FLAC__StreamDecoder* d) {
    DecImpl* m = D(d);
    if (m->state == DS_UNINITIALIZED) return 1;
    // stream_decoder.h:1145-1150: false when MD5 checking is on and the MD5 of what was delivered differs from STREAMINFO's
    bool md5_failed = false;
    if (m->md5_active) { uint8_t got[16]; m->md5.final(got); md5_failed = memcmp(got, m->stored_md5, 16) != 0; }
    if (m->file && m->file != stdin) fclose(m->file);
    m->file = nullptr; m->in.clear(); m->ready.clear();
    m->read_cb = nullptr; m->write_cb = nullptr; m->error_cb = nullptr; m->meta_cb = nullptr; m->client = nullptr;   // stream_decoder.h: finish resets the callbacks too
    m->seek_cb = nullptr; m->tell_cb = nullptr; m->length_cb = nullptr; m->eof_cb = nullptr;
    m->md5_checking = 0; m->md5_active = false;
    m->filter_defaults();                                              // (finish puts every setting back to its default)
    m->state = DS_UNINITIALIZED;
    return md5_failed ? 0 : 1;
}
static bool input_seek(FLAC__StreamDecoder* d, uint64_t off) {
    DecImpl* m = D(d);
    if (m->file) return m->file != stdin && fseeko(m->file, (off_t)off, SEEK_SET) == 0;
    return m->seek_cb && m->seek_cb(d, off, m->client) == 0;
}
static bool input_length(FLAC__StreamDecoder* d, uint64_t* len) {
    DecImpl* m = D(d);
    if (m->file) { struct stat st; if (m->file == stdin || fstat(fileno(m->file), &st) != 0 || st.st_size <= 0) return false; *len = (uint64_t)st.st_size; return true; }
    return m->length_cb && m->length_cb(d, len, m->client) == 0;
}
// stream_decoder.h:1370-1400: flush drops the buffered input and turns MD5 checking off
FLAC__bool FLAC__stream_decoder_flush(FLAC__StreamDecoder* d) {
    DecImpl* m = D(d);
    if (m->state == DS_UNINITIALIZED) return 0;
    if (!m->metadata_done && m->find_i == 4) { m->metadata_done = true; m->meta_pos = 0; m->meta_skip = false; }   // inside the blocks: the ones not read yet are dropped with the input
    if (m->metadata_done) m->bytes_consumed += m->in.size();           // the decode position moves behind the input that is dropped
    else m->find_reset();                                              // (still looking for the marker: nothing of the dropped input is looked at again)
    m->in.clear(); m->ready.clear(); m->md5_active = false; m->have_last = false;
    m->state = DS_SEARCH_FOR_FRAME_SYNC;
    return 1;
}
// stream_decoder.h:1402-1437: reset = flush + back to the start of a seekable input (stdin cannot rewind: false), metadata is
// read again and MD5 checking starts over
FLAC__bool FLAC__stream_decoder_reset(FLAC__StreamDecoder* d) {
    DecImpl* m = D(d);
    if (m->state == DS_UNINITIALIZED) return 0;
    FLAC__stream_decoder_flush(d);
    if (m->file) { if (m->file == stdin) return 0; if (fseeko(m->file, 0, SEEK_SET) != 0) return 0; }
    else if (m->seek_cb && m->seek_cb(d, 0, m->client) == 1) return 0;          // seekable and the seek fails: reset fails
    m->metadata_done = false; m->eof = false; m->frame_index = 0; m->bytes_consumed = 0; m->next_sample = 0; m->last_blocksize = 0;
    m->meta_reset();
    m->total_samples = 0;                                              // (FLAC__stream_decoder_get_total_samples: 0 until STREAMINFO has been read again)
    m->md5_active = m->md5_checking != 0; m->md5.init();
    m->state = DS_SEARCH_FOR_METADATA;
    return 1;
}

FLAC__bool FLAC__stream_decoder_process_single(FLAC__StreamDecoder* d) {
    DecImpl* m = D(d);
    if (m->state == DS_UNINITIALIZED || m->state == DS_MEMORY_ALLOCATION_ERROR) return 0;
    if (m->state == DS_ABORTED) return 1;                               // (only the call in which the client aborted fails; as END_OF_STREAM, libFLAC)
    if (m->state == DS_END_OF_STREAM) return 1;
    return step(d, false) >= 0;
}
FLAC__bool FLAC__stream_decoder_process_until_end_of_metadata(FLAC__StreamDecoder* d) {
    DecImpl* m = D(d);
    if (m->state == DS_UNINITIALIZED || m->state == DS_MEMORY_ALLOCATION_ERROR) return 0;
    if (m->state == DS_ABORTED) return 1;                               // (only the call in which the client aborted fails; as END_OF_STREAM, libFLAC)
    while (!m->metadata_done && !m->meta_skip && m->state != DS_END_OF_STREAM) if (step(d, false) < 0) return 0;
    return 1;
}
FLAC__bool FLAC__stream_decoder_process_until_end_of_stream(FLAC__StreamDecoder* d) {
    DecImpl* m = D(d);
    if (m->state == DS_UNINITIALIZED || m->state == DS_MEMORY_ALLOCATION_ERROR) return 0;
    if (m->state == DS_ABORTED) return 1;                               // (only the call in which the client aborted fails; as END_OF_STREAM, libFLAC)
    for (;;) {
        const int r = step(d, true);
        if (r < 0) return 0;
        if (r == 0) return 1;
    }
}
FLAC__bool FLAC__stream_decoder_skip_single_frame(FLAC__StreamDecoder* d) {
    DecImpl* m = D(d);
    if (m->state == DS_UNINITIALIZED || m->state == DS_MEMORY_ALLOCATION_ERROR) return 0;
    if (m->state == DS_ABORTED) return 1;                               // (only the call in which the client aborted fails; as END_OF_STREAM, libFLAC)
    FLAC__StreamDecoderWriteCallback keep = m->write_cb;
    m->write_cb = [](const FLAC__StreamDecoder*, const FLAC__Frame*, const FLAC__int32* const*, void*) -> int { return 0; };
    const int r = step(d, false);
    m->write_cb = keep;
    return r >= 0;
}
// stream_decoder.h:1519-1538 (builder/decoder.py:475).  libFLAC bisects the byte range with single-frame reads; here each probe
// reads a window of the input at an estimated byte position and decodes every frame inside it in one GPU batch (headerless mode
// finds the first frame whose header CRC and frame CRC hold, and the frame headers give the stream position), so one or two probes
// decide.  As in libFLAC the frame that holds the target sample is delivered through the write callback before the call returns,
// cut to start at the target; the following process_* calls continue behind it.  Errors met while probing are not reported
// (libFLAC: is_seeking), MD5 checking is off afterwards, and a failed search leaves the decoder in SEEK_ERROR until flush / reset.
FLAC__bool FLAC__stream_decoder_seek_absolute(FLAC__StreamDecoder* d, FLAC__uint64 sample) {
    DecImpl* m = D(d);
    if (m->state != DS_SEARCH_FOR_METADATA && m->state != DS_READ_METADATA && m->state != DS_SEARCH_FOR_FRAME_SYNC &&
        m->state != DS_READ_FRAME && m->state != DS_END_OF_STREAM) return 0;
    if (m->file ? m->file == stdin : !(m->seek_cb && m->tell_cb && m->length_cb && m->eof_cb)) return 0;   // not seekable
    if (m->total_samples > 0 && sample >= m->total_samples) return 0;
    m->md5_active = false;
    uint64_t length = 0;
    if (!input_length(d, &length)) return 0;
    if (!m->metadata_done) {
        m->is_seeking = true;                                          // metadata read on behalf of a seek is not reported
        const bool ok = FLAC__stream_decoder_process_until_end_of_metadata(d) && m->metadata_done;
        m->is_seeking = false;
        if (!ok) return 0;
        if (m->total_samples > 0 && sample >= m->total_samples) return 0;
    }
    uint64_t lo = m->first_frame_offset, hi = length;                  // the target frame starts in [lo, hi)
    uint64_t lo_s = 0, hi_s = m->total_samples;                        // stream positions at those bytes (hi_s == 0: unknown)
    size_t window = 1u << 18;
    for (int iter = 0; iter < 64 && lo < hi; iter++) {
        uint64_t pos = lo;
        if (hi - lo > window) {
            if (hi_s > lo_s && sample >= lo_s && iter < 6) {           // proportional guess, a quarter window early
                const uint64_t g = lo + (uint64_t)((long double)(sample - lo_s) / (long double)(hi_s - lo_s) * (long double)(hi - lo));
                pos = g > lo + window / 4 ? g - window / 4 : lo;
            } else pos = lo + (hi - lo) / 2;
            if (pos >= hi) pos = hi - 1;
        }
        if (!input_seek(d, pos)) break;
        m->in.clear(); m->ready.clear(); m->eof = false; m->have_last = false; m->next_sample = 0; m->last_blocksize = 0;
        m->bytes_consumed = pos;
        while (!m->eof && m->in.size() < window) if (!pull(d, window - m->in.size())) return 0;
        const bool hit_end = m->eof;
        if (!decode_buffered(d)) return 0;
        uint64_t first_s = 0, end_s = 0; bool any = false;
        for (const PendingFrame& pf : m->ready) if (pf.error < 0) { if (!any) first_s = pf.first_sample; any = true; end_s = pf.first_sample + pf.blocksize; }
        if (!any) {                                                    // no whole frame inside the window
            if (!hit_end && window < (64u << 20)) { window *= 4; continue; }
            if (pos == lo) break;
            hi = pos; continue;
        }
        if (sample < first_s) { if (pos == lo) break; hi = pos; hi_s = first_s; continue; }      // pos == lo: the target was never coded (missing frames)
        if (sample >= end_s) { if (hit_end) break; lo = m->bytes_consumed; lo_s = end_s; continue; }
        while (!m->ready.empty()) {
            PendingFrame& pf = m->ready.front();
            if (pf.error < 0 && sample < pf.first_sample + pf.blocksize) break;
            m->ready.pop_front();
        }
        PendingFrame& pf = m->ready.front();
        const uint32_t delta = (uint32_t)(sample - pf.first_sample);
        if (delta) {
            const uint32_t n = pf.blocksize - delta;
            std::vector<int32_t> cut((size_t)n * m->channels);
            for (uint32_t c = 0; c < m->channels; c++) memcpy(&cut[(size_t)c * n], &pf.planar[(size_t)c * pf.blocksize + delta], (size_t)n * 4);
            pf.planar.swap(cut); pf.blocksize = n; pf.first_sample = sample;
        }
        return deliver_one(d) ? 1 : 0;
    }
    m->in.clear(); m->ready.clear();
    m->state = DS_SEEK_ERROR;
    return 0;
}

}  // extern "C"
